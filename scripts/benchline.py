import json,sys
d=json.loads(sys.stdin.readline()); r=d["roofline"]
print("%s value %.0f step %.4f scan %.4f resolve %.4f frac %.4f launches %d" % (sys.argv[1] if len(sys.argv)>1 else "", d["value"], d["ms_per_step"], r["kernel_ms_per_step"], r["resolve_kernels_ms_per_step"], r["frac"], d["gpu_launches"]))
