#!/usr/bin/env python3
"""Per-call latency on ONE 512 KiB buffer (BASELINE configs[0]/[1]: the reference's own bench
case, benches/demod_benchmark.rs: icao_flush + to_mag + demodulate2400 on a capture), through
the host-pointer C ABI, compared with the oracle on one host core.  Run under gpurun."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import dump1090_rs_b200 as d          # noqa: E402
from dump1090_rs_b200 import _ffi     # noqa: E402
from oracle import oracle as O        # noqa: E402

z = np.load(os.path.join(REPO, "tests", "golden", "captures.npz"))
out = {}
ctx = d.Context(0)
L = _ffi.lib()
for name in z.files:
    iq = np.ascontiguousarray(z[name].reshape(-1, 2)[:, ::-1])
    # pinned host copy, as a caller that cares about latency would hold it
    p = L.b200adsb_host_alloc(iq.nbytes)
    C.memmove(p, iq.ctypes.data, iq.nbytes)
    frames = (_ffi.Frame * 4096)()
    n = C.c_size_t(0)

    def routine_fused():
        L.b200adsb_icao_flush(ctx._h)
        L.b200adsb_demod_iq(ctx._h, p, 131072, frames, 4096, C.byref(n))

    def routine_two_step():
        d.icao_filter.icao_flush(ctx)
        mb = d.utils.to_mag(iq, ctx)
        return d.demod_2400.demodulate2400(mb, ctx)

    for fn, key in ((routine_fused, "fused_demod_iq"), (routine_two_step, "to_mag_then_demodulate2400")):
        for _ in range(20):
            fn()
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        out.setdefault(name, {})[key + "_ms_median"] = 1e3 * ts[len(ts) // 2]
    out[name]["frames"] = int(n.value)
    o = O.Oracle()
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        o.demod_iq(iq, flush=True)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    out[name]["oracle_1core_ms_median"] = 1e3 * ts[len(ts) // 2]
    out[name]["speedup_fused_vs_1core"] = out[name]["oracle_1core_ms_median"] / out[name]["fused_demod_iq_ms_median"]
    out[name]["msamples_per_s_fused"] = 131072 / out[name]["fused_demod_iq_ms_median"] / 1e3
    L.b200adsb_host_free(p)
print(json.dumps(out, indent=1))
