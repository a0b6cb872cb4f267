import ctypes as C, os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import dump1090_rs_b200 as d
from dump1090_rs_b200 import _ffi
z = np.load("tests/golden/captures.npz"); iq = np.ascontiguousarray(z[z.files[0]].reshape(-1,2)[:, ::-1])
L = _ffi.lib()
p = L.b200adsb_host_alloc(iq.nbytes); C.memmove(p, iq.ctypes.data, iq.nbytes)
frames = (_ffi.Frame*4096)(); n = C.c_size_t(0)
for T in (0, 472, 856, 1240, 1624, 2008, 3544, 7384):
    ctx = d.Context(0); ctx.set_option(_ffi.OPT_TILE, T); ctx.set_option(_ffi.OPT_PROFILE, 1)
    def fn():
        L.b200adsb_icao_flush(ctx._h); L.b200adsb_demod_iq(ctx._h, p, 131072, frames, 4096, C.byref(n))
    for _ in range(20): fn()
    ctx.timing(reset=True)
    ts=[]
    for _ in range(200):
        t0=time.perf_counter(); fn(); ts.append(time.perf_counter()-t0)
    ts.sort(); tm = ctx.timing()
    print(f"T={T}: median {1e6*ts[100]:.1f} us  scan {1e3*tm['scan_ms']/200:.1f} us  resolve-stage {1e3*tm['resolve_ms']/200:.1f} us frames {n.value}")
    ctx.close()
