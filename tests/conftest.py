import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The tests exercise the built product (libb200adsb.so) and the checker (oracle/): in a fresh
    # checkout neither exists yet (built artefacts are not in the history), so build them once.
    # nvcc cross-compiles for sm_100a without a GPU; the package itself never builds on import.
    so = os.path.join(REPO, "dump1090_rs_b200", "libb200adsb.so")
    orc = os.path.join(REPO, "oracle", "libdump1090_oracle.so")
    if not (os.path.exists(so) and os.path.exists(orc)):
        import __graft_entry__ as g
        g.build()
    else:
        from dump1090_rs_b200 import _ffi
        _ffi.build()          # no-op when the library is newer than every csrc/*.cu, *.cuh and the header


@pytest.fixture(scope="session")
def captures():
    """The reference's three test_iq captures as int16 [131072, 2] in memory order (re, im)
    (file order is im, re: utils.rs:29-31)."""
    z = np.load(os.path.join(GOLDEN, "captures.npz"))
    return {k: np.ascontiguousarray(z[k].reshape(-1, 2)[:, ::-1]) for k in z.files}


@pytest.fixture(scope="session")
def golden_frames():
    with open(os.path.join(GOLDEN, "golden_frames.json")) as f:
        return json.load(f)["captures"]


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.lib()
    return O


def frames_key(frames):
    """What parity means for a frame list: order, position, phase, score and visible bytes."""
    return [(f.get("buffer", 0), f["j"], f["phase"], f["score"], f["msg"].hex()) for f in frames]


def oracle_stream(O, bufs, flush_each=False):
    """Reference semantics for a run of buffers of one stream -> frames with buffer index."""
    o = O.Oracle()
    out = []
    for b, iq in enumerate(bufs):
        for f in o.demod_iq(iq, flush=flush_each):
            f["buffer"] = b
            out.append(f)
    return out, o
