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
