"""Mirror of the reference's `mode_s` module (src/mode_s/mod.rs)."""
import numpy as np

from .context import default_context
from .demod_2400 import MsgLen


def getbits(data: bytes, firstbit_1idx: int, lastbit_1idx: int) -> int:   # src/mode_s/mod.rs:14-30
    from . import _ffi
    buf = bytes(data)
    if lastbit_1idx > 8 * len(buf) or firstbit_1idx < 1:
        raise IndexError("index out of bounds")                          # the reference panics
    return int(_ffi.lib().b200adsb_getbits(buf, firstbit_1idx, lastbit_1idx))


def score_modes_message(msg: bytes, ctx=None):
    """score_modes_message (src/mode_s/mod.rs:34-139): Some((MsgLen, score)) or None.
    Always given the 14-byte buffer by its only caller (demod_2400.rs:156,184)."""
    return (ctx or default_context()).score_modes_message(bytes(msg))


def score_modes_messages(msgs, ctx=None):
    """Batched form: n 14-byte messages scored in order against the context filter."""
    return (ctx or default_context()).score_modes_messages(msgs)
