"""Mirror of the reference's `mode_s` module (src/mode_s/mod.rs)."""
import numpy as np

from .context import default_context
from .demod_2400 import MsgLen


def getbits(data: bytes, firstbit_1idx: int, lastbit_1idx: int) -> int:   # src/mode_s/mod.rs:14-30
    v = int.from_bytes(bytes(data), "big")
    total = 8 * len(data)
    width = lastbit_1idx - firstbit_1idx + 1
    return (v >> (total - lastbit_1idx)) & ((1 << width) - 1)


def score_modes_message(msg: bytes, ctx=None):
    """score_modes_message (src/mode_s/mod.rs:34-139): Some((MsgLen, score)) or None.
    Always given the 14-byte buffer by its only caller (demod_2400.rs:156,184)."""
    if len(msg) * 8 < 56:
        return None
    m = np.zeros(14, dtype=np.uint8)
    m[: min(len(msg), 14)] = np.frombuffer(bytes(msg[:14]), dtype=np.uint8)
    if len(msg) < 14 and (m[0] & 0x80):
        return None                                                       # :48-50
    lens, scores = (ctx or default_context()).score_modes_messages(m)
    if lens[0] == 0:
        return None
    return (MsgLen.Long if lens[0] == 14 else MsgLen.Short, int(scores[0]))


def score_modes_messages(msgs, ctx=None):
    """Batched form: n 14-byte messages scored in order against the context filter."""
    return (ctx or default_context()).score_modes_messages(msgs)
