"""Mirror of the reference's `crc` module (src/crc.rs)."""
import numpy as np

from .context import default_context


def modes_checksum(message: bytes, bits: int, ctx=None) -> int:   # src/crc.rs:263-282
    assert bits % 8 == 0 and bits // 8 >= 3 and len(message) >= bits // 8   # crc.rs:264-267
    return (ctx or default_context()).modes_checksum_one(bytes(message), bits)


def modes_checksum_batch(msgs, bits: int, ctx=None) -> np.ndarray:
    return (ctx or default_context()).modes_checksum(msgs, bits)
