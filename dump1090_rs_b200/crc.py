"""Mirror of the reference's `crc` module (src/crc.rs)."""
import numpy as np

from .context import default_context


def modes_checksum(message: bytes, bits: int, ctx=None) -> int:   # src/crc.rs:263-282
    n = bits // 8
    assert n >= 3                                                  # crc.rs:267
    m = np.zeros(14, dtype=np.uint8)
    m[:n] = np.frombuffer(bytes(message[:n]), dtype=np.uint8)
    return int((ctx or default_context()).modes_checksum(m, bits)[0])


def modes_checksum_batch(msgs, bits: int, ctx=None) -> np.ndarray:
    return (ctx or default_context()).modes_checksum(msgs, bits)
