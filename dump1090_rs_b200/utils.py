"""Mirror of the reference's `utils` module (src/utils.rs)."""
from __future__ import annotations

import time
from dataclasses import dataclass

import numpy as np

from . import _ffi
from .context import default_context


@dataclass
class MagnitudeBuffer:
    """src/lib.rs:29-51: data[326 + 131072] u16, length; samples live at data[326..326+length]."""
    data: np.ndarray
    length: int
    first_sample_timestamp_12mhz: int = 0


def to_mag(data, ctx=None) -> MagnitudeBuffer:
    """utils::to_mag (src/utils.rs:43-58) on the GPU.  `data`: Complex<i16> samples as an
    int16 array of (re, im) pairs in memory order, or a complex array."""
    d, length = (ctx or default_context()).to_mag(data)
    return MagnitudeBuffer(d, length)


def read_test_data(filepath: str) -> np.ndarray:
    """utils::read_test_data (src/utils.rs:23-40): the capture files store im first, then
    re, little endian; returns [0x20000, 2] int16 in memory order (re, im)."""
    raw = np.fromfile(filepath, dtype="<i2", count=2 * 0x20000)
    if raw.size != 2 * 0x20000:
        raise EOFError("capture shorter than 0x20000 samples (the reference unwraps a read error)")
    return np.ascontiguousarray(raw.reshape(-1, 2)[:, ::-1])


def save_test_data(data, name: str | None = None) -> str:
    """utils::save_test_data (src/utils.rs:8-21): writes im then re, little endian, to
    test_<unix ms>.iq."""
    a = np.ascontiguousarray(data, dtype=np.int16).reshape(-1, 2)
    name = name or f"test_{int(time.time() * 1000)}.iq"
    np.ascontiguousarray(a[:, ::-1]).astype("<i2").tofile(name)
    return name
