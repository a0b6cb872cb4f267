"""Mirror of the reference's `demod_2400` module (src/demod_2400.rs)."""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum

from . import _ffi
from .context import default_context
from .utils import MagnitudeBuffer


class MsgLen(Enum):          # src/demod_2400.rs:87-91
    Short = _ffi.MODES_SHORT_MSG_BYTES
    Long = _ffi.MODES_LONG_MSG_BYTES


@dataclass
class ModeSMessage:          # src/demod_2400.rs:93-112
    msglen: MsgLen
    msg: bytes
    score: int
    phase: int
    j: int
    buffer_index: int = 0

    def buffer(self) -> bytes:
        return self.msg[: self.msglen.value]


def _wrap(frames) -> list[ModeSMessage]:
    return [ModeSMessage(MsgLen.Long if f["len"] == 14 else MsgLen.Short, f["msg"], f["score"],
                         f["phase"], f["j"], f["buffer"]) for f in frames]


def demodulate2400(mag: MagnitudeBuffer, ctx=None) -> list[ModeSMessage]:
    """demod_2400::demodulate2400 (src/demod_2400.rs:115-212).  The reference returns
    Result<Vec<_>, &str> that is always Ok (:211); errors here raise."""
    return _wrap((ctx or default_context()).demodulate2400(mag.data, mag.length))


def demod_iq(iq, ctx=None) -> list[ModeSMessage]:
    """to_mag + demodulate2400 fused on the GPU (what main.rs:166-167 does per buffer)."""
    return _wrap((ctx or default_context()).demod_iq(iq))


def demod_iq_batch(iq, n_buffers: int, samples_per_buffer: int, ctx=None, **kw) -> list[ModeSMessage]:
    """A run of consecutive buffers of one stream, identical to n_buffers sequential calls."""
    return _wrap((ctx or default_context()).demod_iq_batch(iq, n_buffers, samples_per_buffer, **kw))
