"""dump1090_rs_b200 -- B200-native (sm_100a CUDA) drop-in for the dump1090_rs hot path:
CS16 IQ -> magnitude -> preamble gate -> 5-phase PPM slicing -> Mode-S CRC-24 -> frames.

Module names mirror the reference crate (libdump1090_rs): utils, demod_2400,
icao_filter, crc, mode_s.  Everything computes on the GPU through libb200adsb.so
(include/b200adsb.h); there is no CPU fallback.
"""
from . import _ffi
from ._ffi import B200AdsbError
from .context import Context, default_context

MODES_MAG_BUF_SAMPLES = _ffi.MODES_MAG_BUF_SAMPLES
MODES_LONG_MSG_BYTES = _ffi.MODES_LONG_MSG_BYTES
MODES_SHORT_MSG_BYTES = _ffi.MODES_SHORT_MSG_BYTES

from . import crc, demod_2400, icao_filter, mode_s, utils  # noqa: E402

__all__ = ["Context", "default_context", "B200AdsbError", "utils", "demod_2400", "icao_filter", "crc",
           "mode_s", "MODES_MAG_BUF_SAMPLES", "MODES_LONG_MSG_BYTES", "MODES_SHORT_MSG_BYTES"]
