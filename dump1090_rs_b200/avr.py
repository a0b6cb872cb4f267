"""AVR text output of the reference binary (dump1090_rs/src/main.rs:174-176): one
`*{hex};\\n` line per frame, what adsb_deku-style clients read from TCP port 30002.
Formatting only; the TCP listener is out of scope."""
from __future__ import annotations

import ctypes as C

from . import _ffi


def format_frames(frames) -> bytes:
    """frames: list of dicts / ModeSMessage-like objects exposing the visible bytes."""
    arr = (_ffi.Frame * max(len(frames), 1))()
    for i, f in enumerate(frames):
        msg = f["msg"] if isinstance(f, dict) else f.buffer()
        arr[i].len = len(msg)
        for k, b in enumerate(msg):
            arr[i].msg[k] = b
    need = C.c_size_t(0)
    L = _ffi.lib()
    L.b200adsb_format_avr(arr, len(frames), None, 0, C.byref(need))
    buf = C.create_string_buffer(need.value + 1)
    rc = L.b200adsb_format_avr(arr, len(frames), buf, need.value, C.byref(need))
    if rc != 0:
        raise _ffi.B200AdsbError(rc, "format_avr")
    return buf.raw[: need.value]
