"""ctypes binding of the CPU oracle (oracle/dump1090_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package
(dump1090_rs_b200), which must fail loudly when its CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdump1090_oracle.so")

MAG_BUF_SAMPLES = 131072
TRAILING_SAMPLES = 326
MAG_DATA_LEN = TRAILING_SAMPLES + MAG_BUF_SAMPLES

K_NONE, K_PAR_SHORT, K_DF11_IID0, K_DF11_IID, K_DF17, K_DF18, K_PAR_LONG = range(7)
ADSB_NT = 1 << 25


class MagBuf(C.Structure):
    _fields_ = [("data", C.c_uint16 * MAG_DATA_LEN), ("length", C.c_size_t)]


class Filter(C.Structure):
    _fields_ = [("a", C.c_uint32 * 4096), ("b", C.c_uint32 * 4096), ("full_events", C.c_uint64)]


class Frame(C.Structure):
    _fields_ = [
        ("msg", C.c_uint8 * 14),
        ("len", C.c_uint8),
        ("phase", C.c_uint8),
        ("score", C.c_int32),
        ("j", C.c_uint32),
        ("signal_level", C.c_double),
    ]


class Record(C.Structure):
    _fields_ = [("j", C.c_uint32), ("w", C.c_uint32 * 5)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dump1090_oracle.c")
    hdr = os.path.join(_HERE, "dump1090_oracle.h")
    exe = os.path.join(_HERE, "orc_bench")
    stale = (not os.path.exists(_SO)) or (not os.path.exists(exe)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_crc_table.restype = C.POINTER(C.c_uint32)
        L.orc_modes_checksum.restype = C.c_uint32
        L.orc_modes_checksum.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_filter_flush.argtypes = [C.POINTER(Filter)]
        L.orc_icao_hash.restype = C.c_uint32
        L.orc_icao_hash.argtypes = [C.c_uint32]
        L.orc_filter_add.argtypes = [C.POINTER(Filter), C.c_uint32]
        L.orc_filter_test.restype = C.c_int
        L.orc_filter_test.argtypes = [C.POINTER(Filter), C.c_uint32]
        L.orc_getbits.restype = C.c_size_t
        L.orc_getbits.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.orc_score_modes_message.restype = C.c_int
        L.orc_score_modes_message.argtypes = [
            C.POINTER(Filter), C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_to_mag.restype = C.c_int
        L.orc_to_mag.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(MagBuf)]
        L.orc_to_mag_carry.restype = C.c_int
        L.orc_to_mag_carry.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(MagBuf)]
        L.orc_mag_one.restype = C.c_uint16
        L.orc_mag_one.argtypes = [C.c_int16, C.c_int16]
        L.orc_check_preamble.restype = C.c_int
        L.orc_check_preamble.argtypes = [
            C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_slice_phase.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.orc_gate.restype = C.c_int
        L.orc_gate.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_demodulate2400.restype = C.c_size_t
        L.orc_demodulate2400.argtypes = [
            C.POINTER(Filter), C.POINTER(MagBuf), C.POINTER(Frame), C.c_size_t]
        L.orc_classify.restype = C.c_uint32
        L.orc_classify.argtypes = [C.c_void_p]
        L.orc_demod_records.restype = C.c_size_t
        L.orc_demod_records.argtypes = [C.POINTER(MagBuf), C.POINTER(Record), C.c_size_t]
        L.orc_routine.restype = C.c_size_t
        L.orc_routine.argtypes = [
            C.POINTER(Filter), C.c_void_p, C.c_size_t, C.POINTER(Frame), C.c_size_t, C.c_int]
        L.orc_bench.restype = C.c_double
        L.orc_bench.argtypes = [
            C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _iq_ptr(iq: np.ndarray):
    """iq: int16 array of (re, im) pairs in memory order, shape [n, 2] or [2n]."""
    a = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1)
    assert a.size % 2 == 0
    return a, a.ctypes.data_as(C.c_void_p), a.size // 2


class Oracle:
    """One independent stream: a filter plus the reference routines."""

    def __init__(self):
        self.L = lib()
        self.filter = Filter()
        self.L.orc_filter_flush(C.byref(self.filter))

    # icao_filter.rs
    def icao_flush(self):
        self.L.orc_filter_flush(C.byref(self.filter))

    def icao_filter_add(self, addr: int):
        self.L.orc_filter_add(C.byref(self.filter), addr)

    def icao_filter_test(self, addr: int) -> bool:
        return bool(self.L.orc_filter_test(C.byref(self.filter), addr))

    def members(self) -> set[int]:
        a = np.ctypeslib.as_array(self.filter.a)
        return set(int(x) for x in a[a != 0])

    # utils.rs
    def to_mag(self, iq: np.ndarray) -> MagBuf:
        a, p, n = _iq_ptr(iq)
        mb = MagBuf()
        rc = self.L.orc_to_mag(p, n, C.byref(mb))
        if rc != 0:
            raise IndexError("to_mag: more than 131072 samples (reference panics, lib.rs:48)")
        return mb

    # demod_2400.rs
    def demodulate2400(self, mb: MagBuf, cap: int = 65536):
        out = (Frame * cap)()
        n = self.L.orc_demodulate2400(C.byref(self.filter), C.byref(mb), out, cap)
        assert n <= cap
        return [
            dict(msg=bytes(out[i].msg[: out[i].len]), j=out[i].j, phase=out[i].phase,
                 score=out[i].score)
            for i in range(n)
        ]

    def demod_iq_carry(self, iq: np.ndarray):
        """Stream continuity (not reference behaviour): the previous buffers' last 326 samples
        fill the leading slots."""
        a, p, n = _iq_ptr(iq)
        tail = getattr(self, "_tail", None)
        if tail is None:
            tail = np.zeros((TRAILING_SAMPLES, 2), dtype=np.int16)
        mb = MagBuf()
        assert self.L.orc_to_mag_carry(tail.ctypes.data_as(C.c_void_p), p, n, C.byref(mb)) == 0
        self._tail = np.ascontiguousarray(np.concatenate([tail, a.reshape(-1, 2)])[-TRAILING_SAMPLES:])
        return self.demodulate2400(mb)

    def demod_iq(self, iq: np.ndarray, flush: bool = False):
        if flush:
            self.icao_flush()
        return self.demodulate2400(self.to_mag(iq))

    def records(self, mb: MagBuf, cap: int = 1 << 17):
        out = (Record * cap)()
        n = self.L.orc_demod_records(C.byref(mb), out, cap)
        assert n <= cap
        return [(out[i].j, [int(out[i].w[k]) for k in range(5)]) for i in range(n)]


def mag_array(mb: MagBuf) -> np.ndarray:
    return np.ctypeslib.as_array(mb.data).copy()


def magbuf_from_array(data: np.ndarray, length: int) -> MagBuf:
    mb = MagBuf()
    d = np.ascontiguousarray(data, dtype=np.uint16)
    assert d.size == MAG_DATA_LEN
    C.memmove(mb.data, d.ctypes.data, d.nbytes)
    mb.length = length
    return mb


def modes_checksum(msg: bytes, bits: int) -> int:
    b = (C.c_uint8 * len(msg)).from_buffer_copy(msg)
    return int(lib().orc_modes_checksum(b, bits))


def crc_table() -> np.ndarray:
    return np.ctypeslib.as_array(lib().orc_crc_table(), shape=(256,)).copy()


def icao_hash(a: int) -> int:
    return int(lib().orc_icao_hash(a))


def bench(iq: np.ndarray, n_buffers: int, spb: int, iters: int, threads: int, flush_each: bool):
    """Times the reference routine in a separate process (oracle/orc_bench): host threads
    inside a Python process that has numpy/torch loaded do not run in parallel in this
    image, a plain C process does.  Returns (seconds, frames)."""
    import json
    import tempfile
    a = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1)
    assert a.size == 2 * n_buffers * spb
    build()
    exe = os.path.join(_HERE, "orc_bench")
    with tempfile.NamedTemporaryFile(suffix=".iq", dir=os.environ.get("TMPDIR", "/tmp")) as f:
        a.tofile(f)
        f.flush()
        out = subprocess.run([exe, f.name, str(n_buffers), str(spb), str(iters), str(threads),
                              str(int(flush_each))], check=True, capture_output=True, text=True).stdout
    r = json.loads(out.strip().splitlines()[-1])
    return float(r["sec"]), int(r["frames"])
