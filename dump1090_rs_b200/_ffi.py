"""ctypes binding of libb200adsb.so (include/b200adsb.h).

The library is the product: if it is missing or CUDA is unavailable every call
raises -- there is no CPU path behind this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
SO_PATH = os.path.join(_HERE, "libb200adsb.so")
CSRC = os.path.join(_HERE, "csrc")

MODES_MAG_BUF_SAMPLES = 131072  # src/lib.rs:22
TRAILING_SAMPLES = 326          # src/lib.rs:24
MAG_DATA_LEN = TRAILING_SAMPLES + MODES_MAG_BUF_SAMPLES
MODES_LONG_MSG_BYTES = 14       # src/lib.rs:25
MODES_SHORT_MSG_BYTES = 7       # src/lib.rs:26
ICAO_FILTER_ADSB_NT = 1 << 25   # src/icao_filter.rs:6

OK, ERR_BAD_ARG, ERR_CAPACITY, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_EVENTS = 0, -1, -2, -3, -4, -5, -6
OPT_TILE, OPT_POOL_SHIFT, OPT_PROFILE, OPT_H2D_CHUNK, OPT_CARRY = 1, 2, 3, 4, 5

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
    "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
]


class Frame(C.Structure):
    _fields_ = [
        ("msg", C.c_uint8 * 14),
        ("len", C.c_uint8),
        ("phase", C.c_uint8),
        ("score", C.c_int16),
        ("reserved", C.c_uint16),
        ("j", C.c_uint32),
        ("buffer", C.c_uint32),
    ]


assert C.sizeof(Frame) == 28


class Timing(C.Structure):
    _fields_ = [
        ("scan_ms", C.c_double), ("resolve_ms", C.c_double), ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double), ("scan_launches", C.c_uint64), ("other_launches", C.c_uint64),
        ("samples", C.c_uint64), ("candidates", C.c_uint64),
    ]


class B200AdsbError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__(f"{where}: status {status} ({detail})")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/b200adsb.cu for sm_100a into the in-tree shared library."""
    import glob
    hdr = os.path.join(_REPO, "include", "b200adsb.h")
    deps = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))) + [hdr]
    stale = (not os.path.exists(SO_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(SO_PATH) for p in deps if os.path.exists(p))
    if force or stale:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-o", SO_PATH, os.path.join(CSRC, "b200adsb.cu")]
        subprocess.check_call(cmd)
    return SO_PATH


_lib = None

_PROTOS = {
    "b200adsb_version": (C.c_int, []),
    "b200adsb_strerror": (C.c_char_p, [C.c_int]),
    "b200adsb_last_error": (C.c_char_p, [C.c_void_p]),
    "b200adsb_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "b200adsb_ctx_destroy": (None, [C.c_void_p]),
    "b200adsb_ctx_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "b200adsb_ctx_sync": (C.c_int, [C.c_void_p]),
    "b200adsb_timing_get": (C.c_int, [C.c_void_p, C.POINTER(Timing), C.c_int]),
    "b200adsb_host_alloc": (C.c_void_p, [C.c_size_t]),
    "b200adsb_host_free": (None, [C.c_void_p]),
    "b200adsb_to_mag": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]),
    "b200adsb_demodulate2400": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_size_t)]),
    "b200adsb_demod_iq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t)]),
    "b200adsb_demod_iq_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                          C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                          C.c_void_p]),
    "b200adsb_demod_cu8_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                           C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                           C.c_void_p]),
    "b200adsb_cu8_to_cs16": (C.c_int16, [C.c_uint8]),
    "b200adsb_demod_iq_batch_dev_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200adsb_demod_iq_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                              C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                              C.c_void_p]),
    "b200adsb_demod_iq_batch_submit": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                 C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200adsb_demod_iq_batch_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "b200adsb_scan_batch_dev_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                C.c_void_p, C.c_uint64, C.c_uint64]),
    "b200adsb_resolve_batch_dev_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200adsb_scan_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                          C.c_void_p, C.c_uint64, C.c_uint64]),
    "b200adsb_events_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "b200adsb_events_export_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200adsb_events_import_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "b200adsb_events_pack_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "b200adsb_events_import_packed_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]),
    "b200adsb_events_symm_words": (C.c_size_t, [C.c_size_t, C.c_size_t]),
    "b200adsb_events_push_symm_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64,
                                                C.c_uint32]),
    "b200adsb_events_import_symm_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                  C.c_uint64]),
    "b200adsb_frames_pack_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                           C.c_size_t]),
    "b200adsb_frames_merge_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_size_t, C.c_void_p]),
    "b200adsb_frames_symm_bytes": (C.c_size_t, [C.c_size_t, C.c_size_t]),
    "b200adsb_frames_push_symm_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                                C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, C.c_void_p]),
    "b200adsb_frames_merge_symm_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64,
                                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200adsb_resolve_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                             C.c_void_p]),
    "b200adsb_icao_flush": (C.c_int, [C.c_void_p]),
    "b200adsb_icao_hash": (C.c_uint32, [C.c_uint32]),
    "b200adsb_icao_filter_add": (C.c_int, [C.c_void_p, C.c_uint32]),
    "b200adsb_icao_filter_test": (C.c_int, [C.c_void_p, C.c_uint32]),
    "b200adsb_icao_snapshot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200adsb_icao_restore": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "b200adsb_modes_checksum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "b200adsb_score_modes_messages": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "b200adsb_format_avr": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200adsb_modes_checksum_one": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32)]),
    "b200adsb_score_modes_message": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200adsb_getbits": (C.c_uint32, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "b200adsb_async_acknowledge": (C.c_int, [C.c_void_p]),
    "b200adsb_debug_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200adsb_debug_crc_tabs": (C.c_int, [C.c_void_p]),
    "b200adsb_debug_crc_lane_tabs": (C.c_int, [C.c_void_p]),
    "b200adsb_debug_mag_sweep": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
}

EXPORTS = tuple(_PROTOS)


def lib():
    """Load libb200adsb.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        so = os.environ.get("B200ADSB_LIB", SO_PATH)   # A/B experiments load a variant build
        if not os.path.exists(so):
            raise ImportError(
                f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the CUDA library is the product; there is no CPU fallback)")
        L = C.CDLL(so)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
