"""Context: one ordered stream of IQ buffers with one ICAO filter on one GPU.

The reference keeps its filter in two process-wide tables (src/icao_filter.rs:8-9);
here that state lives in GPU memory inside a b200adsb_ctx.  `default_context()`
plays the role of the reference's process-wide state for the module-level
functions in utils / demod_2400 / icao_filter / crc / mode_s.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import B200AdsbError, Frame, Timing


def _as_iq(iq) -> np.ndarray:
    """Accepts int16 [n,2] / [2n] in memory order (re, im) or complex arrays."""
    a = np.asarray(iq)
    if np.iscomplexobj(a):
        out = np.empty((a.size, 2), dtype=np.int16)
        out[:, 0] = a.real.astype(np.int16).reshape(-1)
        out[:, 1] = a.imag.astype(np.int16).reshape(-1)
        return out.reshape(-1)
    return np.ascontiguousarray(a, dtype=np.int16).reshape(-1)


class Context:
    def __init__(self, device: int = 0, stream: int | None = None):
        self._L = _ffi.lib()
        h = C.c_void_p()
        rc = self._L.b200adsb_ctx_create(C.byref(h), int(device), C.c_void_p(stream or 0))
        if rc != 0:
            raise B200AdsbError(rc, "b200adsb_ctx_create",
                                "no usable CUDA device: this package has no CPU fallback")
        self._h = h
        self.device = device

    # -------------------------------------------------------------- plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._L.b200adsb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, where: str, allow=()):
        if rc != 0 and rc not in allow:
            raise B200AdsbError(rc, where, (self._L.b200adsb_strerror(rc) or b"").decode() + "; " +
                                (self._L.b200adsb_last_error(self._h) or b"").decode())
        return rc

    def set_option(self, opt: int, value: int):
        self._check(self._L.b200adsb_ctx_set_option(self._h, opt, value), "set_option")

    def sync(self):
        self._check(self._L.b200adsb_ctx_sync(self._h), "sync")

    def timing(self, reset: bool = False) -> dict:
        t = Timing()
        self._check(self._L.b200adsb_timing_get(self._h, C.byref(t), int(reset)), "timing_get")
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    @staticmethod
    def _frames(arr, n) -> list[dict]:
        return [dict(msg=bytes(arr[i].msg[: arr[i].len]), len=int(arr[i].len), phase=int(arr[i].phase),
                     score=int(arr[i].score), j=int(arr[i].j), buffer=int(arr[i].buffer))
                for i in range(n)]

    # -------------------------------------------------------------- utils.rs / demod_2400.rs
    def to_mag(self, iq):
        a = _as_iq(iq)
        n = a.size // 2
        if n > _ffi.MODES_MAG_BUF_SAMPLES:
            raise IndexError("index out of bounds: more than 131072 samples (src/lib.rs:48)")
        data = np.zeros(_ffi.MAG_DATA_LEN, dtype=np.uint16)
        length = C.c_size_t(0)
        self._check(self._L.b200adsb_to_mag(self._h, a.ctypes.data, n, data.ctypes.data, C.byref(length)),
                    "to_mag")
        return data, int(length.value)

    def demodulate2400(self, data: np.ndarray, length: int, cap: int = 4096) -> list[dict]:
        d = np.ascontiguousarray(data, dtype=np.uint16)
        if d.size != _ffi.MAG_DATA_LEN:
            raise ValueError("MagnitudeBuffer.data must have 326+131072 entries")
        while True:
            out = (Frame * cap)()
            n = C.c_size_t(0)
            rc = self._check(self._L.b200adsb_demodulate2400(self._h, d.ctypes.data, int(length), out, cap,
                                                             C.byref(n)), "demodulate2400",
                             allow=(_ffi.ERR_CAPACITY,))
            if rc == 0:
                return self._frames(out, n.value)
            raise B200AdsbError(rc, "demodulate2400", f"{n.value} frames > cap {cap}")

    def demod_iq(self, iq, cap: int = 4096) -> list[dict]:
        a = _as_iq(iq)
        out = (Frame * cap)()
        n = C.c_size_t(0)
        self._check(self._L.b200adsb_demod_iq(self._h, a.ctypes.data, a.size // 2, out, cap, C.byref(n)),
                    "demod_iq")
        return self._frames(out, n.value)

    def demod_iq_batch(self, iq, n_buffers: int, samples_per_buffer: int, stride: int | None = None,
                       lengths=None, cap: int = 1 << 16, want_counts: bool = False):
        a = _as_iq(iq)
        stride = samples_per_buffer if stride is None else stride
        out = (Frame * cap)()
        n = C.c_size_t(0)
        cnt = np.zeros(max(n_buffers, 1), dtype=np.uint32)
        ln = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.uint32)
        self._check(self._L.b200adsb_demod_iq_batch(
            self._h, a.ctypes.data, n_buffers, samples_per_buffer, stride,
            None if ln is None else ln.ctypes.data, out, cap, C.byref(n),
            cnt.ctypes.data if want_counts else None), "demod_iq_batch")
        fr = self._frames(out, n.value)
        return (fr, cnt[:n_buffers]) if want_counts else fr

    def demod_cu8_batch(self, iq_u8, n_buffers: int, samples_per_buffer: int, cap: int = 1 << 16):
        """Opt-in 8-bit ingest (b200adsb_demod_cu8_batch): uint8 (I, Q) pairs, expanded to CS16 on the device."""
        a = np.ascontiguousarray(iq_u8, dtype=np.uint8).reshape(-1)
        out = (Frame * cap)()
        n = C.c_size_t(0)
        self._check(self._L.b200adsb_demod_cu8_batch(self._h, a.ctypes.data, n_buffers, samples_per_buffer,
                                                     samples_per_buffer, None, out, cap, C.byref(n), None),
                    "demod_cu8_batch")
        return self._frames(out, n.value)

    def demod_cu8_batch_ptr(self, iq_ptr: int, n_buffers: int, spb: int, stride: int, out_ptr: int, cap: int) -> int:
        n = C.c_size_t(0)
        self._check(self._L.b200adsb_demod_cu8_batch(self._h, iq_ptr, n_buffers, spb, stride, None, out_ptr, cap,
                                                     C.byref(n), None), "demod_cu8_batch")
        return int(n.value)

    # raw-pointer forms (device memory owned by the caller, e.g. torch tensors)
    def demod_iq_batch_ptr(self, iq_ptr: int, n_buffers: int, spb: int, stride: int, out_ptr: int,
                           cap: int, lengths_ptr: int = 0, counts_ptr: int = 0, host: bool = False) -> int:
        n = C.c_size_t(0)
        fn = self._L.b200adsb_demod_iq_batch if host else self._L.b200adsb_demod_iq_batch_dev
        self._check(fn(self._h, iq_ptr, n_buffers, spb, stride, lengths_ptr or None, out_ptr, cap,
                       C.byref(n), counts_ptr or None), "demod_iq_batch(_dev)")
        return int(n.value)

    def demod_iq_batch_async_ptr(self, iq_ptr: int, n_buffers: int, spb: int, stride: int, out_ptr: int,
                                 cap: int, result_ptr: int, lengths_ptr: int = 0) -> None:
        """Enqueue-only batch (b200adsb_demod_iq_batch_dev_async): result_ptr -> 4 x uint32 on the
        device {frames, overflow flags, candidates, frames > cap}; read it after a stream sync."""
        self._check(self._L.b200adsb_demod_iq_batch_dev_async(self._h, iq_ptr, n_buffers, spb, stride,
                                                              lengths_ptr or None, out_ptr, cap, result_ptr),
                    "demod_iq_batch_dev_async")

    def scan_batch_dev(self, iq_ptr: int, n_buffers: int, spb: int, stride: int, first_ordinal: int,
                       ordinal_stride: int, lengths_ptr: int = 0):
        self._check(self._L.b200adsb_scan_batch_dev(self._h, iq_ptr, n_buffers, spb, stride,
                                                    lengths_ptr or None, first_ordinal, ordinal_stride),
                    "scan_batch_dev")

    def scan_batch_dev_async(self, iq_ptr: int, n_buffers: int, spb: int, stride: int, first_ordinal: int,
                             ordinal_stride: int, lengths_ptr: int = 0):
        self._check(self._L.b200adsb_scan_batch_dev_async(self._h, iq_ptr, n_buffers, spb, stride,
                                                          lengths_ptr or None, first_ordinal, ordinal_stride),
                    "scan_batch_dev_async")

    def resolve_batch_dev_async(self, out_ptr: int, cap: int, result_ptr: int):
        self._check(self._L.b200adsb_resolve_batch_dev_async(self._h, out_ptr, cap, result_ptr),
                    "resolve_batch_dev_async")

    def events_count(self) -> int:
        n = C.c_size_t(0)
        self._check(self._L.b200adsb_events_count(self._h, C.byref(n)), "events_count")
        return int(n.value)

    def events_export_dev(self, pairs_ptr: int, cap: int) -> int:
        n = C.c_size_t(0)
        self._check(self._L.b200adsb_events_export_dev(self._h, pairs_ptr, cap, C.byref(n)), "events_export")
        return int(n.value)

    def events_import_dev(self, pairs_ptr: int, n: int):
        self._check(self._L.b200adsb_events_import_dev(self._h, pairs_ptr, n), "events_import")

    def events_pack_dev(self, rows_ptr: int, rows_cap: int):
        self._check(self._L.b200adsb_events_pack_dev(self._h, rows_ptr, rows_cap), "events_pack")

    def events_import_packed_dev(self, gathered_ptr: int, n_ranks: int, rows_per_rank: int, skip_rank: int):
        self._check(self._L.b200adsb_events_import_packed_dev(self._h, gathered_ptr, n_ranks, rows_per_rank,
                                                              skip_rank), "events_import_packed")

    def events_push_symm_dev(self, peer_bufs_dev_ptr: int, rank: int, n_ranks: int, rows_per_rank: int, epoch: int,
                             force_flags: int = 0):
        self._check(self._L.b200adsb_events_push_symm_dev(self._h, peer_bufs_dev_ptr, rank, n_ranks, rows_per_rank,
                                                          epoch, force_flags), "events_push_symm")

    def events_import_symm_dev(self, local_buf_ptr: int, rank: int, n_ranks: int, rows_per_rank: int, epoch: int):
        self._check(self._L.b200adsb_events_import_symm_dev(self._h, local_buf_ptr, rank, n_ranks, rows_per_rank,
                                                            epoch), "events_import_symm")

    def frames_pack_dev(self, frames_ptr: int, block_ptr: int, rows_cap: int, count: int = 0, count_ptr: int = 0,
                        stream: int = 0):
        self._check(self._L.b200adsb_frames_pack_dev(self._h, stream or None, frames_ptr or None, count_ptr or None,
                                                     count, block_ptr, rows_cap), "frames_pack")

    def frames_merge_dev(self, gathered_ptr: int, n_ranks: int, rows_cap: int, out_ptr: int, cap: int, n_out_ptr: int,
                         stream: int = 0):
        self._check(self._L.b200adsb_frames_merge_dev(self._h, stream or None, gathered_ptr, n_ranks, rows_cap,
                                                      out_ptr or None, cap, n_out_ptr), "frames_merge")

    def frames_push_symm_dev(self, frames_ptr: int, peer_bufs_dev_ptr: int, rank: int, n_ranks: int, rows_cap: int,
                             epoch: int, ticket_ptr: int, count: int = 0, count_ptr: int = 0, stream: int = 0):
        self._check(self._L.b200adsb_frames_push_symm_dev(self._h, stream or None, frames_ptr or None,
                                                          count_ptr or None, count, peer_bufs_dev_ptr, rank, n_ranks,
                                                          rows_cap, epoch, ticket_ptr), "frames_push_symm")

    def frames_merge_symm_dev(self, local_buf_ptr: int, n_ranks: int, rows_cap: int, epoch: int, out_ptr: int, cap: int,
                              n_out_ptr: int, stream: int = 0):
        self._check(self._L.b200adsb_frames_merge_symm_dev(self._h, stream or None, local_buf_ptr, n_ranks, rows_cap,
                                                           epoch, out_ptr or None, cap, n_out_ptr), "frames_merge_symm")

    def resolve_batch_dev(self, out_ptr: int, cap: int, counts_ptr: int = 0) -> int:
        n = C.c_size_t(0)
        self._check(self._L.b200adsb_resolve_batch_dev(self._h, out_ptr, cap, C.byref(n), counts_ptr or None),
                    "resolve_batch_dev")
        return int(n.value)

    # -------------------------------------------------------------- icao_filter.rs
    def icao_flush(self):
        self._check(self._L.b200adsb_icao_flush(self._h), "icao_flush")

    def icao_filter_add(self, addr: int):
        self._check(self._L.b200adsb_icao_filter_add(self._h, addr & 0xFFFFFFFF), "icao_filter_add")

    def icao_filter_test(self, addr: int) -> bool:
        rc = self._L.b200adsb_icao_filter_test(self._h, addr & 0xFFFFFFFF)
        if rc < 0:
            self._check(rc, "icao_filter_test")
        return bool(rc)

    def icao_snapshot(self) -> list[int]:
        keys = np.zeros(4096, dtype=np.uint32)
        n = C.c_size_t(0)
        self._check(self._L.b200adsb_icao_snapshot(self._h, keys.ctypes.data, 4096, C.byref(n)), "icao_snapshot")
        return [int(k) for k in keys[: n.value]]

    def icao_restore(self, keys):
        k = np.ascontiguousarray(list(keys), dtype=np.uint32)
        self._check(self._L.b200adsb_icao_restore(self._h, k.ctypes.data if k.size else None, k.size),
                    "icao_restore")

    # -------------------------------------------------------------- crc.rs / mode_s
    def modes_checksum(self, msgs, bits: int) -> np.ndarray:
        m = np.ascontiguousarray(msgs, dtype=np.uint8).reshape(-1, 14)
        out = np.zeros(m.shape[0], dtype=np.uint32)
        self._check(self._L.b200adsb_modes_checksum(self._h, m.ctypes.data, m.shape[0], bits, out.ctypes.data),
                    "modes_checksum")
        return out

    def modes_checksum_one(self, message: bytes, bits: int) -> int:
        out = C.c_uint32(0)
        self._check(self._L.b200adsb_modes_checksum_one(self._h, message, len(message), bits, C.byref(out)),
                    "modes_checksum_one")
        return int(out.value)

    def score_modes_message(self, msg: bytes):
        from .demod_2400 import MsgLen
        ln, sc = C.c_int(0), C.c_int(0)
        self._check(self._L.b200adsb_score_modes_message(self._h, msg, len(msg), C.byref(ln), C.byref(sc)),
                    "score_modes_message")
        if ln.value == 0:
            return None
        return (MsgLen.Long if ln.value == 14 else MsgLen.Short, int(sc.value))

    def debug_records_np(self, cap: int = 1 << 18):
        """Stage-1 records of the pending batch (between scan_batch_dev and resolve_batch_dev) as arrays:
        buffers [n] and records [n, 6] = {j, w4..w8}, in (buffer, j) order."""
        bufs = np.zeros(cap, dtype=np.uint32)
        rec = np.zeros((cap, 6), dtype=np.uint32)
        n = C.c_size_t(0)
        self._check(self._L.b200adsb_debug_records(self._h, bufs.ctypes.data, rec.ctypes.data, cap, C.byref(n)),
                    "debug_records")
        return bufs[: n.value], rec[: n.value]

    def debug_records(self, cap: int = 1 << 18):
        """The same as a list [(buffer, j, [w4..w8])]."""
        bufs, rec = self.debug_records_np(cap)
        return [(int(bufs[i]), int(rec[i, 0]), [int(x) for x in rec[i, 1:]]) for i in range(len(bufs))]

    def async_acknowledge(self):
        self._check(self._L.b200adsb_async_acknowledge(self._h), "async_acknowledge")

    def score_modes_messages(self, msgs):
        m = np.ascontiguousarray(msgs, dtype=np.uint8).reshape(-1, 14)
        lens = np.zeros(m.shape[0], dtype=np.uint8)
        scores = np.zeros(m.shape[0], dtype=np.int32)
        self._check(self._L.b200adsb_score_modes_messages(self._h, m.ctypes.data, m.shape[0],
                                                          lens.ctypes.data, scores.ctypes.data),
                    "score_modes_messages")
        return lens, scores


_default: Context | None = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context(0)
    return _default
