"""One stream dealt round-robin over several GPUs (BASELINE.json configs[4]).

The sample-domain work has no dependency between IQ buffers (utils.rs:44, lib.rs:47-50), so
buffer g of the stream goes to rank g % world.  The one thing the reference shares between
buffers is the ICAO address filter (icao_filter.rs:8-9).  Its evolution is order-free once
every rank knows, for every address, the earliest stream position that adds it
(SURVEY.md A.6), so the exchange step is one tiny all-gather of (key, ordinal) pairs
between the scan and the resolve stages -- a few entries per buffer.

torch.distributed is the plumbing only (NCCL on GPUs; the same code runs on gloo/CPU
tensors in tests/test_sharded_gloo.py).
"""
from __future__ import annotations


def owner_of(global_buffer: int, world: int) -> tuple[int, int]:
    """(rank, local index) of stream buffer g under round-robin dealing."""
    return global_buffer % world, global_buffer // world


def local_buffers(n_total: int, world: int, rank: int) -> list[int]:
    """Global indices of the buffers rank `rank` owns, in local order."""
    return list(range(rank, n_total, world))


def exchange_events(pairs, count: int, group=None):
    """All-gather the ranks' ICAO add-events.

    pairs: int64 tensor [cap, 2] holding (key, ordinal) rows, the first `count` valid, on
    the device the backend works with.  Returns (remote, n_remote): the other ranks' valid
    rows concatenated ([m, 2]).  With world size 1 returns an empty tensor.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return pairs[:0], 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cnt = torch.tensor([count], dtype=torch.int64, device=pairs.device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    counts = [int(c.item()) for c in cnts]
    width = max(max(counts), 1)
    send = torch.zeros((width, 2), dtype=torch.int64, device=pairs.device)
    send[:count] = pairs[:count]
    gathered = [torch.zeros_like(send) for _ in range(world)]
    dist.all_gather(gathered, send, group=group)
    rows = [gathered[r][:counts[r]] for r in range(world) if r != rank and counts[r]]
    if not rows:
        return pairs[:0], 0
    remote = torch.cat(rows).contiguous()
    return remote, remote.shape[0]


class ShardedDemodulator:
    """This rank's share of a sharded stream: scan -> exchange -> resolve on its Context, then
    (optionally) the frame gather that turns the ranks' outputs into the reference's single ordered
    stream (dump1090_rs/src/main.rs:166-200).

    The exchanges are fixed-size all-gathers on the context's stream with no host round trip: row 0 of
    each rank's event block carries its event count and its bad-batch flags
    (b200adsb_events_pack_dev / b200adsb_events_import_packed_dev), row 0 of its frame block the frame
    count (b200adsb_frames_pack_dev / b200adsb_frames_merge_dev).  A batch that fails on one rank
    (candidate pool, event table or exchange buffer overflow) fails on every rank and is committed on
    none, so the ranks' filters stay identical and every rank takes the same path through the
    collectives."""

    FRAME_BYTES = 28

    def __init__(self, ctx, rank: int, world: int, group=None, event_rows: int = 1024, frame_rows: int = 2048,
                 exchange: str = "auto"):
        """exchange: "symm" = the event exchange as peer stores into symmetric memory from the library's own
        kernels (b200adsb_events_push_symm_dev / _import_symm_dev; no NCCL call between scan and resolve),
        "nccl" = pack + all_gather_into_tensor + import, "auto" = symm when torch can set it up."""
        import torch
        import torch.distributed as dist

        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.event_rows, self.frame_rows = event_rows, frame_rows
        dev = torch.device("cuda", ctx.device)
        self.rows = torch.zeros((event_rows, 2), dtype=torch.int64, device=dev)
        self.gathered = torch.zeros((world * event_rows, 2), dtype=torch.int64, device=dev)
        self.fblock = torch.zeros(((frame_rows + 1) * self.FRAME_BYTES,), dtype=torch.uint8, device=dev)
        self.fgathered = torch.zeros((world * (frame_rows + 1) * self.FRAME_BYTES,), dtype=torch.uint8, device=dev)
        self.n_out = torch.zeros((2,), dtype=torch.int32, device=dev)
        self.position = 0          # stream position (in global buffers) of the next batch
        self.exchange = "nccl"
        self.epoch = 0
        if world > 1 and exchange in ("auto", "symm"):
            try:
                self._setup_symm(dev, group)
                self.exchange = "symm"
            except Exception as e:      # noqa: BLE001 -- any failure means "not available here"
                if exchange == "symm":
                    raise
                self.symm_error = repr(e)[:200]
        if world > 1:                   # every rank must have taken the same decision
            flag = torch.tensor([1 if self.exchange == "symm" else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            self.exchange = "symm" if int(flag.item()) == 1 else "nccl"
        # The frame gather depends on nothing but its batch's resolve, and only the emitter waits for it: it
        # runs on a side stream with a communicator of its own, off the scan -> exchange -> resolve critical path
        # (on the main stream it cost 0.075 ms per step at N = 8, profiles/r2).
        self.fstream = torch.cuda.Stream(device=dev)
        self.fgroup = dist.new_group(backend=dist.get_backend(group)) if world > 1 else None
        self.ev_resolved = torch.cuda.Event()
        self.ev_packed = torch.cuda.Event()

    def _setup_symm(self, dev, group) -> None:
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from . import _ffi
        words = int(_ffi.lib().b200adsb_events_symm_words(self.world, self.event_rows))
        g = group if group is not None else dist.group.WORLD
        self.symm = symm_mem.empty(words, dtype=torch.int64, device=dev)
        self.symm.zero_()
        torch.cuda.synchronize(dev)
        self.symm_hdl = symm_mem.rendezvous(self.symm, g)
        self.peer_bufs_dev = int(self.symm_hdl.buffer_ptrs_dev)
        # the frame gather's blocks, the same way
        fbytes = int(_ffi.lib().b200adsb_frames_symm_bytes(self.world, self.frame_rows))
        self.fsymm = symm_mem.empty((fbytes + 7) // 8, dtype=torch.int64, device=dev)
        self.fsymm.zero_()
        torch.cuda.synchronize(dev)
        self.fsymm_hdl = symm_mem.rendezvous(self.fsymm, g)
        self.fpeer_bufs_dev = int(self.fsymm_hdl.buffer_ptrs_dev)
        self.fticket = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.ev_merged = [torch.cuda.Event(), torch.cuda.Event()]
        self.fepoch = 0
        dist.barrier(group=group)       # nobody pushes before every buffer is zeroed

    def _exchange_events(self, poisoned: bool = False) -> None:
        import torch.distributed as dist

        if self.world <= 1:
            return
        if self.exchange == "symm":
            # the frame blocks alternate between two parities: the gather that follows this batch reuses the
            # blocks of the gather before last, so this rank's merge of that one has to be through before its
            # peers can get past this exchange (they cannot push frames of this batch before they have its events).
            # With one gather per batch the in-order side streams already guarantee it
            # (tests/test_exchange_protocol.py); the wait also covers callers that gather only some batches.
            import torch
            torch.cuda.current_stream().wait_event(self.ev_merged[(self.fepoch + 1) & 1])
            self.epoch += 1
            self.ctx.events_push_symm_dev(self.peer_bufs_dev, self.rank, self.world, self.event_rows, self.epoch,
                                          force_flags=2 if poisoned else 0)
            if not poisoned:
                self.ctx.events_import_symm_dev(self.symm.data_ptr(), self.rank, self.world, self.event_rows, self.epoch)
            return
        if poisoned:
            # this rank's scan failed before the exchange: still take part in the collective, with a block
            # that makes every other rank fail the batch too (count beyond the block, EV_OVF flag)
            self.rows.zero_()
            self.rows[0, 0] = 1 << 40
            self.rows[0, 1] = 2
        else:
            self.ctx.events_pack_dev(self.rows.data_ptr(), self.event_rows)
        dist.all_gather_into_tensor(self.gathered, self.rows, group=self.group)
        if not poisoned:
            self.ctx.events_import_packed_dev(self.gathered.data_ptr(), self.world, self.event_rows, self.rank)

    def last_event_counts(self) -> list[int]:
        """Events each rank published in the last exchange (diagnostics; synchronises)."""
        if self.world <= 1:
            return [0]
        if self.exchange == "symm":
            p = self.epoch & 1
            base = 2 * self.world
            return [int(self.symm[base + 2 * self.event_rows * (p * self.world + r)].item()) for r in range(self.world)]
        return [int(self.gathered[r * self.event_rows, 0].item()) for r in range(self.world)]

    def step(self, iq_ptr: int, n_local: int, spb: int, stride: int, out_ptr: int, cap: int,
             n_total: int | None = None, counts_ptr: int = 0) -> int:
        """Demodulates this rank's n_local buffers of a batch of n_total stream buffers
        (default n_local * world); frames (with local buffer indices) go to out_ptr.
        Raises B200AdsbError on EVERY rank when the batch failed on any of them."""
        from ._ffi import B200AdsbError

        n_total = n_local * self.world if n_total is None else n_total
        err = None
        try:
            self.ctx.scan_batch_dev(iq_ptr, n_local, spb, stride, self.position + self.rank, self.world)
        except B200AdsbError as e:      # (the failed scan has already ended its pending batch)
            err = e
        self._exchange_events(poisoned=err is not None)
        if err is not None:
            raise err
        n = self.ctx.resolve_batch_dev(out_ptr, cap, counts_ptr)
        self.position += n_total
        return n

    def step_async(self, iq_ptr: int, n_local: int, spb: int, stride: int, out_ptr: int, cap: int,
                   result_ptr: int, n_total: int | None = None) -> None:
        """Enqueue-only step: scan, event exchange and resolve are queued on the context's stream
        (which must be torch's current stream, so that the NCCL all-gather is ordered with them);
        the outcome {frames, failure flags, candidates, frames > cap} lands in result_ptr
        (4 x uint32 on the device; flags as documented for b200adsb_demod_iq_batch_dev_async --
        identical zero / non-zero on every rank).  Batches execute in call order."""
        n_total = n_local * self.world if n_total is None else n_total
        self.ctx.scan_batch_dev_async(iq_ptr, n_local, spb, stride, self.position + self.rank, self.world)
        self._exchange_events()
        self.ctx.resolve_batch_dev_async(out_ptr, cap, result_ptr)
        self.position += n_total

    def gather_frames(self, frames_ptr: int, out_ptr: int, cap: int, count: int = 0, result_ptr: int = 0):
        """The ranks' frames of the last step -> the single stream in (global buffer, j) order with global
        buffer indices, on EVERY rank (any of them can be the emitting one).  Enqueue-only, on the side stream
        `self.fstream`: pack waits for the step's resolve; the main stream only waits for the pack (so that the
        next batch may overwrite the frame array).  count: frames this rank holds (synchronous step) or
        result_ptr: the step_async outcome (its word 0 is the count).  Returns the device tensor
        n_out = [frames, overflow]; synchronise `fstream` (or the device) before reading it or the output."""
        import torch
        import torch.distributed as dist

        main = torch.cuda.current_stream()
        self.ev_resolved.record(main)
        self.fstream.wait_event(self.ev_resolved)
        fs = self.fstream.cuda_stream
        if self.world > 1 and self.exchange == "symm":
            # peer stores into every rank's block + flag, then a merge that waits for the flags: no collective
            # library kernel competes with the next scan for the SMs
            self.fepoch += 1
            self.ctx.frames_push_symm_dev(frames_ptr, self.fpeer_bufs_dev, self.rank, self.world, self.frame_rows,
                                          self.fepoch, self.fticket.data_ptr(), count=count, count_ptr=result_ptr,
                                          stream=fs)
            self.ev_packed.record(self.fstream)
            main.wait_event(self.ev_packed)
            self.ctx.frames_merge_symm_dev(self.fsymm.data_ptr(), self.world, self.frame_rows, self.fepoch, out_ptr, cap,
                                           self.n_out.data_ptr(), stream=fs)
            self.ev_merged[self.fepoch & 1].record(self.fstream)
            return self.n_out
        self.ctx.frames_pack_dev(frames_ptr, self.fblock.data_ptr(), self.frame_rows, count=count, count_ptr=result_ptr,
                                 stream=fs)
        self.ev_packed.record(self.fstream)
        main.wait_event(self.ev_packed)
        if self.world > 1:
            with torch.cuda.stream(self.fstream):
                dist.all_gather_into_tensor(self.fgathered, self.fblock, group=self.fgroup)
            src = self.fgathered
        else:
            src = self.fblock
        self.ctx.frames_merge_dev(src.data_ptr(), self.world, self.frame_rows, out_ptr, cap, self.n_out.data_ptr(),
                                  stream=fs)
        return self.n_out
