#!/usr/bin/env python3
"""bench.py -- BASELINE.json metric on B200: IQ Msamples/s through the demodulation hot path.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun)
  python bench.py --impl reference ...                   (CPU restatement of the reference)

A step = one pass of the hot path (IQ -> frames) over one batch of synthetic 2.4 Msps CS16
buffers of 131,072 samples (512 KiB).  The workload is ONE stream of global buffers 0, 1, 2, ...
(rtl-like noise, torch generator seed 1090; the first few buffers carry injected DF17 traffic from a
shared aircraft pool so that parity at any N exercises the cross-rank filter exchange):
  N = 1   BASELINE configs[2]: 1000 buffers back to back, device resident.
  N > 1   configs[4] style: the stream dealt round-robin over the ranks, 1000 buffers per GPU (weak
          scaling, the same per-GPU load as N = 1), ICAO add-event all-gather between scan and resolve,
          frame gather into the single (buffer, j)-ordered stream inside the timed step.
Sub-records of the same JSON line: `configs0` (one capture through b200adsb_demod_iq, median latency,
benches/demod_benchmark.rs:7-12), `configs3` (1/10/100 injected messages per buffer), `strong_8192`
(configs[4] as written: 8192 buffers in total over the N GPUs), `h2d_ceiling`, `iq_scatter`.

`value`  : device-resident inputs, CUDA-event timed, max over ranks.
`e2e`    : same batch through the host-buffer C-ABI call (pinned host IQ in, frames out),
           H2D/D2H inside the timed region.
`roofline`: scan kernel, 4 algorithmic bytes per sample / CUDA-event kernel time, against
           MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: oracle/ (C restatement; the Rust reference cannot be built here) on one host
           core over a bounded sample of the same batch.
`parity_on_sample`: the frames of the first global buffers (all ranks' shares, gathered) against the
           single-stream oracle, at every N.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

SAMPLES = 131072
METRIC = "iq_msamples_per_s"
UNIT = "Msamples/s"
SEED = 1090
BUFFERS_PER_GPU = 1000          # BASELINE configs[2]
SIGNAL_PER_RANK = 8             # injected-traffic buffers at the head of the stream, per rank
SIGNAL_MSGS = 24
SIGNAL_POOL = 8                 # aircraft addresses shared by all signal buffers (=> cross-rank filter events)


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML (nvidia_ml_py) is
    polled in a thread (a query takes well under a millisecond, the timed region only
    milliseconds); falls back to nvidia-smi when NVML is unavailable."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.source = None
        self._stop = threading.Event()
        self._t = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(index))
            self._nv = pynvml
            self.source = "nvml"
        except Exception:
            self._h = None
            self.source = "nvidia-smi"

    def _poll_nvml(self):
        nv = self._nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        if not self.mx:          # a constant of the board: one query
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for b, name in self.BITS.items():
            if bits & b:
                self.reasons.add(name)

    def _poll_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        for line in out.strip().splitlines():
            r = [x.strip() for x in line.split(",")]
            self.sm.append(float(r[0]))
            self.mx.append(float(r[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._h is not None:
                    self._poll_nvml()
                else:
                    self._poll_smi()
            except Exception:
                pass
            self._stop.wait(0.0005 if self._h is not None else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"], "source": self.source}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.source}


def physical_index(index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        parts = vis.split(",")
        if index < len(parts) and parts[index].strip().isdigit():
            return int(parts[index])
    return index


def bind_to_gpu_numa(index: int) -> dict:
    """Pin this process to the CPUs next to its GPU BEFORE any pinned allocation: cudaHostAlloc pages land on
    the caller's NUMA node, and a rank whose staging buffers sit on the far socket copies at a fraction of
    the PCIe rate (round 1: per-GPU H2D fell from 54 to 23 GB/s at N = 8)."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        dev = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}"
        with open(dev + "/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        try:
            with open(dev + "/numa_node") as f:
                info["numa_node"] = int(f.read().strip())
        except Exception:
            pass
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus))
    except Exception as e:          # best effort: the bench still runs unbound
        info["error"] = str(e)[:80]
    return info


def stream_noise(n_global: int, take, device, chunk: int = 64):
    """Global buffers 0..n_global-1 of the bench stream (rtl-like noise, one torch generator seeded SEED, chunks
    of 64 global buffers); returns the buffers `take` selects (a list of global indices, ascending) as one
    int16 tensor [len(take), SAMPLES, 2] on `device`.  Every rank and the reference arm call this with the
    same n_global-independent prefix property: buffer g does not depend on how many buffers follow it."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(SEED)
    want = list(take)
    out = torch.empty((len(want), SAMPLES, 2), dtype=torch.int16, device=device)
    pos = 0
    for g0 in range(0, n_global, chunk):
        f = torch.randn((chunk, SAMPLES, 2), generator=g, device=device, dtype=torch.float32) * 5.5 + 127.4
        sel = []
        while pos + len(sel) < len(want) and want[pos + len(sel)] < g0 + chunk:
            sel.append(want[pos + len(sel)] - g0)
        if sel:
            u8 = torch.clamp(torch.round(f[sel]), 0, 255)
            out[pos:pos + len(sel)] = torch.trunc((u8 - 127.4) / 128.0 * 32767.0).to(torch.int16)
            pos += len(sel)
        if pos >= len(want):
            break
    return out


def signal_buffers(n: int):
    """The injected-traffic buffers that replace global buffers 0..n-1 (numpy, identical everywhere)."""
    from dump1090_rs_b200 import synth
    return synth.make_batch(SEED, n, msgs_per_buffer=SIGNAL_MSGS, icao_pool=SIGNAL_POOL)


def frames_to_tuples(raw):
    """uint8 [n, 28] frame rows -> (buffer, j, phase, score, hex)"""
    import numpy as np
    return [(int(r[24:28].view(np.uint32)[0]), int(r[20:24].view(np.uint32)[0]), int(r[15]),
             int(r[16:18].view(np.int16)[0]), bytes(r[: r[14]]).hex()) for r in raw]


def run_reference(args) -> int:
    """--impl reference: the reference's CPU algorithm (oracle port; Rust itself cannot be built
    in this image) on all host threads over THE SAME 1000-buffer batch as the GPU arm at N = 1 (one step =
    one pass over it, buffers dealt to the threads, each thread a private filter as independent receivers
    would be); also the 1-thread figure the reference's own bench reports."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import torch
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    nb = args.buffers or BUFFERS_PER_GPU
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    batch = stream_noise(nb, range(nb), dev).cpu().numpy()
    k = min(SIGNAL_PER_RANK, nb)
    batch[:k] = signal_buffers(k)
    for _ in range(max(args.warmup, 1)):
        O.bench(batch[: 4 * cores], min(4 * cores, nb), SAMPLES, 1, cores, False)
    t = 0.0
    for _ in range(args.steps):
        sec, _fr = O.bench(batch, nb, SAMPLES, 1, cores, False)
        t += sec
    value = nb * SAMPLES * args.steps / t / 1e6
    n1 = min(64, nb)
    O.bench(batch[:n1], n1, SAMPLES, 1, 1, False)
    sec1, _ = O.bench(batch[:n1], n1, SAMPLES, 3, 1, False)
    one = n1 * 3 * SAMPLES / sec1 / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32 (+f32 magnitude)",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": workload_text(nb, max(args.gpus, 1)), "buffers_per_gpu": nb, "samples_per_buffer": SAMPLES,
                   "host_threads": cores, "data_generated_on": dev,
                   "sample": "one GPU's share of the stream per step (rank 0's batch at N = 1): the CPU arm is one host"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"the whole {nb}-buffer batch, one pass per step x {args.steps} steps, {cores} threads "
                                   "(buffers dealt to the threads, private filter each), C restatement of "
                                   "to_mag+demodulate2400 (rustc unavailable)",
                         "one_thread_value": one,
                         "one_thread_sample": f"first {n1} buffers x 3 passes, 1 thread (the shape of benches/demod_benchmark.rs)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_text(nb: int, world: int) -> str:
    base = (f"synthetic 2.4Msps CS16 stream, rtl-like noise (sigma 5.5 LSB of 8 bit, torch generator seed {SEED}), "
            f"{nb} x 512KiB buffers per GPU per step; the first {SIGNAL_PER_RANK} buffers per GPU carry {SIGNAL_MSGS} injected DF17 "
            f"each from a pool of {SIGNAL_POOL} aircraft; ")
    return base + ("BASELINE configs[2]" if world == 1 else
                   "configs[4]-style round-robin shards (same per-GPU load as N=1) + ICAO event all-gather + frame gather")


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--buffers", type=int, default=0, help=f"buffers per rank (default {BUFFERS_PER_GPU})")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--sync-steps", action="store_true", help="time the synchronous ABI call (one host round trip per step)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (configs0, configs3, strong_8192, ...)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "symm", "nccl"],
                    help="N>1 event exchange: peer stores into symmetric memory from the library's kernels, or NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa(local)

    import numpy as np
    import torch
    import dump1090_rs_b200 as d
    from dump1090_rs_b200 import _ffi, sharded

    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    nb = args.buffers or BUFFERS_PER_GPU

    # ---- data: this rank's share of the stream (global buffer g = b*world + rank)
    n_sig = min(SIGNAL_PER_RANK, nb) * world
    iq = stream_noise(nb * world, range(rank, nb * world, world), dev)
    sig = signal_buffers(n_sig)
    iq[: n_sig // world] = torch.from_numpy(np.ascontiguousarray(sig[rank::world])).to(dev)

    # a real (non-default) stream shared by torch and the library: the timing events below are
    # recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = d.Context(local, stream.cuda_stream)
    ctx.set_option(_ffi.OPT_PROFILE, 1)
    if args.tile:
        ctx.set_option(_ffi.OPT_TILE, args.tile)
    cap = 1 << 16
    frames = torch.zeros((cap, 28), dtype=torch.uint8, device=dev)
    merged = torch.zeros((cap, 28), dtype=torch.uint8, device=dev)
    n_frames = [0]
    sh = sharded.ShardedDemodulator(ctx, rank, world, frame_rows=2048, exchange=args.exchange)

    # The timed steps are queued back to back with the enqueue-only entry points: the batch outcome {frames,
    # failure flags, ...} stays on the device and is checked after the timed region; warm-up steps use the
    # synchronous call (which also sizes the candidate pool).  --sync-steps times the synchronous calls instead.
    results = torch.zeros((max(args.steps, 1), 4), dtype=torch.int32, device=dev)
    step_no = [0]
    use_async = [False]

    def run_step(iq_t, nbuf, queued):
        ctx.icao_flush()
        if world == 1 and queued:
            r = results[step_no[0] % results.shape[0]]
            ctx.demod_iq_batch_async_ptr(iq_t.data_ptr(), nbuf, SAMPLES, SAMPLES, frames.data_ptr(), cap, r.data_ptr())
            step_no[0] += 1
        elif world == 1:
            n_frames[0] = ctx.demod_iq_batch_ptr(iq_t.data_ptr(), nbuf, SAMPLES, SAMPLES, frames.data_ptr(), cap)
        elif queued:
            # scan -> all-gather of ICAO add-events (NCCL) -> resolve -> all-gather of frames -> ordered stream
            sh.position = 0
            r = results[step_no[0] % results.shape[0]]
            sh.step_async(iq_t.data_ptr(), nbuf, SAMPLES, SAMPLES, frames.data_ptr(), cap, r.data_ptr())
            sh.gather_frames(frames.data_ptr(), merged.data_ptr(), cap, result_ptr=r.data_ptr())
            step_no[0] += 1
        else:
            sh.position = 0
            n_frames[0] = sh.step(iq_t.data_ptr(), nbuf, SAMPLES, SAMPLES, frames.data_ptr(), cap)
            sh.gather_frames(frames.data_ptr(), merged.data_ptr(), cap, count=n_frames[0])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(iq_t, nbuf, steps, queued, sample_clocks=True):
        step_no[0] = 0
        barrier()              # ranks enter together (rank 0 may come from seconds of CPU-side work)
        for _ in range(args.warmup):
            run_step(iq_t, nbuf, False)
        barrier()
        ctx.timing(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk_:
            barrier()
            e0.record()
            for _ in range(steps):
                run_step(iq_t, nbuf, queued)
            e1.record()
            barrier()
        ms_ = e0.elapsed_time(e1)
        ctx.sync()
        tim_ = ctx.timing(reset=True)
        ok = True
        if queued:
            res = results[:min(steps, results.shape[0])].cpu().numpy()
            ok = bool((res[:, 1] == 0).all() and (res[:, 3] == 0).all() and (res[:, 0] == n_frames[0]).all())
            if not ok:
                print("bench: a queued batch failed or disagreed with the synchronous call: %r" % (res.tolist(),),
                      file=sys.stderr)
        if dist is not None:   # every rank must take the same path
            flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            ok = int(flag.item()) == 0
            t = torch.tensor([ms_], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_, tim_, clk_, ok

    def measure(iq_t, nbuf, steps):
        """(ms, timing, clocks, queued?) of `steps` steps, queued unless that cannot be verified"""
        queued = not args.sync_steps
        ms_, tim_, clk_, ok = timed_region(iq_t, nbuf, steps, queued)
        if not ok:             # never report an unverified number: time the synchronous call instead
            queued = False
            ms_, tim_, clk_, ok = timed_region(iq_t, nbuf, steps, False)
        return ms_, tim_, clk_, queued

    if world > 1 and sh.exchange == "symm":
        # one trial step: if the peer-memory exchange does not work on this box (every rank then fails the
        # batch together), the NCCL exchange takes over
        barrier()
        bad = 0
        try:
            run_step(iq, nb, False)
            ctx.sync()
        except _ffi.B200AdsbError:
            bad = 1
        flag = torch.tensor([bad], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()):
            sh.exchange = "nccl"
            ctx.async_acknowledge()
    ms, tim, clk, queued_steps = measure(iq, nb, args.steps)
    total_samples = nb * SAMPLES * world
    value = total_samples * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (scan): algorithmic bytes = 4 B per IQ sample
    peak, peak_src = peaks()

    def roofline_of(tim_, steps, ms_):
        launches = max(tim_["scan_launches"], 1)
        achieved = 4.0 * tim_["samples"] / (tim_["scan_ms"] * 1e-3) / 1e9 if tim_["scan_ms"] > 0 else 0.0
        traffic = None
        try:
            with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
                per = json.load(f).get("dram_bytes_per_sample", None)
                if per is not None:
                    traffic = per * tim_["samples"] / launches
        except Exception:
            pass
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "kernel": "scan7_kernel<false, 7384, true>", "kernel_ms_per_launch": tim_["scan_ms"] / launches,
                "algorithmic_bytes_per_launch": 4.0 * tim_["samples"] / launches, "peak_source": peak_src,
                "kernel_share_of_step": (tim_["scan_ms"] / ms_) if ms_ else None,
                "kernel_ms_per_step": tim_["scan_ms"] / steps,
                "resolve_kernels_ms_per_step": tim_["resolve_ms"] / steps}

    roofline = roofline_of(tim, args.steps, ms)
    launches_main = int(tim["scan_launches"] + tim["other_launches"])

    # ---- parity on the head of the stream, at every N: one synchronous step, frames gathered into the single
    # ordered stream, the first n_sig global buffers against the sequential oracle
    run_step(iq, nb, False)
    ctx.sync()
    sh.fstream.synchronize()
    if world > 1:
        n_out = sh.n_out.cpu().numpy()
        stream_frames = frames_to_tuples(merged[: int(n_out[0])].cpu().numpy())
        remote_events = int(sum(c for r, c in enumerate(sh.last_event_counts()) if r != rank))
        gather_overflow = int(n_out[1])
    else:
        stream_frames = frames_to_tuples(frames[: n_frames[0]].cpu().numpy())
        remote_events, gather_overflow = 0, 0
    parity = None
    if rank == 0 and not args.no_cpu:
        from oracle import oracle as O
        o = O.Oracle()
        ref = []
        for g in range(n_sig):
            ref += [(g, f["j"], f["phase"], f["score"], f["msg"].hex()) for f in o.demod_iq(sig[g])]
        got = [t for t in stream_frames if t[0] < n_sig]
        parity = {"ok": got == ref and gather_overflow == 0, "global_buffers": n_sig, "frames": len(ref),
                  "frames_whole_step": len(stream_frames), "events_from_other_ranks": remote_events}

    # ---- and deep inside the batch (N = 1): stage-1 records of buffers sampled across the whole launch --
    # survivor set and every (j, try-phase) classification -- against the oracle (frames are too rare in noise
    # to say anything about buffer 500)
    if rank == 0 and world == 1 and not args.no_cpu and parity is not None:
        try:
            from oracle import oracle as O
            pick = sorted(set(range(0, nb, max(nb // 11, 1))) | {nb - 1})[:13]
            ctx.scan_batch_dev(iq.data_ptr(), nb, SAMPLES, SAMPLES, 0, 1)
            rb, rr = ctx.debug_records_np(cap=max(nb * 2600, 1 << 16))
            ctx.resolve_batch_dev(frames.data_ptr(), cap)
            host = iq[pick].cpu().numpy()
            o = O.Oracle()
            deep_ok, n_rec = True, 0
            norm = lambda w: [0 if (x >> 29) == 0 else x for x in w]
            for k, b in enumerate(pick):
                got = [(int(r[0]), norm([int(x) for x in r[1:]])) for r in rr[rb == b]]
                ref = [(j, norm(w)) for j, w in o.records(o.to_mag(host[k]), cap=1 << 17)]
                deep_ok = deep_ok and got == ref
                n_rec += len(ref)
            parity["sampled_buffers"] = pick
            parity["stage1_records_checked"] = n_rec
            parity["stage1_records_ok"] = deep_ok
            parity["ok"] = parity["ok"] and deep_ok
        except Exception as e:    # noqa: BLE001 -- report, do not lose the line
            try:                  # (a scan left pending would block the context for the sub-records)
                ctx.resolve_batch_dev(frames.data_ptr(), cap)
            except Exception:     # noqa: BLE001
                pass
            parity["stage1_records_ok"] = False
            parity["stage1_records_error"] = str(e)[:120]
            parity["ok"] = False

    sub = {}
    # ---- e2e: host-buffer C-ABI call, H2D + D2H inside the timed region; and the H2D-only ceiling beside it
    e2e = None
    if not args.no_e2e:
        host_iq = torch.empty((nb, SAMPLES, 2), dtype=torch.int16, pin_memory=True)
        host_iq.copy_(iq)
        host_frames = torch.zeros((cap, 28), dtype=torch.uint8, pin_memory=True)
        ectx = d.Context(local)          # private stream; its own filter
        nf = 0
        for _ in range(2):
            ectx.icao_flush()
            nf = ectx.demod_iq_batch_ptr(host_iq.data_ptr(), nb, SAMPLES, SAMPLES, host_frames.data_ptr(), cap, host=True)
        barrier()
        t0 = time.perf_counter()
        k_e2e = max(3, args.steps // 5)
        for _ in range(k_e2e):
            ectx.icao_flush()
            nf = ectx.demod_iq_batch_ptr(host_iq.data_ptr(), nb, SAMPLES, SAMPLES, host_frames.data_ptr(), cap, host=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # ceiling: the same pinned batch through plain async copies, nothing else
        stage = torch.empty_like(iq)
        cs = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(cs):
            stage.copy_(host_iq, non_blocking=True)
        cs.synchronize()
        barrier()
        t1 = time.perf_counter()
        with torch.cuda.stream(cs):
            for _ in range(k_e2e):
                stage.copy_(host_iq, non_blocking=True)
        cs.synchronize()
        dt_copy = time.perf_counter() - t1
        if dist is not None:
            t = torch.tensor([dt, dt_copy], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_copy = float(t[0].item()), float(t[1].item())
        e2e = {"value": total_samples * k_e2e / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": nb * SAMPLES * 4, "d2h_bytes_per_step": nf * 28 + 64,
               "steps": k_e2e, "frames_per_step": nf}
        sub["h2d_ceiling"] = {"value": total_samples * k_e2e / dt_copy / 1e6, "unit": UNIT,
                              "gbytes_per_s_per_gpu": nb * SAMPLES * 4 * k_e2e / dt_copy / 1e9,
                              "e2e_fraction_of_ceiling": dt_copy / dt, "numa": numa,
                              "what": "the same pinned host batch through cudaMemcpyAsync alone (max over ranks)"}
        if not args.no_sub:
            # opt-in 8-bit ingest (RTL-SDR delivers u8; the reference receives SoapySDR's CS16 expansion of it):
            # the same batch as u8 pairs, expanded on the device -- half the bytes over PCIe, identical frames
            L = _ffi.lib()
            lut = torch.tensor([L.b200adsb_cu8_to_cs16(v) for v in range(256)], dtype=torch.int16, device=dev)
            inv = torch.zeros(65536, dtype=torch.uint8, device=dev)
            inv[lut.to(torch.int64) + 32768] = torch.arange(256, dtype=torch.uint8, device=dev)
            host_u8 = torch.empty((nb, SAMPLES, 2), dtype=torch.uint8, pin_memory=True)
            host_u8.copy_(inv[iq.to(torch.int64) + 32768])
            ectx.icao_flush()
            nf8 = ectx.demod_cu8_batch_ptr(host_u8.data_ptr(), nb, SAMPLES, SAMPLES, host_frames.data_ptr(), cap)
            barrier()
            t2 = time.perf_counter()
            for _ in range(k_e2e):
                ectx.icao_flush()
                nf8 = ectx.demod_cu8_batch_ptr(host_u8.data_ptr(), nb, SAMPLES, SAMPLES, host_frames.data_ptr(), cap)
            torch.cuda.synchronize()
            dt8 = time.perf_counter() - t2
            if dist is not None:
                t = torch.tensor([dt8], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt8 = float(t.item())
            sub["e2e_cu8_ingest"] = {"value": total_samples * k_e2e / dt8 / 1e6, "unit": UNIT,
                                     "h2d_bytes_per_step": nb * SAMPLES * 2, "frames_per_step": nf8,
                                     "same_frames_as_cs16": nf8 == nf,
                                     "what": "opt-in b200adsb_demod_cu8_batch: the batch as unsigned 8-bit pairs (what an RTL-SDR "
                                             "delivers), expanded to CS16 on the device with SoapySDR's conversion; NOT the "
                                             "headline e2e, which moves the reference's CS16"}
            del host_u8
        ectx.close()
        del host_iq, stage

    # ---- CPU baseline (rank 0): oracle on a bounded sample of the same batch, one core
    cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import oracle as O
        ns = min(32, nb)
        sample = iq[:ns].cpu().numpy()
        iters = 120          # ~10 s of single-core work
        O.bench(sample, ns, SAMPLES, 1, 1, False)
        sec, _fr = O.bench(sample, ns, SAMPLES, iters, 1, False)
        cpu = {"value": ns * SAMPLES * iters / sec / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first {ns} buffers of rank 0's batch x {iters} passes, 1 thread, C restatement "
                         "(oracle/; rustc unavailable so the Rust reference cannot be built)",
               "host_cpus": os.cpu_count(), "parity_on_sample": None if parity is None else parity["ok"]}

    if not args.no_sub:
        # ---- configs[4] as written: 8192 buffers in total over the N GPUs (strong scaling)
        del iq
        torch.cuda.empty_cache()
        nbs = 8192 // world
        iq_s = stream_noise(nbs * world, range(rank, nbs * world, world), dev)
        steps_s = max(5, args.steps // 5)
        ms_s, tim_s, _clk, q_s = measure(iq_s, nbs, steps_s)
        sub["strong_8192"] = {"value": nbs * world * SAMPLES * steps_s / (ms_s * 1e-3) / 1e6, "unit": UNIT,
                              "ms_per_step": ms_s / steps_s, "buffers_total": nbs * world, "buffers_per_gpu": nbs,
                              "steps": steps_s, "queued": q_s, "scaling": "strong",
                              "roofline_frac": roofline_of(tim_s, steps_s, ms_s)["frac"],
                              "frames_per_step_this_rank": n_frames[0]}
        del iq_s
        torch.cuda.empty_cache()

        if world > 1:
            # ---- buffer hand-off from one source (K6): rank 0 scatters 64 buffers to every rank over NVLink
            per = 64
            recv = torch.empty((per, SAMPLES), dtype=torch.int32, device=dev)      # one CS16 pair per word (NCCL has no int16)
            src = [torch.empty_like(recv) for _ in range(world)] if rank == 0 else None
            for _ in range(2):
                dist.scatter(recv, src, src=0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 10
            for _ in range(reps):
                dist.scatter(recv, src, src=0)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sc_ms = float(t.item()) / reps
            out_bytes = (world - 1) * per * SAMPLES * 4
            sub["iq_scatter"] = {"ms": sc_ms, "gbytes_per_s_out_of_rank0": out_bytes / (sc_ms * 1e-3) / 1e9,
                                 "nvlink_reference_gbs": 770.0, "frac_of_reference": out_bytes / (sc_ms * 1e-3) / 1e9 / 770.0,
                                 "msamples_per_s": world * per * SAMPLES / (sc_ms * 1e-3) / 1e6,
                                 "what": f"torch.distributed.scatter (NCCL) of {per} buffers per rank from rank 0; reference = "
                                         "measured peer copy 770 GB/s per direction (B200_PROFILING.md)"}
            del recv, src

        if rank == 0:
            from dump1090_rs_b200 import synth
            # ---- configs[3]: injected DF17 at 1 / 10 / 100 per buffer (16 distinct buffers repeated over the batch)
            sub["configs3"] = {}
            for m in (1, 10, 100):
                try:                      # (a sub-record must never take the main line down with it)
                    inj = torch.from_numpy(synth.make_batch(SEED, 16, msgs_per_buffer=m)).to(dev)
                    n3 = min(nb, 1000)
                    iq_m = inj.repeat((n3 + 15) // 16, 1, 1)[:n3].contiguous()
                    c3 = d.Context(local, stream.cuda_stream)
                    c3.set_option(_ffi.OPT_PROFILE, 1)
                    res3 = torch.zeros((10, 4), dtype=torch.int32, device=dev)
                    for _ in range(3):
                        c3.icao_flush()
                        nfm = c3.demod_iq_batch_ptr(iq_m.data_ptr(), n3, SAMPLES, SAMPLES, frames.data_ptr(), cap)
                    torch.cuda.synchronize()
                    c3.timing(reset=True)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for k in range(10):
                        c3.icao_flush()
                        c3.demod_iq_batch_async_ptr(iq_m.data_ptr(), n3, SAMPLES, SAMPLES, frames.data_ptr(), cap, res3[k].data_ptr())
                    e1.record()
                    torch.cuda.synchronize()
                    c3.sync()
                    t3 = c3.timing(reset=True)
                    r3 = res3.cpu().numpy()
                    ms3 = e0.elapsed_time(e1)
                    sub["configs3"][str(m)] = {
                        "value": n3 * SAMPLES * 10 / (ms3 * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms3 / 10,
                        "frames_per_step": nfm, "verified": bool((r3[:, 1] == 0).all() and (r3[:, 0] == nfm).all()),
                        "scan_ms_per_step": t3["scan_ms"] / 10, "resolve_ms_per_step": t3["resolve_ms"] / 10,
                        "roofline_frac": 4.0 * t3["samples"] / (t3["scan_ms"] * 1e-3) / 1e9 / peak if t3["scan_ms"] else None,
                        "buffers": n3}
                    c3.close()
                    del iq_m, inj
                except Exception as e:    # noqa: BLE001
                    sub["configs3"][str(m)] = {"error": str(e)[:120]}
            # ---- configs[0]: the cargo bench '01' case, one capture per call (benches/demod_benchmark.rs:7-12,23)
            try:
                z = np.load(os.path.join(REPO, "tests", "golden", "captures.npz"))
                cap_iq = np.ascontiguousarray(z["test_1641427457780"].reshape(-1, 2)[:, ::-1])
                c0 = d.Context(local)
                lat = []
                for k in range(230):
                    t0 = time.perf_counter()
                    c0.icao_flush()
                    fr0 = c0.demod_iq(cap_iq)
                    if k >= 30:
                        lat.append(time.perf_counter() - t0)
                med = statistics.median(lat)
                rec0 = {"capture": "test_iq/test_1641427457780.iq", "call": "icao_flush + b200adsb_demod_iq (host pointer in, frames out)",
                        "median_ms": med * 1e3, "p90_ms": sorted(lat)[int(0.9 * len(lat))] * 1e3,
                        "value": SAMPLES / med / 1e6, "unit": UNIT, "frames": len(fr0), "iterations": len(lat),
                        "reference_readme_ms": 3.695}
                if not args.no_cpu:
                    from oracle import oracle as O
                    O.bench(cap_iq[None], 1, SAMPLES, 20, 1, True)
                    sec0, _ = O.bench(cap_iq[None], 1, SAMPLES, 200, 1, True)
                    rec0["cpu_port_ms"] = sec0 / 200 * 1e3
                    rec0["speedup_vs_cpu_port_1_thread"] = (sec0 / 200) / med
                sub["configs0"] = rec0
                c0.close()
            except Exception as e:
                sub["configs0"] = {"error": str(e)[:120]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32 (+f32 magnitude)",
            "data": "synthetic",
            "config": {"workload": workload_text(nb, world),
                       "buffers_per_gpu": nb, "samples_per_buffer": SAMPLES,
                       "l2": f"inputs {nb * SAMPLES * 4 / 2**20:.0f} MiB per GPU > 126 MB L2 (no flush needed)",
                       "frames_per_step_rank0": n_frames[0],
                       "step_call": ("synchronous ABI calls (host round trips inside every step)" if not queued_steps else
                                     ("b200adsb_demod_iq_batch_dev_async" if world == 1 else
                                      "b200adsb_scan_batch_dev_async + event exchange (" + sh.exchange + ") + "
                                      "b200adsb_resolve_batch_dev_async + frame gather (" + sh.exchange + ") on a side stream")
                                     + " (steps queued back to back, outcomes checked after the timed region)")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "parity_on_sample": None if parity is None else parity["ok"], "parity": parity,
            "gpu_launches": launches_main,
            "clocks": clk.summary(),
        }
        line.update(sub)
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
