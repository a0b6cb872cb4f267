#!/usr/bin/env python3
"""bench.py -- BASELINE.json metric on B200: IQ Msamples/s through the demodulation hot path.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun)
  python bench.py --impl reference ...                   (CPU restatement of the reference)

A step = one pass of the hot path (IQ -> frames) over one batch of synthetic 2.4 Msps CS16
buffers of 131,072 samples (512 KiB).  N=1: BASELINE configs[2] (1000 buffers back to back,
device resident).  N>1: configs[4] style, the stream dealt round-robin over the ranks
(1024 buffers per rank = 8192 at N=8; weak scaling), with the ICAO add-event all-gather
between scan and resolve so that the sharded run equals the single stream.

`value`  : device-resident inputs, CUDA-event timed, max over ranks.
`e2e`    : same batch through the host-buffer C-ABI call (pinned host IQ in, frames out),
           H2D/D2H inside the timed region.
`roofline`: scan kernel, 4 algorithmic bytes per sample / CUDA-event kernel time, against
           MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: oracle/ (C restatement; the Rust reference cannot be built here) on one host
           core over a bounded sample of the same batch; also checks frame parity on it.
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
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv = pynvml
            self.source = "nvml"
        except Exception:
            self._h = None
            self.source = "nvidia-smi"

    def _poll_nvml(self):
        nv = self._nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
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


def run_reference(args) -> int:
    """--impl reference: the reference's CPU algorithm (oracle port; Rust itself cannot be built
    in this image) on all host threads, independent streams one per thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from dump1090_rs_b200 import synth
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    nb = max(cores * 4, 16)
    passes = 8            # each thread goes over its buffers several times per step: amortises thread start-up
    batch = synth.make_batch(1090, min(nb, 32))
    reps = (nb + batch.shape[0] - 1) // batch.shape[0]
    batch = np.concatenate([batch] * reps)[:nb]
    for _ in range(max(args.warmup, 1)):
        O.bench(batch, nb, SAMPLES, 1, cores, False)
    t = 0.0
    for _ in range(args.steps):
        sec, _fr = O.bench(batch, nb, SAMPLES, passes, cores, False)
        t += sec
    value = nb * passes * SAMPLES * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32 (+f32 magnitude)",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": f"synthetic 2.4Msps CS16 rtl-like noise, {nb} x 512KiB buffers x {passes} passes per step "
                               f"(bounded sample of BASELINE configs[2]), {cores} independent streams"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{nb} buffers x {passes} passes x {args.steps} steps, {cores} threads, C restatement of "
                                   "to_mag+demodulate2400 (rustc unavailable)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--buffers", type=int, default=0, help="buffers per rank (default 1000 at N=1, 1024 at N>1)")
    ap.add_argument("--msgs", type=int, default=0, help="injected DF17 per buffer in the first 16 buffers")
    ap.add_argument("--msgs-all", action="store_true", help="repeat the 16 signal buffers over the whole batch (configs[3])")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--sync-steps", action="store_true", help="time the synchronous ABI call (one host round trip per step)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import dump1090_rs_b200 as d
    from dump1090_rs_b200 import _ffi, sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nb = args.buffers or (1000 if world == 1 else 1024)

    # ---- data: this rank's share of the stream (global buffer g = b*world + rank)
    iq = synth.noise_batch_torch(1090 + rank, nb, device=dev)
    if args.msgs:
        k = min(16, nb)
        inj = synth.make_batch(1090, k, msgs_per_buffer=args.msgs, first_index=rank * 1000)
        iq[:k] = torch.from_numpy(inj).to(dev)
        if args.msgs_all:
            for b0 in range(k, nb, k):
                iq[b0:b0 + k] = iq[:min(k, nb - b0)]
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
    n_frames = [0]

    sh = sharded.ShardedDemodulator(ctx, rank, world) if world > 1 else None

    # N = 1: the timed steps are queued back to back with the enqueue-only entry point
    # (b200adsb_demod_iq_batch_dev_async): the batch outcome {frames, overflow flags, ...} stays on
    # the device and is checked after the timed region; warm-up steps use the synchronous call
    # (which also sizes the candidate pool).  --sync-steps times the synchronous call instead.
    results = torch.zeros((max(args.steps, 1), 4), dtype=torch.int32, device=dev)
    step_no = [0]
    use_async = [False]

    def step_device():
        ctx.icao_flush()
        if world == 1 and use_async[0]:
            ctx.demod_iq_batch_async_ptr(iq.data_ptr(), nb, SAMPLES, SAMPLES, frames.data_ptr(), cap,
                                         results[step_no[0] % results.shape[0]].data_ptr())
            step_no[0] += 1
        elif world == 1:
            n_frames[0] = ctx.demod_iq_batch_ptr(iq.data_ptr(), nb, SAMPLES, SAMPLES, frames.data_ptr(), cap)
        elif use_async[0]:
            sh.position = 0
            sh.step_async(iq.data_ptr(), nb, SAMPLES, SAMPLES, frames.data_ptr(), cap,
                          results[step_no[0] % results.shape[0]].data_ptr())
            step_no[0] += 1
        else:
            # scan -> all-gather of ICAO add-events (NCCL) -> resolve
            sh.position = 0
            n_frames[0] = sh.step(iq.data_ptr(), nb, SAMPLES, SAMPLES, frames.data_ptr(), cap)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.timing(reset=True)
    def timed_region(queued: bool):
        use_async[0] = queued
        step_no[0] = 0
        ctx.timing(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk_:
            barrier()
            e0.record()
            for _ in range(args.steps):
                step_device()
            e1.record()
            barrier()
        ms_ = e0.elapsed_time(e1)
        ctx.sync()
        tim_ = ctx.timing(reset=True)
        ok = True
        if queued:
            res = results.cpu().numpy()
            ok = bool((res[:, 1] == 0).all() and (res[:, 3] == 0).all() and (res[:, 0] == n_frames[0]).all())
            if not ok:
                print("bench: a queued batch overflowed or disagreed with the synchronous call: %r" % (res.tolist(),),
                      file=sys.stderr)
        use_async[0] = False
        return ms_, tim_, clk_, ok

    queued_steps = not args.sync_steps
    ms, tim, clk, ok = timed_region(queued_steps)
    if dist is not None:   # every rank must take the same path
        flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        ok = int(flag.item()) == 0
    if not ok:             # never report an unverified number: time the synchronous call instead
        queued_steps = False
        ms, tim, clk, ok = timed_region(False)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_samples = nb * SAMPLES * world
    value = total_samples * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (scan): algorithmic bytes = 4 B per IQ sample
    SCAN_KERNEL = {"6": "scan_kernel<false>"}.get(
        os.environ.get("B200ADSB_SCAN", "7"), "scan7_kernel<false>")
    peak, peak_src = peaks()
    # (the scan kernel is launched once per chunk of tiles; sum over the timed region)
    scan_launches = max(tim["scan_launches"], 1)
    scan_ms = tim["scan_ms"] / scan_launches
    alg_bytes = 4.0 * tim["samples"] / scan_launches
    achieved = 4.0 * tim["samples"] / (tim["scan_ms"] * 1e-3) / 1e9 if tim["scan_ms"] > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
            traffic = tj.get("dram_bytes_per_sample", None)
            if traffic is not None:
                traffic = traffic * tim["samples"] / scan_launches
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "kernel": SCAN_KERNEL, "kernel_ms_per_launch": scan_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "kernel_share_of_step": (tim["scan_ms"] / ms) if ms else None,
                "kernel_ms_per_step": tim["scan_ms"] / args.steps,
                "decode_kernel_ms_per_step": tim.get("decode_ms", 0.0) / args.steps,
                "resolve_kernels_ms_per_step": tim["resolve_ms"] / args.steps}

    # ---- e2e: host-buffer C-ABI call, H2D + D2H inside the timed region
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
        k_e2e = max(3, args.steps // 2)
        for _ in range(k_e2e):
            ectx.icao_flush()
            nf = ectx.demod_iq_batch_ptr(host_iq.data_ptr(), nb, SAMPLES, SAMPLES, host_frames.data_ptr(), cap, host=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": total_samples * k_e2e / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": nb * SAMPLES * 4, "d2h_bytes_per_step": nf * 28 + 64,
               "steps": k_e2e, "frames_per_step": nf}
        ectx.close()
        del host_iq

    # ---- CPU baseline (rank 0, N=1): oracle on a bounded sample + parity check on it
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        ns = min(32, nb)
        sample = iq[:ns].cpu().numpy()
        # parity on the sample (first buffers of the stream; filter state is a prefix property)
        o = O.Oracle()
        ref = []
        for b in range(min(8, ns)):
            ref += [(b, f["j"], f["phase"], f["score"], f["msg"].hex()) for f in o.demod_iq(sample[b])]
        raw = frames[: n_frames[0]].cpu().numpy()
        got = [(int(r[24:28].view(np.uint32)[0]), int(r[20:24].view(np.uint32)[0]), int(r[15]),
                int(r[16:18].view(np.int16)[0]), bytes(r[: r[14]]).hex()) for r in raw
               if int(r[24:28].view(np.uint32)[0]) < min(8, ns)]
        parity = got == ref
        iters = 120          # ~10 s of single-core work
        O.bench(sample, ns, SAMPLES, 1, 1, False)
        sec, _fr = O.bench(sample, ns, SAMPLES, iters, 1, False)
        cpu = {"value": ns * SAMPLES * iters / sec / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first {ns} buffers of the batch x {iters} passes, 1 thread, C restatement "
                         "(oracle/; rustc unavailable so the Rust reference cannot be built)",
               "host_cpus": os.cpu_count(), "parity_on_sample": parity}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u16/i32 (+f32 magnitude)",
            "data": "synthetic",
            "config": {"workload": (f"synthetic 2.4Msps CS16 rtl-like noise (sigma 5.5 LSB of 8 bit), {nb} x 512KiB "
                                    f"buffers per GPU per step, {'BASELINE configs[2]' if world == 1 else 'configs[4] round-robin shards + ICAO event all-gather'}"),
                       "buffers_per_gpu": nb, "samples_per_buffer": SAMPLES, "injected_msgs": args.msgs, "injected_in_all_buffers": bool(args.msgs_all),
                       "l2": f"inputs {nb * SAMPLES * 4 / 2**20:.0f} MiB per GPU > 126 MB L2 (no flush needed)",
                       "frames_per_step": n_frames[0],
                       "step_call": ("synchronous ABI calls (host round trips inside every step)" if not queued_steps else
                                     ("b200adsb_demod_iq_batch_dev_async" if world == 1 else
                                      "b200adsb_scan_batch_dev_async + events all-gather + b200adsb_resolve_batch_dev_async")
                                     + " (steps queued back to back, outcomes checked after the timed region)")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(tim["scan_launches"] + tim["other_launches"]),
            "clocks": clk.summary(),
        }
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
