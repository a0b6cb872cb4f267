"""Worker of tests/test_sharded_gpu.py (launched under torchrun, one rank per GPU): the sharded CUDA + NCCL
path -- scan, packed ICAO event all-gather, resolve, frame gather -- against the single-stream oracle,
including a batch that fails on ONE rank (candidate pool overflow) and its recovery."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import dump1090_rs_b200 as d                      # noqa: E402
from dump1090_rs_b200 import _ffi, sharded, synth  # noqa: E402
from oracle import oracle as O                     # noqa: E402

SPB = 131072


def tuples(raw):
    return [(int(r[24:28].view(np.uint32)[0]), int(r[20:24].view(np.uint32)[0]), int(r[15]),
             int(r[16:18].view(np.int16)[0]), bytes(r[: r[14]]).hex()) for r in raw]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = d.Context(local, stream.cuda_stream)
    sh = sharded.ShardedDemodulator(ctx, rank, world, frame_rows=1024, exchange=os.environ.get("B200ADSB_EXCHANGE", "auto"))
    cap = 8192
    frames = torch.zeros((cap, 28), dtype=torch.uint8, device=dev)
    merged = torch.zeros((cap, 28), dtype=torch.uint8, device=dev)
    res = torch.zeros((4, 4), dtype=torch.int32, device=dev)

    # three batches of one stream, 3 buffers per rank each; every buffer carries traffic of the same 6 aircraft
    nloc = 3
    batches = [synth.make_batch(300 + k, nloc * world, msgs_per_buffer=30, icao_pool=6) for k in range(3)]
    # batch 1: global buffer 1 (rank 1's first) is a dense periodic pattern whose survivors overflow the pool
    pat = np.array([3, 8, 11, 4, 11, 1, 4, 9, 8, 1, 11])
    levels = np.tile(pat, SPB // len(pat) + 1)[:SPB]
    batches[1] = batches[1].copy()
    batches[1][1, :, 0] = (250 * levels + 100).astype(np.int16)
    batches[1][1, :, 1] = 0
    mine = [torch.from_numpy(np.ascontiguousarray(b[rank::world])).to(dev) for b in batches]

    # reference: the single stream through the oracle (rank 0 checks)
    ref = []
    if rank == 0:
        o = O.Oracle()
        g = 0
        for b in batches:
            for k in range(b.shape[0]):
                ref.append([(k, f["j"], f["phase"], f["score"], f["msg"].hex()) for f in o.demod_iq(b[k])])
                g += 1
        per_batch = [sum(ref[3 * world * i: 3 * world * (i + 1)], []) for i in range(3)]
        members = o.members()

    def merged_frames():
        sh.fstream.synchronize()          # the gather runs on the side stream
        n = sh.n_out.cpu().numpy()
        assert n[1] == 0, n
        return tuples(merged[: int(n[0])].cpu().numpy())

    # batch 0: synchronous step (sizes the pool for ordinary traffic), gathered stream == oracle
    n0 = sh.step(mine[0].data_ptr(), nloc, SPB, SPB, frames.data_ptr(), cap)
    sh.gather_frames(frames.data_ptr(), merged.data_ptr(), cap, count=n0)
    ctx.sync()
    if rank == 0:
        assert merged_frames() == per_batch[0], "batch 0 (synchronous) differs from the oracle"
    snap0 = sorted(ctx.icao_snapshot())

    # batches 1 and 2 queued: batch 1 overflows the pool on rank 1 only -> failed on EVERY rank, batch 2 skipped
    for k in (1, 2):
        sh.step_async(mine[k].data_ptr(), nloc, SPB, SPB, frames.data_ptr(), cap, res[k].data_ptr())
    ctx.sync()
    r = res.cpu().numpy()
    flags = torch.tensor([int(r[1, 1]), int(r[2, 1])], dtype=torch.int64, device=dev)
    allf = [torch.zeros_like(flags) for _ in range(world)]
    dist.all_gather(allf, flags)
    allf = [t.cpu().tolist() for t in allf]
    assert all(f[0] != 0 for f in allf), ("batch 1 must fail on every rank", allf)
    assert all(f[1] & 8 for f in allf), ("batch 2 must be skipped on every rank", allf)
    assert allf[1][0] & 1 and allf[0][0] & 16, ("rank 1 overflowed its pool, rank 0 learned it from the exchange", allf)
    assert sorted(ctx.icao_snapshot()) == snap0, "a failed batch reached the filter"

    # recovery: the same two batches through the synchronous step (grows the pool, acknowledges the failure)
    sh.position = nloc * world
    got = []
    for k in (1, 2):
        n = sh.step(mine[k].data_ptr(), nloc, SPB, SPB, frames.data_ptr(), cap)
        sh.gather_frames(frames.data_ptr(), merged.data_ptr(), cap, count=n)
        ctx.sync()
        got.append(merged_frames())
    if rank == 0:
        assert got[0] == per_batch[1], "batch 1 after recovery differs from the oracle"
        assert got[1] == per_batch[2], "batch 2 after recovery differs from the oracle"
        assert set(ctx.icao_snapshot()) == members, "filter differs from the oracle's"
    # every rank ends with the same filter
    snap = torch.tensor(sorted(ctx.icao_snapshot()) + [0] * 64, dtype=torch.int64, device=dev)[:64]
    alls = [torch.zeros_like(snap) for _ in range(world)]
    dist.all_gather(alls, snap)
    assert all(torch.equal(alls[0], t) for t in alls), "filters differ between ranks"

    # and an enqueue-only batch commits again, with the gather queued behind it
    ctx.icao_flush()
    sh.position = 0
    sh.step_async(mine[0].data_ptr(), nloc, SPB, SPB, frames.data_ptr(), cap, res[0].data_ptr())
    sh.gather_frames(frames.data_ptr(), merged.data_ptr(), cap, result_ptr=res[0].data_ptr())
    ctx.sync()
    assert int(res.cpu().numpy()[0, 1]) == 0
    if rank == 0:
        assert merged_frames() == per_batch[0], "enqueue-only batch + gather differs from the oracle"
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_OK exchange", sh.exchange, "frames", [len(p) for p in per_batch])


if __name__ == "__main__":
    main()
