"""Model check of the peer-memory exchanges (dump1090_rs_b200/sharded.py, kernels.cuh::events_push_symm_kernel /
events_import_symm_kernel / frames_push_symm_kernel / frames_merge_kernel): blocks alternate between two
parities, flags carry epochs.  A random scheduler runs every interleaving the stream and flag dependencies
allow; a reader must always find the epoch it waited for (no peer may have overwritten the block yet).

Per rank, in program order on the main stream, for batch k = 1, 2, ...:
    wait_merged(k-2)          main waits for its own side-stream merge of gather k-2     (sharded._exchange_events;
                              redundant when every batch is gathered, see the last test)
    push_events(k)            stores into every rank's block [k&1][me], then flag
    import_events(k)          waits for all flags >= k, reads blocks [k&1][*]
    resolve(k)
and on the side stream: push_frames(k) after resolve(k); merge_frames(k) after all frame flags >= k.
"""
import random

import pytest


def simulate(world: int, batches: int, seed: int, guard: bool = True):
    rng = random.Random(seed)
    ev_block = [[[0] * world for _ in range(2)] for _ in range(world)]    # [owner][parity][src] = epoch stored
    ev_flag = [[[0] * world for _ in range(2)] for _ in range(world)]
    fr_block = [[[0] * world for _ in range(2)] for _ in range(world)]
    fr_flag = [[[0] * world for _ in range(2)] for _ in range(world)]
    main_pc = [0] * world          # index into the main-stream program
    side_pc = [0] * world
    resolved = [0] * world         # last batch resolved (main stream)
    merged = [0] * world           # last gather merged (side stream)
    main_prog = [(op, k) for k in range(1, batches + 1) for op in ("wait_merged", "push_events", "import_events", "resolve")]
    side_prog = [(op, k) for k in range(1, batches + 1) for op in ("push_frames", "merge_frames")]

    def enabled(r, stream):
        if stream == "main":
            if main_pc[r] >= len(main_prog):
                return False
            op, k = main_prog[main_pc[r]]
            if op == "wait_merged":
                return (not guard) or merged[r] >= k - 2
            if op == "import_events":
                return all(ev_flag[r][k & 1][s] >= k for s in range(world))
            return True
        if side_pc[r] >= len(side_prog):
            return False
        op, k = side_prog[side_pc[r]]
        if op == "push_frames":
            return resolved[r] >= k
        return all(fr_flag[r][k & 1][s] >= k for s in range(world))

    violations = 0
    while True:
        todo = [(r, s) for r in range(world) for s in ("main", "side") if enabled(r, s)]
        if not todo:
            break
        r, stream = rng.choice(todo)
        if stream == "main":
            op, k = main_prog[main_pc[r]]
            main_pc[r] += 1
            if op == "push_events":
                for q in range(world):
                    ev_block[q][k & 1][r] = k
                    ev_flag[q][k & 1][r] = k
            elif op == "import_events":
                violations += sum(ev_block[r][k & 1][s] != k for s in range(world))
            elif op == "resolve":
                resolved[r] = k
        else:
            op, k = side_prog[side_pc[r]]
            side_pc[r] += 1
            if op == "push_frames":
                for q in range(world):
                    fr_block[q][k & 1][r] = k
                    fr_flag[q][k & 1][r] = k
            else:
                violations += sum(fr_block[r][k & 1][s] != k for s in range(world))
                merged[r] = k
    done = all(pc == len(main_prog) for pc in main_pc) and all(pc == len(side_prog) for pc in side_pc)
    return violations, done


@pytest.mark.parametrize("world", [2, 3, 8])
def test_two_parities_suffice_with_the_merge_guard(world):
    for seed in range(60):
        violations, done = simulate(world, batches=12, seed=seed)
        assert done, "the protocol deadlocked"
        assert violations == 0, (world, seed)


def test_in_order_side_streams_already_bound_the_lead():
    """Even without the explicit guard no schedule overwrites an unread block: a rank's gather k+2 follows its
    own merge of gather k+1 on the (in-order) side stream, which needed every peer's push k+1, which follows
    that peer's merge k.  The guard in sharded.py is belt and braces (it also covers callers that skip
    gathers on some batches)."""
    for world in (2, 3):
        for seed in range(150):
            violations, done = simulate(world, batches=12, seed=seed, guard=False)
            assert done and violations == 0, (world, seed)
