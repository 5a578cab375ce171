"""Model check of the multi-GPU frame handshake: every rank stores its pixels straight into one of two
frames owned by rank 0 and signals through two stream-ordered counters per frame,
    wait(free[k] >= use)  trace  add(done[k])            on every rank
    wait(done[k] >= N * (use + 1))  consume  add(free[k])  on rank 0,
where `use` counts the previous uses of buffer k. The model runs N ranks under random interleavings (each rank's operations stay in
stream order) and checks what the protocol has to guarantee: no deadlock, a frame is consumed only after ALL ranks wrote it, and no
rank overwrites a buffer that rank 0 has not consumed yet. CPU-only; the GPU implementation is rt_flag_add / rt_flag_wait_ge."""
import random

import pytest


def _ops(rank, world, steps):
    ops = []
    for s in range(steps):
        k, use = s & 1, s >> 1
        ops.append(("wait_free", k, use))
        ops.append(("trace", k, s))
        ops.append(("add_done", k))
        if rank == 0:
            ops.append(("wait_done", k, world * (use + 1)))
            ops.append(("consume", k, s))
            ops.append(("add_free", k))
    return ops


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_handshake_model(world, seed):
    rng = random.Random(1000 * world + seed)
    steps = 9
    queues = [_ops(r, world, steps) for r in range(world)]
    pc = [0] * world
    done, free = [0, 0], [0, 0]
    written = [dict() for _ in range(2)]        # buffer -> {rank: step last written}
    consumed = [-1, -1]                         # buffer -> last step consumed
    progress = 0
    while any(pc[r] < len(queues[r]) for r in range(world)):
        runnable = []
        for r in range(world):
            if pc[r] >= len(queues[r]):
                continue
            op = queues[r][pc[r]]
            if op[0] == "wait_free" and free[op[1]] < op[2]:
                continue
            if op[0] == "wait_done" and done[op[1]] < op[2]:
                continue
            runnable.append(r)
        assert runnable, f"deadlock: pc={pc} done={done} free={free}"
        r = rng.choice(runnable)
        op = queues[r][pc[r]]
        pc[r] += 1
        progress += 1
        if op[0] == "trace":
            k, s = op[1], op[2]
            # the previous frame in this buffer (step s - 2) must have been consumed before anybody overwrites it
            assert s < 2 or consumed[k] >= s - 2, f"rank {r} overwrites buffer {k} (step {s}) before step {s - 2} was consumed"
            written[k][r] = s
        elif op[0] == "add_done":
            done[op[1]] += 1
        elif op[0] == "consume":
            k, s = op[1], op[2]
            assert all(written[k].get(q) == s for q in range(world)), f"frame {s} consumed before every rank wrote it: {written[k]}"
            consumed[k] = s
        elif op[0] == "add_free":
            free[op[1]] += 1
    assert consumed == [steps - 1 if (steps - 1) % 2 == 0 else steps - 2, steps - 1 if (steps - 1) % 2 == 1 else steps - 2]
    assert progress == sum(len(q) for q in queues)
