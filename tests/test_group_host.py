"""The render group's rendezvous, barrier and host-frame handshake (csrc/rt_group.cu) on CPU: world_size 2 and 3, real processes,
the real library — only the pixels are stand-ins (no CUDA device here, so the group is created without a context). The GPU side of
the same protocol is tests/test_group_gpu.py."""
import ctypes as C
import multiprocessing as mp
import os
import time

import numpy as np
import pytest

import partition_model as partition

W, H, BR = 96, 52, 8           # 52 rows: 7 bands, the last one short (4 rows) -> ragged shares
FRAMES = 7


def _pixel_pattern(rows, frame_no):
    yy = rows[:, None].astype(np.int64)
    xx = np.arange(W)[None, :]
    v = (yy * W + xx) * 2654435761 + frame_no * 97
    return np.stack([(v >> s) & 0xFF for s in (0, 8, 16, 24)], axis=-1).astype(np.uint8)


def _worker(rank, world, name, q, slow_rank):
    try:
        from build_up_phase_b200 import rtcore
        g = rtcore.Group(None, name, rank, world, W, H)
        assert (g.L.rt_group_rank(g.h), g.L.rt_group_world(g.h)) == (rank, world)
        rows = partition.local_rows_of_part(H, BR, rank, world)
        rows = rows[rows >= 0]
        seen = []
        for f in range(FRAMES):
            if rank == slow_rank:
                time.sleep(0.01 * (f % 3))                           # ranks drift apart: the handshake has to hold them together
            p = g.host_frame_begin()
            frame = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(H, W, 4))
            frame[rows] = _pixel_pattern(rows, f)
            out = g.host_frame_end()
            if rank == 0:
                full = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(H, W, 4))
                if f % 2 == 0:
                    time.sleep(0.02)                                 # rank 0's caller is slow to consume: nobody may overwrite the frame meanwhile
                seen.append(bool(np.array_equal(full, _pixel_pattern(np.arange(H), f))))
            else:
                assert out is None
        g.barrier()
        g.barrier()
        g.close()
        q.put((rank, seen, None))
    except Exception as e:       # noqa: BLE001
        q.put((rank, [], repr(e)))


@pytest.mark.parametrize("world,slow", [(2, 1), (3, 0), (3, 2)])
def test_host_frame_handshake_real_processes(world, slow):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = f"pytest-{os.getpid()}-{world}-{slow}"
    ps = [ctx.Process(target=_worker, args=(r, world, name, q, slow)) for r in range(world)]
    for p in ps:
        p.start()
    res = {}
    for _ in range(world):
        r, seen, err = q.get(timeout=120)
        res[r] = (seen, err)
    for p in ps:
        p.join(timeout=30)
    assert all(res[r][1] is None for r in res), res
    assert res[0][0] == [True] * FRAMES
    assert not any(f.startswith("rtcore.pytest-") for f in os.listdir("/dev/shm")), "the group must not leave objects behind in /dev/shm"


def test_group_of_one_and_bad_arguments():
    from build_up_phase_b200 import rtcore
    g = rtcore.Group(None, f"solo-{os.getpid()}", 0, 1, 16, 8)
    p = g.host_frame_begin()
    assert p
    assert g.host_frame_end() == p
    p2 = g.host_frame_begin()
    assert p2 != p                                                    # two frames alternate
    g.host_frame_end()
    g.barrier()
    g.close()
    with pytest.raises(rtcore.RtError):
        rtcore.Group(None, "bad", 2, 2, 16, 8)                        # rank out of range
