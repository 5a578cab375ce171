"""Host-side logic of the N>1 path on CPU: world_size-2/3 gloo gather of packed band buffers."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import partition_model as partition


@pytest.mark.parametrize("h,br,n", [(2160, 8, 8), (250, 8, 3), (17, 4, 2), (8, 8, 4), (1, 4, 2)])
def test_every_row_owned_once(h, br, n):
    seen = np.zeros(h, dtype=int)
    for p in range(n):
        rows = partition.local_rows_of_part(h, br, p, n)
        assert len(rows) == partition.packed_rows(h, br, n)
        seen[rows[rows >= 0]] += 1
    assert np.all(seen == 1)


def test_matches_c_abi(rt):
    L = rt.load()
    for (w, h, br, n) in [(3840, 2160, 8, 8), (322, 250, 8, 3), (7680, 4320, 8, 4), (5, 1, 4, 2)]:
        assert int(L.rt_rows_packed_pixels(w, h, br, n)) == partition.packed_pixels(w, h, br, n)


def _worker(rank, world, port, w, h, br, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = partition.local_rows_of_part(h, br, rank, world)
    packed = np.zeros((len(rows), w, 4), dtype=np.uint8)
    ok = rows >= 0
    # stand-in for the trace kernel: pixel value = f(global pixel index)
    yy = rows[ok][:, None].astype(np.int64)
    xx = np.arange(w)[None, :]
    idx = yy * w + xx
    packed[ok] = np.stack([(idx >> s) & 0xFF for s in (0, 8, 16, 24)], axis=-1).astype(np.uint8)
    t = torch.from_numpy(packed)
    gathered = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    if rank == 0:
        allp = np.stack([g.numpy() for g in gathered])
        frame = partition.unpack(allp, w, h, br, world)
        ref_idx = np.arange(h * w, dtype=np.int64).reshape(h, w)
        ref = np.stack([(ref_idx >> s) & 0xFF for s in (0, 8, 16, 24)], axis=-1).astype(np.uint8)
        q.put(bool(np.array_equal(frame, ref)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,port", [(2, 29611), (3, 29613)])
def test_gloo_gather_unpack(world, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 67, 53, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
