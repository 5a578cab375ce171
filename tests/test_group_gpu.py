"""The render group on the GPU (csrc/rt_group.cu): real processes, one rt_context each — on a one-GPU box they all use cuda:0, which
exercises the same CUDA-IPC mappings, stream-ordered flags, shared pinned host frame and BLAS pulls as one process per GPU does.
Every assembled frame must equal the frame one context renders alone, bit for bit (which the parity tests tie to the oracle)."""
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 322, 250          # 32 bands, the last one short (2 rows); not a multiple of the tile size either


def _scene():
    from build_up_phase_b200 import scenes
    return scenes.instanced_scene(n_side=4, quads=12, width=W, height=H, bounces=1)


def _worker(rank, world, name, q, n_dev):
    try:
        import torch
        from build_up_phase_b200 import rtcore, scenes
        dev = rank % n_dev
        torch.cuda.set_device(dev)
        scene = _scene()
        ctx = rtcore.Context(dev)
        g = rtcore.Group(ctx, name, rank, world, W, H)
        # cfg5 in miniature: the BLASes are built by their owners (blas % world) and pulled by everybody else
        blases = []
        for b, geoms in enumerate(scene.blases):
            owner = b % world
            mine = ctx.build_blas(geoms) if owner == rank else None
            blases.append(g.share_blas(b % 64, owner, mine))
            if b % 64 == 63:
                g.share_finish()
        share_ms = g.share_finish()
        tlas = ctx.build_tlas(scene.instances, blases)
        ctx.set_hit_records(scene.hit_records); ctx.set_miss_color(scene.miss_color)
        cam = ctx.camera(scene.camera_pos, scene.yfov_deg)
        out = {"share_ms": share_ms}
        frames = []
        # device frame on rank 0, every rank's kernel storing into it; synchronous, then pipelined over two streams
        for i in range(3):
            p = g.trace(tlas, cam, W, H, 1, rtcore.GROUP_OUT_DEVICE)
            if rank == 0:
                frames.append(rtcore.device_view(p, W * H * 4, f"cuda:{dev}").view(H, W, 4).cpu().numpy().copy())
        ptrs = []
        for i in range(5):
            ptrs.append(g.trace(tlas, cam, W, H, 1, rtcore.GROUP_OUT_DEVICE | rtcore.GROUP_PIPELINE))
        g.sync()
        if rank == 0:
            frames.append(rtcore.device_view(ptrs[-1], W * H * 4, f"cuda:{dev}").view(H, W, 4).cpu().numpy().copy())
            frames.append(rtcore.device_view(ptrs[-2], W * H * 4, f"cuda:{dev}").view(H, W, 4).cpu().numpy().copy())
        # shared pinned host frame, every rank copying its own bands
        for i in range(3):
            f = g.trace_host(tlas, cam, W, H, 1)
            if rank == 0:
                frames.append(f.copy())
        # pipelined host output: two frames in flight, call k returns frame k - 1; a different camera per frame shows which frame came back
        cams = [ctx.camera((0.3 * i, -0.2 * i, 10.0 + i), scene.yfov_deg) for i in range(5)]
        piped = []
        for i in range(5):
            f = g.trace_host(tlas, cams[i], W, H, 1, pipeline=True)
            if rank == 0:
                assert (f is None) == (i == 0)
                if f is not None:
                    piped.append(f.copy())
        f = g.flush_host(W, H)
        if rank == 0:
            piped.append(f.copy())
        assert g.flush_host(W, H) is None                                  # nothing pending any more
        # a smaller frame than the group's maximum
        f = g.trace_host(tlas, cam, 160, 96, 0)
        small = f.copy() if rank == 0 else None
        if rank == 0:
            ctx1 = rtcore.Context(dev)
            sh = rtcore.SceneHandles(ctx1, scene)
            ref, _, _ = sh.trace(want_hits=False)
            ref_small, _, _ = ctx1.trace(sh.tlas, sh.cam, 160, 96, 0)
            out["equal"] = [bool(np.array_equal(fr, ref)) for fr in frames]
            out["piped_equal"] = [bool(np.array_equal(piped[i], ctx1.trace(sh.tlas, cams[i], W, H, 1)[0])) for i in range(5)]
            out["piped_distinct"] = len({fr.tobytes() for fr in piped})
            sh.free(); ctx1.close()
            out["small_equal"] = bool(np.array_equal(small, ref_small))
            out["nonblack"] = int((ref[..., :3].max(axis=-1) > 60).sum())
        g.barrier()
        tlas.free()
        for b, h in enumerate(blases):
            h.free()
        g.close()
        ctx.close()
        q.put((rank, out, None))
    except Exception as e:       # noqa: BLE001
        import traceback
        q.put((rank, {}, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_group_frames_equal_single_context(world):
    import torch
    n_dev = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = f"pytest-gpu-{os.getpid()}-{world}"
    ps = [ctx.Process(target=_worker, args=(r, world, name, q, n_dev)) for r in range(world)]
    for p in ps:
        p.start()
    res = {}
    for _ in range(world):
        r, out, err = q.get(timeout=600)
        res[r] = (out, err)
    for p in ps:
        p.join(timeout=60)
    assert all(res[r][1] is None for r in res), "\n".join(str(res[r][1]) for r in res if res[r][1])
    out = res[0][0]
    assert out["nonblack"] > 5000
    assert out["equal"] == [True] * 8, out           # 3 device frames, 2 pipelined, 3 host frames
    assert out["small_equal"]
    assert out["piped_equal"] == [True] * 5 and out["piped_distinct"] == 5, out      # frame k - 1 really is frame k - 1
    print("group", world, out, {r: res[r][0].get("share_ms") for r in res})


def test_cpp_multi_gpu_host_program(oracle, tmp_path):
    """host/sample_scene_mgpu.cpp — the reference's main() as one process per GPU above rt_group_* (no Python, no NCCL): the sample
    scene's frame equals the oracle's; a scene file (.rtscene) of an instanced scene with a bounce renders to the same CRC at 1, 2 and
    3 processes, through the shared pinned host frame, through rank 0's device frame, and with BLASes built by their owners only and
    pulled over NVLink by the others."""
    import re
    import subprocess
    from build_up_phase_b200 import build as b, scenes
    exe = b.build_host_sample("sample_scene_mgpu")
    out = tmp_path / "f.ppm"

    def run(*args):
        p = subprocess.run([exe, str(out), *args], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr + p.stdout
        m = re.search(r"crc32 ([0-9a-f]{8})", p.stdout)
        assert m, p.stdout
        return m.group(1), p.stdout

    crc1, _ = run("--gpus", "1", "--size", "600", "400")
    raw = out.read_bytes()
    hdr = b"P6\n600 400\n255\n"
    img = np.frombuffer(raw[len(hdr):], dtype=np.uint8).reshape(400, 600, 3)
    o = oracle.OracleScene(scenes.sample_scene(600, 400))
    ref = o.trace(mode=oracle.MODE_BRUTE)[0]
    o.close()
    assert np.abs(img.astype(np.int16) - ref[:, :, :3].astype(np.int16)).max() <= 1
    assert run("--gpus", "2", "--size", "600", "400")[0] == crc1
    assert run("--gpus", "3", "--size", "600", "400", "--device-frame")[0] == crc1
    scene = _scene()
    path = tmp_path / "inst.rtscene"
    scenes.save_scene(scene, str(path))
    c1, text = run("--gpus", "1", "--scene", str(path))
    assert f"{scene.triangle_count} triangles in 16 BLAS, 16 instances, 1 bounce" in text
    for args in (("--gpus", "2"), ("--gpus", "3", "--frames", "4"), ("--gpus", "2", "--device-frame", "--frames", "5"), ("--gpus", "3", "--split-build"),
                 ("--gpus", "2", "--split-build", "--device-frame")):
        assert run("--scene", str(path), *args)[0] == c1, args
    # the frame the C++ host wrote is the frame the Python binding renders
    from build_up_phase_b200 import rtcore
    with rtcore.Context(0) as ctx:
        sh = rtcore.SceneHandles(ctx, scene)
        ref, _, _ = sh.trace(want_hits=False)
        sh.free()
    raw = out.read_bytes()
    hdr = f"P6\n{W} {H}\n255\n".encode()
    img = np.frombuffer(raw[len(hdr):], dtype=np.uint8).reshape(H, W, 3)
    assert np.array_equal(img, ref[:, :, :3])
