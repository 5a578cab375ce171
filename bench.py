#!/usr/bin/env python
"""bench.py — the driver-facing benchmark of the hot path (BASELINE.json metric).

A "step" is one vkCmdTraceRaysKHR-equivalent pass over the full frame of the workload (primary rays +
one diffuse bounce), with the acceleration structures already resident in HBM. At N GPUs the frame is
split into interleaved 8-scanline bands (scene replicated); every rank's trace kernel stores its pixels straight into
rank 0's framebuffer over NVLink (--gather p2p, default) or the packed bands are gathered with NCCL and unpacked.

  python bench.py                         # N=1, inst10m (BASELINE configs[3], the config the metric is quoted on)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N
  python bench.py --impl reference        # the CPU arm: the scalar oracle on all host cores, bounded sample

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from build_up_phase_b200 import scenes  # noqa: E402

BLOCK_ROWS = 8
B_TRI_BUILD = 356.0          # algorithmic bytes per triangle of the LBVH build (SURVEY §8(d), 32-bit-key contract figure)


SOUP_SIZES = {"soup100m": (100_000_000, 7680, 4320), "soup10m": (10_000_000, 3840, 2160), "soup1m": (1_000_000, 1920, 1080)}


def make_workload(name: str, width: int | None, height: int | None, only_parts=None, soup_split: str = "slab"):
    if name == "inst10m":
        s = scenes.instanced_scene(32, 70, 3840, 2160, 1)
    elif name == "inst640k":                       # CPU-container-sized variant of the same generator
        s = scenes.instanced_scene(8, 70, 1920, 1080, 1)
    elif name == "tess1m":
        s = scenes.tess_scene(1000, 500, 3840, 2160, 1)
    elif name == "sample":
        s = scenes.sample_scene(1920, 1080)
    elif name in SOUP_SIZES:
        n, w, h = SOUP_SIZES[name]
        s = scenes.soup_scene(n, w, h, 0, only_parts=only_parts, split=soup_split)
    else:
        raise SystemExit(f"unknown workload {name}")
    if width:
        s.width = width
    if height:
        s.height = height
    return s


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def b_ray_bytes(st: dict, pixels: int) -> float:
    """B_ray = 64*nodes + 56*triangles + 64*instances + 4 per pixel."""
    return 64.0 * st["nodes_visited"] + 56.0 * st["triangles_tested"] + 64.0 * st["instances_entered"] + 4.0 * pixels


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle): only here and in cpu_baseline may bench.py execute anything under oracle/
# ------------------------------------------------------------------------------------------------
def cpu_arm(scene, target_seconds: float, steps: int, warmup: int):
    import oracle_binding as ob
    t0 = time.time()
    o = ob.OracleScene(scene, build_bvh=True)
    setup_s = time.time() - t0
    w, h = scene.width, scene.height
    # calibrate the row stride so that one step is ~target_seconds of CPU work
    probe_rows = (0, h, max(1, h // 16))
    t0 = time.time()
    _, _, _, st = o.trace(rows=probe_rows, want_hits=False)
    dt = time.time() - t0
    rays_probe = st["rays_primary"] + st["rays_secondary"]
    rate = rays_probe / max(dt, 1e-9)
    total_rays_est = rays_probe * (h / max(1, len(range(*probe_rows))))
    stride = max(1, int(np.ceil(total_rays_est / max(rate * target_seconds, 1.0))))
    rows = (0, h, stride)
    times, rays = [], 0
    for i in range(warmup + steps):
        t0 = time.time()
        _, _, _, st = o.trace(rows=rows, want_hits=False)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
            rays = st["rays_primary"] + st["rays_secondary"]
    ms = 1000.0 * float(np.mean(times))
    return {"mrays_per_s": rays / (ms * 1e-3) / 1e6, "ms_per_step": ms, "rays_per_step": rays, "rows": rows, "threads": ob.num_threads(),
            "setup_s": setup_s, "oracle": o}


def cpu_build_baseline(scene, max_tris: int = 1_500_000):
    """CPU LBVH build (std::stable_sort + Karras + atomic refit, OpenMP) of a bounded subset of the BLASes."""
    import oracle_binding as ob
    tris, secs = 0, 0.0
    for geoms in scene.blases:
        n = sum(g.triangle_count for g in geoms)
        if tris and tris + n > max_tris:
            break
        secs += ob.time_blas_build(geoms)
        tris += n
    return {"mtris_per_s": tris / max(secs, 1e-9) / 1e6, "triangles": tris}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = make_workload(args.workload, args.width, args.height, soup_split=args.soup_split)
    steps = args.steps if args.steps else 2
    warmup = args.warmup if args.warmup is not None else 1
    res = cpu_arm(scene, args.cpu_seconds, steps, warmup)
    line = {
        "impl": "reference", "metric": "Mrays/s (primary+secondary rays per second)", "value": res["mrays_per_s"], "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(scene, args, args.gpus),
        "cpu_baseline": {"value": res["mrays_per_s"], "unit": "Mrays/s", "cores": res["threads"], "kind": "port",
                         "sample": f"rows {res['rows'][0]}..{res['rows'][1]} step {res['rows'][2]} of the {scene.width}x{scene.height} frame "
                                   f"({res['rays_per_step']} rays/step), oracle LBVH traversal, OpenMP"},
        "e2e": {"value": res["mrays_per_s"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def ncu_traffic(kind: str, workload: str):
    """Measured DRAM bytes per launch from the committed ncu summary (profiles/ncu_traffic.json, written by
    tools/summarise_profile.py from one `ncu --set full` capture); None when there is no capture for this workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kind]
        if t["workload"] != workload:
            return None, None
        return t.get("dram_bytes_per_frame", t.get("dram_bytes_per_build")), t["source"]
    except Exception:
        return None, None


def ncu_issue(workload: str):
    """The bound that actually limits the trace kernels (they are latency/issue bound, not HBM bound): issue-slot utilisation, lanes
    per instruction and cache hit rates of both stages from the committed ncu capture; None without one."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k_trace"]
        if t["workload"] != workload:
            return None
        return {k: t[k] for k in ("issue_slot_utilisation_pct", "lanes_per_instruction", "l1_hit_pct", "l2_hit_pct", "source") if k in t}
    except Exception:
        return None


def workload_config(scene, args, n_gpus):
    n_tris = SOUP_SIZES[args.workload][0] if args.workload in SOUP_SIZES else scene.triangle_count
    return {"workload": f"{args.workload}: {scene.name}, {n_tris} triangles in {len(scene.blases)} BLAS, "
                        f"{len(scene.instances)} instances, {scene.width}x{scene.height} primary + {scene.bounces} diffuse bounce",
            "triangles": n_tris, "instances": len(scene.instances), "width": scene.width, "height": scene.height,
            "bounces": scene.bounces, "partition": f"{BLOCK_ROWS}-scanline bands interleaved over {n_gpus} GPU(s), scene replicated"
            + ("" if n_gpus == 1 else {"p2p": "; every rank stores its pixels straight into rank 0's framebuffer over NVLink (CUDA IPC), frame barrier: " + ("stream-ordered counters in that buffer" if args.barrier == "flags" else "NCCL all-reduce"),
                                       "allgather": "; packed bands, NCCL all-gather, unpack on rank 0",
                                       "gather": "; packed bands, NCCL gather to rank 0, unpack"}[args.gather]),
            "l2_policy": "inputs larger than L2 (BVH nodes + triangles >> 126 MB); no flush between iterations"
                         if n_tris * 112 > 2 * 126e6 else "scene fits in L2: numbers are L2-resident (parity config)"}


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from build_up_phase_b200 import rtcore

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps = args.steps if args.steps else 20
    warmup = args.warmup if args.warmup is not None else 3
    if warmup < 3:
        warmup = 3

    is_soup = args.workload in SOUP_SIZES
    my_parts = [p for p in range(scenes.SOUP_PARTS) if p % world == rank] if is_soup else None
    scene = make_workload(args.workload, args.width, args.height, only_parts=my_parts, soup_split=args.soup_split)
    W, H, bounces = scene.width, scene.height, scene.bounces
    ctx = rtcore.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    # ---- scene upload (untimed) and acceleration-structure build (timed separately: Mtri/s) ----
    def upload(geoms):
        dg = []
        for g in geoms:
            v = torch.from_numpy(np.ascontiguousarray(g.vertices)).to(dev)
            i = torch.from_numpy(np.ascontiguousarray(g.indices).view(np.int32)).to(dev) if g.indices is not None else None
            t = torch.from_numpy(np.ascontiguousarray(g.transform)).to(dev) if g.transform is not None else None
            dg.append(scenes.Geometry(v, i, t))
        return dg

    n_tris_total = SOUP_SIZES[args.workload][0] if is_soup else scene.triangle_count
    bcast_ms = 0.0
    if not is_soup:
        dev_blases = [upload(geoms) for geoms in scene.blases]
        torch.cuda.synchronize()
        build_ms = []
        blases = None
        for rep in range(args.build_reps + 1):
            if blases is not None:
                for b in blases:
                    b.free()
            blases = ctx.build_blas_batch(dev_blases, device=True) if len(dev_blases) > 1 else [ctx.build_blas(dev_blases[0], device=True)]
            if rep > 0:
                build_ms.append(ctx.build_timing())
        bt = min(build_ms, key=lambda t: t["total_ms"])
        build_note = "one batched build of all BLASes on every GPU (scene replicated)"
    else:
        # cfg5: the soup is 8 index ranges = 8 BLASes; rank r builds parts p with p % N == r, then every BLAS
        # blob is broadcast (NCCL over NVLink) from its builder and adopted by the other ranks.
        own = {}
        phase_keys = ("total_ms", "setup_ms", "morton_ms", "sort_ms", "hierarchy_ms", "refit_ms")
        bt = {k: 0.0 for k in phase_keys}
        for p in my_parts:
            dg = upload(scene.blases[p])
            torch.cuda.synchronize()
            best, b = None, None
            for rep in range(args.build_reps + 1):
                if b is not None:
                    b.free()
                b = ctx.build_blas(dg, device=True)
                t = ctx.build_timing()
                if rep > 0 or args.build_reps == 0:
                    best = t if best is None or t["total_ms"] < best["total_ms"] else best
            own[p] = b
            for k in phase_keys:
                bt[k] += best[k]
            del dg
        bt["primitives"] = n_tris_total
        infos = {p: own[p].info() for p in my_parts}
        blases = [None] * scenes.SOUP_PARTS
        for p in my_parts:
            blases[p] = own[p]
        if world > 1:
            tms = torch.tensor([bt[k] for k in phase_keys], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)          # the slowest rank's build defines the build time
            bt.update(dict(zip(phase_keys, [float(x) for x in tms.tolist()])))
            meta = [None] * scenes.SOUP_PARTS
            for p in range(scenes.SOUP_PARTS):
                obj = [None]
                if p in infos:
                    i = infos[p]
                    obj = [dict(triangle_count=i.triangle_count, node_count=i.node_count, root_ref=i.root_ref, max_depth=i.max_depth,
                                lo=list(i.bounds_lo), hi=list(i.bounds_hi), storage_bytes=i.storage_bytes)]
                dist.broadcast_object_list(obj, src=p % world)
                meta[p] = obj[0]
            bufs = {}
            for p in range(scenes.SOUP_PARTS):
                if p in own:
                    bufs[p] = rtcore.device_view(infos[p].device_storage, int(infos[p].storage_bytes), dev)
                else:
                    bufs[p] = torch.empty(int(meta[p]["storage_bytes"]), dtype=torch.uint8, device=dev)
            dist.barrier(); torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            for p in range(scenes.SOUP_PARTS):
                dist.broadcast(bufs[p], src=p % world)
            b1.record()
            torch.cuda.synchronize()
            bc = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
            dist.all_reduce(bc, op=dist.ReduceOp.MAX)
            bcast_ms = float(bc.item())
            for p in range(scenes.SOUP_PARTS):
                if p not in own:
                    m = meta[p]
                    info = rtcore.RtBlasInfo()
                    info.triangle_count, info.node_count, info.root_ref, info.max_depth = m["triangle_count"], m["node_count"], m["root_ref"], m["max_depth"]
                    for k in range(3):
                        info.bounds_lo[k], info.bounds_hi[k] = m["lo"][k], m["hi"][k]
                    info.storage_bytes = m["storage_bytes"]
                    blases[p] = ctx.import_blas(info, bufs[p])
            del bufs
        build_note = (f"{scenes.SOUP_PARTS} BLASes of {n_tris_total // scenes.SOUP_PARTS} triangles, {len(my_parts)} built per GPU, "
                      f"blobs broadcast with NCCL ({bcast_ms:.2f} ms); build time = slowest rank's builds + broadcast")
    tlas = ctx.build_tlas(scene.instances, blases)
    tlas_t = ctx.build_timing()
    ctx.set_hit_records(scene.hit_records)
    ctx.set_miss_color(scene.miss_color)
    cam = ctx.camera(scene.camera_pos, scene.yfov_deg)

    # ---- buffers ----
    px_packed = ctx.rows_packed_pixels(W, H, BLOCK_ROWS, world)
    mode = args.gather if world > 1 else "single"
    shared, shared_ptrs, token = [], [], None
    if world > 1:
        packed = torch.zeros((px_packed, 4), dtype=torch.uint8, device=dev)
        if mode == "gather":
            gathered = torch.zeros((world, px_packed, 4), dtype=torch.uint8, device=dev) if rank == 0 else None
            gather_list = [gathered[r] for r in range(world)] if rank == 0 else None
        elif mode == "allgather":
            gathered = torch.zeros((world, px_packed, 4), dtype=torch.uint8, device=dev)
        else:
            # two framebuffers on rank 0 (double buffer: frame k may still be read while frame k + 1 is written), mapped by every rank
            handles = [None, None]
            ok = 1
            try:
                if os.environ.get("RTCORE_BENCH_NO_IPC"):
                    raise RuntimeError("disabled by RTCORE_BENCH_NO_IPC (test hook of the fallback)")
                if rank == 0:
                    for k in range(2):
                        ptr, h = ctx.frame_share_create(W * H * 4 + 256)      # + the frame's "done" and "free" counters
                        shared_ptrs.append(ptr); handles[k] = h
            except Exception as e:      # noqa: BLE001
                sys.stderr.write(f"[bench] shared framebuffer unavailable on rank 0 ({e}); falling back to the NCCL gather\n")
                ok = 0
            dist.broadcast_object_list(handles, src=0)
            try:
                if rank != 0 and ok and handles[0] is not None:
                    shared_ptrs = [ctx.frame_share_open(h) for h in handles]
                elif rank != 0:
                    ok = 0
            except Exception as e:      # noqa: BLE001
                sys.stderr.write(f"[bench] rank {rank} cannot map rank 0's framebuffer ({e}); falling back to the NCCL gather\n")
                ok = 0
            agree = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            if int(agree.item()) == 0:          # CUDA IPC not usable here: every rank takes the packed-bands + NCCL path instead
                for p_ in shared_ptrs:
                    try:
                        (ctx.frame_share_free if rank == 0 else ctx.frame_share_close)(p_)
                    except Exception:   # noqa: BLE001
                        pass
                shared_ptrs = []
                mode = "gather"
                args.gather = "gather"
                gathered = torch.zeros((world, px_packed, 4), dtype=torch.uint8, device=dev) if rank == 0 else None
                gather_list = [gathered[r] for r in range(world)] if rank == 0 else None
            else:
                if rank == 0:
                    shared = [rtcore.device_view(p, W * H * 4, dev).view(H, W, 4) for p in shared_ptrs]
                token = torch.zeros(1, dtype=torch.int32, device=dev)
    frame = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
    host_frame = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
    host_np = host_frame.numpy()
    step_no = [0]
    frame_uses = [0, 0]

    def step_multi(consume=None):
        """One frame at N > 1; on rank 0 `consume(frame)` is enqueued once the frame is complete. Returns the frame tensor on rank 0."""
        if mode == "p2p":
            k, use = step_no[0] & 1, step_no[0] >> 1
            step_no[0] += 1
            if args.barrier == "flags":
                fu = frame_uses[k]                              # how often THIS path (whole-frame `done` counter) has used buffer k
                frame_uses[k] += 1
                done_ctr, free_ctr = shared_ptrs[k] + W * H * 4, shared_ptrs[k] + W * H * 4 + 4
                ctx.flag_wait_ge(free_ctr, use)                 # rank 0 has handed buffer k back `use` times: safe to overwrite it
                ctx.trace_rows(tlas, cam, W, H, bounces, BLOCK_ROWS, rank, world, shared_ptrs[k], device=True, async_=True, full_frame=True)
                ctx.flag_add(done_ctr)                          # this rank's pixels of frame k have landed in rank 0's memory
                if rank == 0:
                    ctx.flag_wait_ge(done_ctr, world * (fu + 1))    # ... and so have everybody else's: the frame is complete
                    if consume:
                        consume(shared[k])
                    ctx.flag_add(free_ctr)
                    return shared[k]
                return None
            ctx.trace_rows(tlas, cam, W, H, bounces, BLOCK_ROWS, rank, world, shared_ptrs[k], device=True, async_=True, full_frame=True)
            dist.all_reduce(token)          # frame-complete barrier on the stream: every rank's stores have landed
            if rank == 0 and consume:
                consume(shared[k])
            return shared[k] if rank == 0 else None
        ctx.trace_rows(tlas, cam, W, H, bounces, BLOCK_ROWS, rank, world, packed, device=True, async_=True)
        if mode == "gather":
            dist.gather(packed, gather_list, dst=0)
        else:
            dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1))
        if rank == 0:
            ctx.unpack_rows(gathered, W, H, BLOCK_ROWS, world, frame)
            if consume:
                consume(frame)
            return frame
        return None

    # ---- e2e at N > 1 (p2p + counters): row-chunk pipeline. Packed rows [a, b) of every rank together are image rows [a*N, b*N).
    chunk_bounds, chunk_uses, ctx2, copy_stream = None, [0, 0], None, None
    if mode == "p2p" and args.barrier == "flags" and args.e2e_chunks > 1:
        local_rows = px_packed // W
        nc, wsum, acc, chunk_bounds = args.e2e_chunks, args.e2e_chunks * (args.e2e_chunks + 1) // 2, 0, [0]
        for c in range(nc):
            acc += nc - c
            end = local_rows if c == nc - 1 else min(local_rows, (local_rows * acc // wsum + 7) // 8 * 8)
            if end > chunk_bounds[-1]:
                chunk_bounds.append(end)
        if rank == 0:
            copy_stream = torch.cuda.Stream(device=dev)
            ctx2 = rtcore.Context(local_rank)           # only enqueues flag waits / adds on the copy stream
            ctx2.set_stream(copy_stream.cuda_stream)

    def step_e2e_chunked():
        k, use = step_no[0] & 1, step_no[0] >> 1
        step_no[0] += 1
        cu = chunk_uses[k]
        chunk_uses[k] += 1
        tail = shared_ptrs[k] + W * H * 4
        ctx.flag_wait_ge(tail + 4, use)
        for c in range(len(chunk_bounds) - 1):
            a, b = chunk_bounds[c], chunk_bounds[c + 1]
            ctx.trace_rows_range(tlas, cam, W, H, bounces, BLOCK_ROWS, rank, world, a, b - a, shared_ptrs[k], full_frame=True)
            ctx.flag_add(tail + 8 + 4 * c)
            if rank == 0:
                ctx2.flag_wait_ge(tail + 8 + 4 * c, world * (cu + 1))      # on the copy stream: every rank's rows of this chunk have landed
                g0, g1 = a * world, min(b * world, H)
                with torch.cuda.stream(copy_stream):
                    host_frame[g0:g1].copy_(shared[k][g0:g1], non_blocking=True)
        if rank == 0:
            ctx2.flag_add(tail + 4)                     # buffer handed back once its last rows are on the host
        torch.cuda.synchronize()

    def step_device():
        if world == 1:
            ctx.trace_device(tlas, cam, W, H, bounces, frame, async_=True)
        else:
            step_multi()

    def step_e2e():
        # reference-facing call with HOST buffers: camera struct in (16 B), RGBA8 framebuffer out (pinned host memory)
        if world == 1:
            ctx.trace(tlas, cam, W, H, bounces, rgba_out=host_np)
        elif chunk_bounds is not None:
            step_e2e_chunked()
        else:
            step_multi(consume=lambda f: host_frame.copy_(f, non_blocking=True))
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- one stats pass: ray counts + traversal counters for the byte model (untimed, slower kernel) ----
    if world == 1:
        ctx.trace_device(tlas, cam, W, H, bounces, frame, stats=True)
    else:
        ctx.trace_rows(tlas, cam, W, H, bounces, BLOCK_ROWS, rank, world, packed, device=True, stats=True)
    st = ctx.trace_stats()
    cnt = torch.tensor([st[k] for k in ("rays_primary", "rays_secondary", "nodes_visited", "triangles_tested", "instances_entered",
                                        "primary_hits", "secondary_hits", "near_edge_hits")], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    tot = dict(zip(("rays_primary", "rays_secondary", "nodes_visited", "triangles_tested", "instances_entered",
                    "primary_hits", "secondary_hits", "near_edge_hits"), [int(x) for x in cnt.tolist()]))
    total_rays = tot["rays_primary"] + tot["rays_secondary"]

    # ---- timed region: device-resident ----
    for _ in range(warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    barrier()
    l1 = ctx.launch_count()
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / steps

    # trace-kernel-only time on this rank (CUDA events on the launching stream, inside the library)
    k_ms = []
    for _ in range(min(steps, 10)):
        if world == 1:
            ctx.trace_device(tlas, cam, W, H, bounces, frame)
        else:
            ctx.trace_rows(tlas, cam, W, H, bounces, BLOCK_ROWS, rank, world, packed, device=True)
        k_ms.append(ctx.trace_ms())
    kern = torch.tensor([float(np.mean(k_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kern, op=dist.ReduceOp.MAX)
    kernel_ms = float(kern.item())

    # ---- timed region: end to end through the host-buffer call ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, steps // 2)
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1000.0 / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None

    # checksums of the full frame and of the hit records (primary + secondary): variants of the kernels must reproduce them exactly
    crc = None
    if world == 1:
        import zlib
        prim_c = torch.zeros((H, W, 7), dtype=torch.int32, device=dev)
        sec_c = torch.zeros((H, W, 7), dtype=torch.int32, device=dev)
        ctx.trace_device(tlas, cam, W, H, bounces, frame, prim_c, sec_c)
        crc = {"rgba": zlib.crc32(frame.cpu().numpy().tobytes()), "primary_hits": zlib.crc32(prim_c.cpu().numpy().tobytes()),
               "secondary_hits": zlib.crc32(sec_c.cpu().numpy().tobytes())}
        del prim_c, sec_c
    else:
        import zlib
        # the assembled frame of the multi-GPU path must be the single-GPU frame, bit for bit
        step_multi(consume=lambda f: host_frame.copy_(f, non_blocking=True))
        torch.cuda.synchronize()
        ctx.sync()                          # also surfaces a device watchdog (flag wait timed out) as an error
        if rank == 0:
            crc = {"rgba": zlib.crc32(host_frame.numpy().tobytes()), "gather": mode, "barrier": args.barrier if mode == "p2p" else "nccl"}

    if rank == 0:
        hbm, peak_src = peaks()
        algo_bytes = b_ray_bytes(tot, W * H)
        # per launch on one GPU: this rank's share of the frame (1/world of the bytes) over its kernel time
        achieved = algo_bytes / world / (kernel_ms * 1e-3) / 1e9
        n_tris = n_tris_total
        build_total_ms = bt["total_ms"] + bcast_ms
        build_gbs = n_tris * B_TRI_BUILD / (build_total_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic("k_trace", args.workload) if world == 1 else (None, None)
        btraffic, btraffic_src = ncu_traffic("build", args.workload)
        line = {
            "metric": "Mrays/s (primary+secondary rays per second)", "value": total_rays / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(scene, args, world),
            "rays_per_step": total_rays, "rays_primary": tot["rays_primary"], "rays_secondary": tot["rays_secondary"],
            "trace_kernel_ms": kernel_ms,
            "e2e": {"value": total_rays / (float(e2e_ms.item()) * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": float(e2e_ms.item()),
                    "h2d_bytes_per_step": 16, "d2h_bytes_per_step": W * H * 4,
                    "note": "rt_trace with a pinned host framebuffer: camera struct in, RGBA8 frame out" if world == 1 else
                            ("every rank traces its share in %d shrinking row chunks straight into rank 0's frame; rank 0 copies the finished image rows "
                             "to pinned host memory while the next chunk is traced" % (len(chunk_bounds) - 1) if chunk_bounds is not None else
                             "frame assembled on rank 0, then copied to pinned host memory")},
            "gpu_launches": int(l1 - l0),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                         "traffic_source": traffic_src, "kernel": "k_trace (stage 0 + stage 1 of one frame)", "peak_source": peak_src,
                         "note": "achieved = ALGORITHMIC bytes (served mostly by L1/L2) over HBM peak, as SURVEY 8(d) defines it; the kernels are bound by "
                                 "issue slots x SIMD divergence and L1/L2 latency: see `sm_issue`",
                         "sm_issue": ncu_issue(args.workload),
                         "bytes_model": "64*nodes + 56*triangles + 64*instances + 4*pixels (SURVEY 8d), counters from the RT_TRACE_STATS pass",
                         "algorithmic_bytes_per_launch": algo_bytes / world,
                         "per_ray": {"nodes": tot["nodes_visited"] / total_rays, "triangles": tot["triangles_tested"] / total_rays,
                                     "instances": tot["instances_entered"] / total_rays, "bytes": algo_bytes / total_rays}},
            "build": {"metric": "LBVH build Mtri/s (first setup kernel .. last refit kernel, CUDA events)",
                      "value": n_tris / (build_total_ms * 1e-3) / 1e6, "unit": "Mtri/s", "ms": build_total_ms, "phases_ms": bt,
                      "broadcast_ms": bcast_ms, "note": build_note, "tlas_ms": tlas_t["total_ms"],
                      "roofline": {"bound": "hbm", "achieved": build_gbs, "peak": hbm, "unit": "GB/s", "frac": build_gbs / hbm,
                                   "bytes_per_triangle": B_TRI_BUILD, "traffic": btraffic, "traffic_source": btraffic_src}},
            "traversal": tot,
            "crc32": crc,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            res = cpu_arm(scene, args.cpu_seconds, 1, 0)
            cb = cpu_build_baseline(scene)
            line["cpu_baseline"] = {"value": res["mrays_per_s"], "unit": "Mrays/s", "cores": res["threads"], "kind": "port",
                                    "sample": f"rows 0..{H} step {res['rows'][2]} of the frame ({res['rays_per_step']} rays), oracle LBVH traversal; "
                                              f"CPU LBVH build of {cb['triangles']} triangles: {cb['mtris_per_s']:.2f} Mtri/s",
                                    "build_mtris_per_s": cb["mtris_per_s"]}
            # parity of the rows the CPU traced, same run (hit ids bit-exact, RGBA within 1 LSB)
            from parity import compare_hits, compare_rgba
            prim = torch.zeros((H, W, 7), dtype=torch.int32, device=dev)
            sec = torch.zeros((H, W, 7), dtype=torch.int32, device=dev)
            ctx.trace_device(tlas, cam, W, H, bounces, frame, prim, sec)
            rows = slice(*res["rows"])
            rgba_o, prim_o, sec_o, _ = res["oracle"].trace(rows=res["rows"], want_hits=True)
            gp = prim.cpu().numpy().view(rtcore.HIT_DTYPE).reshape(H, W)
            gs = sec.cpu().numpy().view(rtcore.HIT_DTYPE).reshape(H, W)
            rp, rs = compare_hits(gp, prim_o, rows), compare_hits(gs, sec_o, rows)
            rc = compare_rgba(frame.cpu().numpy(), rgba_o, rows)
            line["parity"] = {"rows_checked": len(range(*res["rows"])), "primary": {k: rp[k] for k in ("rays", "hits", "id_mismatches", "near_edge_rays", "t_bit_mismatches")},
                              "secondary": {k: rs[k] for k in ("rays", "hits", "id_mismatches", "near_edge_rays", "t_bit_mismatches")}, "rgba": rc}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        if ctx2 is not None:
            ctx2.close()
        if mode == "p2p":               # peers unmap first, then the owner frees
            shared = []
            if rank != 0:
                for p_ in shared_ptrs:
                    ctx.frame_share_close(p_)
            dist.barrier()
            if rank == 0:
                for p_ in shared_ptrs:
                    ctx.frame_share_free(p_)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="inst10m")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--build-reps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per oracle step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "allgather", "gather"],
                    help="N > 1: how the bands reach rank 0. p2p = every rank's trace kernel stores its pixels straight into rank 0's "
                         "framebuffer over NVLink (CUDA IPC mapping) + one tiny NCCL all-reduce as the frame-complete barrier; "
                         "allgather / gather = packed bands + NCCL collective + unpack kernel on rank 0")
    ap.add_argument("--barrier", default="flags", choices=["flags", "nccl"],
                    help="--gather p2p: how a frame is declared complete. flags = stream-ordered counters in rank 0's shared frame "
                         "(every rank adds 1 after its trace, rank 0 waits for N; a second counter hands the buffer back): no collective "
                         "on the step; nccl = a 4-byte all-reduce per frame")
    ap.add_argument("--e2e-chunks", type=int, default=1,
                    help="N > 1, p2p + flags: the e2e step traces every rank's share in this many shrinking row chunks with one 'done' counter "
                         "each, and rank 0 copies the finished image rows of ALL ranks to the host while the next chunk is traced. Measured "
                         "with 3 chunks (profiles/README.md r02k): N = 2 e2e 4475 vs 4398 Mrays/s, N = 8 8293 vs 8979 (the per-rank chunks get "
                         "too small) -> default 1 = one chunk, frame copied after the frame barrier")
    ap.add_argument("--soup-split", default="slab", choices=["slab", "index"],
                    help="cfg5 soup: 8 BLASes as x-slabs of the volume (default) or as index ranges of a fully mixed soup")
    args = ap.parse_args()
    # torchrun exports OMP_NUM_THREADS=1 to its children; the CPU legs (reference arm, cpu_baseline) are specified to use
    # all host threads, so undo that before the OpenMP runtime of the oracle library is loaded.
    if os.environ.get("OMP_NUM_THREADS") == "1" and "RT_KEEP_OMP" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
