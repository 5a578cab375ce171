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
        s = scenes.soup_scene(n, w, h, 1, only_parts=only_parts, split=soup_split)
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
    """Fallback for the issue-slot roofline when the live counter pass (count_instructions) is unavailable: the committed capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k_trace"]
        if t["workload"] != workload:
            return None
        return {k: t[k] for k in ("issue_slot_utilisation_pct", "lanes_per_instruction", "l1_hit_pct", "l2_hit_pct", "warp_instructions_per_frame", "source") if k in t}
    except Exception:
        return None


def count_instructions(args):
    """Live counters for the bound that really limits the trace kernels (issue slots): one frame of THIS workload with THIS library under
    `ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum` (counts only — no time measured under the profiler is
    used anywhere). Returns {"warp_inst": .., "thread_inst": .., "kernels": ..} per frame, or None when ncu cannot run here."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu) or os.environ.get("RTCORE_BENCH_NO_NCU"):
        return None
    cmd = [ncu, "--metrics", "smsp__inst_executed.sum,smsp__thread_inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum", "--clock-control", "none", "-k", "regex:k_trace",
           "--csv", sys.executable, os.path.join(ROOT, "tools", "frame_once.py"), "--workload", args.workload]
    if args.width:
        cmd += ["--width", str(args.width)]
    if args.height:
        cmd += ["--height", str(args.height)]
    try:
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        if p.returncode != 0:
            return None
        import csv
        import io
        rows = [r for r in csv.reader(io.StringIO(p.stdout)) if len(r) > 6]
        hdr = next(r for r in rows if "Metric Name" in r)
        ni, vi, ki = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
        out = {"warp_inst": 0.0, "thread_inst": 0.0, "lsu_wavefronts": 0.0, "kernels": set()}
        for r in rows:
            if r is hdr or len(r) <= max(ni, vi, ki):
                continue
            v = float(r[vi].replace(",", ""))
            if r[ni] == "smsp__inst_executed.sum":
                out["warp_inst"] += v
                out["kernels"].add(r[ki][:40])
            elif r[ni] == "smsp__thread_inst_executed.sum":
                out["thread_inst"] += v
            elif r[ni] == "l1tex__data_pipe_lsu_wavefronts.sum":
                out["lsu_wavefronts"] += v
        out["kernels"] = len(out["kernels"])
        return out if out["warp_inst"] > 0 else None
    except Exception:       # noqa: BLE001
        return None


def workload_config(scene, args, n_gpus):
    n_tris = SOUP_SIZES[args.workload][0] if args.workload in SOUP_SIZES else scene.triangle_count
    how = {"p2p": "; every rank's trace kernel stores its pixels straight into rank 0's frame over NVLink (rt_group_*: CUDA IPC + stream-ordered "
                  "counters, no collective)" + (", consecutive frames alternate over two streams" if getattr(args, "pipeline", True) else ""),
           "allgather": "; packed bands, NCCL all-gather, unpack on rank 0",
           "gather": "; packed bands, NCCL gather to rank 0, unpack"}
    return {"workload": f"{args.workload}: {scene.name}, {n_tris} triangles in {len(scene.blases)} BLAS, "
                        f"{len(scene.instances)} instances, {scene.width}x{scene.height} primary + {scene.bounces} diffuse bounce",
            "triangles": n_tris, "instances": len(scene.instances), "width": scene.width, "height": scene.height,
            "bounces": scene.bounces, "partition": f"{BLOCK_ROWS}-scanline bands interleaved over {n_gpus} GPU(s), scene replicated"
            + ("" if n_gpus == 1 else how[args.gather]),
            "l2_policy": "inputs larger than L2 (BVH nodes + triangles >> 126 MB); no flush between iterations"
                         if n_tris * 112 > 2 * 126e6 else "scene fits in L2: numbers are L2-resident (parity config)"}


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import zlib

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
        dist.init_process_group("nccl", device_id=dev)       # plumbing only: timing reductions, scene-wide counters, the NCCL comparison modes
    steps = args.steps if args.steps else 20
    warmup = args.warmup if args.warmup is not None else 3
    if warmup < 3:
        warmup = 3

    def rmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def rall(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    is_soup = args.workload in SOUP_SIZES
    want_cpu = not args.no_cpu_baseline
    # cfg5: the soup is 8 BLASes. Which parts a rank needs on the HOST: its own for the split build; all of them for the replicated
    # build (measured beside it) and, on rank 0, for the CPU arm / parity.
    soup_modes = ("split", "replicated") if (is_soup and args.soup_build == "both") else ((args.soup_build,) if is_soup else ())
    if world == 1 and is_soup:
        soup_modes = ("replicated",)
    my_parts = [p for p in range(scenes.SOUP_PARTS) if p % world == rank] if is_soup else None
    need_all = is_soup and ("replicated" in soup_modes or (rank == 0 and want_cpu))
    scene = make_workload(args.workload, args.width, args.height, only_parts=None if need_all else my_parts, soup_split=args.soup_split)
    if args.bounces is not None:
        scene.bounces = args.bounces
    W, H, bounces = scene.width, scene.height, scene.bounces
    ctx = rtcore.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    group = None
    if world > 1 or args.pipeline:
        gname = f"bench-{os.environ.get('MASTER_PORT', '0')}-{os.environ.get('TORCHELASTIC_RUN_ID', 'solo')}-{os.getppid() if world > 1 else os.getpid()}"
        group = rtcore.Group(ctx, gname, rank, world, W, H)

    # ---- scene upload (untimed) and acceleration-structure build (timed separately: Mtri/s) ----
    def upload(geoms):
        dg = []
        for g in geoms:
            v = torch.from_numpy(np.ascontiguousarray(g.vertices)).to(dev)
            i = torch.from_numpy(np.ascontiguousarray(g.indices).view(np.int32)).to(dev) if g.indices is not None else None
            t = torch.from_numpy(np.ascontiguousarray(g.transform)).to(dev) if g.transform is not None else None
            dg.append(scenes.Geometry(v, i, t))
        return dg

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    phase_keys = ("total_ms", "setup_ms", "morton_ms", "sort_ms", "hierarchy_ms", "refit_ms")
    n_tris_total = SOUP_SIZES[args.workload][0] if is_soup else scene.triangle_count
    build_variants = {}
    if not is_soup:
        dev_blases = [upload(geoms) for geoms in scene.blases]
        torch.cuda.synchronize()
        build_ms = []
        blases = None
        for rep in range(args.build_reps + 1):
            if blases is not None:
                for b in blases:
                    b.free()
            bf = rtcore.RT_BUILD_ALLOW_COMPACTION if args.compact else 0
            blases = ctx.build_blas_batch(dev_blases, device=True, flags=bf) if len(dev_blases) > 1 else [ctx.build_blas(dev_blases[0], device=True, flags=bf)]
            if rep > 0 or args.build_reps == 0:
                build_ms.append(ctx.build_timing())
        bt = min(build_ms, key=lambda t: t["total_ms"])
        build_total_ms, share_ms = bt["total_ms"], 0.0
        build_note = "one batched build of all BLASes on every GPU (scene replicated)"
    else:
        # cfg5 ("BLASes are built per GPU and broadcast"). Two ways to get the 8 BLASes onto every GPU, measured side by side:
        #   split:      rank r builds parts p % N == r; every part is PULLED by the other ranks from its builder's memory over NVLink
        #               (rt_group_share_blas: CUDA-IPC mapping + copy engine on a separate stream, overlapping the puller's next builds).
        #               Wall time of the whole phase, slowest rank (host clock: the pulls run beside the builds).
        #   replicated: every rank builds all 8 parts itself; sum of the builds' CUDA-event times, slowest rank.
        dev_parts = {p: upload(scene.blases[p]) for p in (range(scenes.SOUP_PARTS) if "replicated" in soup_modes else my_parts)}
        torch.cuda.synchronize()
        blases = [None] * scenes.SOUP_PARTS
        bt = {k: 0.0 for k in phase_keys}
        best_own = {}
        for p in sorted(dev_parts):                      # warm the allocator / scratch, and the per-part kernel times
            b = None
            for rep in range(max(1, args.build_reps)):
                if b is not None:
                    b.free()
                b = ctx.build_blas(dev_parts[p], device=True)
                t = ctx.build_timing()
                best_own[p] = t if p not in best_own or t["total_ms"] < best_own[p]["total_ms"] else best_own[p]
            b.free()
        batched = None
        if "replicated" in soup_modes:
            rep_ms = sum(best_own[p]["total_ms"] for p in range(scenes.SOUP_PARTS))
            build_variants["replicated"] = {"ms": rmax(rep_ms), "note": "every GPU builds all 8 parts itself (sum of the 8 builds' CUDA-event times, slowest rank)"}
            # the same 8 BLASes as ONE rt_build_blas_batch (one set of launches over all triangles; the BLAS id rides in the sort key)
            best_batch = None
            for rep in range(max(1, args.build_reps)):
                if batched is not None:
                    for b in batched:
                        b.free()
                batched = ctx.build_blas_batch([dev_parts[p] for p in range(scenes.SOUP_PARTS)], device=True)
                t = ctx.build_timing()
                best_batch = t if best_batch is None or t["total_ms"] < best_batch["total_ms"] else best_batch
            build_variants["replicated_batched"] = {"ms": rmax(best_batch["total_ms"]),
                                                    "note": "every GPU builds all 8 parts itself in one rt_build_blas_batch (CUDA-event time, slowest rank)"}
        if "split" in soup_modes and world > 1:
            best = None
            for rep in range(2):
                for b in blases:
                    if b is not None:
                        b.free()
                blases = [None] * scenes.SOUP_PARTS
                barrier()
                t0 = time.perf_counter()
                own_ms = 0.0
                for p in range(scenes.SOUP_PARTS):
                    owner = p % world
                    mine = None
                    if owner == rank:
                        mine = ctx.build_blas(dev_parts[p], device=True)
                        own_ms += ctx.build_timing()["total_ms"]
                    blases[p] = group.share_blas(p, owner, mine)
                pull_ms = group.share_finish()           # waits for this rank's pulls, then a barrier of the group
                wall = (time.perf_counter() - t0) * 1000.0
                cur = {"ms": rmax(wall), "own_builds_ms": rmax(own_ms), "pull_ms": rmax(pull_ms), "rank0_host_ms": group.share_host_ms()}
                best = cur if best is None or cur["ms"] < best["ms"] else best
            gb = sum(int(blases[p].info().storage_bytes) for p in range(scenes.SOUP_PARTS) if p % world != rank) / 1e9
            best["pulled_gb_per_rank"] = gb
            best["pull_gbs"] = gb / max(best["pull_ms"], 1e-9) * 1e3
            best["note"] = (f"{len(my_parts)} of 8 parts built per GPU, the others pulled from their builders over NVLink while the next part is built "
                            f"(host wall clock of the phase, slowest rank)")
            build_variants["split"] = best
        for k in phase_keys:
            bt[k] = sum(best_own[p][k] for p in best_own)
        bt["primitives"] = n_tris_total
        fastest = min(build_variants, key=lambda m: build_variants[m]["ms"])
        if fastest == "replicated_batched":
            for k in phase_keys:
                bt[k] = best_batch[k]
        if batched is not None and (fastest == "replicated_batched" or blases[0] is None):
            for b in blases:                             # the frame is traced through the batch's handles
                if b is not None:
                    b.free()
            blases = list(batched)
        elif batched is not None:
            for b in batched:
                b.free()
        if blases[0] is None:
            for p in range(scenes.SOUP_PARTS):
                blases[p] = ctx.build_blas(dev_parts[p], device=True)
        build_total_ms = build_variants[fastest]["ms"]
        share_ms = build_variants.get("split", {}).get("pull_ms", 0.0)
        build_note = f"8 BLASes of {n_tris_total // scenes.SOUP_PARTS} triangles; reported = the faster way ({fastest}); see build.variants"
        del dev_parts
    compaction = None
    if args.compact and not is_soup:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b0, b1 = ctx.compact_blas(blases[0])            # the whole batch
        torch.cuda.synchronize()
        compaction = {"bytes_before": b0, "bytes_after": b1, "host_ms": (time.perf_counter() - t0) * 1e3,
                      "note": "rt_compact_blas after the timed build (VK_COPY_ACCELERATION_STRUCTURE_MODE_COMPACT_KHR); not part of build.value"}
    tlas = ctx.build_tlas(scene.instances, blases)
    tlas_t = ctx.build_timing()
    tlas_cold_ms = tlas_t["total_ms"]                  # the first TLAS build also pays the lazy loading of its kernels
    for _ in range(max(1, args.build_reps)):             # the per-frame path (rt_update_tlas): same kernels, warm
        ctx.update_tlas(tlas, scene.instances, blases)
        t = ctx.build_timing()
        if t["total_ms"] < tlas_t["total_ms"]:
            tlas_t = t
    ctx.set_hit_records(scene.hit_records)
    ctx.set_miss_color(scene.miss_color)
    cam = ctx.camera(scene.camera_pos, scene.yfov_deg)

    # ---- buffers ----
    px_packed = ctx.rows_packed_pixels(W, H, BLOCK_ROWS, world)
    mode = args.gather if world > 1 else "single"
    frame = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
    host_frame = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
    host_np = host_frame.numpy()
    if world > 1:
        packed = torch.zeros((px_packed, 4), dtype=torch.uint8, device=dev)
        if mode == "gather":
            gathered = torch.zeros((world, px_packed, 4), dtype=torch.uint8, device=dev) if rank == 0 else None
            gather_list = [gathered[r] for r in range(world)] if rank == 0 else None
        elif mode == "allgather":
            gathered = torch.zeros((world, px_packed, 4), dtype=torch.uint8, device=dev)
    gflags = rtcore.GROUP_OUT_DEVICE | (rtcore.GROUP_PIPELINE if args.pipeline else rtcore.GROUP_ASYNC)

    def step_nccl(consume=None):
        ctx.trace_rows(tlas, cam, W, H, bounces, BLOCK_ROWS, rank, world, packed, device=True, async_=True)
        if mode == "gather":
            dist.gather(packed, gather_list, dst=0)
        else:
            dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1))
        if rank == 0:
            ctx.unpack_rows(gathered, W, H, BLOCK_ROWS, world, frame)
            if consume:
                consume(frame)

    def step_device():
        if group is not None and mode in ("single", "p2p"):
            group.trace(tlas, cam, W, H, bounces, gflags)
        elif world == 1:
            ctx.trace_device(tlas, cam, W, H, bounces, frame, async_=True)
        else:
            step_nccl()

    def finish_device():
        if group is not None:
            group.join()                 # stream-ordered: the context stream (torch's) waits for the group's second stream

    def step_e2e():
        # reference-facing call with HOST buffers: camera struct in (16 B), RGBA8 framebuffer out (pinned host memory)
        if world == 1:
            ctx.trace(tlas, cam, W, H, bounces, rgba_out=host_np)
            return host_np
        if mode == "p2p":
            return group.trace_host(tlas, cam, W, H, bounces)        # every rank copies its bands over its own PCIe link; rank 0 gets the frame
        step_nccl(consume=lambda f: host_frame.copy_(f, non_blocking=True))
        torch.cuda.synchronize()
        return host_np

    # ---- one stats pass: ray counts + traversal counters for the byte model (untimed, slower kernel) ----
    if world == 1:
        ctx.trace_device(tlas, cam, W, H, bounces, frame, stats=True)
    else:
        ctx.trace_rows(tlas, cam, W, H, bounces, BLOCK_ROWS, rank, world, packed, device=True, stats=True)
    st = ctx.trace_stats()
    stat_keys = ("rays_primary", "rays_secondary", "nodes_visited", "triangles_tested", "instances_entered", "primary_hits", "secondary_hits", "near_edge_hits")
    cnt = torch.tensor([st[k] for k in stat_keys], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    tot = dict(zip(stat_keys, [int(x) for x in cnt.tolist()]))
    total_rays = tot["rays_primary"] + tot["rays_secondary"]

    # ---- timed region: device-resident ----
    for _ in range(warmup):
        step_device()
    finish_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_device()
    finish_device()
    e1.record()
    barrier()
    l1 = ctx.launch_count()
    ms_step = rmax(e0.elapsed_time(e1)) / steps

    # trace-kernel-only time of ONE frame on every rank (CUDA events on the launching stream, inside the library)
    k_ms = []
    for _ in range(min(steps, 10)):
        if world == 1:
            ctx.trace_device(tlas, cam, W, H, bounces, frame)
        else:
            ctx.trace_rows(tlas, cam, W, H, bounces, BLOCK_ROWS, rank, world, packed, device=True)
        k_ms.append(ctx.trace_ms())
    per_rank_kernel_ms = rall(float(np.mean(k_ms)))
    kernel_ms = max(per_rank_kernel_ms)

    # ---- timed region: end to end through the host-buffer call ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, steps // 2)
    for _ in range(e2e_steps):
        e2e_frame = step_e2e()
    barrier()
    e2e_ms = rmax((time.perf_counter() - t0) * 1000.0 / e2e_steps)
    e2e_crc = zlib.crc32(np.ascontiguousarray(e2e_frame).tobytes()) if rank == 0 else None
    # ---- the same with TWO frames in flight (rt_group_trace, RT_GROUP_OUT_HOST | RT_GROUP_PIPELINE; a one-rank group at N = 1): call k
    # enqueues frame k and returns frame k - 1, so a frame's device->host copies and the ranks' handshake overlap the next frame's trace.
    # Every frame's camera upload and 4 B/pixel download are inside the timed region; this is the reported e2e, the synchronous
    # call-per-frame figure stays beside it as e2e.sync_value ----
    e2e_sync_ms, e2e_pipe_crc = e2e_ms, None
    e2e_piped = group is not None and mode in ("p2p", "single") and args.pipeline
    if e2e_piped:
        for _ in range(2):
            group.trace_host(tlas, cam, W, H, bounces, pipeline=True)
        group.flush_host(W, H)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            group.trace_host(tlas, cam, W, H, bounces, pipeline=True)
        last = group.flush_host(W, H)
        barrier()
        e2e_ms = rmax((time.perf_counter() - t0) * 1000.0 / e2e_steps)
        if rank == 0:
            e2e_pipe_crc = zlib.crc32(np.ascontiguousarray(last).tobytes())
    clocks = sampler.stop() if sampler else None

    # checksums of the full frame and of the hit records (primary + secondary): variants of the kernels must reproduce them exactly.
    # Rank 0 holds a replica of the scene, so it also renders the whole frame ALONE: the frame the N GPUs assembled must equal it.
    crc = None
    if rank == 0:
        prim_c = torch.zeros((H, W, 7), dtype=torch.int32, device=dev)
        sec_c = torch.zeros((H, W, 7), dtype=torch.int32, device=dev)
        ctx.trace_device(tlas, cam, W, H, bounces, frame, prim_c, sec_c)
        solo = frame.cpu().numpy()
        gp = prim_c.cpu().numpy().view(rtcore.HIT_DTYPE).reshape(H, W)
        gs = sec_c.cpu().numpy().view(rtcore.HIT_DTYPE).reshape(H, W)
        crc = {"rgba": zlib.crc32(solo.tobytes()), "primary_hits": zlib.crc32(gp.tobytes()), "secondary_hits": zlib.crc32(gs.tobytes()),
               "e2e_frame_rgba": e2e_crc, "e2e_pipelined_frame_rgba": e2e_pipe_crc}
        del prim_c, sec_c
    if world > 1:
        if mode == "p2p":
            ptr = group.trace(tlas, cam, W, H, bounces, rtcore.GROUP_OUT_DEVICE)
            if rank == 0:
                crc["assembled_rgba"] = zlib.crc32(rtcore.device_view(ptr, W * H * 4, dev).cpu().numpy().tobytes())
        else:
            step_nccl()
            torch.cuda.synchronize()
            if rank == 0:
                crc["assembled_rgba"] = zlib.crc32(frame.cpu().numpy().tobytes())
        if rank == 0:
            crc["gather"] = mode
            crc["assembled_equals_single_gpu_frame"] = bool(crc["assembled_rgba"] == crc["rgba"] == crc["e2e_frame_rgba"])
    elif rank == 0:
        crc["assembled_equals_single_gpu_frame"] = bool(crc["e2e_frame_rgba"] == crc["rgba"])
    ctx.sync()                          # also surfaces a device watchdog (flag wait timed out) as an error

    if rank == 0:
        hbm, peak_src = peaks()
        algo_bytes = b_ray_bytes(tot, W * H)
        # per launch on one GPU: this rank's share of the frame (1/world of the bytes) over its kernel time
        achieved = algo_bytes / world / (kernel_ms * 1e-3) / 1e9
        n_tris = n_tris_total
        build_gbs = n_tris * B_TRI_BUILD / (build_total_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic("k_trace", args.workload) if world == 1 else (None, None)
        btraffic, btraffic_src = ncu_traffic("build", args.workload)
        # ---- the bound that binds: SM issue slots. Warp-instructions of one frame (live ncu counter pass of the same library and
        # workload; the committed capture only if ncu cannot run here) over kernel time x 4 schedulers x SMs x measured SM clock ----
        sm_count = ctx.device_info()["sm_count"]
        clock_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        live = count_instructions(args) if (world == 1 and not args.no_issue_counters) else None
        committed = ncu_issue(args.workload)
        warp_inst = live["warp_inst"] if live else (committed or {}).get("warp_instructions_per_frame")
        issue_peak = sm_count * 4 * clock_mhz * 1e6 / 1e9                     # G warp-instructions / s
        issue = None
        if warp_inst:
            issue_ach = warp_inst / world / (kernel_ms * 1e-3) / 1e9
            issue = {"bound": "issue", "achieved": issue_ach, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": issue_ach / issue_peak,
                     "warp_instructions_per_frame": warp_inst,
                     "lanes_per_instruction": (live["thread_inst"] / live["warp_inst"]) if live else None,
                     "simd_ideal_frac": (live["thread_inst"] / 32.0 / world / (kernel_ms * 1e-3) / 1e9 / issue_peak) if live else None,
                     "counters": "live: ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum on one frame of this run's library and workload"
                                 if live else "committed capture (profiles/ncu_traffic.json): ncu could not run in this process environment",
                     "sm_clock_mhz": clock_mhz, "sm_count": sm_count}
        hbm_roof = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                    "traffic_source": traffic_src, "peak_source": peak_src,
                    "note": "SURVEY 8(d) figure: ALGORITHMIC bytes (served mostly by L1/L2) over the HBM peak; measured DRAM traffic is `traffic`",
                    "bytes_model": "64*nodes + 56*triangles + 64*instances + 4*pixels (SURVEY 8d), counters from the RT_TRACE_STATS pass of this run",
                    "algorithmic_bytes_per_launch": algo_bytes / world}
        # ---- the other unit the node fetches load: the LSU data pipe of L1TEX, one wavefront per cycle and SM (ncu: 75-82 % of peak on
        # inst10m, above the issue-slot utilisation). Same live counter pass; reported as the primary roofline when its fraction is the larger
        l1tex = None
        if live and live.get("lsu_wavefronts"):
            wf_ach = live["lsu_wavefronts"] / world / (kernel_ms * 1e-3) / 1e9
            wf_peak = sm_count * clock_mhz * 1e6 / 1e9                           # G wavefronts / s: one per cycle and SM
            l1tex = {"bound": "l1tex-lsu", "achieved": wf_ach, "peak": wf_peak, "unit": "Gwavefronts/s", "frac": wf_ach / wf_peak,
                     "wavefronts_per_frame": live["lsu_wavefronts"],
                     "counters": "live: ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum on one frame of this run's library and workload; "
                                 "peak = 1 wavefront per cycle per SM at the SM clock sampled during the timed region",
                     "sm_clock_mhz": clock_mhz, "sm_count": sm_count}
        if issue and l1tex and l1tex["frac"] > issue["frac"]:
            roofline = dict(l1tex)
            roofline["sm_issue"] = issue
        elif issue:
            roofline = dict(issue)
            if l1tex:
                roofline["l1tex_lsu"] = l1tex
        else:
            roofline = dict(hbm_roof)
        roofline.update({"kernel": "k_trace (stage 0 + stage 1 of one frame)", "traffic": traffic, "traffic_source": traffic_src,
                         "hbm_algorithmic": hbm_roof,
                         "per_ray": {"nodes": tot["nodes_visited"] / total_rays, "triangles": tot["triangles_tested"] / total_rays,
                                     "instances": tot["instances_entered"] / total_rays, "bytes": algo_bytes / total_rays}})
        line = {
            "metric": "Mrays/s (primary+secondary rays per second)", "value": total_rays / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(scene, args, world),
            "rays_per_step": total_rays, "rays_primary": tot["rays_primary"], "rays_secondary": tot["rays_secondary"],
            "trace_kernel_ms": kernel_ms, "trace_kernel_ms_per_rank": {"min": min(per_rank_kernel_ms), "max": max(per_rank_kernel_ms), "all": per_rank_kernel_ms},
            "e2e": {"value": total_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 16, "d2h_bytes_per_step": W * H * 4,
                    "sync_value": total_rays / (e2e_sync_ms * 1e-3) / 1e6, "sync_ms_per_step": e2e_sync_ms,
                    "frames_in_flight": 2 if e2e_piped else 1,
                    "note": ("rt_group_trace(RT_GROUP_OUT_HOST | RT_GROUP_PIPELINE), two frames in flight: every rank traces its bands and copies them over its own "
                             "PCIe link into the group's shared pinned host frame while the next frame is traced; camera struct in (16 B), RGBA8 frame out, both "
                             "inside the timed region for every frame; sync_value = one blocking host-buffer call per frame "
                             + ("(rt_trace)" if world == 1 else "(rt_group_trace, RT_GROUP_OUT_HOST)")) if e2e_piped else
                            ("rt_trace with a pinned host framebuffer: camera struct in, RGBA8 frame out" if world == 1 else
                             "frame assembled on rank 0 (NCCL), then copied to pinned host memory")},
            "gpu_launches": int(l1 - l0),
            "roofline": roofline,
            "build": {"metric": "LBVH build Mtri/s (first setup kernel .. last refit kernel, CUDA events)",
                      "value": n_tris / (build_total_ms * 1e-3) / 1e6, "unit": "Mtri/s", "ms": build_total_ms, "phases_ms": bt,
                      "share_ms": share_ms, "note": build_note, "tlas_ms": tlas_t["total_ms"], "tlas_first_build_ms": tlas_cold_ms, "variants": build_variants or None,
                      "roofline": {"bound": "hbm", "achieved": build_gbs, "peak": hbm, "unit": "GB/s", "frac": build_gbs / hbm,
                                   "bytes_per_triangle": B_TRI_BUILD, "traffic": btraffic, "traffic_source": btraffic_src}},
            "traversal": tot,
            "compaction": compaction,
            "crc32": crc,
            "clocks": clocks,
        }
        if want_cpu:
            # the CPU arm and the parity of the rows it traced, same run, at EVERY N (rank 0 holds the whole scene):
            # hit ids / t / u / v bit-exact, RGBA8 within 1 LSB
            res = cpu_arm(scene, args.cpu_seconds, 1, 0)
            cb = cpu_build_baseline(scene)
            line["cpu_baseline"] = {"value": res["mrays_per_s"], "unit": "Mrays/s", "cores": res["threads"], "kind": "port",
                                    "sample": f"rows 0..{H} step {res['rows'][2]} of the frame ({res['rays_per_step']} rays), oracle LBVH traversal; "
                                              f"CPU LBVH build of {cb['triangles']} triangles: {cb['mtris_per_s']:.2f} Mtri/s",
                                    "build_mtris_per_s": cb["mtris_per_s"]}
            from parity import compare_hits, compare_rgba
            rows = slice(*res["rows"])
            rgba_o, prim_o, sec_o, _ = res["oracle"].trace(rows=res["rows"], want_hits=True)
            rp, rs = compare_hits(gp, prim_o, rows), compare_hits(gs, sec_o, rows)
            rc = compare_rgba(solo, rgba_o, rows)
            keys = ("rays", "hits", "id_mismatches", "near_edge_rays", "t_bit_mismatches", "u_bit_mismatches", "v_bit_mismatches")
            line["parity"] = {"rows_checked": len(range(*res["rows"])), "primary": {k: rp[k] for k in keys}, "secondary": {k: rs[k] for k in keys}, "rgba": rc,
                              "frame_checked": "rank 0's own full-frame render; the frame the N GPUs assembled and the e2e host frame have the same CRC-32: "
                                               + str(crc["assembled_equals_single_gpu_frame"])}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if group is not None:
        group.close()
    if world > 1:
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
                    help="N > 1: how the bands reach rank 0. p2p = the render group of the C ABI (rt_group_*): every rank's trace kernel stores "
                         "its pixels straight into rank 0's framebuffer over NVLink, stream-ordered counters as the frame handshake, no collective; "
                         "allgather / gather = packed bands + NCCL collective + unpack kernel on rank 0 (the comparison the north_star names)")
    ap.add_argument("--no-pipeline", dest="pipeline", action="store_false",
                    help="device-timed steps: do NOT alternate consecutive frames over two streams (RT_GROUP_PIPELINE). Pipelined, frame k + 1's "
                         "first rays fill the tail of frame k's persistent kernels - the fixed per-frame cost that limits scaling at 8 GPUs")
    ap.add_argument("--bounces", type=int, default=None, help="override the workload's bounce count (0 or 1)")
    ap.add_argument("--soup-build", default="both", choices=["both", "split", "replicated"],
                    help="cfg5 at N > 1: split = per-GPU builds + NVLink pulls of the other parts, replicated = every GPU builds all parts; both = "
                         "measure both, report the faster")
    ap.add_argument("--compact", action="store_true", help="build with RT_BUILD_ALLOW_COMPACTION and trace the compacted BLASes (rt_compact_blas)")
    ap.add_argument("--no-issue-counters", action="store_true", help="skip the live ncu instruction-count pass (N = 1) behind roofline.frac")
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
