#!/usr/bin/env python3
"""bench.py -- the SdfKit hot path on B200: SdfExpr.ToSdf() -> Voxels sampling -> MarchingCubes meshing.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid N] [--scene readme|csg50|sphere]

One "step" = one pass of the hot path over one grid of the README RepeatXY scene (BASELINE.json): sample every
voxel (distance + colour, clip to bounds) into HBM, then mesh it (classify -> scan -> compact -> emit).  The SDF
is analytic, so the step has no input arrays: "inputs resident in HBM" is the compiled SDF module.
  N = 1: 1024^3 (the size BASELINE.json's metric is quoted on; 17.2 GB of voxels, fits one B200).
  N > 1: weak scaling -- an n^3 grid with n^3 ~= N * 1024^3 (2048^3 at N = 8, BASELINE config 4), z-slab sharded:
         every rank samples its slices + halo, classifies, the ranks all-gather their (vertices, triangles)
         counts over NCCL, and each emits its part of the mesh at the resulting global offsets.
`value` = voxels of the whole job / step time (device events, max over ranks).  `e2e` = the same metric through
the public API Sdf.ToMesh (host result: the mesh is copied back to host memory every step).
--impl reference times the CPU restatement of the reference (the oracle) on a bounded sample of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "voxel_samples_per_s"
UNIT = "voxels/s"
WEAK_GRID = {1: 1024, 2: 1280, 4: 1624, 8: 2048}       # n^3 ~= N * 1024^3 (1280 = 10 full 128-voxel chunks per row)


def workload_name(scene, n, world):
    """The same string in our line and in the reference arm's line: both measure this workload."""
    return "README RepeatXY scene (%s): SdfExpr -> %d^3 Voxels (clip) + MarchingCubes, z-slab sharded over %d GPU(s)" % (scene, n, world)


def scene_by_name(name):
    from sdfkit_b200 import scenes
    return {"readme": scenes.readme_scene, "csg50": scenes.csg50, "sphere": scenes.sphere, "perf": scenes.perf_scene}[name]()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.5)        # (an nvidia-smi query holds a driver lock for milliseconds: keep them rare)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_baseline(scene_name, n_sample, steps, warmup):
    """The CPU restatement of the reference path (oracle) on an n_sample^3 grid of the same scene and bounds:
    sampling on all host cores in 2048-sample batches (Voxels.cs:88), clip, single-threaded marching cubes
    (MarchingCubes.cs:39-92).  Returns (voxels/s, tris/s, detail)."""
    import numpy as np
    import oracle
    expr, mn, mx = scene_by_name(scene_name)
    cores = os.cpu_count() or 1
    sdf = oracle.compile_sdf(expr.Lower())
    mn, mx = np.float32(mn), np.float32(mx)
    times = []
    ntris = 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        v, c = oracle.sample(sdf, mn, mx, n_sample, n_sample, n_sample, threads=cores)
        t1 = time.perf_counter()
        oracle.clip(v, mn, mx)
        m = oracle.marching_cubes(v, c, mn, mx)
        t2 = time.perf_counter()
        ntris = len(m.triangles)
        if it >= warmup:
            times.append((t1 - t0, t2 - t1))
    ts = statistics.median([a for a, _ in times])
    tm = statistics.median([b for _, b in times])
    nvox = n_sample ** 3
    return nvox / (ts + tm), ntris / tm, {
        "sample_voxels_per_s": nvox / ts, "mesh_tris_per_s": ntris / tm, "mesh_cells_per_s": (n_sample - 1) ** 3 / tm,
        "sample_s": ts, "mesh_s": tm, "cores": cores}


def cpu_baseline_expr(expr, mn, mx, n):
    """cpu_baseline for an arbitrary SdfExpr (one pass, no warm-up): the CPU leg of tools/run_configs.py's table."""
    import numpy as np
    import oracle
    cores = os.cpu_count() or 1
    sdf = oracle.compile_sdf(expr.Lower())
    mn, mx = np.float32(mn), np.float32(mx)
    t0 = time.perf_counter()
    v, c = oracle.sample(sdf, mn, mx, n, n, n, threads=cores)
    t1 = time.perf_counter()
    oracle.clip(v, mn, mx)
    m = oracle.marching_cubes(v, c, mn, mx)
    t2 = time.perf_counter()
    return {"cpu_grid": n, "cpu_cores": cores, "cpu_samples_per_s": n ** 3 / (t1 - t0), "cpu_tris_per_s": len(m.triangles) / (t2 - t1),
            "cpu_cells_per_s": (n - 1) ** 3 / (t2 - t1), "cpu_step_voxels_per_s": n ** 3 / (t2 - t0), "cpu_triangles": len(m.triangles)}


def cpu_baseline_render(expr, w, h):
    """The CPU restatement of RayMarcher.Render (row bands on all cores, RayMarcher.cs:50-61) on a w x h image."""
    import oracle
    from sdfkit_b200 import numerics, scenes
    cores = os.cpu_count() or 1
    sdf = oracle.compile_sdf(expr.Lower())
    view = numerics.create_look_at(*scenes.CAMERA)
    t0 = time.perf_counter()
    oracle.render(sdf, w, h, view=view, bands=cores)
    t = time.perf_counter() - t0
    return {"cpu_image": [w, h], "cpu_cores": cores, "cpu_render_ms": t * 1e3, "cpu_pixels_per_s": w * h / t}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_s = args.cpu_n
    val, tris, d = cpu_baseline(args.scene, n_s, max(1, args.steps), max(0, min(args.warmup, 1)))
    n = args.n or WEAK_GRID.get(args.gpus, 1024)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (d["sample_s"] + d["mesh_s"]), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.scene, n, args.gpus), "grid": [n, n, n]},
        "tris_per_s": tris,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": d["cores"], "kind": "port",
                         "sample": "same scene and bounds at %d^3 (1/%d of the voxels); sampling on %d threads, marching cubes "
                                   "single-threaded like the reference; C++ restatement of SdfKit's CPU path, g++ -O2 "
                                   "-ffp-contract=off (the .NET reference cannot run here)" % (n_s, max(1, round((n / n_s) ** 3)), d["cores"]),
                         "detail": d},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    # stdout carries exactly one JSON line: library chatter (e.g. NCCL's version banner) is sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SdfKit GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sdfkit_b200 as sk
    from sdfkit_b200 import dist as skd

    n = args.n or WEAK_GRID.get(world, int(round(1024 * world ** (1 / 3) / 8)) * 8)
    expr, mn, mx = scene_by_name(args.scene)
    ctx = sk.Context(local)
    t0 = time.perf_counter()
    sdf = sk.GpuSdf(expr, ctx=ctx)
    jit_s = time.perf_counter() - t0
    spr = args.slabs_per_rank or 1
    job = skd.ShardedMesher(sdf, mn, mx, n, n, n, rank, world, spr, clip=True, balanced=(world > 1 and not args.uniform_slabs))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        counts = job.sample_classify()
        allc = skd.all_gather_int64(counts, device=dev) if world > 1 else counts[None]
        offs, tot = job.offsets(allc)
        job.emit(offs)
        return tot

    for _ in range(max(args.warmup, 3)):
        tot = step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ctx.launch_count()
    sample_ms, stage = [], {"classify_ms": [], "scan_ms": [], "compact_ms": [], "emit_ms": []}
    host = {"sample_classify": 0.0, "allgather": 0.0, "emit": 0.0}
    barrier()
    ctx.mark(0)
    w0 = time.perf_counter()
    for _ in range(args.steps):
        h0 = time.perf_counter()
        counts = np.zeros((spr, 2), dtype=np.int64)
        k1 = 0.0
        for k, s_ in enumerate(job.slabs):
            if s_.ke > s_.kb:
                ctx.mark(2)
                s_.sample()
                ctx.mark(3)
                counts[k] = s_.classify()                     # synchronises: the event pair is complete
                k1 += ctx.elapsed(2, 3)
        h1 = time.perf_counter()
        allc = skd.all_gather_int64(counts, device=dev) if world > 1 else counts[None]
        offs, tot = job.offsets(allc)
        h2 = time.perf_counter()
        job.emit(offs)
        h3 = time.perf_counter()
        host["sample_classify"] += (h1 - h0) * 1e3
        host["allgather"] += (h2 - h1) * 1e3
        host["emit"] += (h3 - h2) * 1e3
        sample_ms.append(k1)
        st = job.stats()
        for k in stage:
            stage[k].append(st[k])
    ctx.mark(1)
    barrier()
    wall = time.perf_counter() - w0
    dev_ms = ctx.elapsed(0, 1)
    launches = ctx.launch_count() - l0
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = t.tolist()
    ms_per_step = dev_ms / args.steps
    nvox_total = n ** 3
    value = nvox_total / (ms_per_step * 1e-3)
    ntris_total = int(tot[1])

    # dram__bytes_read + dram__bytes_write of one K1 launch from the committed ncu capture (same workload only)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01s2_k1_traffic.json")))
        if world == 1 and n == 1024 and args.scene == "readme":
            traffic = float(tj["traffic_bytes"])
    except Exception:
        pass
    # ---- roofline of the dominant kernel (K1 sdfk_k_sample): 16 algorithmic bytes written per voxel, 0 read
    hbm, peak_src = peaks()
    slab_vox = sum(n * n * (s_.z1 - s_.z0) for s_ in job.slabs if s_.ke > s_.kb)   # voxels this rank's K1 launches write per step (incl. halo slices)
    k1_ms = statistics.mean(sample_ms)
    achieved = 16.0 * slab_vox / (k1_ms * 1e-3) / 1e9
    mesh_ms = {k: statistics.mean(v) for k, v in stage.items()}
    cells = (n - 1) * (n - 1) * (n - 1)
    # per-rank view of the step (device stage times + host-side wall per phase), gathered to rank 0
    mine = [k1_ms] + [mesh_ms[k] for k in ("classify_ms", "scan_ms", "compact_ms", "emit_ms")] + [host[k] / args.steps for k in ("sample_classify", "allgather", "emit")]
    per_rank = torch.tensor(mine, dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.empty_like(per_rank) for _ in range(world)]
        dist.all_gather(allr, per_rank)
        per_rank = torch.stack(allr)
    else:
        per_rank = per_rank[None]
    per_rank = per_rank.cpu().numpy()
    names = ["sample_ms", "classify_ms", "scan_ms", "compact_ms", "emit_ms", "host_sample_classify_ms", "host_allgather_ms", "host_emit_ms"]
    rank_stages = {nm: [round(float(x), 4) for x in per_rank[:, i]] for i, nm in enumerate(names)}
    mesh_total_ms = float(per_rank[:, 1:5].sum(axis=1).max())

    # ---- the Sdf.ToMesh variant of the step: distance-only voxels (4 B/voxel), colours evaluated at the created vertices
    fused = None
    if not args.no_fused:
        fjob = skd.ShardedMesher(sdf, mn, mx, n, n, n, rank, world, spr, clip=True,
                                 balanced=(world > 1 and not args.uniform_slabs), colors=False)

        def fstep():
            counts = fjob.sample_classify()
            allc = skd.all_gather_int64(counts, device=dev) if world > 1 else counts[None]
            offs, tot_ = fjob.offsets(allc)
            fjob.emit(offs)
        for _ in range(3):
            fstep()
        barrier()
        ctx.mark(4)
        for _ in range(args.steps):
            fstep()
        ctx.mark(5)
        barrier()
        tf = torch.tensor([ctx.elapsed(4, 5) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fused = {"ms_per_step": tf.item(), "value": nvox_total / (tf.item() * 1e-3), "unit": UNIT,
                 "note": "same step with distance-only voxels (what Sdf.ToMesh runs): 4 B/voxel written, vertex colours from the SDF"}
        fjob.close()

    # ---- e2e through the public API: Sdf.ToMesh(min, max, n, n, n) -> Mesh in host memory, every step
    e2e = None
    if not args.no_e2e:
        barrier()
        e_steps = max(3, min(args.steps, 10))
        e_times = []
        d2h = 0
        if world == 1:
            for _ in range(2):                                   # warm: device + pinned host pools reach steady state
                mesh = sdf.ToMesh(mn, mx, n, n, n)               # (two result sets are alive while `mesh = sdf.ToMesh()` runs)
            torch.cuda.synchronize()
            e0 = time.perf_counter()
            for _ in range(e_steps):
                t_ = time.perf_counter()
                mesh = sdf.ToMesh(mn, mx, n, n, n)                # returns when the whole mesh is in host memory
                e_times.append(time.perf_counter() - t_)
            torch.cuda.synchronize()
            e_s = (time.perf_counter() - e0) / e_steps
            d2h = mesh.Vertices.nbytes * 3 + mesh.Triangles.nbytes + 24
        else:
            # N ranks: every rank delivers ITS share of the mesh (global indices from the count all-gather) into its own
            # page-locked host memory over its own PCIe link (sdfk_mesh_emit_host: emit in sub-ranges, copies overlapped);
            # the job's mesh is the concatenation of the shares in rank order
            # (slabs balanced for this path: an active cell also costs its ~60 bytes over PCIe, dist.ACTIVE_CELL_COST_E2E)
            ejob = skd.ShardedMesher(sdf, mn, mx, n, n, n, rank, world, spr, clip=True,
                                     balanced=(world > 1 and not args.uniform_slabs), colors=False,
                                     active_cell_cost=skd.ACTIVE_CELL_COST_E2E)

            def e2e_step():
                counts = ejob.sample_classify()
                allc = skd.all_gather_int64(counts, device=dev)
                offs, _ = ejob.offsets(allc)
                parts = ejob.emit_host(offs)
                return sum(m.Vertices.nbytes * 3 + m.Triangles.nbytes + 24 for m in parts)
            for _ in range(3):
                e2e_step()
            barrier()
            e0 = time.perf_counter()
            for _ in range(e_steps):
                t_ = time.perf_counter()
                d2h = e2e_step()
                e_times.append(time.perf_counter() - t_)
            barrier()
            e_s = (time.perf_counter() - e0) / e_steps
            ejob.close()
            tb = torch.tensor([float(d2h)], dtype=torch.float64, device=dev)
            dist.all_reduce(tb)
            d2h = int(tb.item())
        te = torch.tensor([e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_s = te.item()
        e2e = {"value": nvox_total / e_s, "unit": UNIT, "ms_per_step": e_s * 1e3, "steps": e_steps,
               "ms_per_step_median_rank0": statistics.median(e_times) * 1e3, "h2d_bytes_per_step": 256,
               "d2h_bytes_per_step": int(d2h),
               "note": ("Sdf.ToMesh through the host API (sdfk_sdf_to_mesh_host: z-slabs pipelined, mesh parts streamed to page-locked "
                        "host memory while the next slabs are computed)" if world == 1 else
                        "per rank: distance-only sampling + classify, NCCL all-gather of the counts, chunked emit with every part "
                        "streamed to the rank's own page-locked host memory (sdfk_mesh_emit_host); d2h bytes summed over ranks") +
                       "; the SDF is analytic so the only host->device bytes are kernel parameters; the whole mesh (vertices, "
                       "colours, normals, triangles) lands in host memory every step"}

    clocks = sampler.stop() if sampler else None          # sampled over all timed regions above (main loop, fused, e2e)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        val, tris, d = cpu_baseline(args.scene, args.cpu_n, 3, 1)
        cpu = {"value": val, "unit": UNIT, "cores": d["cores"], "kind": "port",
               "sample": "same scene and bounds at %d^3 (1/%d of the voxels); sampling on %d threads, marching cubes single-"
                         "threaded like the reference; C++ restatement, g++ -O2 -ffp-contract=off" % (
                             args.cpu_n, max(1, round((n / args.cpu_n) ** 3)), d["cores"]),
               "tris_per_s": tris, "detail": d}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.scene, n, world),
                       "grid": [n, n, n], "slabs_per_rank": spr, "slab_layers": [list(l) for l in job.layers], "sdf_nodes": sdf.lowered.node_count, "sdf_flops_per_sample": sdf.lowered.flops,
                       "l2": "working set %.1f GB per GPU >> 126 MB L2, no flush needed" % (16.0 * slab_vox / 1e9),
                       "parity_mode": "IEEE f32/f64, no FMA contraction (bit-exact vs the CPU oracle)"},
            "tris_per_s": ntris_total / (ms_per_step * 1e-3), "triangles": ntris_total, "vertices": int(tot[0]),
            "stages_ms": dict(sample_ms=k1_ms, **mesh_ms),
            "mesh": {"tris_per_s": ntris_total / max(mesh_total_ms * 1e-3, 1e-9), "cells_per_s": cells / max(mesh_total_ms * 1e-3, 1e-9),
                     "classify_gbs_rank0": 4.0 * slab_vox / (mesh_ms["classify_ms"] * 1e-3) / 1e9,
                     "note": "meshing stages only (K2-K4), slowest rank"},
            "per_rank_ms": rank_stages,
            "roofline": {"kernel": "sdfk_k_sample", "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": 16.0 * slab_vox, "launch_ms": k1_ms},
            "wall_ms_per_step": wall_ms / args.steps, "jit_compile_s": jit_s, "gpu_launches": int(launches),
            "clocks": clocks, "e2e": e2e, "fused_to_mesh": fused, "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    job.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=0, help="grid size override (default: 1024 per GPU-equivalent)")
    ap.add_argument("--scene", default="readme")
    ap.add_argument("--cpu-n", type=int, default=384, help="grid size of the bounded CPU-baseline sample (384^3: ~4 s per step on the box)")
    ap.add_argument("--slabs-per-rank", type=int, default=0, help="z-slabs dealt round-robin to every rank (default 1)")
    ap.add_argument("--uniform-slabs", action="store_true", help="equal-thickness z-slabs instead of the cost-balanced plan")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
