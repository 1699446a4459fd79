#!/usr/bin/env python3
"""Runs the five BASELINE.json configurations on one B200 and prints one JSON line each + a markdown table
(committed as RESULTS.md).  GPU numbers: CUDA events on the library's stream, best of `--reps` after 3 warm-ups.
CPU numbers: the oracle (C++ restatement of the reference's CPU path) on a bounded sample of the same workload.

    python tools/run_configs.py [--reps 5] [--skip-cpu] [--fp32-peak]
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def gpu_mesh_config(sk, skd, ctx, name, expr, mn, mx, n, reps):
    sdf = sk.GpuSdf(expr, ctx=ctx)
    slab = skd.SlabMesher(sdf, mn, mx, n, n, n, 0, skd.cells_along(n, 1), clip=True)
    best = None
    for it in range(3 + reps):
        ctx.mark(0)
        slab.sample()
        ctx.mark(1)
        nv, nt = slab.classify()
        slab.emit(0, 0)
        ctx.mark(2)
        t_sample, t_total = ctx.elapsed(0, 1), ctx.elapsed(0, 2)
        st = slab.mesh.stats()
        if it >= 3 and (best is None or t_total < best["total_ms"]):
            best = {"total_ms": t_total, "sample_ms": t_sample, "mesh_ms": t_total - t_sample, "stages": st, "nv": nv, "nt": nt}
    slab.close()
    nvox = n ** 3
    hbm = peaks()
    out = {
        "config": name, "grid": n, "sdf_nodes": sdf.lowered.node_count, "sdf_flops_per_sample": sdf.lowered.flops,
        "vertices": best["nv"], "triangles": best["nt"], "active_cells": best["stages"]["active_cells"],
        "sample_ms": best["sample_ms"], "mesh_ms": best["mesh_ms"], "total_ms": best["total_ms"],
        "samples_per_s": nvox / (best["sample_ms"] * 1e-3), "sample_gbs": 16.0 * nvox / (best["sample_ms"] * 1e-3) / 1e9,
        "sample_frac_hbm": 16.0 * nvox / (best["sample_ms"] * 1e-3) / 1e9 / hbm,
        "tris_per_s": best["nt"] / (best["mesh_ms"] * 1e-3), "cells_per_s": (n - 1) ** 3 / (best["mesh_ms"] * 1e-3),
        "classify_gbs": 4.0 * nvox / (best["stages"]["classify_ms"] * 1e-3) / 1e9,
        "classify_frac_hbm": 4.0 * nvox / (best["stages"]["classify_ms"] * 1e-3) / 1e9 / hbm,
        "stages_ms": {k: best["stages"][k] for k in ("classify_ms", "scan_ms", "compact_ms", "emit_ms")},
        "step_voxels_per_s": nvox / (best["total_ms"] * 1e-3),
    }
    return out, sdf


def cpu_mesh_config(expr, mn, mx, n):
    """CPU column of the table: bench.py's cpu_baseline leg (the only code outside tests/ that may run the oracle)."""
    import bench
    return bench.cpu_baseline_expr(expr, mn, mx, n)


def gpu_render_config(sk, ctx, name, expr, w, h, reps):
    import torch
    from sdfkit_b200 import _native as N, numerics, scenes
    sdf = sk.GpuSdf(expr, ctx=ctx)
    rm = sk.RayMarcher(w, h, sdf)
    rm.ViewTransform = numerics.create_look_at(*scenes.CAMERA)
    cam, ivp = rm.camera()
    buf = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
    times = []
    for it in range(3 + reps):
        ctx.mark(0)
        N.check(N.lib().sdfk_render_device(ctx.handle, sdf.handle, w, h, N.fptr(cam), N.fptr(ivp), 1.0, 100.0, 40, 0, h, C.c_void_p(buf.data_ptr())))
        ctx.mark(1)
        if it >= 3:
            times.append(ctx.elapsed(0, 1))
    for _ in range(2):                      # warm: page-locked image buffers reach steady state
        img = rm.Render()
    t0 = time.perf_counter()
    for _ in range(reps):
        img = rm.Render()
    e2e = (time.perf_counter() - t0) / reps
    ms = min(times)
    evals = 46 * w * h
    ops = evals * sdf.lowered.flops + 70 * w * h
    return {"config": name, "image": [w, h], "render_ms": ms, "pixels_per_s": w * h / (ms * 1e-3), "sdf_evals_per_s": evals / (ms * 1e-3),
            "fp32_ops_per_s": ops / (ms * 1e-3), "sdf_flops_per_sample": sdf.lowered.flops, "e2e_ms": e2e * 1e3,
            "e2e_pixels_per_s": w * h / e2e, "checksum": float(img.Array.sum())}, sdf


def cpu_render_config(expr, w, h):
    import bench
    return bench.cpu_baseline_render(expr, w, h)


FP32_SRC = r"""
extern "C" __global__ void fp32_chain(float* out, int iters)
{
    float a = threadIdx.x * 1e-3f + 1.0f, b = 1.0001f, c = 0.9999f, d = a + 1.0f, e = a + 2.0f, f = a + 3.0f, g = a + 4.0f, h = a + 5.0f;
    for (int i = 0; i < iters; i++) {   // 8 independent chains, alternating FMUL / FADD (no FMA: compiled with --fmad=false)
        a = a * b; d = d * b; e = e * b; f = f * b; g = g * b; h = h * b;
        a = a + c; d = d + c; e = e + c; f = f + c; g = g + c; h = h + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f + g + h;
}
"""


def fp32_peak():
    """Non-FMA FP32 lane-op rate (FMUL + FADD) on this GPU, via torch's NVRTC-free path: compile with nvcc at run time."""
    import subprocess
    import tempfile
    import torch
    d = tempfile.mkdtemp()
    cu = os.path.join(d, "fp32.cu")
    open(cu, "w").write(FP32_SRC)
    cubin = os.path.join(d, "fp32.cubin")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-O3", "-fmad=false", "-o", cubin, cu], check=True)
    from cuda.bindings import driver as cu_drv  # cuda-python
    torch.zeros(1, device="cuda")
    err, mod = cu_drv.cuModuleLoadData(open(cubin, "rb").read())
    err, fn = cu_drv.cuModuleGetFunction(mod, b"fp32_chain")
    out = torch.empty(148 * 16 * 256, dtype=torch.float32, device="cuda")
    iters = 20000
    args = (np.array([out.data_ptr()], dtype=np.uint64), np.array([iters], dtype=np.int32))
    argp = np.array([a.ctypes.data for a in args], dtype=np.uint64)
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        cu_drv.cuLaunchKernel(fn, 148 * 16, 1, 1, 256, 1, 1, 0, torch.cuda.current_stream().cuda_stream, argp.ctypes.data, 0)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    ops = 148 * 16 * 256 * iters * 12.0
    return ops / (best * 1e-3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--fp32-peak", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import sdfkit_b200 as sk
    from sdfkit_b200 import dist as skd, scenes
    ctx = sk.Context(0)
    rows = []
    if args.fp32_peak:
        try:
            r = fp32_peak()
            print(json.dumps({"fp32_nofma_lane_ops_per_s": r}))
            rows.append({"config": "fp32 peak (FMUL+FADD chains, no FMA)", "fp32_ops_per_s": r})
        except Exception as ex:   # the measurement is optional
            print(json.dumps({"fp32_peak_error": repr(ex)}))
    mesh_cfgs = [
        ("1: Sphere(0.5) 64^3", scenes.sphere(), 64, 64),
        ("2: README RepeatXY 512^3", scenes.readme_scene(), 512, 256),
        ("2': README RepeatXY 1024^3 (bench.py workload)", scenes.readme_scene(), 1024, 256),
        ("3: CSG-50 1024^3", scenes.csg50(), 1024, 192),
    ]
    for name, (expr, mn, mx), n, ncpu in mesh_cfgs:
        if args.only and args.only not in name:
            continue
        r, _ = gpu_mesh_config(sk, skd, ctx, name, expr, mn, mx, n, args.reps)
        if not args.skip_cpu:
            r.update(cpu_mesh_config(expr, mn, mx, ncpu))
        print(json.dumps(r))
        rows.append(r)
    for name, expr, w, h, wc, hc in [("5: ToImage README scene 1920x1080", scenes.readme_scene()[0], 1920, 1080, 480, 270),
                                     ("5': ToImage Perf scene (spheres + boxes) 1920x1080", scenes.perf_scene()[0], 1920, 1080, 480, 270)]:
        if args.only and args.only not in name:
            continue
        r, _ = gpu_render_config(sk, ctx, name, expr, w, h, args.reps)
        if not args.skip_cpu:
            r.update(cpu_render_config(expr, wc, hc))
        print(json.dumps(r))
        rows.append(r)
    # markdown
    print("\n| config | GPU | CPU restatement (bounded sample) |")
    print("|---|---|---|")
    for r in rows:
        if "samples_per_s" in r:
            g = "sample %.3g voxels/s (%.2f ms, %.0f GB/s = %.2f of HBM peak); mesh %.3g tris/s, %.3g cells/s (%.2f ms; classify %.0f GB/s = %.2f); %d tris" % (
                r["samples_per_s"], r["sample_ms"], r["sample_gbs"], r["sample_frac_hbm"], r["tris_per_s"], r["cells_per_s"], r["mesh_ms"],
                r["classify_gbs"], r["classify_frac_hbm"], r["triangles"])
            c = "" if "cpu_grid" not in r else "%d^3 on %d cores: sample %.3g voxels/s; mesh (1 thread) %.3g tris/s, %.3g cells/s" % (
                r["cpu_grid"], r["cpu_cores"], r["cpu_samples_per_s"], r["cpu_tris_per_s"], r["cpu_cells_per_s"])
        elif "render_ms" in r:
            g = "%.3f ms = %.3g pixels/s = %.3g SDF evals/s = %.3g FP32 op/s; e2e (host image) %.2f ms" % (
                r["render_ms"], r["pixels_per_s"], r["sdf_evals_per_s"], r["fp32_ops_per_s"], r["e2e_ms"])
            c = "" if "cpu_image" not in r else "%dx%d on %d cores: %.0f ms = %.3g pixels/s" % (
                r["cpu_image"][0], r["cpu_image"][1], r["cpu_cores"], r["cpu_render_ms"], r["cpu_pixels_per_s"])
        else:
            g, c = "%.3g lane-op/s" % r["fp32_ops_per_s"], ""
        print("| %s | %s | %s |" % (r["config"], g, c))


if __name__ == "__main__":
    main()
