#!/usr/bin/env python3
"""A/B table of the JIT kernels: K1 (sample, 16 B/voxel), K1d (distance-only) at n^3 and K5 (1080p render) for the bench scenes.
usage: python tools/time_kernels.py [n]     (SDFK_PLAIN_BODY=1 / SDFK_JIT_DEFINES=... select the variant)"""
import os
import sys
import ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdfkit_b200 as sk
from sdfkit_b200 import _native as N, dist as skd, numerics, scenes
from bench import scene_by_name

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(os.environ.get("REPS", "5"))
ctx = sk.Context(0)
tag = "plain" if os.environ.get("SDFK_PLAIN_BODY") == "1" else "guards"
for scene in ("readme", "csg50", "perf"):
    expr, mn, mx = scene_by_name(scene)
    sdf = sk.GpuSdf(expr, ctx=ctx)
    row = []
    for colors in (True, False):
        s = skd.SlabMesher(sdf, mn, mx, n, n, n, 0, n - 1, True, 0.0, 1, colors)
        for _ in range(2):
            s.sample()
        ts = []
        for _ in range(reps):
            ctx.mark(0)
            s.sample()
            ctx.mark(1)
            ts.append(ctx.elapsed(0, 1))
        row.append(min(ts))
        s.close()
    w, h = 1920, 1080
    rm = sk.RayMarcher(w, h, sdf)
    rm.ViewTransform = numerics.create_look_at(*scenes.CAMERA)
    cam, ivp = rm.camera()
    buf = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
    ts = []
    for it in range(3 + reps):
        ctx.mark(0)
        N.check(N.lib().sdfk_render_device(ctx.handle, sdf.handle, w, h, N.fptr(cam), N.fptr(ivp), 1.0, 100.0, 40, 0, h, C.c_void_p(buf.data_ptr())))
        ctx.mark(1)
        if it >= 3:
            ts.append(ctx.elapsed(0, 1))
    print("%-7s %-7s K1 %8.3f ms   K1d %8.3f ms   K5 %7.4f ms   (%d flops/sample)" % (tag, scene, row[0], row[1], min(ts), sdf.lowered.flops), flush=True)
