#!/usr/bin/env python3
"""Wall time of Sdf.ToMesh (pipelined host mesh) for several slab counts.  usage: python tools/time_tomesh.py [n] [scene]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdfkit_b200 as sk
from bench import scene_by_name
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
expr, mn, mx = scene_by_name(sys.argv[2] if len(sys.argv) > 2 else "readme")
ctx = sk.Context(0)
sdf = sk.GpuSdf(expr, ctx=ctx)
for slabs in [int(x) for x in os.environ.get("SLABS", "1,4,8,16").split(",")]:
    for _ in range(3):
        m = sdf.ToMesh(mn, mx, n, n, n, slabs=slabs)
    ctx.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        m = sdf.ToMesh(mn, mx, n, n, n, slabs=slabs)
    dt = (time.perf_counter() - t0) / reps
    print("slabs=%d  ToMesh %.3f ms  (%d vertices, %d triangles)" % (slabs, dt * 1e3, len(m.Vertices), len(m.Triangles) // 3), flush=True)
