#!/usr/bin/env python3
"""Feasibility probe: does marching cubes (latency-bound gathers) hide under the sampling kernel (HBM-bound) when both run
at the same time on one GPU?  Two contexts = two streams on device 0: A re-samples its voxels while B meshes its own."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdfkit_b200 as sk
from sdfkit_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
expr, mn, mx = scenes.readme_scene()
A, B = sk.Context(0), sk.Context(0)
sa, sb = sk.GpuSdf(expr, ctx=A), sk.GpuSdf(expr, ctx=B)
va, vb = sa.ToVoxels(mn, mx, n, n, n), sb.ToVoxels(mn, mx, n, n, n)
A.synchronize(); B.synchronize()
def t_sample():
    t0 = time.perf_counter(); va.Resample(sa, clip=True); A.synchronize(); return (time.perf_counter() - t0) * 1e3
def t_mesh():
    t0 = time.perf_counter(); g = sk.MarchingCubes.CreateGpuMesh(vb); dt = (time.perf_counter() - t0) * 1e3; g.destroy(); return dt
def t_both():
    t0 = time.perf_counter(); va.Resample(sa, clip=True); g = sk.MarchingCubes.CreateGpuMesh(vb); tm = (time.perf_counter() - t0) * 1e3
    A.synchronize(); dt = (time.perf_counter() - t0) * 1e3; g.destroy(); return dt, tm
for _ in range(3): t_sample(); t_mesh(); t_both()
s = min(t_sample() for _ in range(5)); m = min(t_mesh() for _ in range(5))
b = [t_both() for _ in range(5)]
print("sample alone %.3f ms, mesh alone %.3f ms, sum %.3f; together %.3f ms (mesh part done at %.3f)" % (s, m, s + m, min(x[0] for x in b), min(x[1] for x in b)))
