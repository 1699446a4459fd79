#!/usr/bin/env python3
"""The multi-device context with ONE physical GPU listed several times ("virtual devices": every entry has its own streams,
pools and worker thread): does meshing slab i under the sampling of slab i+1 pay on a single GPU?
usage: python tools/time_virtual.py [n] [ndev ...]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdfkit_b200 as sk
from bench import scene_by_name

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
devs = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
expr, mn, mx = scene_by_name(os.environ.get("SCENE", "readme"))
for nd in devs:
    ctx = sk.Context(devices=[0] * nd) if nd > 1 else sk.Context(0)
    sdf = sk.GpuSdf(expr, ctx=ctx)
    ts = []
    for it in range(8):
        t0 = time.perf_counter()
        m = sdf.ToMesh(mn, mx, n, n, n)
        ts.append((time.perf_counter() - t0) * 1e3)
    td = [float("nan")] * 8
    vox = sdf.ToVoxels(mn, mx, n, n, n)
    for it in range(8):
        t0 = time.perf_counter()
        vox.Resample(sdf, clip=True)
        gm = sk.MarchingCubes.CreateGpuMesh(vox)
        td[it] = (time.perf_counter() - t0) * 1e3
        gm.destroy()
    vox.Dispose()
    print("%d virtual devices on GPU 0  %d^3: ToMesh(host) best %.3f median %.3f ms | device step best %.3f median %.3f ms | %d tris" % (
        nd, n, min(ts[2:]), sorted(ts[2:])[3], min(td[2:]), sorted(td[2:])[3], len(m.Triangles) // 3), flush=True)
    del m
    sdf.Dispose()
    ctx.close()
