import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdfkit_b200 as sk
from bench import scene_by_name
expr, mn, mx = scene_by_name("readme")
sdf = sk.GpuSdf(expr, ctx=sk.Context(0))
n = 1024
for it in range(9):
    t0 = time.perf_counter()
    mesh = sdf.ToMesh(mn, mx, n, n, n)
    sys.stderr.write("== call %d: %.3f ms\n" % (it, (time.perf_counter() - t0) * 1e3))
