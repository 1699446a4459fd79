#!/usr/bin/env python3
"""Times the sampling kernels (K1 full / K1d distance-only) with and without the sign planes, and the mesh stages.
usage: python tools/time_sample.py [n] [scene]"""
import os
import sys
import ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sdfkit_b200 as sk
from sdfkit_b200 import _native as N, dist as skd
from bench import scene_by_name

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
scene = sys.argv[2] if len(sys.argv) > 2 else "readme"
reps = int(os.environ.get("REPS", "5"))
expr, mn, mx = scene_by_name(scene)
ctx = sk.Context(0)
sdf = sk.GpuSdf(expr, ctx=ctx)
for colors in (True, False):
    for opt in (0, 1):
        ctx.set_option(N.OPT_SIGN_PLANES, opt)
        s = skd.SlabMesher(sdf, mn, mx, n, n, n, 0, n - 1, True, 0.0, 1, colors)
        for _ in range(2):
            s.sample()
        ctx.mark(0)
        for _ in range(reps):
            s.sample()
        ctx.mark(1)
        t = ctx.elapsed(0, 1) / reps
        s.classify()
        s.emit(0, 0)
        st = s.mesh.stats()
        print("colors=%d signs=%d sample %.3f ms  classify %.3f scan %.3f compact %.3f emit %.3f" % (
            colors, opt, t, st["classify_ms"], st["scan_ms"], st["compact_ms"], st["emit_ms"]), flush=True)
        s.close()
