#!/usr/bin/env python3
"""Wall time of Sdf.ToImage through the public API (host image), like Perf/Program.cs:43-65.
usage: python tools/time_toimage.py [scene] [w h]     (SDFK_RENDER_BANDS=n: row bands of the render -> copy pipeline)"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sdfkit_b200 as sk
from sdfkit_b200 import scenes
from bench import scene_by_name

scene = sys.argv[1] if len(sys.argv) > 1 else "readme"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
sdf = sk.GpuSdf(scene_by_name(scene)[0], ctx=sk.Context(0))
ts = []
for it in range(13):
    t0 = time.perf_counter()
    img = sdf.ToImage(w, h, *scenes.CAMERA)
    ts.append((time.perf_counter() - t0) * 1e3)
ts = ts[3:]
print("%s %dx%d ToImage: best %.3f median %.3f ms  (bands=%s, checksum %.1f)" % (
    scene, w, h, min(ts), sorted(ts)[len(ts) // 2], os.environ.get("SDFK_RENDER_BANDS", "auto"), float(np.asarray(img.Array, dtype=np.float64).sum())), flush=True)
