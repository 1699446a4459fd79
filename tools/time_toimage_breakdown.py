#!/usr/bin/env python3
"""Where the wall time of Sdf.ToImage goes: the C call alone (sdfk_render into a pre-allocated page-locked image), a plain
D2H copy of the same bytes, and the full Python call.  usage: python tools/time_toimage_breakdown.py [scene]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import sdfkit_b200 as sk
from sdfkit_b200 import _native as N, numerics, scenes
from bench import scene_by_name

scene = sys.argv[1] if len(sys.argv) > 1 else "readme"
w, h = 1920, 1080
ctx = sk.Context(0)
sdf = sk.GpuSdf(scene_by_name(scene)[0], ctx=ctx)
rm = sk.RayMarcher(w, h, sdf)
rm.ViewTransform = numerics.create_look_at(*scenes.CAMERA)
cam, ivp = rm.camera()
out = N.PinnedPool.empty((h, w, 3), np.float32)

def med(f, reps=15):
    ts = []
    for _ in range(reps + 3):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    ts = sorted(ts[3:])
    return ts[0], ts[len(ts) // 2]

c_call = lambda: N.check(N.lib().sdfk_render(ctx.handle, sdf.handle, w, h, N.fptr(cam), N.fptr(ivp), 1.0, 100.0, 40, 0, h, N.fptr(out)))
print("C call sdfk_render (host image):   best %.3f median %.3f ms" % med(c_call))
d = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
hp = torch.from_numpy(out)
def copy():
    hp.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("plain D2H copy of 24.9 MB (pinned): best %.3f median %.3f ms" % med(copy))
print("rm.Render():                       best %.3f median %.3f ms" % med(lambda: rm.Render()))
print("sdf.ToImage(...):                  best %.3f median %.3f ms" % med(lambda: sdf.ToImage(w, h, *scenes.CAMERA)))
print("camera matrices only:              best %.3f median %.3f ms" % med(lambda: rm.camera()))
print("PinnedPool.empty only:             best %.3f median %.3f ms" % med(lambda: N.PinnedPool.empty((h, w, 3), np.float32)))
