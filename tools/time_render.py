#!/usr/bin/env python3
"""Times K5 (sdfk_k_render) on the device for one scene.  usage: python tools/time_render.py [readme|perf|csg50] [w h]"""
import os
import sys
import ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdfkit_b200 as sk
from sdfkit_b200 import _native as N, numerics, scenes
from bench import scene_by_name

scene = sys.argv[1] if len(sys.argv) > 1 else "readme"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
reps = int(os.environ.get("REPS", "10"))
expr = scene_by_name(scene)[0]
ctx = sk.Context(0)
sdf = sk.GpuSdf(expr, ctx=ctx)
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
flops = sdf.lowered.flops
print("%s %dx%d: best %.4f ms median %.4f ms; %d flops/eval -> %.3e IEEE op/s" % (
    scene, w, h, min(ts), sorted(ts)[len(ts) // 2], flops, (46 * flops + 70) * w * h / (min(ts) * 1e-3)))
