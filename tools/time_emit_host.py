#!/usr/bin/env python3
"""Wall time of the per-rank e2e path of a multi-GPU job, emulated on one GPU: sample (distances) + classify + chunked emit
streamed to host (sdfk_mesh_emit_host).  usage: python tools/time_emit_host.py [n] [world] [rank]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sdfkit_b200 as sk
from sdfkit_b200 import dist as skd
from bench import scene_by_name
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
expr, mn, mx = scene_by_name("readme")
ctx = sk.Context(0)
sdf = sk.GpuSdf(expr, ctx=ctx)
job = skd.ShardedMesher(sdf, mn, mx, n, n, n, rank, world, 1, clip=True, balanced=world > 1, colors=False)
print("layers", job.layers)
for chunks in (1, 0, 16):
    def step():
        t0 = time.perf_counter()
        counts = job.sample_classify()
        t1 = time.perf_counter()
        offs, tot = job.offsets(np.broadcast_to(counts[None], (world,) + counts.shape).copy() if world == 1 else np.stack([counts] * world))
        parts = job.emit_host(offs, chunks)
        t2 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, sum(len(p.Vertices) for p in parts)
    for _ in range(3):
        step()
    ts = [step() for _ in range(5)]
    print("chunks=%d  sample+classify %.3f ms  emit_host %.3f ms  total %.3f ms  (%d vertices)" % (
        chunks, np.mean([t[0] for t in ts]), np.mean([t[1] for t in ts]), np.mean([t[0] + t[1] for t in ts]), ts[0][2]), flush=True)
