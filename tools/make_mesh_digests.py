#!/usr/bin/env python3
"""Writes tests/golden/mesh_digests.json: sha256 of the four arrays of the single-GPU mesh (Sdf.ToMesh, pipelined z-slabs) of
the bench scene at the grids bench.py runs (1024^3 at N = 1 / strong scaling, 1280^3, 1624^3, 2048^3 in the weak-scaling
runs).  bench.py's parity_check holds the N-rank job's rank-ordered output to these digests.  Run on one B200:
    python tools/make_mesh_digests.py [out.json] [grid ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sdfkit_b200 as sk  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "mesh_digests.json")
grids = [int(a) for a in sys.argv[2:]] or [256, 1024, 1280, 1624, 2048]
res = {"_how": "tools/make_mesh_digests.py on one B200: sha256 of Mesh.Vertices / Colors / Normals / Triangles bytes of sdf.ToMesh(min, max, n, n, n) "
               "(clip on, iso 0, step 1); the same arrays are bit-identical to the CPU oracle's at every size the oracle is run on (tests/)"}
for scene in ("readme",):
    expr, mn, mx = bench.scene_by_name(scene)
    sdf = expr.ToSdf()
    res[scene] = {}
    for n in grids:
        m = sdf.ToMesh(mn, mx, n, n, n)
        res[scene][str(n)] = {"vertices": int(len(m.Vertices)), "triangles": int(len(m.Triangles) // 3), "sha256": bench.mesh_sha(m)}
        print(scene, n, res[scene][str(n)]["vertices"], res[scene][str(n)]["triangles"], flush=True)
        del m
json.dump(res, open(out_path, "w"), indent=1)
