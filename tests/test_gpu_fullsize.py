"""Parity at BASELINE.json's full sizes (512^3 config 2, 1024^3 bench size), where the CPU oracle would take minutes:
size-independent properties of the domain + spot checks against the oracle on sub-volumes.

  * sampled slabs of the 1024^3 grid (both walls and the middle) are bit-identical to the oracle's SDF at the same positions;
  * the mesh is a closed 2-manifold (every edge in exactly two triangles, opposite orientations) with Euler characteristic
    2 per sphere (25 spheres -> 50), vertex count ~ n^2, AABB inside the bounds;
  * Voxels.ToMesh, fused Sdf.ToMesh and the pipelined multi-slab Sdf.ToMesh give the same arrays bit for bit."""
import ctypes as C
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def manifold_stats(tris, nverts):
    t = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
    a = np.concatenate([t[:, 0], t[:, 1], t[:, 2]])
    b = np.concatenate([t[:, 1], t[:, 2], t[:, 0]])
    directed = a * nverts + b
    assert len(np.unique(directed)) == len(directed), "a directed edge occurs twice: inconsistent orientation"
    und = np.minimum(a, b) * nverts + np.maximum(a, b)
    uniq, cnt = np.unique(und, return_counts=True)
    return len(uniq), int(cnt.min()), int(cnt.max())


@pytest.mark.parametrize("n,expect_tris", [(512, 1950636), (1024, 7809020)])
def test_readme_scene_full_size_mesh_properties(n, expect_tris):
    import sdfkit_b200 as sk
    from sdfkit_b200 import scenes
    expr, mn, mx = scenes.readme_scene()
    sdf = expr.ToSdf()
    mesh = sdf.ToMesh(mn, mx, n, n, n)
    nv, nt = len(mesh.Vertices), len(mesh.Triangles) // 3
    assert nt == expect_tris                                      # measured in round 1 (RESULTS.md); ~ 4x per doubling of n
    assert abs(nv / (244192 * (n / 256) ** 2) - 1) < 0.01         # vertices ~ n^2 (256^3: 244,192)
    nedges, cmin, cmax = manifold_stats(mesh.Triangles, nv)
    assert (cmin, cmax) == (2, 2), "not a closed 2-manifold"
    assert nv - nedges + nt == 2 * 25                             # 5 x 5 spheres
    assert int(mesh.Triangles.min()) == 0 and int(mesh.Triangles.max()) == nv - 1
    assert np.all(mesh.Min >= np.float32(mn)) and np.all(mesh.Max <= np.float32(mx))
    assert np.allclose(mesh.Center, 0, atol=2e-3)
    assert np.isfinite(mesh.Normals).all() and np.allclose(np.linalg.norm(mesh.Normals[::97], axis=1), 1, atol=1e-5)
    assert mesh.Colors.min() >= 0.5666 and mesh.Colors.max() <= 0.9 + 1e-6      # colour range of the scene
    # every path to the same mesh: materialised voxels, one slab, many slabs
    ref = (sha(mesh.Vertices), sha(mesh.Colors), sha(mesh.Normals), sha(mesh.Triangles))
    for slabs in (1, 5):
        m2 = sdf.ToMesh(mn, mx, n, n, n, slabs=slabs)
        assert (sha(m2.Vertices), sha(m2.Colors), sha(m2.Normals), sha(m2.Triangles)) == ref, "slabs=%d" % slabs
        del m2
    vox = sdf.ToVoxels(mn, mx, n, n, n)
    m3 = vox.ToMesh()
    assert (sha(m3.Vertices), sha(m3.Colors), sha(m3.Normals), sha(m3.Triangles)) == ref, "Voxels.ToMesh"
    vox.Dispose()


@pytest.mark.parametrize("scene,z0", [("readme", 0), ("readme", 510), ("readme", 1020), ("csg50", 400)])
def test_sampled_slabs_of_the_1024_grid_match_the_oracle(oracle, scene, z0):
    """4 slices (4.2 M voxels) of the 1024^3 grid, sampled as a z-slab, against the oracle's SDF at the same positions."""
    import sdfkit_b200 as sk
    from sdfkit_b200 import _native as N, numerics
    from bench import scene_by_name
    expr, mn, mx = scene_by_name(scene)
    sdf = expr.ToSdf()
    n, nzl = 1024, 4
    vmin, vmax = numerics.vec3(mn), numerics.vec3(mx)
    h = C.c_void_p()
    N.check(N.lib().sdfk_voxels_sample_slab(sdf.ctx.handle, sdf.handle, N.fptr(vmin), N.fptr(vmax), n, n, n, 1, z0, z0 + nzl, C.byref(h)))
    vals = np.empty((n, n, nzl), dtype=np.float32)
    cols = np.empty((n, n, nzl, 3), dtype=np.float32)
    N.check(N.lib().sdfk_voxels_export(h, N.fptr(vals), N.fptr(cols)))
    N.lib().sdfk_voxels_destroy(h)
    f = np.float32
    d = ((vmax - vmin) / f(n)).astype(f)                                          # Voxels.cs:32-34
    m0 = (vmin + f(0.5) * d).astype(f)                                            # Voxels.cs:81
    ix, iy, iz = np.meshgrid(np.arange(n), np.arange(n), np.arange(z0, z0 + nzl), indexing="ij")
    pts = np.stack([m0[0] + ix.astype(f) * d[0], m0[1] + iy.astype(f) * d[1], m0[2] + iz.astype(f) * d[2]], axis=-1).astype(f)
    out = oracle.eval_sdf(sdf.lowered, pts.reshape(-1, 3)).reshape(n, n, nzl, 4)
    wall = (ix == 0) | (ix == n - 1) | (iy == 0) | (iy == n - 1) | (iz == 0) | (iz == n - 1)
    expect = np.where(wall, (vmax[0] - vmin[0]) / f(n), out[..., 3]).astype(f)    # ClipToBounds (Voxels.cs:133-167)
    assert np.array_equal(vals.view(np.uint32), expect.view(np.uint32)), "distances"
    assert np.array_equal(cols.view(np.uint32), np.ascontiguousarray(out[..., :3]).view(np.uint32)), "colours (untouched by the clip)"
