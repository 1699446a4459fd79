"""Pins the CPU oracle against every known-answer value the reference's own tests hold for the hot
path.  Each test cites the reference test it restates."""
import numpy as np
import pytest

from oracle import sdf_numpy as S
from sdfkit_b200 import numerics

f32 = np.float32


def _dist_only(fn):
    """Opaque Sdfs.* lambdas only assign .W (Sdf.cs:135,153,211); colours stay zero."""
    return lambda p: fn(p)[:, 3]


def _mesh(oracle, fn, vmin, vmax, n, clip, writes_color=False):
    sdf = oracle.numpy_sdf(_dist_only(fn) if not writes_color else fn, writes_color=writes_color)
    vmin, vmax = numerics.vec3(vmin), numerics.vec3(vmax)
    values, colors = oracle.sample(sdf, vmin, vmax, n, n, n)
    if clip:
        oracle.clip(values, vmin, vmax)
    return oracle.marching_cubes(values, colors, vmin, vmax)


def test_create_mesh_sphere_1248(oracle):
    # Tests/SdfTests.cs:28-39  Sdfs.Sphere(0.5).ToMesh([-1,1]^3, 32^3) (clip on by default)
    m = _mesh(oracle, S.sphere(0.5), -1, 1, 32, clip=True)
    assert len(m.vertices) == 1248


def test_solid_sphere_expr_1248(oracle):
    # Tests/SdfTests.cs:42-52  SdfExprs.Solid(p => p.Length() - r).ToSdf().ToMesh(...)  -- via the lowering
    from sdfkit_b200.exprs import SdfExprs
    low = SdfExprs.Solid(lambda p: p.Length() - 0.5).Lower()
    m = oracle.to_mesh(low, numerics.vec3(-1), numerics.vec3(1), 32, 32, 32)
    assert len(m.vertices) == 1248
    assert np.allclose(m.colors, 1.0, atol=1e-6)


def test_colored_spheres_104(oracle):
    # Tests/MarchingCubesTests.cs:11-28
    r = f32(1)
    fn = S.union(S.translate(S.with_color(S.sphere(r * f32(0.4)), (1.0, 0.2, 0.3)), (-1, 0, 0)),
                 S.translate(S.with_color(S.sphere(r * f32(0.2)), (0.1, 1.0, 0.3)), (1, 0, 0)))
    m = _mesh(oracle, fn, -3, 3, 32, clip=False, writes_color=True)
    assert len(m.vertices) == 104
    assert len(m.colors) == 104
    assert m.colors[0][0] > 0.5


@pytest.mark.parametrize("r,half,n,expect,tol", [(1.0, 1.5, 5, 54, 0.3), (2.0, 2.5, 10, 312, 0.2)])
def test_sphere_5_and_10(oracle, r, half, n, expect, tol):
    # Tests/MarchingCubesTests.cs:30-62
    m = _mesh(oracle, S.sphere(r), -half, half, n, clip=False)
    assert len(m.vertices) == expect
    assert numerics.length3(m.center) < 1e-6
    assert abs(m.size[0] / 2 - r) < tol


def test_unclipped_sphere10_empty(oracle):
    # Tests/MarchingCubesTests.cs:64-79
    m = _mesh(oracle, S.sphere(2.0), -1, 1, 10, clip=False)
    assert len(m.vertices) == 0 and len(m.triangles) == 0


def test_clipped_sphere10_384(oracle):
    # Tests/MarchingCubesTests.cs:81-98
    m = _mesh(oracle, S.sphere(2.0), -1, 1, 10, clip=True)
    assert len(m.vertices) == 384
    assert numerics.length3(m.center) < 1e-6
    assert abs(m.size[0] - 2.0) < 1e-1


def test_box10_384(oracle):
    # Tests/MarchingCubesTests.cs:100-115
    m = _mesh(oracle, S.box(2.0), -2.5, 2.5, 10, clip=False)
    assert len(m.vertices) == 384
    assert numerics.length3(m.center) < 1e-6
    assert abs(m.size[0] / 2 - 2.0) < 3e-1


def test_cylinder50_7456(oracle):
    # Tests/MarchingCubesTests.cs:117-138  Sdfs.Cylinder(1,3) == SdfExprs.Cylinder(1,3).ToSdf() (Sdf.cs:200-201)
    sdf = oracle.numpy_sdf(S.cylinder(1, 3))
    mn, mx = numerics.vec3(-1.5, -3.5, -1.5), numerics.vec3(1.5, 3.5, 1.5)
    values, colors = oracle.sample(sdf, mn, mx, 50, 50, 50)
    m = oracle.marching_cubes(values, colors, mn, mx)
    assert len(m.vertices) == 7456
    assert np.all(np.abs(m.center) < 1e-6)
    assert abs(m.size[0] / 2 - 1) < 1e-1


def test_sphere128_progress_72240(oracle):
    # Tests/MarchingCubesTests.cs:140-171
    seen = []
    sdf = oracle.numpy_sdf(_dist_only(S.sphere(3.0)), writes_color=False)
    mn, mx = numerics.vec3(f32(-3.1)), numerics.vec3(f32(3.1))
    values, colors = oracle.sample(sdf, mn, mx, 128, 128, 128)
    m = oracle.marching_cubes(values, colors, mn, mx, progress=seen.append)
    assert len(m.vertices) == 72240
    assert all(0.0 <= f <= 1.0 for f in seen)
    assert any(f < 1e-6 for f in seen) and any(1.0 - f < 1e-6 for f in seen)
    assert numerics.length3(m.center) < 1e-6
    assert abs(m.size[0] / 2 - 3.0) < 0.1


# ---------------------------------------------------------------- Voxels (Tests/VolumeTests.cs, Tests/SdfTests.cs)

def test_one_is_centered(oracle):
    # Tests/VolumeTests.cs:39-58
    seen = []
    def fn(p):
        seen.append(p.copy())
        return np.ones((len(p), 4), dtype=np.float32)
    v, _ = oracle.sample(oracle.numpy_sdf(fn), numerics.vec3(-1), numerics.vec3(1), 1, 1, 1)
    assert v[0, 0, 0] == 1.0
    assert np.all(np.abs(np.concatenate(seen)) < 1e-3)


def test_three_has_center(oracle):
    # Tests/VolumeTests.cs:60-80
    seen = []
    def fn(p):
        seen.append(p.copy())
        return np.ones((len(p), 4), dtype=np.float32)
    oracle.sample(oracle.numpy_sdf(fn), numerics.vec3(-1), numerics.vec3(1), 3, 3, 3)
    pts = np.concatenate(seen)
    assert np.any(np.sqrt((pts * pts).sum(1)) < 1e-3)


def test_sphere_5_center_value(oracle):
    # Tests/VolumeTests.cs:82-106
    v, _ = oracle.sample(oracle.numpy_sdf(S.sphere(0.5)), numerics.vec3(-1), numerics.vec3(1), 5, 5, 5)
    assert abs(v[2, 2, 2] + 0.5) < 1e-3


def test_sphere_128_batch70(oracle):
    # Tests/VolumeTests.cs:108-135: every batch has 70 points except one of 22
    v, _, bs = oracle.sample(oracle.numpy_sdf(S.sphere(0.5)), numerics.vec3(-1), numerics.vec3(1), 128, 128, 128,
                             batch_size=70, threads=4, return_batch_sizes=True)
    assert abs(v[63, 63, 63] + 0.5) < 2e-2
    assert sorted(set(bs.tolist())) == [22, 70] and int((bs == 22).sum()) == 1


def test_to_voxels_128(oracle):
    # Tests/SdfTests.cs:12-26 (clip on: the probed voxel is interior, unaffected)
    v, _ = oracle.to_voxels(oracle.numpy_sdf(S.sphere(0.5)), numerics.vec3(-1), numerics.vec3(1), 128, 128, 128, threads=4)
    assert abs(v[63, 63, 63] + 0.5) < 2e-2
    assert v[0, 5, 5] == f32(2.0) / f32(128)


# ---------------------------------------------------------------- RayMarcher (Tests/RayMarcherTests.cs)

def test_sphere_depth(oracle):
    # :10-26
    img = oracle.render_depth(oracle.numpy_sdf(_dist_only(S.sphere(1.0)), writes_color=False), 50, 30)
    assert img.shape == (30, 50)
    assert abs(img[15, 25] - 4.0) < 1e-2
    assert img[0, 0] > 9.0


def test_box_depth(oracle):
    # :28-42
    img = oracle.render_depth(oracle.numpy_sdf(_dist_only(S.box(1.0)), writes_color=False), 50, 30)
    assert abs(img[15, 25] - 4.0) < 1e-2
    assert img[0, 0] > 9.0


def test_cylinder_repeat_depth(oracle):
    # :44-60  SdfExprs.Cylinder(r, 2r).RepeatX(4r), checked through BOTH the numpy restatement and the lowering
    from sdfkit_b200.exprs import SdfExprs
    r = f32(0.25)
    a = oracle.render_depth(oracle.numpy_sdf(S.repeat(S.cylinder(r, r * 2), sx=4 * r)), 50, 30)
    b = oracle.render_depth(SdfExprs.Cylinder(r, r * 2).RepeatX(4 * r).Lower(), 50, 30)
    for img in (a, b):
        assert abs(img[15 - 2, 25] - (5 - r)) < 1e-1
        assert img[0, 0] > 9.0
    assert np.array_equal(a, b)


def test_plane_depth(oracle):
    # :62-75  Sdfs.PlaneXY()
    img = oracle.render_depth(oracle.numpy_sdf(_dist_only(S.plane((0, 0, 1), 0.0)), writes_color=False), 50, 30)
    assert abs(img[15, 25] - 5.0) < 1e-2
    assert img[0, 0] < 9.0
