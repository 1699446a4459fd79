"""GPU tests of the multi-GPU layer behind the C ABI (sdfk_ctx_create_multi, csrc/sdfk_multi.inl) and of the round-2
(f)-row features.  On a one-GPU box the job context lists device 0 several times: every entry still gets its own
stream, pools and worker thread, so the slab planner, the count exchange, the global offsets and the per-device copies
into ONE host result are exercised exactly as on N GPUs.  With >= 2 GPUs the same tests spread over real devices, and
test_two_rank_torchrun_job launches the one-process-per-GPU NCCL path of bench.py.
Everything is compared bit for bit with the single-GPU result and with the CPU oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


@pytest.fixture(scope="module")
def sk():
    import sdfkit_b200
    return sdfkit_b200


def device_list(n):
    import torch
    have = torch.cuda.device_count()
    return [k % have for k in range(n)]


@pytest.fixture(scope="module")
def ctx1(sk):
    c = sk.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=[2, 3, 5])
def mctx(sk, request):
    c = sk.Context(devices=device_list(request.param))
    assert c.device_count() == request.param
    yield c
    c.close()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_mesh(a, b, what):
    assert len(a.Vertices) == len(b.Vertices) and len(a.Triangles) == len(b.Triangles), what + ": counts differ"
    assert np.array_equal(np.asarray(a.Triangles), np.asarray(b.Triangles)), what + ": triangles differ"
    for name in ("Vertices", "Colors", "Normals"):
        assert np.array_equal(bits(getattr(a, name)), bits(getattr(b, name))), "%s: %s differ" % (what, name)
    if len(a.Vertices):
        assert np.array_equal(bits(a.Min), bits(b.Min)) and np.array_equal(bits(a.Max), bits(b.Max)), what + ": AABB differs"


def scene(name):
    from sdfkit_b200 import scenes
    return {"sphere": scenes.sphere, "readme": scenes.readme_scene, "perf": scenes.perf_scene, "csg50": scenes.csg50}[name]()


CASES = [("readme", (96, 96, 96), 1, 0.0), ("sphere", (64, 64, 64), 1, 0.0), ("perf", (50, 37, 29), 1, 0.0),
         ("csg50", (96, 96, 48), 1, 0.0), ("readme", (70, 52, 33), 2, 0.0), ("readme", (64, 64, 64), 1, 0.1),
         ("sphere", (40, 40, 3), 1, 0.0)]          # last: fewer cell layers than devices


@pytest.mark.parametrize("name,dims,step,iso", CASES)
def test_multi_to_mesh_equals_single_gpu_and_oracle(sk, oracle, ctx1, mctx, name, dims, step, iso):
    """SdfEx.ToMesh on N devices lands ONE host mesh identical to the single-GPU one (and to the oracle's)."""
    expr, mn, mx = scene(name)
    nx, ny, nz = dims
    one = sk.GpuSdf(expr, ctx=ctx1).ToMesh(mn, mx, nx, ny, nz, isoValue=iso, step=step)
    many = sk.GpuSdf(expr, ctx=mctx).ToMesh(mn, mx, nx, ny, nz, isoValue=iso, step=step)
    same_mesh(many, one, "%s %s step %d iso %g on %d devices" % (name, dims, step, iso, mctx.device_count()))
    om = oracle.to_mesh(sk.GpuSdf(expr, ctx=ctx1).lowered, f32(mn), f32(mx), nx, ny, nz, iso=iso, step=step)
    assert np.array_equal(np.asarray(many.Triangles).reshape(-1, 3), om.triangles)
    assert np.array_equal(bits(many.Vertices), bits(om.vertices)) and np.array_equal(bits(many.Colors), bits(om.colors))
    assert np.array_equal(bits(many.Normals), bits(om.normals))


@pytest.mark.parametrize("cap", [1, 5, 17])
def test_multi_to_mesh_with_several_sub_slabs_per_device(sk, ctx1, mctx, monkeypatch, cap):
    """A device's layer range is meshed in sub-slabs (32-bit cell ids bound one job to 2^32 cells: 2048^3 on 2 devices needs
    them); forced here on a small grid."""
    expr, mn, mx = scene("readme")
    one = sk.GpuSdf(expr, ctx=ctx1).ToMesh(mn, mx, 80, 64, 72)
    monkeypatch.setenv("SDFK_SUBSLAB_CAP", str(cap))
    many = sk.GpuSdf(expr, ctx=mctx).ToMesh(mn, mx, 80, 64, 72)
    same_mesh(many, one, "sub-slab cap %d on %d devices" % (cap, mctx.device_count()))
    stepped = sk.GpuSdf(expr, ctx=mctx).ToMesh(mn, mx, 80, 64, 72, step=2, isoValue=0.05)
    monkeypatch.delenv("SDFK_SUBSLAB_CAP")
    same_mesh(stepped, sk.GpuSdf(expr, ctx=ctx1).ToMesh(mn, mx, 80, 64, 72, step=2, isoValue=0.05), "sub-slabs, step 2")


@pytest.mark.parametrize("name,dims", [("readme", (96, 96, 96)), ("perf", (50, 37, 29)), ("sphere", (40, 40, 3)), ("readme", (260, 9, 7))])
def test_sharded_voxels_and_mesh_equal_single_gpu(sk, oracle, ctx1, mctx, name, dims):
    """sdfk_voxels_sample on a multi context shards by z-slab; export, clip, resample and sdfk_mesh_create on the shards give
    the single-GPU (= oracle) result."""
    from sdfkit_b200 import _native as N
    import ctypes as C
    expr, mn, mx = scene(name)
    nx, ny, nz = dims
    s1, sm = sk.GpuSdf(expr, ctx=ctx1), sk.GpuSdf(expr, ctx=mctx)
    v1 = s1.ToVoxels(mn, mx, nx, ny, nz)
    vm = sm.ToVoxels(mn, mx, nx, ny, nz)
    nparts = C.c_int()
    lay = (C.c_int * 128)()
    N.check(N.lib().sdfk_voxels_layers(vm.handle, lay, 64, C.byref(nparts)))
    assert nparts.value == mctx.device_count()
    cuts = [(lay[2 * k], lay[2 * k + 1]) for k in range(nparts.value)]
    assert cuts[0][0] == 0 and cuts[-1][1] == max(nz - 1, 0) and all(cuts[k][1] == cuts[k + 1][0] for k in range(len(cuts) - 1))
    assert np.array_equal(bits(vm.Values), bits(v1.Values)) and np.array_equal(bits(vm.Colors), bits(v1.Colors))
    ov, oc = oracle.to_voxels(s1.lowered, f32(mn), f32(mx), nx, ny, nz, clip_to_bounds=True, threads=4)
    assert np.array_equal(bits(vm.Values), bits(ov)) and np.array_equal(bits(vm.Colors), bits(oc))
    same_mesh(vm.ToMesh(), v1.ToMesh(), "%s %s sharded mesh" % (name, dims))
    same_mesh(vm.ToMesh(0.05), v1.ToMesh(0.05), "%s %s sharded mesh, iso 0.05" % (name, dims))
    # unclipped + explicit ClipToBounds (classifies from the distances), then an in-place resample
    u1, um = s1.ToVoxels(mn, mx, nx, ny, nz, clipToBounds=False), sm.ToVoxels(mn, mx, nx, ny, nz, clipToBounds=False)
    u1.ClipToBounds()
    um.ClipToBounds()
    assert np.array_equal(bits(um.Values), bits(v1.Values))
    same_mesh(um.ToMesh(), u1.ToMesh(), "clip after sampling")
    um.Resample(sm, clip=True)
    same_mesh(um.ToMesh(), v1.ToMesh(), "resample")
    with pytest.raises(N.SdfkError):
        vm.ToMesh(0.0, 2)                     # sharded voxels hold one halo slice: step 1 only (documented)


def test_multi_render_row_bands(sk, oracle, ctx1, mctx, tmp_path):
    from sdfkit_b200 import numerics, scenes
    expr = scenes.readme_scene()[0]
    view = numerics.create_look_at(*scenes.CAMERA)
    imgs = []
    for ctx in (ctx1, mctx):
        rm = sk.RayMarcher(161, 91, sk.GpuSdf(expr, ctx=ctx))
        rm.ViewTransform = view
        p = str(tmp_path / ("img%d.tga" % len(imgs)))
        rm.RenderTga(p)
        imgs.append((rm.Render().Array.copy(), rm.RenderDepth().Array.copy(), open(p, "rb").read(), rm.RenderDepthGray(1.0, 9.0).copy()))
    for a, b in zip(imgs[0], imgs[1]):
        assert (a == b) if isinstance(a, bytes) else np.array_equal(a.view(np.uint8), b.view(np.uint8))
    ref = oracle.render(sk.GpuSdf(expr, ctx=ctx1).lowered, 161, 91, view=view, bands=2)
    assert np.array_equal(bits(imgs[1][0]), bits(ref))


def test_plan_layers_native(sk, ctx1):
    from sdfkit_b200 import dist as skd
    expr, mn, mx = scene("readme")
    sdf = sk.GpuSdf(expr, ctx=ctx1)
    for parts in (1, 2, 4, 8):
        cuts = skd.plan_layers(sdf, mn, mx, 256, 256, 256, parts)
        assert cuts[0][0] == 0 and cuts[-1][1] == 255 and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert all(ke > kb for kb, ke in cuts)
    # the README scene's surface sits in the middle of z: the middle slabs must be thinner than the outer ones
    cuts = skd.plan_layers(sdf, mn, mx, 256, 256, 256, 4)
    th = [ke - kb for kb, ke in cuts]
    assert th[1] < th[0] and th[2] < th[3], th
    assert skd.plan_layers(sdf, mn, mx, 64, 64, 3, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


# ------------------------------------------------------------------------------------------ (f)-row features

def test_depth_tga_on_device_equals_save_depth_tga(sk, tmp_path):
    """FloatData.SaveDepthTga (VectorData.cs:244-276): device byte conversion == host conversion, byte for byte."""
    from sdfkit_b200 import numerics, scenes
    rm = sk.RayMarcher(200, 120, scenes.readme_scene()[0].ToSdf())
    rm.ViewTransform = numerics.create_look_at(*scenes.CAMERA)
    for near, far in ((1.0, 9.0), (2.0, 6.5), (0.5, 100.0)):
        a, b = str(tmp_path / "a.tga"), str(tmp_path / "b.tga")
        rm.RenderDepth().SaveDepthTga(a, near, far)
        rm.RenderDepthTga(b, near, far)
        da, db = open(a, "rb").read(), open(b, "rb").read()
        assert len(da) == 18 + 200 * 120 and da == db, (near, far)
        px = np.frombuffer(da[18:], dtype=np.uint8)
        assert px.min() == 0 and px.max() > 100          # background and near surface both present


def test_with_color_matches_oracle(sk, oracle):
    """SdfEx.WithColor (Sdf.cs:101-115): same distances, every colour replaced -- on voxels, mesh and image."""
    expr, mn, mx = scene("readme")
    sdf = expr.ToSdf()
    red = sdf.WithColor(1.0, 0.25, 0.125)
    pts = np.random.default_rng(5).uniform(-3, 3, (4099, 3)).astype(f32)
    base, out = sdf(pts), red(pts)
    assert np.array_equal(bits(out[:, 3]), bits(base[:, 3]))
    assert np.array_equal(out[:, :3], np.broadcast_to(f32([1.0, 0.25, 0.125]), (len(pts), 3)))
    assert np.array_equal(bits(out), bits(oracle.eval_sdf(red.lowered, pts)))
    m0, m1 = sdf.ToMesh(mn, mx, 64, 64, 64), red.ToMesh(mn, mx, 64, 64, 64)
    assert np.array_equal(bits(m0.Vertices), bits(m1.Vertices)) and np.array_equal(m0.Triangles, m1.Triangles)
    om = oracle.to_mesh(red.lowered, f32(mn), f32(mx), 64, 64, 64)
    assert np.array_equal(bits(m1.Colors), bits(om.colors))
    assert np.allclose(m1.Colors, [1.0, 0.25, 0.125], atol=1e-6)
    grey = sdf.WithColor((0.5, 0.5, 0.5))                  # WithColor(Vector3)
    assert np.array_equal(grey(pts)[:, :3], np.full((len(pts), 3), 0.5, dtype=f32))


def test_voxel_indexer_writes_reach_the_mesh(sk, oracle):
    """The reference's Values is live storage (Voxels.cs:42-64): a write through the indexers must show in the next mesh;
    the exported arrays themselves are read-only snapshots, out-of-range indices throw."""
    expr, mn, mx = scene("sphere")
    vox = expr.ToSdf().ToVoxels(mn, mx, 24, 24, 24)
    before = vox.ToMesh()
    with pytest.raises(ValueError):
        vox.Values[3, 3, 3] = 1.0
    with pytest.raises(IndexError):
        vox[-1, 0, 0]
    with pytest.raises(IndexError):
        vox[0, 24, 0] = 1.0
    vox[3, 4, 5] = -0.75                                   # a new inside voxel far from the sphere: 6 new crossings
    vox[(0.0, 0.0, 0.0)] = 0.5                             # Vector3 indexer: the centre voxel becomes outside
    assert vox[3, 4, 5] == f32(-0.75) and vox[12, 12, 12] == f32(0.5)
    after = vox.ToMesh()
    assert len(after.Vertices) > len(before.Vertices)
    om = oracle.marching_cubes(np.array(vox.Values), np.array(vox.Colors), f32(mn), f32(mx))
    assert np.array_equal(np.asarray(after.Triangles).reshape(-1, 3), om.triangles)
    assert np.array_equal(bits(after.Vertices), bits(om.vertices)) and np.array_equal(bits(after.Colors), bits(om.colors))
    vox.ClipToBounds()                                     # re-import keeps later device-side edits consistent
    assert vox[3, 4, 5] == f32(-0.75) and vox[0, 0, 0] == vox.Size[0] / f32(24)


def test_sdf_may_be_destroyed_before_its_distance_only_voxels(sk):
    """Distance-only voxels evaluate vertex colours from the SDF at meshing time: the library keeps the module alive."""
    import gc
    from sdfkit_b200.voxels import Voxels
    expr, mn, mx = scene("readme")
    sdf = expr.ToSdf()
    ref = sdf.ToMesh(mn, mx, 48, 48, 48)
    vox = Voxels._sample(sdf, mn, mx, 48, 48, 48, clip=True, colors=False)
    tmp = expr.ToSdf()
    vox.Resample(tmp, clip=True)
    vox._sdf = None                                        # drop the Python keep-alives: only the C-side reference remains
    tmp.Dispose()
    del tmp, sdf
    gc.collect()
    m = vox.ToMesh()
    assert np.array_equal(bits(m.Colors), bits(ref.Colors)) and np.array_equal(m.Triangles, ref.Triangles)


@pytest.mark.parametrize("dims", [(192, 160, 224), (100, 3, 301)])
def test_chunked_voxel_export(sk, oracle, dims):
    """Voxels.Values / Colors leave the device in several transposed x-chunks (64 MB staging): every chunk lands in place."""
    expr, mn, mx = scene("readme")
    nx, ny, nz = dims
    sdf = expr.ToSdf()
    vox = sdf.ToVoxels(mn, mx, nx, ny, nz)
    ov, oc = oracle.to_voxels(sdf.lowered, f32(mn), f32(mx), nx, ny, nz, clip_to_bounds=True, threads=8)
    assert np.array_equal(bits(vox.Colors), bits(oc)) and np.array_equal(bits(vox.Values), bits(ov))


# ------------------------------------------------------------------------------------------ the N-rank NCCL path

def test_two_rank_torchrun_job(sk):
    """bench.py under torch.distributed.run with 2 ranks (one process per GPU, NCCL count all-gather): the rank-ordered
    concatenation of the shares must be digest-equal to the single-GPU mesh, the slab boundaries must match the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "1", "--grid", "256",
           "--no-cpu", "--no-configs", "--no-strong"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    pc = line["parity_check"]
    assert pc["ok"] is True, pc
    assert pc["mesh_vs_single_gpu"]["equal"] is True and pc["slab_boundaries_vs_oracle"]["equal"] is True
    assert pc["totals_vs_allgather"]["equal"] is True
