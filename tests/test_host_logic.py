"""CPU tests of the host-side logic: the C-ABI library loads and exports every declared symbol, the SdfExpr
lowering agrees with an independent numpy restatement, NVRTC accepts every scene, the System.Numerics
restatement, and the sharding arithmetic (incl. a world_size-2 gloo run)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def test_library_exports_every_declared_symbol():
    from sdfkit_b200 import _native as N
    header = open(os.path.join(ROOT, "include", "sdfk.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    declared = set(re.findall(r"\b(sdfk_[a-z0-9_]+)\s*\(", header)) - {"sdfk_progress_fn"}
    assert declared, "no declarations parsed"
    assert declared == set(N.SIGNATURES), sorted(declared ^ set(N.SIGNATURES))
    L = N.lib()                      # binds (dlsym) every name in SIGNATURES
    assert L.sdfk_version() >= 100
    for name in declared:
        assert getattr(L, name) is not None


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sdfkit_b200 as sk
    with pytest.raises(sk.SdfkError, match="no CPU fallback"):
        sk.Context()
    with pytest.raises(sk.SdfkError):
        sk.SdfExprs.Sphere(0.5).ToSdf()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sdfkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")) and fn != "jit_embed.h":
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "oracle/" not in text and "sdfk_oracle" not in text, fn


def _scenes():
    from oracle import sdf_numpy as S
    from sdfkit_b200 import scenes
    return {
        "sphere": (scenes.sphere()[0], S.sphere(0.5)),
        "readme": (scenes.readme_scene()[0], S.readme_scene()),
        "perf": (scenes.perf_scene()[0], S.perf_scene()),
        "csg50": (scenes.csg50()[0], S.csg50(scenes.csg50_parts())),
    }


@pytest.mark.parametrize("name", ["sphere", "readme", "perf", "csg50"])
def test_lowering_matches_numpy_restatement(oracle, name):
    """The lowered dialect text (compiled by g++ here, by NVRTC on the GPU) against the numpy restatement written
    independently from SdfExpr.cs -- bit for bit."""
    expr, ref = _scenes()[name]
    rng = np.random.default_rng(1)
    pts = np.concatenate([rng.uniform(-4, 4, (20000, 3)), rng.normal(0, 0.6, (20000, 3)), np.zeros((1, 3))]).astype(np.float32)
    a = oracle.eval_sdf(expr.Lower(), pts)
    b = ref(pts)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", ["sphere", "readme", "perf", "csg50"])
def test_nvrtc_accepts_every_scene(name):
    """NVRTC compiles for sm_100a without a GPU: the JIT source (prelude + body + kernels) must build."""
    from sdfkit_b200 import _native as N
    expr, _ = _scenes()[name]
    low = expr.Lower()
    body = low.body.encode()
    n = C.c_size_t()
    N.check(N.lib().sdfk_sdf_check(body, len(body), C.byref(n)))
    assert n.value > 1000
    assert low.flops > 0
    if name == "csg50":
        assert low.node_count == 50


def test_nvrtc_reports_bad_source():
    from sdfkit_b200 import _native as N
    body = b"    return not_a_function(p);\n"
    rc = N.lib().sdfk_sdf_check(body, len(body), None)
    assert rc == -3
    assert b"not_a_function" in N.lib().sdfk_last_error()


def test_exprs_semantics():
    from sdfkit_b200.exprs import MathF, SdfExprs, Vector3, lower
    # Union: strict <, ties pick b (SdfExpr.cs:63-66)
    import oracle
    a = SdfExprs.Sphere(0.5, (1, 0, 0))
    b = SdfExprs.Sphere(0.5, (0, 1, 0))
    out = oracle.eval_sdf(lower(SdfExprs.Union(a, b)), np.float32([[0.3, 0.1, 0.2]]))
    assert tuple(out[0, :3]) == (0.0, 1.0, 0.0)
    # opaque callables cannot be lowered
    with pytest.raises(TypeError):
        lower(lambda p: p)
    # a symbolic value used in Python control flow is an error, not a silent constant
    with pytest.raises(TypeError):
        lower(SdfExprs.Solid(lambda p: 1.0 if p.X > 0 else 2.0))
    # closure constants are folded in float32
    low = lower(SdfExprs.Solid(lambda p: p.X * (f32(0.1) + f32(0.2))))
    assert float(f32(0.1) + f32(0.2)).hex() in low.body


def test_numerics_camera_and_mesh_transforms():
    from sdfkit_b200 import numerics as nm
    view = nm.create_look_at((0, 0, 5), (0, 0, 0), (0, 1, 0))
    assert np.allclose(view, np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, -5, 1]], dtype=np.float32))
    cam, ivp = nm.camera_matrices(view, 50, 30, 60.0, 1.0, 100.0)
    assert np.allclose(cam, [0, 0, 5])
    proj = nm.create_perspective_fov(np.float32(np.pi / 3), np.float32(50 / 30), 1.0, 100.0)
    assert np.allclose(nm.multiply(nm.multiply(view, proj), ivp), np.eye(4), atol=1e-4)
    assert abs(proj[2, 2] - (100.0 / (1.0 - 100.0))) < 1e-6 and proj[2, 3] == -1 and abs(proj[3, 2] - (100.0 / -99.0)) < 1e-6
    M, Nn = nm.mesh_transforms((-1, -1, -1), (1, 1, 1), 64, 64, 64)
    s = f32(2.0) / f32(63)
    assert M[0, 0] == s and M[3, 0] == f32(-63 / 2.0) * s + f32(0)
    assert np.allclose(Nn[:3, :3], np.eye(3) / s, rtol=1e-6)
    a = np.float32(np.random.default_rng(0).uniform(-1, 1, (4, 4)))
    assert np.allclose(nm.multiply(a, nm.invert(a)), np.eye(4), atol=1e-4)
    assert nm.invert(np.zeros((4, 4), np.float32)) is None


def test_sharding_arithmetic():
    from sdfkit_b200 import dist
    assert dist.cells_along(64, 1) == 63 and dist.cells_along(5, 2) == 2 and dist.cells_along(4, 2) == 1
    assert dist.cells_along(1, 1) == 0 and dist.cells_along(2, 1) == 1
    assert dist.partition(63, 4) == [(0, 16), (16, 32), (32, 48), (48, 63)]
    assert dist.partition(3, 8) == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5
    assert dist.slab_slices(0, 16, 1, 64) == (0, 18)       # bottom slab: no ghost below, one above
    assert dist.slab_slices(16, 32, 1, 64) == (15, 34)
    assert dist.slab_slices(48, 63, 1, 64) == (47, 64)     # top slab
    assert dist.slab_slices(2, 4, 2, 12) == (2, 11)
    excl, tot = dist.exclusive_offsets([[10, 20], [0, 0], [5, 7]])
    assert excl.tolist() == [[0, 0], [10, 20], [10, 20]] and tot.tolist() == [15, 27]
    # cost-balanced contiguous cuts
    assert dist.weighted_partition([1] * 8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    wp = dist.weighted_partition([1, 1, 1, 1, 10, 10, 1, 1, 1, 1], 3)
    assert wp[0][0] == 0 and wp[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(wp, wp[1:])) and all(b > a for a, b in wp)
    assert wp[1] in ((4, 5), (4, 6), (5, 6))          # the heavy layers get a thin slab
    assert dist.weighted_partition([5, 5], 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    # round-robin slabs: slab g belongs to rank g % world; offsets follow the slab order g
    class _J(dist.ShardedMesher):
        def __init__(self, rank, world, spr):
            self.rank, self.world, self.spr = rank, world, spr
            self.slab_ids = [rank + world * s for s in range(spr)]
    counts = np.arange(2 * 3 * 2).reshape(2, 3, 2)            # [world=2, spr=3, 2]; slab g=r+2s has counts[r, s]
    offs, tot = _J(1, 2, 3).offsets(counts)
    order = [counts[g % 2, g // 2] for g in range(6)]
    cum = np.concatenate([[np.zeros(2, int)], np.cumsum(order, axis=0)[:-1]])
    assert offs.tolist() == [cum[1].tolist(), cum[3].tolist(), cum[5].tolist()] and tot.tolist() == counts.reshape(-1, 2).sum(0).tolist()


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from sdfkit_b200 import dist as skd
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
counts = [(7, 11), (3, 5)]
excl, tot = skd.all_gather_counts(*counts[rank])
assert excl.tolist() == [[0, 0], [7, 11]] and tot.tolist() == [10, 16], (excl, tot)
allc = skd.all_gather_int64(np.array([[rank, 1], [10 + rank, 2]]))
assert allc.tolist() == [[0, 1, 10, 2], [1, 1, 11, 2]], allc
local = torch.arange(counts[rank][0] * 3, dtype=torch.float32).reshape(-1, 3) + 100 * rank
out = skd.gather_rows(local, [c[0] for c in counts])
if rank == 0:
    assert out.shape == (10, 3) and out[7, 0].item() == 100.0 and out[6, 2].item() == 20.0
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_count_all_gather_and_mesh_gather_gloo_world2(tmp_path):
    """The N>1 host path on CPU: world_size 2, gloo -- count all-gather -> offsets, variable-length gather to rank 0."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


@pytest.mark.parametrize("name", ["sphere", "readme", "perf", "csg50"])
def test_packed_body_compiles_offline_and_never_adds_packed_products(name):
    """The packed (f32x2) body: NVRTC accepts it for sm_100a without a GPU, the scalar body is left untouched, and no packed
    add/sub ever takes a product as operand (ptxas would contract the pair into FFMA2 even under --fmad=false)."""
    import ctypes as C
    import re
    from sdfkit_b200 import _native as N, scenes
    from sdfkit_b200.exprs import PACKED_MARKER, lower
    expr = {"sphere": scenes.sphere, "readme": scenes.readme_scene, "perf": scenes.perf_scene, "csg50": scenes.csg50}[name]()[0]
    low = lower(expr, fast_div=lambda c: True, packed=True)
    assert "sk2_" not in low.body and low.body == lower(expr).body
    products = set(re.findall(r"const sk_f2 (v\d+) = sk2_mul\(", low.body2))
    for line in low.body2.splitlines():
        m = re.match(r"\s*const sk_f2 v\d+ = sk2_(add|sub)\((\w+), (\w+)\);", line)
        if m:
            assert m.group(2) not in products and m.group(3) not in products, line
    text = (low.body + PACKED_MARKER + "\n" + low.body2).encode()
    n = C.c_size_t()
    N.check(N.lib().sdfk_sdf_check(text, len(text), C.byref(n)))
    assert n.value > 10000


@pytest.mark.parametrize("name", ["sphere", "readme", "perf", "csg50"])
def test_shared_guard_device_forms(name):
    """The default device forms (two-point evaluator, row-of-voxels evaluator): NVRTC accepts them for sm_100a without a GPU;
    the scalar body the oracle compiles is untouched by them; in the grid form only z-dependent operations get a shared guard
    (x/y-only ones stay plain instructions the compiler hoists out of the z loop); every verified constant division and every
    sqrt appears exactly as often as in the scalar body times the number of points."""
    import ctypes as C
    import re
    from sdfkit_b200 import _native as N, scenes
    from sdfkit_b200.exprs import GRID_M, GRID_MARKER, PACKED_MARKER, lower
    expr = {"sphere": scenes.sphere, "readme": scenes.readme_scene, "perf": scenes.perf_scene, "csg50": scenes.csg50}[name]()[0]
    low = lower(expr, fast_div=lambda c: True)
    assert low.body == lower(expr).body and "sk_sqrt_core" not in low.body and "sk_divc_core" not in low.body
    nsqrt, ndiv = low.op_counts.get("sqrt", 0), low.op_counts.get("div", 0)
    count = lambda text, pat: len(re.findall(pat, text))
    # pair form: every sqrt of both points is either in a group (core + IEEE redo) or plain
    st = low.guard_stats["pair"]
    assert st["sqrt_grouped"] + st["sqrt_plain"] == 2 * nsqrt and st["div_grouped"] + st["div_plain"] <= 2 * ndiv
    assert count(low.pair_body, r"sk_sqrt_core\(") == st["sqrt_grouped"]
    assert count(low.pair_body, r"sk_divc_core\(") == st["div_grouped"]
    # grid form: y/z-only work is emitted once, x-dependent work GRID_M times
    sg = low.guard_stats["grid"]
    assert "#define SDFK_GRID_M %d" % GRID_M in low.grid_text
    assert sg["sqrt_grouped"] + sg["sqrt_plain"] <= GRID_M * nsqrt
    if name == "readme":
        assert sg == {"sqrt_groups": 1, "sqrt_grouped": 4, "div_groups": 0, "div_grouped": 0, "sqrt_plain": 0, "div_plain": 0}
        assert count(low.grid_text, r" / ") == 2 * GRID_M + 2          # x divisions per voxel, y divisions once: all plain (hoistable)
    text = low.device_text().encode()
    n = C.c_size_t()
    N.check(N.lib().sdfk_sdf_check(text, len(text), C.byref(n)))
    assert n.value > 10000
    assert low.decls == "" and "ctab" not in text.decode()
    low = lower(expr, fast_div=lambda c: True, color_table=True)       # opt-in rewrite (measured slower: off by default)
    text = low.device_text().encode()
    N.check(N.lib().sdfk_sdf_check(text, len(text), C.byref(n)))
    if name == "csg50":
        # the union of coloured primitives: the colour decision tree became a table (12 primitives + the subtracted sphere) and an
        # integer carried through the same comparisons; no colour select is left in the device forms
        assert low.decls.count("{") == 1 + 13 and "sdfk_ctab[13]" in low.decls
        assert low.pair_body.count("const int i") == 2 * 12 and low.grid_text.count("const int i") == GRID_M * 12
        assert low.pair_body.count("sk_sel(") == 2 * 12          # the distance selects remain: one per union / subtract
        assert low.body.count("sk_sel(") == 4 * 12 and "ctab" not in low.body
    else:
        assert low.decls == ""
    # a host that sends only the scalar body (the C# shim) gets the library's default device forms
    N.check(N.lib().sdfk_sdf_check(low.body.encode(), len(low.body), C.byref(n)))


def test_cost_balanced_plan_weights():
    """weighted_partition: heavier layers get thinner slabs; the e2e weight (PCIe bytes per active cell) concentrates them more."""
    from sdfkit_b200 import dist
    w_compute = np.ones(100) + dist.ACTIVE_CELL_COST * np.r_[np.zeros(40), np.full(20, 0.01), np.zeros(40)]
    w_e2e = np.ones(100) + dist.ACTIVE_CELL_COST_E2E * np.r_[np.zeros(40), np.full(20, 0.01), np.zeros(40)]
    a, b = dist.weighted_partition(w_compute, 4), dist.weighted_partition(w_e2e, 4)
    assert a[0][0] == 0 and a[-1][1] == 100 and b[0][0] == 0 and b[-1][1] == 100
    thick = lambda parts: [hi - lo for lo, hi in parts]
    assert min(thick(b)) < min(thick(a)) <= 25


def test_mirror_covers_the_reference_api_surface():
    """Every public member of the reference's hot-path types (tests/golden/reference_api_surface.json, extracted from the C#
    sources by tests/golden/make_api_surface.py) exists under the same name in the Python mirror."""
    import json
    import os
    import sdfkit_b200 as sk
    surf = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_api_surface.json")))
    where = {"SdfConfig": [sk.SdfConfig], "SdfEx": [sk.GpuSdf], "SdfExprs": [sk.SdfExprs], "SdfExprEx": [sk.SdfExpr],
             "SdfIndexedInput": [sk.SdfIndexedInput], "Voxels": [sk.Voxels], "MarchingCubes": [sk.MarchingCubes], "Mesh": [sk.Mesh],
             "RayMarcher": [sk.RayMarcher]}
    instance_attrs = {"Voxels": {"Colors", "Values", "DX", "DY", "DZ", "NX", "NY", "NZ", "Min", "Max"},
                      "Mesh": {"Vertices", "Colors", "Normals", "Triangles", "Min", "Max"},
                      "SdfIndexedInput": {"Position", "Index"},
                      "RayMarcher": {"DepthIterations", "FarPlaneDistance", "NearPlaneDistance", "VerticalFieldOfViewDegrees", "ViewTransform"}}
    missing = []
    for typ, info in surf.items():
        for name in info["members"]:
            if name == "this[]":
                name = "__getitem__"
            if name in instance_attrs.get(typ, ()):
                continue                      # set per instance in __init__ (checked on real objects by the GPU tests)
            if not any(hasattr(c, name) for c in where[typ]):
                missing.append("%s.%s" % (typ, name))
    assert not missing, missing


def test_mesh_transform_and_measure():
    """Mesh.Transform / Measure (Mesh.cs:30-64) on host arrays: row-vector convention, normals by the inverse transpose."""
    import sdfkit_b200 as sk
    from sdfkit_b200 import numerics
    v = np.float32([[0, 0, 0], [1, 0, 0], [0, 2, 0]])
    n = np.float32([[0, 0, 1], [0, 0, 1], [1, 0, 0]])
    m = sk.Mesh(v.copy(), np.ones_like(v), n.copy(), np.int32([0, 1, 2]))
    assert np.array_equal(m.Min, [0, 0, 0]) and np.array_equal(m.Max, [1, 2, 0])
    M = numerics.multiply(numerics.create_scale(2, 3, 4), numerics.create_translation(10, 20, 30))
    m.Transform(M)
    assert np.allclose(m.Vertices, [[10, 20, 30], [12, 20, 30], [10, 26, 30]])
    assert np.allclose(m.Normals, [[0, 0, 1], [0, 0, 1], [1, 0, 0]])
    assert np.allclose(m.Min, [10, 20, 30]) and np.allclose(m.Max, [12, 26, 30]) and np.allclose(m.Center, [11, 23, 30])


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU restatement timed on host cores) runs without a GPU and prints exactly one JSON line
    with the keys the contract names.  The labels are honest: `config` names the grid that was really measured (a bounded
    sample), `sample_of` the workload it stands for, `same_config` says they differ, `warmup` is what was really done."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-n", "48"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "voxel_samples_per_s" and d["unit"] == "voxels/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, root)
    import bench
    assert d["config"]["grid"] == [48, 48, 48] and "48^3" in d["config"]["workload"] and "1024" not in d["config"]["workload"]
    assert d["config"]["same_config"] is False
    assert d["config"]["sample_of"]["workload"] == bench.workload_name("readme", 1024, 1) and d["config"]["sample_of"]["grid"] == [1024] * 3
    assert d["warmup"] == 0 and d["steps"] == 1 and d["cpu_baseline"]["detail"]["grid"] == [48, 48, 48]
    assert abs(d["config"]["sample_of"]["voxel_fraction"] - (48 / 1024) ** 3) < 1e-12


def test_only_the_allowed_places_touch_the_oracle():
    """The oracle is test infrastructure: outside tests/ only __graft_entry__ (build + smoke) and bench.py (cpu_baseline /
    --impl reference legs) may import it -- not the product package, not the tools."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for sub in ("sdfkit_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for fn in files:
                if fn.endswith(".py"):
                    text = open(os.path.join(dirpath, fn)).read()
                    if re.search(r"^\s*(import oracle|from oracle)", text, re.M):
                        offenders.append(os.path.join(sub, fn))
    assert not offenders, offenders
