"""Committed golden vectors (tests/golden/).

  reference_known_answers.json  what the reference's own NUnit tests assert for this path + the counts an independent
                                restatement produced during the survey;
  oracle_fixtures.json          digests of the oracle's outputs on seeded cases (make_oracle_fixtures.py).

CPU tests: the oracle reproduces both files.  GPU tests: the CUDA path reproduces the fixtures without needing the oracle."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_oracle_fixtures as G  # noqa: E402

FIX = json.load(open(os.path.join(HERE, "golden", "oracle_fixtures.json")))
REF = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))


def digest(a, dtype):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dtype).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------------ CPU: oracle vs goldens

def test_fixture_file_is_what_the_generator_produces(oracle):
    """The committed fixtures are exactly what the committed generator produces from the current oracle."""
    assert G.generate() == FIX


def test_oracle_reproduces_survey_probe_counts(oracle):
    """Counts from the survey's independent restatement, incl. the white-noise field that
    exercises every ambiguous Lewiner branch."""
    P = REF["survey_probe"]
    for key, fx in (("sphere_0.5_64", "sphere64"), ("readme_64", "readme64"), ("readme_128", "readme128")):
        assert FIX["meshes"][fx]["vertices"] == P[key]["vertices"] and FIX["meshes"][fx]["triangles"] == P[key]["triangles"], key
    assert FIX["noise"]["noise24"]["triangles"] == P["white_noise_24"]["triangles"]
    # first vertex of config 1 in index space (30, 28, 15.959793): undo the T.S.T mesh transform of MarchingCubes.cs:85-90
    v0 = np.float32(FIX["meshes"]["sphere64"]["first_vertex"])
    idx = (v0 - np.float32(0)) / np.float32(2.0 / 63) + np.float32(31.5)
    assert np.allclose(idx, P["sphere_0.5_64"]["first_vertex_index_space"], atol=2e-4)


def test_oracle_readme_256_probe_counts(oracle):
    from sdfkit_b200 import scenes
    expr, mn, mx = scenes.readme_scene()
    v, c = oracle.to_voxels(expr.Lower(), np.float32(mn), np.float32(mx), 256, 256, 256, threads=os.cpu_count() or 4)
    m = oracle.marching_cubes(v, c, np.float32(mn), np.float32(mx))
    P = REF["survey_probe"]["readme_256"]
    assert (len(m.vertices), m.triangles.reshape(-1, 3).shape[0]) == (P["vertices"], P["triangles"])


# ------------------------------------------------------------------------------------------------ GPU: CUDA path vs goldens

def check_mesh(mesh, rec, what):
    assert len(mesh.Vertices) == rec["vertices"] and len(mesh.Triangles) == 3 * rec["triangles"], what
    assert digest(mesh.Triangles, np.int32) == rec["sha256_triangles"], what + ": triangle indices / order"
    assert digest(mesh.Vertices, np.float32) == rec["sha256_vertices"], what + ": vertex positions"
    assert digest(mesh.Colors, np.float32) == rec["sha256_colors"], what + ": vertex colours"
    assert digest(mesh.Normals, np.float32) == rec["sha256_normals"], what + ": normals"


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FIX["meshes"]))
def test_gpu_meshes_match_fixtures(name):
    rec = FIX["meshes"][name]
    expr, mn, mx = G.scene(rec["scene"])
    sdf = expr.ToSdf()
    nx, ny, nz = rec["dims"]
    vox = sdf.ToVoxels(mn, mx, nx, ny, nz, clipToBounds=rec["clip"])
    assert digest(vox.Values, np.float32) == rec["sha256_values"], name + ": distances"
    assert digest(vox.Colors, np.float32) == rec["sha256_voxel_colors"], name + ": voxel colours"
    check_mesh(vox.ToMesh(rec["iso"], rec["step"]), rec, name + " Voxels.ToMesh")
    check_mesh(sdf.ToMesh(mn, mx, nx, ny, nz, clipToBounds=rec["clip"], isoValue=rec["iso"], step=rec["step"]), rec, name + " Sdf.ToMesh")
    check_mesh(sdf.ToMesh(mn, mx, nx, ny, nz, clipToBounds=rec["clip"], isoValue=rec["iso"], step=rec["step"], slabs=3), rec, name + " Sdf.ToMesh 3 slabs")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FIX["noise"]))
def test_gpu_white_noise_matches_fixtures(name):
    import sdfkit_b200 as sk
    rec = FIX["noise"][name]
    v, c = G.noise_field(rec["n"], rec["seed"])
    vox = sk.Voxels(v, c, (-1, -1, -1), (1, 1, 1))
    check_mesh(vox.ToMesh(), rec, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FIX["images"]))
def test_gpu_images_match_fixtures(name):
    from sdfkit_b200 import scenes
    rec = FIX["images"][name]
    expr, _, _ = G.scene(rec["scene"])
    img = expr.ToSdf().ToImage(rec["w"], rec["h"], *scenes.CAMERA)
    assert digest(img.Array, np.float32) == rec["sha256_rgb"], name


def test_dialect_text_of_the_benchmark_scenes_is_pinned():
    """tests/golden/dialect_bodies.json: the scalar dialect body of the four BASELINE scenes.  The C# SdfExprLowering
    (csharp/SdfKit.B200, never compiled here) has to emit the same operations in the same order for the same trees; this test
    holds the Python lowering to the committed text, so a change of the dialect is a visible diff of the fixture."""
    import json
    import os
    from sdfkit_b200 import scenes
    from sdfkit_b200.exprs import lower
    fix = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dialect_bodies.json")))
    for name, fn in (("sphere", scenes.sphere), ("readme", scenes.readme_scene), ("perf", scenes.perf_scene), ("csg50", scenes.csg50)):
        low = lower(fn()[0])
        assert low.body == fix[name]["body"], name
        assert low.flops == fix[name]["flops"] and low.node_count == fix[name]["node_count"]
    assert fix["readme"]["flops"] == 23 and fix["csg50"]["flops"] == 205 and fix["csg50"]["node_count"] == 50


def test_case13_never_takes_an_impossible_subconfiguration(oracle):
    """MarchingCubes.cs:364-366: a case-13 cell whose six face tests index one of the 18 `-1` entries of subconfig13 emits
    nothing in the reference; the GPU formulation (creator = first sharing cell in visiting order) reports SDFK_ERR_INTERNAL
    if a neighbour needed a vertex from such a cell.  The six tests are functions of the same eight corner
    values, and Lewiner's table marks exactly the combinations that no eight values can produce: on white noise -- thousands
    of case-13 cells with every reachable face-test pattern -- the branch is never taken."""
    rng = np.random.default_rng(13)
    n13 = nimp = 0
    for _ in range(2):
        v = rng.standard_normal((64, 64, 64)).astype(np.float32)
        m = oracle.marching_cubes(v, np.zeros((64, 64, 64, 3), np.float32), transform=False, debug=True)
        is13 = (m.cell_index == 165) | (m.cell_index == 90)          # the two checkerboard cube indices = Lewiner case 13
        n13 += int(is13.sum())
        nimp += int((is13 & (m.cell_ntris == 0)).sum())
    assert n13 > 3000 and nimp == 0
