"""Independent numpy float32 restatement of the reference's SDF formulas -- TEST INFRASTRUCTURE.

Written directly from SdfKit/SdfExpr.cs:16-201, SdfKit/Sdf.cs:118-341 and
SdfKit/VectorData.cs:697-698,860-861 -- NOT from sdfkit_b200/exprs.py -- so that the expression
lowering (tracer + emitter + prelude) is checked against a second, independently written
evaluation.  Every function maps points float32[n,3] -> float32[n,4] = (r,g,b,d); all arithmetic
is elementwise IEEE binary32 (numpy never contracts a*b+c).
"""
import numpy as np

f32 = np.float32


def _len3(x, y, z):
    return np.sqrt((x * x + y * y) + z * z)        # Vector3.Length: sqrt(Dot), (xx+yy)+zz


def _out(n, color, d):
    o = np.empty((n, 4), dtype=np.float32)
    o[:, 0], o[:, 1], o[:, 2] = f32(color[0]), f32(color[1]), f32(color[2])
    o[:, 3] = d
    return o


def _vecmax(a, b):
    return np.where(a > b, a, b)                   # Vector3.Max component


def _vecmin(a, b):
    return np.where(a < b, a, b)


def _mathmax(a, b):
    # Math.Max(float, float): NaN propagates; +0 beats -0 (irrelevant for the callers' outputs)
    r = np.where(b < a, a, b)
    r = np.where(np.isnan(a), a, r)
    both_zero = (a == 0) & (b == 0)
    return np.where(both_zero, np.where(np.signbit(b), a, b), r).astype(np.float32)


def mod(a, b):
    """VectorOps.Mod: a - b*floor(a/b)."""
    b = f32(b)
    return a - b * np.floor(a / b)


def sphere(r, color=(1, 1, 1)):
    r = f32(r)
    return lambda p: _out(len(p), color, _len3(p[:, 0], p[:, 1], p[:, 2]) - r)


def box(bounds, color=(1, 1, 1)):
    b = np.broadcast_to(np.asarray(bounds, dtype=np.float32), (3,))

    def fn(p):
        q = np.abs(p) - b
        qp = _vecmax(q, f32(0))
        qn = _vecmin(q, f32(0))
        d = _len3(qp[:, 0], qp[:, 1], qp[:, 2]) + _mathmax(_mathmax(qn[:, 0], qn[:, 1]), qn[:, 2])
        return _out(len(p), color, d)
    return fn


def cylinder(r, h, color=(1, 1, 1)):
    r, h = f32(r), f32(h)

    def fn(p):
        d = _mathmax(np.sqrt(p[:, 0] * p[:, 0] + p[:, 2] * p[:, 2]) - r, np.abs(p[:, 1]) - h)
        return _out(len(p), color, d)
    return fn


def plane(normal, dist):
    """Sdfs.Plane (Sdf.cs:203-214): Vector3.Dot(p, normal) + distanceFromOrigin; colour untouched (zeros)."""
    nrm = np.asarray(normal, dtype=np.float32)
    return lambda p: _out(len(p), (0, 0, 0), ((p[:, 0] * nrm[0] + p[:, 1] * nrm[1]) + p[:, 2] * nrm[2]) + f32(dist))


def union(a, b):
    def fn(p):
        da, db = a(p), b(p)
        return np.where((da[:, 3] < db[:, 3])[:, None], da, db)     # strict <, ties pick b
    return fn


def subtract(a, b):
    """EXTENSION mirrored from SdfExprs.Subtract (not in the reference)."""
    def fn(p):
        da, db = a(p), b(p)
        nb = -db[:, 3]
        alt = db.copy()
        alt[:, 3] = nb
        return np.where((da[:, 3] > nb)[:, None], da, alt)
    return fn


def translate(sdf, offset):
    """SdfFuncEx.Translate (Sdf.cs:315-326): sdf(p - offset)."""
    off = np.asarray(offset, dtype=np.float32)
    return lambda p: sdf((p - off).astype(np.float32))


def with_color(sdf, color):
    def fn(p):
        o = sdf(p).copy()
        o[:, 0], o[:, 1], o[:, 2] = f32(color[0]), f32(color[1]), f32(color[2])
        return o
    return fn


def _rep(c, s):
    s = f32(s)
    h = s * f32(0.5)
    return mod(c + h, s) - h, np.floor((c + h) / s)


def repeat(sdf, sx=None, sy=None, sz=None, color_mod=None):
    """RepeatX / RepeatY / RepeatXY / RepeatXZ (SdfExpr.cs:149-201); color_mod(index[n,3], mp, d) -> rgb[n,3]."""
    def fn(p):
        mp = p.copy()
        idx = np.zeros_like(p)
        for axis, s in ((0, sx), (1, sy), (2, sz)):
            if s is not None:
                mp[:, axis], idx[:, axis] = _rep(p[:, axis], s)
        d = sdf(mp)
        if color_mod is not None:
            d = d.copy()
            d[:, :3] = color_mod(idx, mp, d)
        return d
    return fn


def readme_color(i, p, d):
    """(i, p, d) => 0.9f*Vector3.One - Vector3.Abs(i)/6f   (README.md:29, Tests/RayMarcherTests.cs:102)"""
    return (f32(0.9) * f32(1.0)) - np.abs(i) / f32(6.0)


def readme_scene():
    """SdfExprs.Sphere(0.5f).RepeatXY(1.125f, 1.125f, colour lambda) -- BASELINE configs 2/4/5."""
    return repeat(sphere(0.5), sx=f32(2.25) * f32(0.5), sy=f32(2.25) * f32(0.5), color_mod=readme_color)


def perf_scene():
    """Perf/Program.cs:5-22: Union(RepeatXY spheres, RepeatXZ boxes)."""
    s = f32(2.25) * f32(0.5)
    boxes = repeat(box(f32(0.5) / f32(2)), sx=s, sz=s, color_mod=readme_color)
    spheres = repeat(sphere(0.5), sx=s, sy=s, color_mod=readme_color)
    return union(spheres, boxes)


def csg50(parts):
    """BASELINE config 3 from the part list of sdfkit_b200.scenes.csg50_parts() (data only)."""
    tree = None
    for kind, args, c, col in parts:
        prim = [sphere, box, cylinder][kind](*args)
        prim = with_color(translate(prim, c), col)
        tree = prim if tree is None else union(tree, prim)
    return repeat(subtract(tree, sphere(0.5)), sx=2.5, sy=2.5)
