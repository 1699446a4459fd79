"""The BASELINE.json workloads as SdfExpr trees."""
import math

import numpy as np

from .exprs import SdfExprs, Vector3

f32 = np.float32


def readme_color(i, p, d):
    # (i, p, d) => 0.9f*Vector3.One - Vector3.Abs(i)/6f        README.md:29, Tests/RayMarcherTests.cs:102
    return 0.9 * Vector3.One - Vector3.Abs(i) / 6.0


def sphere():
    """Config 1: SdfExprs.Sphere(0.5f), bounds [-1,1]^3."""
    return SdfExprs.Sphere(0.5), (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)


def readme_scene():
    """Configs 2/4/5: Sphere(0.5f).RepeatXY(1.125f, 1.125f, colour lambda); bounds = 5x5 whole periods."""
    r = f32(0.5)
    s = f32(2.25) * r
    e = SdfExprs.Sphere(r).RepeatXY(s, s, readme_color)
    return e, (-2.8125, -2.8125, -2.8125), (2.8125, 2.8125, 2.8125)


def perf_scene():
    """Perf/Program.cs:5-22: Union(RepeatXY spheres, RepeatXZ boxes)."""
    r = f32(0.5)
    s = f32(2.25) * r
    boxes = SdfExprs.Box(r / f32(2)).RepeatXZ(s, s, readme_color)
    spheres = SdfExprs.Sphere(r).RepeatXY(s, s, readme_color)
    return SdfExprs.Union(spheres, boxes), (-2.8125, -2.8125, -2.8125), (2.8125, 2.8125, 2.8125)


def csg50_parts():
    """The 12 coloured, translated primitives of config 3 as (kind, args, centre, colour) tuples (float32)."""
    parts = []
    for k in range(12):
        a = 2.0 * math.pi * k / 12.0
        c = (f32(0.6 * math.cos(a)), f32(0.6 * math.sin(a)), f32(0.0))
        col = (f32(0.5 + 0.5 * math.cos(a)), f32(0.5 + 0.5 * math.sin(a)), f32(0.25 + k / 16.0))
        kind = k % 3
        args = [(f32(0.18) + f32(0.01) * f32(k),), (f32(0.15),), (f32(0.10), f32(0.25))][kind]
        parts.append((kind, args, c, col))
    return parts


def csg50():
    """Config 3: 12 primitives (ModifyInput translate + Color) -> 11 Unions -> Subtract(Sphere 0.5) -> RepeatXY(2.5)
    = 50 builder nodes; bounds = 3x3 whole periods.  Subtract is an extension."""
    tree = None
    for kind, args, c, col in csg50_parts():
        prim = [SdfExprs.Sphere, SdfExprs.Box, SdfExprs.Cylinder][kind](*args)
        prim = prim.ModifyInput(lambda p, c=c: p - Vector3(*c)).Color(*col)
        tree = prim if tree is None else SdfExprs.Union(tree, prim)
    tree = SdfExprs.Subtract(tree, SdfExprs.Sphere(0.5)).RepeatXY(2.5, 2.5)
    return tree, (-3.75, -3.75, -3.75), (3.75, 3.75, 3.75)


CAMERA = ((-2.0, 2.0, 4.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0))     # README.md:32-36 / Perf/Program.cs:54-58
