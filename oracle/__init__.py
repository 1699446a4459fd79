"""CPU oracle for the SdfKit hot path -- TEST INFRASTRUCTURE (see oracle/sdfk_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (sdfkit_b200/) never does.
"""
from .oracle import (  # noqa: F401
    OracleMesh, build, clip, compile_sdf, eval_sdf, lib, marching_cubes, numpy_sdf, render, render_depth, sample,
    to_mesh, to_voxels,
)
