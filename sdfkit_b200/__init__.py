"""sdfkit_b200 -- B200-native drop-in for SdfKit's data-parallel hot path (SdfExpr.ToSdf() -> Voxels ->
MarchingCubes, and RayMarcher/ToImage).  Host-side mirror of the reference's public API over libsdfk.so
(C ABI in include/sdfk.h); the CUDA library is required -- there is no CPU fallback."""
from ._native import Context, NotSupportedError, SdfkError  # noqa: F401
from .exprs import (  # noqa: F401
    MathF, SdfExpr, SdfExprs, SdfIndexedInput, SdfMath, Vector3, Vector4, VectorOps,
)
from .raymarcher import FloatData, RayMarcher, Vec3Data  # noqa: F401
from .sdf import GpuSdf, SdfConfig  # noqa: F401
from .voxels import GpuMesh, MarchingCubes, Mesh, Voxels  # noqa: F401
