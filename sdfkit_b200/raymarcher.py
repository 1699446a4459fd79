"""RayMarcher -- host-side mirror of SdfKit/RayMarcher.cs; the per-pixel work runs in the JIT-compiled
sphere-tracing kernel (csrc/jit_kernels.cuh: sdfk_k_render / sdfk_k_render_depth)."""
import numpy as np

from . import _native as N
from . import numerics


class _ImageData:
    """Minimal stand-in for SdfKit's FloatData / Vec3Data (VectorData.cs): Width, Height, Values, [x, y]."""

    def __init__(self, array):
        self.Array = array                      # [h, w] or [h, w, 3] float32
        self.Height, self.Width = array.shape[:2]
        self.Dimensions = 1 if array.ndim == 2 else array.shape[2]

    @property
    def Values(self):
        return self.Array.reshape(-1)

    @property
    def Length(self):
        return self.Array.size

    def __getitem__(self, xy):
        x, y = xy
        return self.Array[y, x]

    def Dispose(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class FloatData(_ImageData):
    def SaveDepthTga(self, path, near, far):
        """FloatData.SaveDepthTga (VectorData.cs:244-276): 8-bit grayscale, top-left origin."""
        v = self.Array.astype(np.float32)
        near, far = np.float32(near), np.float32(far)
        with np.errstate(all="ignore"):
            g = (np.float32(255.0) * (far - v) / (far - near))
        g = np.where(v >= far, 0, np.where(v <= near, 255, np.nan_to_num(g).astype(np.int64) & 0xFF)).astype(np.uint8)
        _write_tga(path, self.Width, self.Height, 3, 8, g.tobytes())


class Vec3Data(_ImageData):
    def SaveTga(self, path):
        """Vec3Data.SaveTga (VectorData.cs:570-619): BGR bytes, (byte)(v*255) truncation with clamping."""
        v = self.Array.astype(np.float32) * np.float32(255.0)
        b = np.where(v <= 0, 0, np.where(v >= 255, 255, np.nan_to_num(v).astype(np.int64) & 0xFF)).astype(np.uint8)
        _write_tga(path, self.Width, self.Height, 2, 24, b[:, :, ::-1].tobytes())


def _write_tga(path, w, h, image_type, bpp, payload):
    import struct
    with open(path, "wb") as f:
        f.write(struct.pack("<BBBHHBHHHHBB", 0, 0, image_type, 0, 0, 0, 0, 0, w, h, bpp, 0b00100000))
        f.write(payload)


class RayMarcher:
    DefaultNearPlaneDistance = 1.0
    DefaultFarPlaneDistance = 100.0
    DefaultVerticalFieldOfViewDegrees = 60.0
    DefaultDepthIterations = 40

    def __init__(self, width, height, sdf, batchSize=2048, maxDegreeOfParallelism=-1):
        from .sdf import require_gpu_sdf
        self.width, self.height = int(width), int(height)
        self.sdf = require_gpu_sdf(sdf)
        self.ViewTransform = numerics.create_look_at((0, 0, 5), (0, 0, 0), (0, 1, 0))   # RayMarcher.cs:22-23
        self.NearPlaneDistance = self.DefaultNearPlaneDistance
        self.FarPlaneDistance = self.DefaultFarPlaneDistance
        self.VerticalFieldOfViewDegrees = self.DefaultVerticalFieldOfViewDegrees
        self.DepthIterations = self.DefaultDepthIterations

    def camera(self):
        """(camera position, inverse(view*projection)) -- RayMarcher.GetCameraRays' matrices (RayMarcher.cs:95-108)."""
        cam, ivp = numerics.camera_matrices(self.ViewTransform, self.width, self.height, self.VerticalFieldOfViewDegrees,
                                            self.NearPlaneDistance, self.FarPlaneDistance)
        return N.f32c(cam), N.f32c(ivp)

    def Render(self, row_begin=0, row_end=None):
        """RayMarcher.Render (RayMarcher.cs:45-64) -> Vec3Data; optional row band for sharded renders."""
        row_end = self.height if row_end is None else row_end
        cam, ivp = self.camera()
        out = N.PinnedPool.empty((row_end - row_begin, self.width, 3), np.float32)     # page-locked, recycled
        N.check(N.lib().sdfk_render(self.sdf.ctx.handle, self.sdf.handle, self.width, self.height, N.fptr(cam), N.fptr(ivp),
                                    float(self.NearPlaneDistance), float(self.FarPlaneDistance), int(self.DepthIterations),
                                    int(row_begin), int(row_end), N.fptr(out)))
        return Vec3Data(out)

    def RenderTga(self, path):
        """Render() + Vec3Data.SaveTga(path) (RayMarcher.cs:45-64, VectorData.cs:570-619) with the float -> BGR byte conversion
        done on the device (sdfk_render_bgr8): the same file, a quarter of the device -> host traffic.  An addition to the
        reference's surface; `Render().SaveTga(path)` writes the identical bytes."""
        import ctypes as C
        cam, ivp = self.camera()
        out = N.PinnedPool.empty((self.height, self.width, 3), np.uint8)
        N.check(N.lib().sdfk_render_bgr8(self.sdf.ctx.handle, self.sdf.handle, self.width, self.height, N.fptr(cam), N.fptr(ivp),
                                         float(self.NearPlaneDistance), float(self.FarPlaneDistance), int(self.DepthIterations),
                                         0, self.height, out.ctypes.data_as(C.POINTER(C.c_ubyte))))
        _write_tga(path, self.width, self.height, 2, 24, out.tobytes())

    def RenderDepth(self, row_begin=0, row_end=None):
        """RayMarcher.RenderDepth (RayMarcher.cs:69-93) -> FloatData."""
        row_end = self.height if row_end is None else row_end
        cam, ivp = self.camera()
        out = N.PinnedPool.empty((row_end - row_begin, self.width), np.float32)
        N.check(N.lib().sdfk_render_depth(self.sdf.ctx.handle, self.sdf.handle, self.width, self.height, N.fptr(cam),
                                          N.fptr(ivp), float(self.NearPlaneDistance), int(self.DepthIterations),
                                          int(row_begin), int(row_end), N.fptr(out)))
        return FloatData(out)

    def RenderDepthGray(self, near, far, row_begin=0, row_end=None):
        """RenderDepth() followed by FloatData.SaveDepthTga's pixel conversion (VectorData.cs:262-273) on the device
        (sdfk_render_depth_gray8): the TGA payload, one byte per pixel -- a quarter of the device -> host traffic."""
        import ctypes as C
        row_end = self.height if row_end is None else row_end
        cam, ivp = self.camera()
        out = N.PinnedPool.empty((row_end - row_begin, self.width), np.uint8)
        N.check(N.lib().sdfk_render_depth_gray8(self.sdf.ctx.handle, self.sdf.handle, self.width, self.height, N.fptr(cam), N.fptr(ivp),
                                                float(self.NearPlaneDistance), int(self.DepthIterations), float(near), float(far),
                                                int(row_begin), int(row_end), out.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return out

    def RenderDepthTga(self, path, near, far):
        """RenderDepth().SaveDepthTga(path, near, far) with the byte conversion on the device: the identical file."""
        _write_tga(path, self.width, self.height, 3, 8, self.RenderDepthGray(near, far).tobytes())
