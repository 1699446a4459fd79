"""Host-side float32 restatement of the System.Numerics calls on the reference's hot path.

The reference (C#) computes its camera and mesh matrices with the .NET BCL, which is not part of
/root/reference.  In the drop-in the C# shim computes these very matrices
with System.Numerics itself and passes them over the C ABI as 16 floats; this module is the
Python host's equivalent, so the native code never re-derives them.

Conventions restated (all arithmetic in IEEE binary32, one rounding per operation):
  * row-vector convention, v' = v . M; matrices are 4x4 row-major (M11 M12 M13 M14 / M21 ...)
  * Matrix4x4.CreateLookAt: right-handed            (call sites Sdf.cs:95, RayMarcher.cs:23)
  * Matrix4x4.CreatePerspectiveFieldOfView: M33 = far/(near-far), M34 = -1, M43 = near*far/(near-far)
                                                    (RayMarcher.cs:101-105)
  * Matrix4x4.Invert: cofactor expansion            (RayMarcher.cs:97,108; Mesh.cs:54)
  * Matrix4x4 multiply, CreateTranslation, CreateScale, Transpose (MarchingCubes.cs:85-90, Mesh.cs:49-55)

Parity note: .NET's hardware-accelerated Invert / multiply may associate sums differently from
the scalar formulas used here (<= 1 ulp effects, far inside the 1e-5 bar); parity unpinned by the
reference beyond the RayMarcher depth goldens and mesh centre/size asserts, which this reproduces.
"""
import math

import numpy as np

f32 = np.float32


def vec3(x, y=None, z=None):
    if y is None:
        a = np.asarray(x, dtype=np.float32).reshape(-1)
        if a.size == 1:
            return np.array([a[0], a[0], a[0]], dtype=np.float32)
        assert a.size == 3
        return a.copy()
    return np.array([x, y, z], dtype=np.float32)


def dot3(a, b):
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def length3(a):
    return f32(np.sqrt(dot3(a, a)))


def normalize3(a):
    """Vector3.Normalize: value / value.Length()."""
    ln = length3(a)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (a / ln).astype(np.float32)


def cross3(a, b):
    return np.array([
        f32(f32(a[1] * b[2]) - f32(a[2] * b[1])),
        f32(f32(a[2] * b[0]) - f32(a[0] * b[2])),
        f32(f32(a[0] * b[1]) - f32(a[1] * b[0])),
    ], dtype=np.float32)


def identity():
    return np.eye(4, dtype=np.float32)


def create_translation(x, y, z):
    m = identity()
    m[3, 0], m[3, 1], m[3, 2] = f32(x), f32(y), f32(z)
    return m


def create_scale(x, y, z):
    m = identity()
    m[0, 0], m[1, 1], m[2, 2] = f32(x), f32(y), f32(z)
    return m


_LOOK_AT_MEMO = {}


def create_look_at(camera_position, camera_target, camera_up):
    """Matrix4x4.CreateLookAt (right-handed).  Memoised on the float32 bits of its nine inputs: the numpy float32 emulation
    costs ~50 us, a tenth of a 1080p Sdf.ToImage call."""
    pos, tgt, up = vec3(camera_position), vec3(camera_target), vec3(camera_up)
    key = pos.tobytes() + tgt.tobytes() + up.tobytes()
    hit = _LOOK_AT_MEMO.get(key)
    if hit is not None:
        return hit.copy()
    m = _create_look_at(pos, tgt, up)
    if len(_LOOK_AT_MEMO) >= 64:
        _LOOK_AT_MEMO.clear()
    _LOOK_AT_MEMO[key] = m.copy()
    return m


def _create_look_at(pos, tgt, up):
    zaxis = normalize3((pos - tgt).astype(np.float32))
    xaxis = normalize3(cross3(up, zaxis))
    yaxis = cross3(zaxis, xaxis)
    m = identity()
    m[0, 0], m[0, 1], m[0, 2] = xaxis[0], yaxis[0], zaxis[0]
    m[1, 0], m[1, 1], m[1, 2] = xaxis[1], yaxis[1], zaxis[1]
    m[2, 0], m[2, 1], m[2, 2] = xaxis[2], yaxis[2], zaxis[2]
    m[3, 0] = -dot3(xaxis, pos)
    m[3, 1] = -dot3(yaxis, pos)
    m[3, 2] = -dot3(zaxis, pos)
    return m


def create_perspective_fov(fov_radians, aspect, near, far):
    """Matrix4x4.CreatePerspectiveFieldOfView."""
    fov_radians, aspect, near, far = f32(fov_radians), f32(aspect), f32(near), f32(far)
    y_scale = f32(f32(1.0) / f32(math.tan(float(f32(fov_radians * f32(0.5))))))
    x_scale = f32(y_scale / aspect)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = x_scale
    m[1, 1] = y_scale
    neg_far_range = f32(-1.0) if np.isposinf(far) else f32(far / f32(near - far))
    m[2, 2] = neg_far_range
    m[2, 3] = f32(-1.0)
    m[3, 2] = f32(near * neg_far_range)
    return m


def multiply(a, b):
    """Matrix4x4 operator*: each element (a1*b1 + a2*b2) + (a3*b3 + a4*b4)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    r = np.zeros((4, 4), dtype=np.float32)
    for i in range(4):
        for j in range(4):
            t0 = f32(f32(a[i, 0] * b[0, j]) + f32(a[i, 1] * b[1, j]))
            t1 = f32(f32(a[i, 2] * b[2, j]) + f32(a[i, 3] * b[3, j]))
            r[i, j] = f32(t0 + t1)
    return r


def transpose(m):
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T)


def invert(mat):
    """Matrix4x4.Invert, cofactor formulation; returns None when the determinant vanishes."""
    m_ = np.asarray(mat, dtype=np.float32)
    a, b, c, d = m_[0]
    e, f, g, h = m_[1]
    i, j, k, l = m_[2]
    m, n, o, p = m_[3]
    with np.errstate(all="ignore"):
        kp_lo = f32(f32(k * p) - f32(l * o))
        jp_ln = f32(f32(j * p) - f32(l * n))
        jo_kn = f32(f32(j * o) - f32(k * n))
        ip_lm = f32(f32(i * p) - f32(l * m))
        io_km = f32(f32(i * o) - f32(k * m))
        in_jm = f32(f32(i * n) - f32(j * m))

        a11 = f32(f32(f32(f * kp_lo) - f32(g * jp_ln)) + f32(h * jo_kn))
        a12 = -f32(f32(f32(e * kp_lo) - f32(g * ip_lm)) + f32(h * io_km))
        a13 = f32(f32(f32(e * jp_ln) - f32(f * ip_lm)) + f32(h * in_jm))
        a14 = -f32(f32(f32(e * jo_kn) - f32(f * io_km)) + f32(g * in_jm))

        det = f32(f32(f32(f32(a * a11) + f32(b * a12)) + f32(c * a13)) + f32(d * a14))
        if abs(float(det)) < 1.401298464324817e-45:   # MathF.Abs(det) < float.Epsilon
            return None
        inv_det = f32(f32(1.0) / det)
        r = np.zeros((4, 4), dtype=np.float32)
        r[0, 0] = f32(a11 * inv_det)
        r[1, 0] = f32(a12 * inv_det)
        r[2, 0] = f32(a13 * inv_det)
        r[3, 0] = f32(a14 * inv_det)

        r[0, 1] = f32(-f32(f32(f32(b * kp_lo) - f32(c * jp_ln)) + f32(d * jo_kn)) * inv_det)
        r[1, 1] = f32(f32(f32(f32(a * kp_lo) - f32(c * ip_lm)) + f32(d * io_km)) * inv_det)
        r[2, 1] = f32(-f32(f32(f32(a * jp_ln) - f32(b * ip_lm)) + f32(d * in_jm)) * inv_det)
        r[3, 1] = f32(f32(f32(f32(a * jo_kn) - f32(b * io_km)) + f32(c * in_jm)) * inv_det)

        gp_ho = f32(f32(g * p) - f32(h * o))
        fp_hn = f32(f32(f * p) - f32(h * n))
        fo_gn = f32(f32(f * o) - f32(g * n))
        ep_hm = f32(f32(e * p) - f32(h * m))
        eo_gm = f32(f32(e * o) - f32(g * m))
        en_fm = f32(f32(e * n) - f32(f * m))

        r[0, 2] = f32(f32(f32(f32(b * gp_ho) - f32(c * fp_hn)) + f32(d * fo_gn)) * inv_det)
        r[1, 2] = f32(-f32(f32(f32(a * gp_ho) - f32(c * ep_hm)) + f32(d * eo_gm)) * inv_det)
        r[2, 2] = f32(f32(f32(f32(a * fp_hn) - f32(b * ep_hm)) + f32(d * en_fm)) * inv_det)
        r[3, 2] = f32(-f32(f32(f32(a * fo_gn) - f32(b * eo_gm)) + f32(c * en_fm)) * inv_det)

        gl_hk = f32(f32(g * l) - f32(h * k))
        fl_hj = f32(f32(f * l) - f32(h * j))
        fk_gj = f32(f32(f * k) - f32(g * j))
        el_hi = f32(f32(e * l) - f32(h * i))
        ek_gi = f32(f32(e * k) - f32(g * i))
        ej_fi = f32(f32(e * j) - f32(f * i))

        r[0, 3] = f32(-f32(f32(f32(b * gl_hk) - f32(c * fl_hj)) + f32(d * fk_gj)) * inv_det)
        r[1, 3] = f32(f32(f32(f32(a * gl_hk) - f32(c * el_hi)) + f32(d * ek_gi)) * inv_det)
        r[2, 3] = f32(-f32(f32(f32(a * fl_hj) - f32(b * el_hi)) + f32(d * ej_fi)) * inv_det)
        r[3, 3] = f32(f32(f32(f32(a * fk_gj) - f32(b * ek_gi)) + f32(c * ej_fi)) * inv_det)
    return r


def transform_point(v, m):
    """Vector3.Transform(position, matrix)."""
    v = vec3(v)
    m = np.asarray(m, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    for j in range(3):
        out[j] = f32(f32(f32(f32(v[0] * m[0, j]) + f32(v[1] * m[1, j])) + f32(v[2] * m[2, j])) + m[3, j])
    return out


_camera_cache = {}


def camera_matrices(view, width, height, fov_degrees, near, far):
    """The matrix part of RayMarcher.GetCameraRays (RayMarcher.cs:95-108).

    Returns (camera_position[3], inverse(view * projection)[4,4]) as float32.  (Memoised like mesh_transforms.)
    """
    view = np.asarray(view, dtype=np.float32).reshape(4, 4)
    key = (view.tobytes(), int(width), int(height), float(fov_degrees), float(near), float(far))
    hit = _camera_cache.get(key)
    if hit is not None:
        return hit[0].copy(), hit[1].copy()
    cam_pos, ivp = _camera_matrices(view, width, height, fov_degrees, near, far)
    if len(_camera_cache) > 256:
        _camera_cache.clear()
    _camera_cache[key] = (cam_pos.copy(), ivp.copy())
    return cam_pos, ivp


def _camera_matrices(view, width, height, fov_degrees, near, far):
    view = np.asarray(view, dtype=np.float32).reshape(4, 4)
    cam_t = invert(view)
    if cam_t is None:
        raise ValueError("view transform is singular")
    cam_pos = transform_point(np.zeros(3, dtype=np.float32), cam_t)
    fov_rad = f32(f32(f32(fov_degrees) * f32(math.pi)) / f32(180.0))
    aspect = f32(f32(width) / f32(height))
    proj = create_perspective_fov(fov_rad, aspect, near, far)
    vp = multiply(view, proj)
    ivp = invert(vp)
    if ivp is None:
        raise ValueError("view-projection transform is singular")
    return cam_pos, ivp


_mesh_transform_cache = {}


def mesh_transforms(vmin, vmax, nx, ny, nz):
    """Index-space -> world transform of MarchingCubes.CreateMesh (MarchingCubes.cs:85-90) and the
    normal transform Mesh.Transform derives from it (Mesh.cs:49-55).  Returns (M, N) float32 4x4.
    (Memoised: the scalar float32 emulation below costs ~0.3 ms, 5 % of a 1024^3 Sdf.ToMesh call.)"""
    vmin, vmax = vec3(vmin), vec3(vmax)
    key = (vmin.tobytes(), vmax.tobytes(), int(nx), int(ny), int(nz))
    hit = _mesh_transform_cache.get(key)
    if hit is not None:
        return hit[0].copy(), hit[1].copy()
    m, n = _mesh_transforms(vmin, vmax, nx, ny, nz)
    if len(_mesh_transform_cache) > 256:
        _mesh_transform_cache.clear()
    _mesh_transform_cache[key] = (m.copy(), n.copy())
    return m, n


def _mesh_transforms(vmin, vmax, nx, ny, nz):
    vmin, vmax = vec3(vmin), vec3(vmax)
    size = (vmax - vmin).astype(np.float32)
    center = ((vmin + vmax).astype(np.float32) * f32(0.5)).astype(np.float32)
    with np.errstate(all="ignore"):
        t1 = create_translation(f32(-(nx - 1)) / f32(2.0), f32(-(ny - 1)) / f32(2.0), f32(-(nz - 1)) / f32(2.0))
        s = create_scale(size[0] / f32(nx - 1), size[1] / f32(ny - 1), size[2] / f32(nz - 1))
    t2 = create_translation(center[0], center[1], center[2])
    m = multiply(multiply(t1, s), t2)
    nt = m.copy()
    nt[3, 0] = nt[3, 1] = nt[3, 2] = f32(0.0)
    nt[3, 3] = f32(1.0)
    inv = invert(nt)
    if inv is None:
        inv = np.full((4, 4), np.nan, dtype=np.float32)   # Matrix4x4.Invert failure fills with NaN
    return m, transpose(inv)
