"""CPU tests of the export formats on the path's far side: Mesh.WriteObj, Vec3Data.SaveTga,
FloatData.SaveDepthTga.  These are pure host code."""
import io
import struct

import numpy as np

from sdfkit_b200.raymarcher import FloatData, Vec3Data
from sdfkit_b200.voxels import Mesh, _net_float

f32 = np.float32


def test_write_obj_format():
    # Mesh.WriteObj (Mesh.cs:66-97): "v x y z" per vertex, "vn x y z" per normal, "f i//i j//j k//k" 1-based
    m = Mesh(np.float32([[0, 0.5, -1.25], [1, 2, 3], [0.1, 1e-7, 123456.0]]), np.zeros((3, 3), f32),
             np.float32([[0, 0, 1], [0, 1, 0], [1, 0, 0]]), np.int32([0, 1, 2, 2, 1, 0]), np.zeros(3, f32), np.ones(3, f32))
    w = io.StringIO()
    m.WriteObj(w)
    lines = w.getvalue().splitlines()
    assert lines[0] == "v 0 0.5 -1.25"
    assert lines[1] == "v 1 2 3"
    assert lines[2] == "v 0.1 1E-07 123456"
    assert lines[3:6] == ["vn 0 0 1", "vn 0 1 0", "vn 1 0 0"]
    assert lines[6:] == ["f 1//1 2//2 3//3", "f 3//3 2//2 1//1"]


def test_net_float_is_shortest_roundtrip():
    for v in [0.1, 1 / 3, 2.5e-5, 1e-5, 9.999999e-6, 1234567.9, 1e15, 3.4028235e38, -0.0, 16777216.0]:
        s = _net_float(f32(v))
        assert f32(float(s.replace("E", "e"))) == f32(v), (v, s)
    assert _net_float(f32(0.1)) == "0.1" and _net_float(f32(-0.0)) == "-0"
    assert _net_float(f32(np.nan)) == "NaN" and _net_float(f32(np.inf)) == "Infinity"


def test_save_tga(tmp_path):
    # Vec3Data.SaveTga (VectorData.cs:570-619): 18-byte header, type 2, 24 bpp, descriptor 0x20, BGR, (byte)(v*255) clamped
    img = np.zeros((2, 3, 3), dtype=np.float32)
    img[0, 0] = (1.0, 0.5, 0.0)
    img[0, 1] = (2.0, -1.0, 0.999)
    img[1, 2] = (0.2, 0.4, 0.6)
    path = tmp_path / "a.tga"
    Vec3Data(img).SaveTga(str(path))
    raw = path.read_bytes()
    hdr = struct.unpack("<BBBHHBHHHHBB", raw[:18])
    assert hdr == (0, 0, 2, 0, 0, 0, 0, 0, 3, 2, 24, 0x20)
    px = np.frombuffer(raw[18:], dtype=np.uint8).reshape(2, 3, 3)
    assert px[0, 0].tolist() == [0, 127, 255]                 # B, G, R with truncation: 0.5*255 = 127.5 -> 127
    assert px[0, 1].tolist() == [254, 0, 255]                 # clamped
    assert px[1, 2].tolist() == [int(f32(0.6) * f32(255)), int(f32(0.4) * f32(255)), int(f32(0.2) * f32(255))]


def test_save_depth_tga(tmp_path):
    # FloatData.SaveDepthTga (VectorData.cs:244-276): type 3, 8 bpp; >= far -> 0, <= near -> 255, else 255*(far-v)/(far-near)
    d = np.float32([[3.0, 10.0, 6.5], [2.0, 11.0, 4.0]])
    path = tmp_path / "d.tga"
    FloatData(d).SaveDepthTga(str(path), 3, 10)
    raw = path.read_bytes()
    hdr = struct.unpack("<BBBHHBHHHHBB", raw[:18])
    assert hdr == (0, 0, 3, 0, 0, 0, 0, 0, 3, 2, 8, 0x20)
    px = np.frombuffer(raw[18:], dtype=np.uint8).reshape(2, 3)
    assert px.tolist() == [[255, 0, int(255.0 * 3.5 / 7.0)], [255, 0, int(255.0 * 6.0 / 7.0)]]
