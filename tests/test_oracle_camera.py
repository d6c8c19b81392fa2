"""Known-answer tests carried over from Camera::unitTest()
(surround360_render/source/render/Camera.cpp:291-374), the only pinned numbers the reference
holds for this path, run against the oracle's camera model; plus rig selection logic."""
import json
import math

import numpy as np
import pytest

import oracle_bind as ob

CAM9 = {"version": 1, "type": "FTHETA", "origin": [-10.51814, 13.00734, -4.22656],
        "forward": [-0.6096207796429852, 0.7538922995778138, -0.24496715221587234],
        "up": [0.7686134846014325, 0.6376793279268061, 0.050974366338976666],
        "right": [0.19502945167097138, -0.15702371237098722, -0.9681462011153862],
        "resolution": [2448, 2048], "focal": [1240, -1240], "id": "cam9", "group": "side"}


@pytest.fixture()
def cam9():
    return ob.Rig(json.dumps({"cameras": [CAM9]}))


def test_pixel_of_forward_is_principal(cam9):
    p = cam9.params(0)
    pos, rot = p[0:3], p[3:12].reshape(3, 3)
    fwd = -rot[2]
    c = cam9.pixel(0, pos + fwd)
    assert abs(c[0] - 2448 / 2) < 1e-10 and abs(c[1] - 2048 / 2) < 1e-10
    assert np.linalg.norm(rot[0] - np.array(CAM9["right"])) <= 1e-3   # Eigen isApprox(right, 1e-3): the camera orthogonalises the rotation
    assert np.allclose(rot @ rot.T, np.eye(3), atol=1e-12)


def test_fov_roundtrip_and_sees(cam9):
    L = ob.lib()
    def rni(px):
        o, d = cam9.rig_ray(0, px)
        return (o + 1e6 * d)
    assert L.o_camera_sees(cam9.h, 0, ob.P(rni([1.0, 1.0])))                 # default fov sees everything in frame
    L.o_camera_set_fov(cam9.h, 0, ob.C.c_double(0.9 * math.pi))
    assert abs(L.o_camera_get_fov(cam9.h, 0) - 0.9 * math.pi) < 1e-10
    L.o_camera_set_fov(cam9.h, 0, ob.C.c_double(0.1 * math.pi))
    assert abs(L.o_camera_get_fov(cam9.h, 0) - 0.1 * math.pi) < 1e-10
    assert not L.o_camera_sees(cam9.h, 0, ob.P(rni([1.0, 1.0])))
    assert L.o_camera_sees(cam9.h, 0, ob.P(rni([1200.0, 1000.0])))
    L.o_camera_set_default_fov(cam9.h, 0)
    assert L.o_camera_sees(cam9.h, 0, ob.P(rni([1.0, 1.0])))


def test_rig_undoes_pixel(cam9):
    pos = cam9.params(0)[0:3]
    d = 3.1
    v = np.array([-2.0, 3.0, -1.0]); v /= np.linalg.norm(v)
    expected = pos + d * v
    o, direction = cam9.rig_ray(0, cam9.pixel(0, expected))
    assert np.allclose(o + d * direction, expected, rtol=1e-12, atol=1e-9)


def test_undistort_undoes_distort(cam9):
    L = ob.lib()
    for d0, d1 in ((0.0, 0.0), (0.20, 0.02)):
        L.o_camera_set_distortion(cam9.h, 0, ob.C.c_double(d0), ob.C.c_double(d1))
        distorted = L.o_camera_distort(cam9.h, 0, ob.C.c_double(3.0))
        assert abs(L.o_camera_undistort(cam9.h, 0, ob.C.c_double(distorted)) - 3.0) < 1e-6


def test_rig17_selection(orig):
    assert orig.n_cameras == 17 and orig.n_side == 14
    assert orig.camera_id(orig.special(0)) == "cam0"       # top: max forward . +Z, axis through the centre
    assert orig.camera_id(orig.special(1)) == "cam15"      # primary bottom (cam16 is 9.8 cm off-axis -> excluded)
    assert orig.camera_id(orig.special(2)) == "cam16"      # secondary bottom = largest axis-to-centre distance
    assert [orig.camera_id(orig.side_index(i)) for i in range(14)] == [f"cam{i + 1}" for i in range(14)]
    assert abs(ob.lib().o_ring_radius(orig.h) - 21.8) < 1e-5
    assert abs(ob.lib().o_approximate_fov(orig.h, 0) - 0.678732) < 1e-5   # atan(1024/1269.58), SURVEY appendix A


def test_geometry_table(orig):
    # SURVEY appendix A
    table = {"3k": (665, 665, 445, 220), "4k": (907, 442, 607, 300), "6k": (1361, 1327, 911, 450), "8k": (1814, 1769, 1214, 600)}
    from surround360_b200.api import PRESETS
    for name, (cw, ch, ov, nn) in table.items():
        g = orig.geometry(ob.make_params(**PRESETS[name]))
        assert (g["camW"], g["camH"], g["overlapW"], g["numNovel"]) == (cw, ch, ov, nn)
    g = orig.geometry(ob.make_params(**PRESETS["8k"]))
    assert abs(g["shiftPx"] - 196.53) < 0.01 and abs(g["slabDisp"] - 196.87) < 0.01


def test_bad_rig_json_is_an_error():
    with pytest.raises(RuntimeError):
        ob.Rig('{"cameras": [{"version": 1, "type": "PINHOLE"}]}')
    with pytest.raises(RuntimeError):
        ob.Rig('{"cameras": []}')
