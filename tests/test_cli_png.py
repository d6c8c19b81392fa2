"""CPU tier: the command-line tool's PNG codec (csrc/cli/png_io.h) against cv2.imread /
cv2.imwrite semantics (CV_LOAD_IMAGE_COLOR: 8-bit BGR, 16-bit keeps the high byte, alpha dropped),
and the tool's flag handling (TRSP:44-70, 717-721)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

cv2 = pytest.importorskip("cv2")
BIN = os.path.join(ROOT, "surround360_b200", "bin", "TestRenderStereoPanorama")


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    d = tmp_path_factory.mktemp("png")
    exe = str(d / "png_roundtrip")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "png_roundtrip.cpp"), "-lz"])
    return exe, d


def read_with_tool(exe, d, png):
    raw = str(d / "out.raw")
    subprocess.check_call([exe, "read", png, raw])
    b = open(raw, "rb").read()
    w, h = np.frombuffer(b[:8], np.int32)
    return np.frombuffer(b[8:], np.uint8).reshape(h, w, 3)


@pytest.mark.parametrize("kind", ["rgb8", "rgb16", "gray8", "gray16", "rgba8"])
def test_png_read_matches_cv2_imread_color(tool, kind):
    exe, d = tool
    rng = np.random.default_rng(3)
    shape = {"rgb8": (37, 53, 3), "rgb16": (37, 53, 3), "gray8": (37, 53), "gray16": (37, 53), "rgba8": (37, 53, 4)}[kind]
    img = rng.integers(0, 65536 if "16" in kind else 256, shape).astype(np.uint16 if "16" in kind else np.uint8)
    path = str(d / f"{kind}.png")
    assert cv2.imwrite(path, img)
    ref = cv2.imread(path, cv2.IMREAD_COLOR)
    assert np.array_equal(read_with_tool(exe, d, path), ref)


@pytest.mark.parametrize("channels", [3, 4])
def test_png_write_is_read_back_by_cv2(tool, channels):
    exe, d = tool
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (29, 41, channels), dtype=np.uint8)
    raw, png = str(d / "in.raw"), str(d / f"w{channels}.png")
    with open(raw, "wb") as f:
        f.write(np.array([41, 29, channels], np.int32).tobytes())
        f.write(img.tobytes())
    subprocess.check_call([exe, "write", raw, png])
    back = cv2.imread(png, cv2.IMREAD_UNCHANGED)
    assert np.array_equal(back, img)


def run_cli(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True)


@pytest.mark.skipif(not os.path.exists(BIN), reason="CLI not built")
def test_cli_flag_surface():
    r = run_cli("--rig_json_file", "x")
    assert r.returncode != 0 and "missing required command line argument: imgs_dir" in r.stderr
    r = run_cli("--no_such_flag=1")
    assert r.returncode != 0 and "unknown command line flag" in r.stderr
    # glog flags the driver script passes are accepted (scripts/batch_process_video.py:31-34)
    r = run_cli("--logbuflevel", "-1", "--stderrthreshold", "0", "--v", "1", "--log_dir", "/tmp")
    assert "missing required command line argument: rig_json_file" in r.stderr


def test_png_bgra_roundtrip_matches_cv2_unchanged(tool):
    """overlap_<i>_{L,R}.png / extended*Spherical_*.png hand-over (TRSP:203-208, 413-415): written as 4-channel PNG,
    read back with cv::imread(..., -1)."""
    exe, d = tool
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (29, 41, 4)).astype(np.uint8)
    raw = str(d / "in4.raw")
    with open(raw, "wb") as f:
        f.write(np.array([41, 29, 4], np.int32).tobytes())
        f.write(img.tobytes())
    png = str(d / "bgra.png")
    subprocess.check_call([exe, "write", raw, png])
    assert np.array_equal(cv2.imread(png, cv2.IMREAD_UNCHANGED), img)       # our writer, OpenCV's reader
    out = str(d / "out4.raw")
    subprocess.check_call([exe, "read4", png, out])
    b = open(out, "rb").read()
    w, h = np.frombuffer(b[:8], np.int32)
    assert np.array_equal(np.frombuffer(b[8:], np.uint8).reshape(h, w, 4), img)
    png2 = str(d / "cv_bgra.png")
    assert cv2.imwrite(png2, img)                                            # OpenCV's writer, our reader
    subprocess.check_call([exe, "read4", png2, out])
    b = open(out, "rb").read()
    assert np.array_equal(np.frombuffer(b[8:], np.uint8).reshape(29, 41, 4), img)
