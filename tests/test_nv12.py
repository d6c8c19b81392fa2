"""Encode hand-off: the numpy restatement (oracle/onv12.py; parity unpinned, see its header) against the definition's
known answers on the CPU tier, and the GPU kernel against it on the GPU tier (bit-exact)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import onv12  # noqa: E402


def test_port_known_answers():
    grey = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 3, 1).reshape(16, 16, 3)
    for m in (0, 1):
        y, uv = onv12.bgr_to_nv12(grey, m)
        assert y.min() == 16 and y.max() == 235 and (uv == 128).all()            # limited range, neutral chroma
        assert np.array_equal(y.reshape(-1), ((np.arange(256) * sum(onv12.COEF[m][:3]) + 128) >> 8) + 16)
    prim = np.zeros((2, 6, 3), np.uint8)
    prim[:, 0:2] = (255, 0, 0); prim[:, 2:4] = (0, 255, 0); prim[:, 4:6] = (0, 0, 255)   # blue, green, red (BGR order)
    y, uv = onv12.bgr_to_nv12(prim, 0)
    assert y[0].tolist() == [41, 41, 144, 144, 82, 82]                              # BT.601 limited: Y of the primaries
    assert uv[0].tolist() == [240, 110, 54, 34, 90, 240]                            # Cb Cr of blue, green, red
    y9, uv9 = onv12.bgr_to_nv12(prim, 1)
    assert y9[0].tolist() == [32, 32, 172, 172, 63, 63] and uv9[0].tolist() == [240, 118, 42, 26, 102, 240]   # BT.709
    odd = np.random.default_rng(0).integers(0, 256, (5, 7, 3), dtype=np.uint8)
    y, uv = onv12.bgr_to_nv12(odd)
    assert y.shape == (5, 7) and uv.shape == (3, 8)


@pytest.mark.gpu
def test_gpu_matches_port(ctx):
    rng = np.random.default_rng(3)
    for (h, w) in [(2, 2), (5, 7), (64, 96), (1024, 2048)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for m in (0, 1):
            y, uv = ctx.frame_to_nv12(img, m)
            wy, wuv = onv12.bgr_to_nv12(img, m)
            assert np.array_equal(y, wy) and np.array_equal(uv, wuv), (h, w, m)


@pytest.mark.gpu
def test_rendered_frame_to_nv12_on_device(ctx):
    """The hand-off as a pipeline uses it: a frame in HBM -> pitched NV12 planes in HBM, on the context's stream."""
    torch = pytest.importorskip("torch")
    h, w = 512, 1024
    img = np.random.default_rng(5).integers(0, 256, (h, w, 3), dtype=np.uint8)
    d = torch.from_numpy(img).cuda()
    y_pitch, uv_pitch = 1280, 1280   # an encoder's aligned pitches
    dy = torch.zeros((h, y_pitch), dtype=torch.uint8, device="cuda")
    duv = torch.zeros((h // 2, uv_pitch), dtype=torch.uint8, device="cuda")
    ctx.frame_to_nv12_dev(d.data_ptr(), w, h, dy.data_ptr(), duv.data_ptr(), matrix=1, y_pitch=y_pitch, uv_pitch=uv_pitch)
    ctx.synchronize()
    wy, wuv = onv12.bgr_to_nv12(img, 1)
    assert np.array_equal(dy.cpu().numpy()[:, :w], wy) and np.array_equal(duv.cpu().numpy()[:, :w], wuv)
    assert not dy.cpu().numpy()[:, w:].any()
