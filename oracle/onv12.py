"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

numpy restatement of the encode hand-off (s360_frame_to_nv12). PARITY UNPINNED: the reference's encode step is the external
ffmpeg tool (`-pix_fmt yuv420p -c:v libx264`, scripts/run_all.py:74-88), whose swscale arithmetic is neither part of /root/reference nor
available in this image; what is restated here is the published 8-bit integer form of the limited-range ITU-R BT.601 / BT.709 matrices
with centre-sited 2 x 2 chroma, which the kernel implements."""
import numpy as np

COEF = {0: (66, 129, 25, -38, -74, 112, 112, -94, -18), 1: (47, 157, 16, -26, -86, 112, 112, -102, -10)}


def bgr_to_nv12(bgr, matrix=0):
    """(h, w, 3) uint8 BGR -> (y (h, w) uint8, uv (ceil(h / 2), 2 * ceil(w / 2)) uint8, Cb Cr interleaved)."""
    yr, yg, yb, ur, ug, ub, vr, vg, vb = COEF[matrix]
    img = np.asarray(bgr).astype(np.int32)
    h, w = img.shape[:2]
    B, G, R = img[..., 0], img[..., 1], img[..., 2]
    y = ((yr * R + yg * G + yb * B + 128) >> 8) + 16
    # 2 x 2 blocks; odd sizes: the last column / row is its own neighbour
    ys = np.minimum(np.arange(0, 2 * ((h + 1) // 2)), h - 1)
    xs = np.minimum(np.arange(0, 2 * ((w + 1) // 2)), w - 1)
    blk = img[ys][:, xs].reshape((h + 1) // 2, 2, (w + 1) // 2, 2, 3)
    m = (blk.sum(axis=(1, 3)) + 2) >> 2
    bm, gm, rm = m[..., 0], m[..., 1], m[..., 2]
    u = ((ur * rm + ug * gm + ub * bm + 128) >> 8) + 128
    v = ((vr * rm + vg * gm + vb * bm + 128) >> 8) + 128
    uv = np.stack([u, v], -1).reshape((h + 1) // 2, 2 * ((w + 1) // 2))
    return y.astype(np.uint8), uv.astype(np.uint8)
