"""Synthetic 17-camera input frames (BASELINE.md "Synthetic inputs").

One procedural world -- a textured sphere at 500 cm plus 12 textured discs at 80-300 cm --
is rendered through the rig's own camera model (pixel -> ray inverse of
surround360_render/source/render/Camera.h:143-150,263-286) so neighbouring cameras agree
up to parallax. Pure numpy; used by bench.py, the tests and the CLI smoke run to make
inputs, never on the hot path.
"""
import json
import numpy as np


def band_limited_noise(h, w, seed, octaves=5):
    """Sum of `octaves` octaves of smoothed uint8 noise, tiled horizontally."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w), np.float32)
    amp = 1.0
    tot = 0.0
    for o in range(octaves):
        gh, gw = max(2, h >> (octaves - o + 1)), max(2, w >> (octaves - o + 1))
        g = rng.integers(0, 256, (gh, gw)).astype(np.float32)
        ys = (np.arange(h) + 0.5) * gh / h - 0.5
        xs = (np.arange(w) + 0.5) * gw / w - 0.5
        y0 = np.floor(ys).astype(int); x0 = np.floor(xs).astype(int)
        fy = (ys - y0)[:, None].astype(np.float32); fx = (xs - x0)[None, :].astype(np.float32)
        fy = fy * fy * (3 - 2 * fy); fx = fx * fx * (3 - 2 * fx)
        y0c = np.clip(y0, 0, gh - 1); y1c = np.clip(y0 + 1, 0, gh - 1)
        x0c = x0 % gw; x1c = (x0 + 1) % gw
        a = g[y0c][:, x0c] * (1 - fx) + g[y0c][:, x1c] * fx
        b = g[y1c][:, x0c] * (1 - fx) + g[y1c][:, x1c] * fx
        acc += amp * (a * (1 - fy) + b * fy)
        tot += amp
        amp *= 0.6
    return acc / tot


class SynthWorld:
    def __init__(self, seed=360, tex_seed=361, n_discs=12, frame=0, frame_seed=362):
        rng = np.random.default_rng(seed)
        self.tex = [band_limited_noise(1024, 2048, tex_seed + c) for c in range(3)]
        ang = rng.uniform(0, 2 * np.pi, n_discs)
        elev = rng.uniform(-0.9, 0.9, n_discs)
        dist = rng.uniform(80, 300, n_discs)
        self.centers = np.stack([dist * np.cos(elev) * np.cos(ang), dist * np.cos(elev) * np.sin(ang), dist * np.sin(elev)], 1)
        self.radii = rng.uniform(15, 45, n_discs) * dist / 150.0
        self.phase = rng.uniform(0, 1, (n_discs, 2))
        if frame:
            mv = np.random.default_rng(frame_seed).normal(0, 1.0, (n_discs, 3))
            self.centers = self.centers + mv * frame * 0.8

    def _tex(self, u, v):
        """u,v in [0,1) -> rgb uint8 (bilinear-free nearest lookup on a 1024x2048 texture)."""
        h, w = self.tex[0].shape
        x = (u * w).astype(np.int64) % w
        y = np.clip((v * h).astype(np.int64), 0, h - 1)
        return np.stack([t[y, x] for t in self.tex], -1)

    def shade(self, origin, dirs):
        """origin (3,), dirs (N,3) unit -> (N,3) uint8 RGB-ish colours (stored as BGR)."""
        n = dirs.shape[0]
        # far sphere radius 500 centred at the rig origin
        dirs = dirs.astype(np.float32)
        b = dirs @ origin.astype(np.float32)
        c = np.float32(origin @ origin - 500.0 ** 2)
        t = -b + np.sqrt(np.maximum(b * b - c, 0))
        p = origin[None, :].astype(np.float32) + dirs * t[:, None]
        u = (np.arctan2(p[:, 1], p[:, 0]) / (2 * np.pi)) % 1.0
        v = np.arccos(np.clip(p[:, 2] / 500.0, -1, 1)) / np.pi
        col = self._tex(u, v)
        best = np.full(n, np.inf, np.float32)
        for i, (cen, rad) in enumerate(zip(self.centers, self.radii)):
            d = cen - origin
            dn = np.linalg.norm(d)
            # bounding-sphere cone test selects candidate rays, exact disc test on the subset
            cosmin = np.sqrt(max(0.0, 1.0 - min(1.0, rad / dn) ** 2))
            cand = np.nonzero(dirs @ (d / dn).astype(dirs.dtype) >= cosmin)[0]
            if cand.size == 0:
                continue
            dc = dirs[cand].astype(np.float64)
            nrm = cen / np.linalg.norm(cen)
            den = dc @ nrm
            tt = (d @ nrm) / np.where(np.abs(den) < 1e-9, 1e-9, den)
            hit = dc * tt[:, None] - d[None, :]
            ok = (tt > 0) & (np.einsum('ij,ij->i', hit, hit) < rad * rad) & (tt < best[cand])
            if not ok.any():
                continue
            ax = np.cross(nrm, [0, 0, 1.0]); ax /= np.linalg.norm(ax) + 1e-12
            ay = np.cross(nrm, ax)
            lu = (hit[ok] @ ax) / (2 * rad) + 0.5 + self.phase[i, 0]
            lv = (hit[ok] @ ay) / (2 * rad) + 0.5
            cc = self._tex((lu * 0.25 + self.phase[i, 1]) % 1.0, np.clip(lv, 0, 0.999))
            cc = (cc * 0.6 + 90 * ((i % 3) == np.arange(3))[None, :]).clip(0, 255)
            col[cand[ok]] = cc
            best[cand[ok]] = tt[ok]
        return col.astype(np.uint8)


def _camera_rays(cam, w, h):
    """numpy restatement of pixel->ray for synthetic-input generation only."""
    fwd = np.array(cam["forward"], float); up = np.array(cam["up"], float); right = np.array(cam["right"], float)
    R = np.stack([right, up, -fwd])
    u, s, vt = np.linalg.svd(R)
    R = u @ vt
    px, py = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5)
    princ = cam.get("principal", [cam["resolution"][0] / 2, cam["resolution"][1] / 2])
    sx = (px - princ[0]) / cam["focal"][0]
    sy = (py - princ[1]) / cam["focal"][1]
    nrm = np.sqrt(sx * sx + sy * sy)
    nrm = np.where(nrm == 0, 1e-12, nrm)
    ang = nrm if cam["type"] == "FTHETA" else np.arctan(nrm)
    k = np.sin(ang) / nrm
    unit = np.stack([k * sx, k * sy, -np.cos(ang)], -1).reshape(-1, 3)
    return (unit @ R).astype(np.float32)  # R^T applied to row vectors


def render_rig_frames(rig_json_path, size=2048, frame=0, world=None):
    """-> list of (size,size,3) uint8 BGR frames in rig-JSON order."""
    rig = json.load(open(rig_json_path))
    world = world or SynthWorld(frame=frame)
    out = []
    for cam in rig["cameras"]:
        w, h = int(cam["resolution"][0]), int(cam["resolution"][1])
        assert w == size and h == size, "synthetic frames follow the rig resolution"
        dirs = _camera_rays(cam, w, h)
        img = world.shade(np.array(cam["origin"], float), dirs).reshape(h, w, 3)
        if cam["type"] == "FTHETA" and "fov" in cam:  # outside the fisheye circle -> black, like a real lens
            fwd = np.array(cam["forward"], float)
            img[(dirs @ fwd).reshape(h, w) < np.cos(cam["fov"])] = 0
        out.append(np.ascontiguousarray(img))
    return out


def pair_512(seed=1):
    """BASELINE config 1: I0 = 5-octave band-limited noise, alpha feather 16 px left/right;
    I1 = I0 warped by u = 4 + 3 sin(2 pi y/512), v = 1.5 sin(2 pi x/256) (bilinear sample)."""
    n = 512
    base = np.stack([band_limited_noise(n + 64, n + 64, seed + c, 5) for c in range(3)], -1)
    I0 = base[32:32 + n, 32:32 + n]
    yy, xx = np.meshgrid(np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32), indexing="ij")
    u = 4 + 3 * np.sin(2 * np.pi * yy / 512); v = 1.5 * np.sin(2 * np.pi * xx / 256)
    sx = xx + u + 32; sy = yy + v + 32
    x0 = np.floor(sx).astype(int); y0 = np.floor(sy).astype(int)
    fx = (sx - x0)[..., None]; fy = (sy - y0)[..., None]
    I1 = (base[y0, x0] * (1 - fx) * (1 - fy) + base[y0, x0 + 1] * fx * (1 - fy) +
          base[y0 + 1, x0] * (1 - fx) * fy + base[y0 + 1, x0 + 1] * fx * fy)
    a = np.full((n, n), 255.0, np.float32)
    ramp = (np.arange(16) + 0.5) / 16 * 255
    a[:, :16] = ramp[None, :]; a[:, -16:] = ramp[::-1][None, :]
    def pack(I):
        return np.ascontiguousarray(np.concatenate([I.clip(0, 255), a[..., None]], -1).astype(np.uint8))
    return pack(I0), pack(I1)


def bayer_frame(h, w, dtype=np.uint16, seed=0):
    """A synthetic raw Bayer frame (pattern-agnostic: every photosite gets the scene value of 'its' colour under GBRG, which is just
    a textured mosaic under any other pattern): smooth colour gradients + band-limited texture + shot-like noise, with a saturated
    patch, a near-black patch below the black level and a few hot pixels so that the ISP's clamps and its `v < 1` test are exercised."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    tex = band_limited_noise(h, w, seed + 11)
    tex = (tex - tex.min()) / max(float(tex.max() - tex.min()), 1e-6)
    r = 0.15 + 0.7 * xx / max(w - 1, 1) * (0.6 + 0.4 * tex)
    g = 0.10 + 0.8 * (0.5 + 0.5 * np.sin(xx / 17.0 + yy / 23.0)) * (0.5 + 0.5 * tex)
    b = 0.12 + 0.7 * yy / max(h - 1, 1) * (0.7 + 0.3 * tex)
    mosaic = np.where((yy % 2 == 1) & (xx % 2 == 0), r, np.where((yy % 2 == 0) & (xx % 2 == 1), b, g))
    mosaic = mosaic + rng.normal(0, 0.01, (h, w)).astype(np.float32) * np.sqrt(np.maximum(mosaic, 0.01))
    mosaic[h // 5:h // 5 + max(h // 8, 4), w // 6:w // 6 + max(w // 7, 4)] = 1.2      # saturated
    mosaic[h // 2:h // 2 + max(h // 9, 4), w // 2:w // 2 + max(w // 6, 4)] *= 0.01    # below the black level
    flat = mosaic.reshape(-1)
    flat[rng.integers(0, flat.size, max(flat.size // 4000, 3))] = 1.0                 # hot pixels
    full = 65535.0 if dtype == np.uint16 else 255.0
    return np.clip(0.02 * full + mosaic * 0.9 * full, 0, full).astype(dtype)


def pack12(samples):
    """(h, w) 12-bit sample values (w even) -> the capture software's packed bytes, three per pixel pair: b0 = even >> 4,
    b1 = (even & 15) | (odd & 15) << 4, b2 = odd >> 4 -- the layout RawConverter::convert12Frame undoes (RawConverter.cpp:34-59)."""
    s = np.asarray(samples).astype(np.uint16)
    assert s.ndim == 2 and s.shape[1] % 2 == 0 and int(s.max(initial=0)) < 4096
    e, o = s[:, 0::2], s[:, 1::2]
    out = np.empty((s.shape[0], s.shape[1] // 2, 3), np.uint8)
    out[..., 0] = e >> 4
    out[..., 1] = (e & 15) | ((o & 15) << 4)
    out[..., 2] = o >> 4
    return out.reshape(-1)


def write_footage_bin(path, frames, bits, serials, timestamp=0, file_index=0, file_count=1):
    """Writes a footage container (camera_isp/BinaryFootageFile.h:18-27, .cpp:178-201): a 32-byte header in a 4096-byte block, then
    frames[f][c] (sample values (h, w): < 256 for 8 bits, < 4096 for 12) packed, frame f of camera c at 4096 + (n_cams * f + c) *
    frame_bytes. Bytes 4..7 of every frame carry the camera's serial number, as the capture software stamps them (Unpacker.cpp:125).
    Returns the packed frames [f][c] as written."""
    n_cams = len(frames[0])
    h, w = np.asarray(frames[0][0]).shape
    hd = np.array([0xfaceb00c, timestamp, file_index, file_count, w, h, bits, n_cams], np.uint32)
    packed_all = []
    with open(path, "wb") as f:
        f.write(hd.tobytes())
        f.write(b"\0" * (4096 - 32))
        for fr in frames:
            row = []
            for c, img in enumerate(fr):
                p = pack12(img) if bits == 12 else np.asarray(img).astype(np.uint8).reshape(-1).copy()
                p[4:8] = np.frombuffer(np.uint32(serials[c]).tobytes(), np.uint8)
                f.write(p.tobytes())
                row.append(p)
            packed_all.append(row)
    return packed_all
