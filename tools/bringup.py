"""GPU bring-up: runs each C-ABI operator against the oracle and prints the differences."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_bind as ob
from surround360_b200 import api, synth

def rep(name, a, b):
    a = a.astype(np.float64); b = b.astype(np.float64); d = np.abs(a - b)
    print(f"{name:44s} max={d.max():.4g} nz={np.count_nonzero(d)}/{d.size} ({100.0*np.count_nonzero(d)/d.size:.4f}%)", flush=True)

rng = np.random.default_rng(0)
ctx = api.Context(0)
RIG = os.path.join(ROOT, "tests/golden/rig17.json")
ctx.load_rig(RIG); orig = ob.Rig.load(RIG)
which = set(sys.argv[1:]) or {"prim", "proj", "flow", "novel", "frame"}

if "prim" in which:
    a = rng.integers(0, 256, (131, 203, 4), dtype=np.uint8)
    rep("resize cubic c4 half", ctx.resizeCubic(a, 101, 65), ob.resize_cubic_u8(a, 101, 65))
    a = rng.integers(0, 256, (120, 333, 3), dtype=np.uint8)
    rep("resize cubic c3", ctx.resizeCubic(a, 320, 120), ob.resize_cubic_u8(a, 320, 120))
    a = rng.integers(0, 256, (64, 210, 4), dtype=np.uint8)
    for off in (49.13, -49.13, 0.5, 209.7):
        rep(f"offset wrap {off}", ctx.offsetHorizontalWrap(a, off), ob.offset_wrap(a, off))
    b = rng.integers(0, 256, (64, 210, 4), dtype=np.uint8); b[:, :50, 3] = 0; b[:, 50:100, 3] = 255
    rep("flatten", ctx.flattenLayersDeghostPreferBase(a, b), ob.flatten_deghost(a, b))
    f = rng.integers(0, 256, (150, 300, 4), dtype=np.uint8); f[..., 3] = (rng.uniform(0, 1, (150, 300)) > 0.01) * 255
    f[:40, :, 3] = 0
    rep("feather", ctx.featherAlphaChannel(f, 31), ob.feather_alpha(f, 31))
    img = rng.integers(0, 256, (90, 140, 3), dtype=np.uint8)
    img = np.ascontiguousarray((np.repeat(np.repeat(img[::6, ::6], 6, 0), 6, 1)[:90, :140]))
    rep("sharpen", ctx.sharpen(img, 0.25), ob.sharpen(img, 0.25))
    eye = rng.integers(0, 256, (128, 256, 3), dtype=np.uint8)
    rep("cubemap video", ctx.cubemap(eye, 64, 64), ob.cubemap(eye, 64, 64))
    rep("cubemap photo", ctx.cubemap(eye, 48, 64, True), ob.cubemap(eye, 48, 64, True))

if "proj" in which:
    src = rng.integers(0, 256, (2048, 2048, 3), dtype=np.uint8)
    hr = float(orig.geometry(ob.make_params())["hRad"])
    for i in (0, 3):
        cam = ctx.side_index(i); d = np.float32(-np.float32(i) / np.float32(14) * np.float32(2.0) * np.pi)
        l, r, t, b = float(np.float32(d + np.float32(hr / 2))), float(np.float32(d - np.float32(hr / 2))), hr / 2, -hr / 2
        g = ctx.bicubicRemapToSpherical(cam, src, 453, 442, l, r, t, b, feather=100)
        o = orig.project_side(cam, src, 453, 442, l, r, t, b, 100)
        rep(f"project side {i}", g, o)

if "flow" in which:
    i0, i1 = synth.pair_512()
    for alg in ("pixflow_low", "pixflow_search_20"):
        for hint in (3, 1):
            a, b = (i0, i1) if hint == 3 else (i1, i0)
            t = time.time(); g = ctx.computeOpticalFlow(alg, a, b, hint=hint); tg = time.time() - t
            t = time.time(); o = ob.pixflow(a, b, alg, hint=hint); to = time.time() - t
            rep(f"flow {alg} hint{hint} (gpu {tg:.2f}s cpu {to:.2f}s)", g, o)
    a = i0[:201, :333].copy(); b = i1[:201, :333].copy()
    rep("flow ragged 333x201", ctx.computeOpticalFlow("pixflow_low", a, b, hint=2), ob.pixflow(a, b, "pixflow_low", hint=2))
    pf = ob.pixflow(a, b, "pixflow_low", hint=2)
    g = ctx.computeOpticalFlow("pixflow_low", a, b, prevFlow=pf, prevI0BGRA=a, prevI1BGRA=b[::-1].copy(), hint=2)
    o = ob.pixflow(a, b, "pixflow_low", hint=2, prev=(pf, a, b[::-1].copy()))
    rep("flow temporal", g, o)

if "novel" in which:
    i0, i1 = synth.pair_512()
    a = i0[:300, :400].copy(); b = i1[:300, :400].copy()
    flr = ob.pixflow(a, b, "pixflow_low", hint=3); frl = ob.pixflow(b, a, "pixflow_low", hint=1)
    gl, gr = ctx.combineLazyNovelViews(a, b, flr, frl, 200, 600, 200, 60.5)
    ol, orr = ob.novel_view_chunk(a, b, flr, frl, 200, 600, 200, 60.5)
    rep("novel chunk L", gl, ol); rep("novel chunk R", gr, orr)
    d = gl.astype(int) - ol.astype(int)
    for ch in range(4): print("  ch", ch, "hist", {int(v): int((d[..., ch] == v).sum()) for v in np.unique(d[..., ch])})

if "frame" in which:
    t = time.time(); frames = synth.render_rig_frames(RIG); print("synth frames", time.time() - t)
    kw = dict(eqr_width=2100, eqr_height=1024, final_eqr_width=2048, final_eqr_height=2048, sharpening=0.25, enable_top=1,
              enable_bottom=1, want_cubemap=1, cubemap_width=256, cubemap_height=256)
    gp = api.default_params(**kw); op = ob.make_params(keep_intermediates=1, **kw)
    ctx.keep_intermediates(True)
    t = time.time(); eq, cube = ctx.renderStereoPanorama(gp, frames); print("gpu frame (incl. plan)", time.time() - t)
    t = time.time(); eq, cube = ctx.renderStereoPanorama(gp, frames); print("gpu frame 2nd", time.time() - t, ctx.stage_ms())
    t = time.time(); r = orig.render_frame(frames, op); print("oracle frame", time.time() - t)
    for i in (0, 7):
        rep(f"projection {i}", ctx.intermediate("projection", i), r.get("projection", i))
        rep(f"flowLtoR {i}", ctx.intermediate("flowLtoR", i), r.get("flowLtoR", i))
        rep(f"flowRtoL {i}", ctx.intermediate("flowRtoL", i), r.get("flowRtoL", i))
    rep("panoL", ctx.intermediate("panoL"), r.get("panoL")); rep("panoR", ctx.intermediate("panoR"), r.get("panoR"))
    rep("topSpherical", ctx.intermediate("topSpherical"), r.get("topSpherical"))
    rep("bottomSpherical", ctx.intermediate("bottomSpherical"), r.get("bottomSpherical"))
    for k in range(4):
        rep(f"poleFlow {k}", ctx.intermediate("poleFlow", k), r.get("poleFlow", k))
        rep(f"poleWarped {k}", ctx.intermediate("poleWarped", k), r.get("poleWarped", k))
    rep("eyeL", ctx.intermediate("eyeL")[..., :3], r.get("eyeL")); rep("eyeR", ctx.intermediate("eyeR")[..., :3], r.get("eyeR"))
    ge, oe = ctx.intermediate("eyeL")[..., :3], r.get("eyeL")
    ys, xs, cs = np.nonzero(ge != oe)
    pl = ctx.intermediate("panoL"); H = ge.shape[0]; padA = (H - pl.shape[0]) // 2
    tw, bw = r.get("poleWarped", 0), r.get("poleWarped", 2)
    for y, x, c in list(zip(ys, xs, cs))[:8]:
        base = pl[y - padA, x] if 0 <= y - padA < pl.shape[0] else np.zeros(4, np.uint8)
        print("  eyeL diff at", (y, x, c), "gpu", ge[y, x], "oracle", oe[y, x], "base", base, "top", tw[y, x], "bot(flipped idx)", bw[H - 1 - y, ge.shape[1] - 1 - x])
    rep("equirect", eq, r.get("equirect")); rep("cubemap", cube, r.get("cubemap"))
    np.save(os.path.join(ROOT, "gpurun_out", "eq_small_gpu.npy"), eq[::4, ::4])
print("launches", ctx.launch_count())
