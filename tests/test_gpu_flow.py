"""GPU parity of PixFlow (s360_flow_compute*) against the CPU oracle: BASELINE config 1
(2-camera 512x512 pair, both presets, LEFT/RIGHT hints), ragged sizes, alpha-gated regions,
temporal regularisation, batch vs single, determinism. Expected bit-exact: the sweep is the
exact wavefront order of the raster scan and every float op is evaluated like the CPU path."""
import numpy as np
import pytest

import oracle_bind as ob
from surround360_b200 import synth

pytestmark = pytest.mark.gpu

FLOW_TOL = 0.0  # px; stated tolerance of the flow parity (bit-exact)


def check(g, o):
    d = np.abs(g - o)
    assert np.isfinite(g).all()
    assert d.max() <= FLOW_TOL, f"max |gpu-oracle| = {d.max()} on {np.count_nonzero(d)} of {d.size} values"


@pytest.fixture(scope="module")
def pair():
    return synth.pair_512()


@pytest.mark.parametrize("alg", ["pixflow_low", "pixflow_search_20"])
@pytest.mark.parametrize("hint", [3, 1])
def test_config1_pair_512(ctx, pair, alg, hint):
    i0, i1 = pair if hint == 3 else pair[::-1]
    check(ctx.computeOpticalFlow(alg, i0, i1, hint=hint), ob.pixflow(i0, i1, alg, hint=hint))


def test_recovers_known_flow(ctx, pair):
    # I1(p) = I0(p + d) with d = (4 + 3 sin(2 pi y/512), 1.5 sin(2 pi x/256)); the flow warps I1 to look
    # like I0 (OpticalFlowInterface.h:26-33): I1(p + f) = I0(p)  =>  f ~ -d in the interior
    i0, i1 = pair
    f = ctx.computeOpticalFlow("pixflow_low", i0, i1, hint=3)
    yy, xx = np.mgrid[64:448, 64:448]
    u = 4 + 3 * np.sin(2 * np.pi * yy / 512)
    v = 1.5 * np.sin(2 * np.pi * xx / 256)
    err = np.hypot(f[64:448, 64:448, 0] + u, f[64:448, 64:448, 1] + v)
    assert np.median(err) < 0.5


@pytest.mark.parametrize("w,h,hint", [(333, 201, 2), (100, 64, 4), (65, 130, 0), (257, 96, 1)])
def test_ragged_sizes_all_hints(ctx, pair, w, h, hint):
    i0, i1 = pair
    a, b = i0[:h, :w].copy(), i1[:h, :w].copy()
    for alg in ("pixflow_low", "pixflow_search_20"):
        check(ctx.computeOpticalFlow(alg, a, b, hint=hint), ob.pixflow(a, b, alg, hint=hint))


def test_alpha_gated_regions(ctx, pair):
    i0, i1 = [x.copy() for x in pair]
    i0[:, 200:260, 3] = 0        # a transparent stripe: gated out of the sweep, filled by diffusion
    i1[100:180, :, 3] = 128
    i0[:30, :, 3] = np.linspace(0, 255, 30).astype(np.uint8)[:, None]
    a, b = i0[:256, :384].copy(), i1[:256, :384].copy()
    check(ctx.computeOpticalFlow("pixflow_low", a, b, hint=3), ob.pixflow(a, b, "pixflow_low", hint=3))


def test_fully_transparent_image(ctx, pair):
    i0, i1 = [x[:128, :160].copy() for x in pair]
    i0[..., 3] = 0
    g = ctx.computeOpticalFlow("pixflow_low", i0, i1, hint=3)
    check(g, ob.pixflow(i0, i1, "pixflow_low", hint=3))
    assert np.abs(g).max() == 0


def test_temporal_regularisation(ctx, pair):
    i0, i1 = pair
    a, b = i0[:201, :333].copy(), i1[:201, :333].copy()
    prev = ob.pixflow(a, b, "pixflow_low", hint=2)
    pb = np.ascontiguousarray(np.roll(b, 3, axis=1))
    g = ctx.computeOpticalFlow("pixflow_low", a, b, prevFlow=prev, prevI0BGRA=a, prevI1BGRA=pb, hint=2)
    o = ob.pixflow(a, b, "pixflow_low", hint=2, prev=(prev, a, pb))
    check(g, o)


def test_batch_equals_single_and_is_deterministic(ctx, pair):
    i0, i1 = pair
    tiles0 = [np.ascontiguousarray(i0[y:y + 200, x:x + 264]) for y in (0, 150, 300) for x in (0, 120, 240)]
    tiles1 = [np.ascontiguousarray(i1[y:y + 200, x:x + 264]) for y in (0, 150, 300) for x in (0, 120, 240)]
    hints = [3, 1, 2, 4, 0, 3, 1, 2, 4]
    fb = ctx.computeOpticalFlowBatch("pixflow_low", tiles0, tiles1, hints)
    fb2 = ctx.computeOpticalFlowBatch("pixflow_low", tiles0, tiles1, hints)
    for k in range(len(tiles0)):
        assert np.array_equal(fb[k], fb2[k])
        assert np.array_equal(fb[k], ctx.computeOpticalFlow("pixflow_low", tiles0[k], tiles1[k], hint=hints[k]))
    check(fb[4], ob.pixflow(tiles0[4], tiles1[4], "pixflow_low", hint=0))


def test_many_bands(ctx, pair):
    # tall image: 2 x 449 working rows -> 15 bands per sweep, exercises the band hand-off
    i0, i1 = pair
    a = np.ascontiguousarray(np.concatenate([i0[:, :96], i0[::-1, :96]], 0)[:898])
    b = np.ascontiguousarray(np.concatenate([i1[:, :96], i1[::-1, :96]], 0)[:898])
    check(ctx.computeOpticalFlow("pixflow_low", a, b, hint=2), ob.pixflow(a, b, "pixflow_low", hint=2))


def test_latency_mode_is_bit_identical(ctx, pair):
    """s360_ctx_set_latency_mode: the sweep variant with the cp.async input / footprint rings (k_sweep<.., 4>) against the oracle -- row
    and column bands, many bands, ragged widths, gated regions, the search preset; and against the default kernel on the same inputs."""
    i0, i1 = pair
    cases = [(i0, i1, "pixflow_low", 3), (i0[:201, :333].copy(), i1[:201, :333].copy(), "pixflow_search_20", 2),
             (i0[:130, :65].copy(), i1[:130, :65].copy(), "pixflow_low", 0)]
    tall0 = np.ascontiguousarray(np.concatenate([i0[:, :96], i0[::-1, :96]], 0)[:898])
    tall1 = np.ascontiguousarray(np.concatenate([i1[:, :96], i1[::-1, :96]], 0)[:898])
    cases.append((tall0, tall1, "pixflow_low", 2))
    g0, g1 = [x.copy() for x in pair]
    g0[:, 200:260, 3] = 0
    g1[100:180, :, 3] = 128
    cases.append((g0[:256, :384].copy(), g1[:256, :384].copy(), "pixflow_low", 3))
    normal = [ctx.computeOpticalFlow(alg, a, b, hint=h) for a, b, alg, h in cases]
    ctx.set_latency_mode(True)
    try:
        for (a, b, alg, h), want in zip(cases, normal):
            got = ctx.computeOpticalFlow(alg, a, b, hint=h)
            assert np.array_equal(got, want)
            check(got, ob.pixflow(a, b, alg, hint=h))
    finally:
        ctx.set_latency_mode(False)
