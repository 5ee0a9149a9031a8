"""GPU: k_cvt_resize (pe_kernels_fused4.cu) -- planar 4:2:0 -> RGBA32 / BGRA32 conversion and the resize of the result in ONE kernel
(BASELINE config 2 without its RGBA intermediate) -- against the oracle's unfused chain (reference converter arithmetic, then the
resize contract), bit for bit: tile-edge geometries, up / down / mixed scaling, every table variant, quirks on / off, YVU420P, single
layers and batches; the launch counter shows that the fused kernel (one launch) is what ran, and the fallback still matches."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pe_testlib as T  # noqa: E402

lb = pytest.importorskip("lives_b200")
pytestmark = pytest.mark.gpu


def _expected(o, y, u, v, w, h, dw, dh, opal, cl, sub, quirks=1):
    order = 0 if opal == 3 else 1
    rgba = np.zeros((h, T.rowstride(w, 4)), np.uint8)
    o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(rgba), rgba.strides[0], order, 1, 0, cl, sub, T.Q_HIGH,
                           quirks, None)
    exp = np.zeros((dh, T.rowstride(dw, 4)), np.uint8)
    o.pe_or_resize_packed(T.ptr(rgba), rgba.strides[0], w, h, T.ptr(exp), exp.strides[0], dw, dh, 4)
    return exp


GEOMS = [(1920, 1080, 1280, 720), (64, 48, 40, 30), (320, 240, 640, 480), (640, 360, 426, 240), (132, 70, 130, 50), (256, 130, 256, 96),
         (200, 100, 300, 66), (8, 4, 12, 6), (1280, 720, 1920, 1080), (644, 362, 500, 300)]


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("opal", [3, 4])
def test_cvt_resize_single_layer(geom, opal):
    o = T.oracle()
    eng = lb.Engine()
    w, h, dw, dh = geom
    rng = np.random.default_rng(w + dw + opal)
    y, u, v = T.make_yuv_planar(rng, w, h, False, True)
    exp = _expected(o, y, u, v, w, h, dw, dh, opal, 0, 1)
    lay = lb.Layer.from_host(eng, 512, w, h, [y, u, v], yuv_clamping=0, yuv_subspace=1)
    before = eng.launch_count
    assert lb.resize_layer(lay, dw, dh, lb.LIVES_INTERP_NORMAL, opal, 0)
    assert eng.launch_count - before == 1, "conversion + resize must leave as one kernel"
    assert (lay.palette, lay.width, lay.height) == (opal, dw, dh)
    got = lay.to_host()[0]
    assert (got[:, :dw * 4] == exp[:, :dw * 4]).all(), (geom, opal, np.argwhere(got[:, :dw * 4] != exp[:, :dw * 4])[:5])
    eng.close()


@pytest.mark.parametrize("cl,sub", [(0, 1), (1, 1), (0, 2), (1, 2)])
@pytest.mark.parametrize("quirks", [True, False])
def test_cvt_resize_table_variants_and_quirks(cl, sub, quirks):
    o = T.oracle()
    eng = lb.Engine(ref_quirks=quirks)
    rng = np.random.default_rng(7 + cl + 2 * sub)
    w, h, dw, dh = 644, 362, 430, 240
    y, u, v = T.make_yuv_planar(rng, w, h, False, cl == 0)
    exp = _expected(o, y, u, v, w, h, dw, dh, 3, cl, sub, int(quirks))
    lay = lb.Layer.from_host(eng, 512, w, h, [y, u, v], yuv_clamping=cl, yuv_subspace=sub)
    before = eng.launch_count
    assert lb.resize_layer(lay, dw, dh, lb.LIVES_INTERP_NORMAL, 3, cl)
    assert eng.launch_count - before == 1
    assert (lay.to_host()[0][:, :dw * 4] == exp[:, :dw * 4]).all()
    eng.close()


def test_cvt_resize_batch_config2_and_fallbacks():
    o = T.oracle()
    eng = lb.Engine()
    rng = np.random.default_rng(22)
    w, h, dw, dh = 1920, 1080, 1280, 720
    frames = [T.make_yuv_planar(rng, w, h, False, True) for _ in range(5)]
    exps = [_expected(o, y, u, v, w, h, dw, dh, 3, 0, 1) for (y, u, v) in frames]
    lays = [lb.Layer.from_host(eng, 512, w, h, list(p), yuv_clamping=0, yuv_subspace=1) for p in frames]
    before = eng.launch_count
    assert lb.resize_layer_batch(lays, dw, dh, lb.LIVES_INTERP_NORMAL, 3, 0) == 5
    assert eng.launch_count - before == 1, "five same-shaped layers = one k_cvt_resize launch"
    for lay, e in zip(lays, exps):
        assert (lay.palette, lay.width, lay.height) == (3, dw, dh)
        assert (lay.to_host()[0][:, :dw * 4] == e[:, :dw * 4]).all()
    # YVU420P: the chroma planes change places on the way in (colourspace.c:12354)
    y, u, v = frames[0]
    lay = lb.Layer.from_host(eng, 513, w, h, [y, v, u], yuv_clamping=0, yuv_subspace=1)
    assert lb.resize_layer(lay, dw, dh, lb.LIVES_INTERP_NORMAL, 3, 0)
    assert (lay.to_host()[0][:, :dw * 4] == exps[0][:, :dw * 4]).all()
    # LIVES_INTERP_BEST (bicubic: negative taps) and a 3 x downscale (more than 4 taps) take the unfused kernels -- same contract
    for interp, (dw2, dh2) in ((lb.LIVES_INTERP_BEST, (1280, 720)), (lb.LIVES_INTERP_NORMAL, (640, 360))):
        lay = lb.Layer.from_host(eng, 512, w, h, list(frames[1]), yuv_clamping=0, yuv_subspace=1)
        before = eng.launch_count
        assert lb.resize_layer(lay, dw2, dh2, interp, 3, 0)
        assert eng.launch_count - before >= 2
        assert (lay.width, lay.height) == (dw2, dh2)
    # the fused path can be switched off: the unfused pair gives the same bytes
    os.environ["PE_NO_CVT_RESIZE"] = "1"
    try:
        lay = lb.Layer.from_host(eng, 512, w, h, list(frames[2]), yuv_clamping=0, yuv_subspace=1)
        before = eng.launch_count
        assert lb.resize_layer(lay, dw, dh, lb.LIVES_INTERP_NORMAL, 3, 0)
        assert eng.launch_count - before == 2
        assert (lay.to_host()[0][:, :dw * 4] == exps[2][:, :dw * 4]).all()
    finally:
        del os.environ["PE_NO_CVT_RESIZE"]
    eng.close()
