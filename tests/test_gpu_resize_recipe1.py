"""OPT-IN GPU check of the second resize recipe (libswscale's coefficient recipe): runs only with
PE_TEST_RECIPE1=1, because the recipe was added after round 1's GPU budget was spent and has not had its GPU pass.  With the variable
set it must be bit-exact against the oracle under the same recipe -- unfused resize, and the fused headline chain through every
fused kernel's envelope -- before recipe 1 may become the default.

    PE_TEST_RECIPE1=1 python -m pytest tests/test_gpu_resize_recipe1.py -m gpu -q
"""
import os

import numpy as np
import pytest

import pe_testlib as T

lb = pytest.importorskip("lives_b200")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PE_TEST_RECIPE1") != "1", reason="opt-in: PE_TEST_RECIPE1=1 (recipe 1 has had no GPU pass yet)")]


@pytest.fixture()
def recipe1():
    o = T.oracle()
    e = lb.Engine()
    e.set_resize_recipe(1)
    o.pe_or_set_resize_recipe(1)
    yield e, o
    o.pe_or_set_resize_recipe(0)
    e.close()


@pytest.mark.parametrize("case", [(64, 48, 32, 24, 3), (64, 48, 96, 72, 4), (130, 50, 77, 34, 4), (1920, 1080, 1280, 720, 4),
                                  (640, 360, 3840, 2160, 4), (100, 100, 100, 50, 3), (3840, 2160, 3840, 1608, 4), (300, 200, 75, 50, 4)])
def test_resize_packed_recipe1(recipe1, case):
    e, o = recipe1
    rng = np.random.default_rng(41)
    sw, sh, dw, dh, ps = case
    pal = 1 if ps == 3 else 3
    src = T.make_packed(rng, sw, sh, ps)
    exp = np.zeros((dh, T.rowstride(dw, ps)), np.uint8)
    o.pe_or_resize_packed(T.ptr(src), src.strides[0], sw, sh, T.ptr(exp), exp.strides[0], dw, dh, ps)
    lay = lb.Layer.from_host(e, pal, sw, sh, [src])
    assert lb.resize_layer(lay, dw, dh, lb.LIVES_INTERP_NORMAL, pal, 0)
    assert (lay.to_host()[0][:, :dw * ps] == exp[:, :dw * ps]).all()


@pytest.mark.parametrize("geom", [(1280, 720, 1280, 720, 536), (3840, 2160, 3840, 2160, 1608), (640, 360, 640, 360, 300),
                                  (320, 240, 320, 240, 236)])
def test_fused_chain_recipe1(recipe1, geom):
    """the headline chain (YUV420P -> RGBA, vertical squeeze, letterbox, alpha-over 0.5, gamma) under recipe 1"""
    e, o = recipe1
    rng = np.random.default_rng(42)
    fw, fh, ow, oh, ih = geom
    y, u, v = T.make_yuv_planar(rng, fw, fh, False, True)
    bg = T.make_packed(rng, ow, oh, 4)
    rgba = np.zeros((fh, T.rowstride(fw, 4)), np.uint8)
    o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), fw, fh, T.ptr(rgba), rgba.strides[0], 0, 1, 0, 0, 1,
                           T.Q_HIGH, 1, None)
    inner = np.zeros((ih, T.rowstride(fw, 4)), np.uint8)
    o.pe_or_resize_packed(T.ptr(rgba), rgba.strides[0], fw, fh, T.ptr(inner), inner.strides[0], fw, ih, 4)
    boxed = np.zeros((oh, T.rowstride(ow, 4)), np.uint8)
    o.pe_or_letterbox_packed(T.ptr(inner), inner.strides[0], fw, ih, T.ptr(boxed), boxed.strides[0], ow, oh, 3)
    exp = bg.copy()
    o.pe_or_alpha_over(T.ptr(exp), exp.strides[0], T.ptr(boxed), boxed.strides[0], 3, ow, oh, 0.5)
    exp[:, 3:ow * 4:4] = 255
    lut = np.zeros(256, np.uint8)
    o.pe_or_gamma_lut8(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut))
    o.pe_or_gamma_apply(T.ptr(exp), exp.strides[0], 3, 0, 0, ow, oh, T.ptr(lut))
    fg_l = lb.Layer.from_host(e, lb.WEED_PALETTE_YUV420P, fw, fh, [y, u, v], yuv_subspace=1)
    bg_l = lb.Layer.from_host(e, lb.WEED_PALETTE_RGBA32, ow, oh, [bg], gamma_type=T.G_LINEAR)
    out_l = lb.Layer.create(e, lb.WEED_PALETTE_RGBA32, ow, oh)
    lb.fused_convert_letterbox_over_gamma(fg_l, bg_l, out_l, fw, ih, 0.5, T.G_LINEAR, T.G_SRGB)
    assert (out_l.to_host()[0][:, :ow * 4] == exp[:, :ow * 4]).all()
