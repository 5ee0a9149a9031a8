"""How far is the resize contract (oracle pe_or_resize_packed == the CUDA path, bit for bit) from the library
the reference actually calls?  Measured against a real libswscale when one is loadable (tests/swscale_ref.py) -- the version is
not the reference's to pin (configure.ac:562), so these are DISTANCE bounds, not parity: resize stays "parity unpinned".

Findings the bounds encode (libswscale 9.1.100, SWS_BILINEAR, whole-frame call):
  * RGBA -> RGBA at horizontal ratios below 2: colour samples 97-98 % equal at config 2's 1.5 and on upscales, 75 % on the
    headline's vertical 2160 -> 1608 squeeze (the tap cut-off of swscale's coefficient recipe matters there), > 99.9 % within +-1
    everywhere; a NOISY alpha channel is only ~50 % equal (within 1): swscale scales alpha less precisely (opaque alpha: exact);
  * at a horizontal downscale of 2 or more swscale computes chroma from every other source pixel (its RGB input goes through
    YUV; chrSrcHSubSample is set when dstW <= srcW / 2 unless SWS_FULL_CHR_H_INP): grey images still match, coloured detail does not
    -- a documented divergence;
  * YUV420P -> RGBA in ONE swscale call (what the reference does for config 2, colourspace.c:14601-14620) uses swscale's own
    YUV -> RGB arithmetic, not LiVES's converter: mean distance ~2 levels to convert_yuv420p_to_rgb_frame + resize."""
import numpy as np
import pytest

import pe_testlib as T
import swscale_ref as S

pytestmark = pytest.mark.skipif(S.load()[0] is None, reason=str(S.load()[1]))


def _textured(rng, w, h, ch, sigma):
    yy, xx = np.mgrid[0:h, 0:w]
    base = (np.sin(xx / 37.0) + np.cos(yy / 23.0)) * 60 + 128
    return np.clip(base[:, :, None] + rng.normal(0, sigma, (h, w, ch)), 0, 255).astype(np.uint8)


def _oracle_resize(src, w, h, dw, dh):
    got = np.zeros((dh, T.rowstride(dw, 4)), np.uint8)
    T.oracle().pe_or_resize_packed(T.ptr(src), src.strides[0], w, h, T.ptr(got), got.strides[0], dw, dh, 4)
    return got


@pytest.mark.parametrize("geom", [(1920, 1080, 1280, 720), (3840, 2160, 3840, 1608), (320, 240, 640, 480), (300, 200, 160, 120),
                                  (200, 100, 200, 100)])
def test_rgba_resize_distance_to_swscale_below_2x(geom):
    w, h, dw, dh = geom
    rng = np.random.default_rng(w + dh)
    src = np.zeros((h, T.rowstride(w, 4)), np.uint8)
    src[:, :w * 4] = _textured(rng, w, h, 4, 25).reshape(h, w * 4)
    ref = S.scale([src], "rgba", w, h, "rgba", dw, dh, T.rowstride(dw, 4))
    d = np.abs(ref[:, :dw * 4].astype(int) - _oracle_resize(src, w, h, dw, dh)[:, :dw * 4].astype(int))
    col = np.ones(dw * 4, bool)
    col[3::4] = False  # a noisy alpha channel takes a less precise scaler inside swscale (~50 % equal, within 1): reported apart
    c, a = d[:, col], d[:, ~col]
    print(geom, "colour: max %d mean %.4f equal %.2f%% within1 %.3f%% | alpha: max %d equal %.2f%%"
          % (c.max(), c.mean(), 100 * (c == 0).mean(), 100 * (c <= 1).mean(), a.max(), 100 * (a == 0).mean()))
    if (w, h) == (dw, dh):
        assert d.max() == 0
    else:
        assert (c == 0).mean() > 0.70 and (c <= 1).mean() > 0.999 and c.max() <= 12 and c.mean() < 0.3 and a.max() <= 12


def test_grey_2x_downscale_matches_and_colour_does_not():
    """dstW <= srcW / 2: same taps (grey content within +-1), but swscale's chroma comes from every other source pixel"""
    w, h, dw, dh = 640, 480, 320, 240
    rng = np.random.default_rng(5)
    grey = np.zeros((h, T.rowstride(w, 4)), np.uint8)
    grey[:, :w * 4] = np.repeat(_textured(rng, w, h, 1, 25), 4, axis=2).reshape(h, w * 4)
    d = np.abs(S.scale([grey], "rgba", w, h, "rgba", dw, dh, T.rowstride(dw, 4))[:, :dw * 4].astype(int)
               - _oracle_resize(grey, w, h, dw, dh)[:, :dw * 4].astype(int))
    assert d.max() <= 2 and (d <= 1).mean() > 0.99, (d.max(), (d <= 1).mean())
    col = np.zeros_like(grey)
    col[:, :w * 4] = rng.integers(0, 256, (h, w * 4), dtype=np.uint8)
    d = np.abs(S.scale([col], "rgba", w, h, "rgba", dw, dh, T.rowstride(dw, 4))[:, :dw * 4].astype(int)
               - _oracle_resize(col, w, h, dw, dh)[:, :dw * 4].astype(int))
    assert d.mean() > 2  # the divergence is real; if a future contract closes it this line goes


def test_config2_one_call_swscale_distance():
    """BASELINE config 2 as the reference runs it (YUV420P 1080p -> RGBA 720p in ONE sws_scale, BT.601 clamped) against our
    convert (reference converter arithmetic) + resize chain: a different YUV -> RGB arithmetic, a couple of levels apart"""
    o = T.oracle()
    fw, fh, dw, dh = 1920, 1080, 1280, 720
    rng = np.random.default_rng(3)
    y, u, v = T.make_yuv_planar(rng, fw, fh, False, True)
    yy, xx = np.mgrid[0:fh, 0:fw]
    y[:, :fw] = np.clip(126 + 70 * np.sin(xx / 91.) * np.cos(yy / 57.) + rng.normal(0, 6, (fh, fw)), 16, 235).astype(np.uint8)
    cy, cx = np.mgrid[0:fh // 2, 0:fw // 2]
    u[:, :fw // 2] = np.clip(128 + 50 * np.sin(cx / 77.) + rng.normal(0, 3, (fh // 2, fw // 2)), 16, 240).astype(np.uint8)
    v[:, :fw // 2] = np.clip(128 + 50 * np.cos(cy / 66.) + rng.normal(0, 3, (fh // 2, fw // 2)), 16, 240).astype(np.uint8)
    ref = S.scale([y, u, v], "yuv420p", fw, fh, "rgba", dw, dh, T.rowstride(dw, 4), yuv=(False, False, False))
    rgba = np.zeros((fh, T.rowstride(fw, 4)), np.uint8)
    o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), fw, fh, T.ptr(rgba), rgba.strides[0], 0, 1, 0, 0, 1,
                           T.Q_HIGH, 1, None)
    got = _oracle_resize(rgba, fw, fh, dw, dh)
    a = ref[:, :dw * 4].astype(int).reshape(dh, dw, 4)[:, :, :3]
    b = got[:, :dw * 4].astype(int).reshape(dh, dw, 4)[:, :, :3]
    d = np.abs(a - b)
    print("config 2: max %d mean %.3f within2 %.1f%% within4 %.1f%%" % (d.max(), d.mean(), 100 * (d <= 2).mean(), 100 * (d <= 4).mean()))
    assert d.mean() < 3 and (d <= 4).mean() > 0.9 and d.max() <= 24


@pytest.mark.parametrize("geom", [(1920, 1080, 1280, 720), (960, 2160, 960, 1608), (320, 240, 640, 480), (300, 200, 160, 120)])
def test_libswscale_coefficient_recipe_is_closer(geom):
    """the opt-in second recipe of the oracle (pe_or_resize_filter_sws: libswscale's own way of cutting, folding and normalising
    the bilinear taps; same two integer passes): on per-channel uniform noise at least 97.5 % (upscale: 93 %) of the colour samples equal
    the library's and no sample, alpha included, is further than 1 away -- the default contract leaves outliers of 6-8 on downscales.  Not the default
    because the CUDA side has not been run with it yet."""
    w, h, dw, dh = geom
    o = T.oracle()
    rng = np.random.default_rng(dw)
    src = np.zeros((h, T.rowstride(w, 4)), np.uint8)
    src[:, :w * 4] = rng.integers(0, 256, (h, w * 4), dtype=np.uint8)
    ref = S.scale([src], "rgba", w, h, "rgba", dw, dh, T.rowstride(dw, 4))[:, :dw * 4].astype(int)
    d0 = np.abs(ref - _oracle_resize(src, w, h, dw, dh)[:, :dw * 4].astype(int))
    o.pe_or_set_resize_recipe(1)
    try:
        d1 = np.abs(ref - _oracle_resize(src, w, h, dw, dh)[:, :dw * 4].astype(int))
    finally:
        o.pe_or_set_resize_recipe(0)
    col = np.ones(dw * 4, bool)
    col[3::4] = False  # the alpha channel goes through a different (less precise) scaler inside swscale: only ~50 % equal on noise
    print(geom, "contract: colour %.2f%% equal, max %d | libswscale recipe: colour %.2f%% equal, max %d; alpha %.2f%% equal, max %d"
          % (100 * (d0[:, col] == 0).mean(), d0[:, col].max(), 100 * (d1[:, col] == 0).mean(), d1[:, col].max(),
             100 * (d1[:, ~col] == 0).mean(), d1[:, ~col].max()))
    assert d1.max() <= 1 and (d1[:, col] == 0).mean() > (0.93 if dw > w else 0.975)
    assert (d1[:, col] == 0).mean() >= (d0[:, col] == 0).mean() - 1e-3 and d1.max() <= d0.max()
