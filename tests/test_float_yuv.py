"""The reference's float ("experimental") YUV -> RGB path: float tables (colourspace.c:1040-1104), clamp0255f (:592), yuv2rgb_float
(:2367).  CPU: oracle == compiled reference (tables bit for bit, all 2^24 YUV triples for both forms and both clampings), product
tables == oracle.  GPU: the CUDA path against the oracle -- bytes equal, float sums 0 ULP apart."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pe_testlib as T  # noqa: E402


def _o():
    o = T.oracle()
    o.pe_or_float_table.argtypes = [T.I, T.I, T.VP]
    o.pe_or_yuv2rgb_float.argtypes = [T.I, T.I, T.VP, T.VP, T.VP, T.VP, C.c_long]
    return o


def _all_triples():
    a = np.arange(1 << 24, dtype=np.uint32)
    return np.stack([(a & 255), (a >> 8) & 255, a >> 16], axis=1).astype(np.uint8)


def _rgb_y(o, cl):
    ty = np.zeros(256, np.int32)
    o.pe_or_conv_table(cl, T.SUB_BT709, 9, T.ptr(ty))
    return ty


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built")
def test_float_tables_and_every_triple_against_the_compiled_reference():
    o = _o()
    r = C.CDLL(os.path.join(T.REF_DIR, "libref_oracle.so"))
    r.ref_init()
    r.ref_get_float_table.argtypes = [T.I, T.I, T.VP]
    r.ref_yuv2rgb_float_bulk.argtypes = [T.I, T.VP, T.VP, C.c_long]
    r.ref_yuv2rgb_floaty_bulk.argtypes = [T.I, T.VP, T.VP, T.VP, C.c_long]
    yuv = _all_triples()
    n = len(yuv)
    for cl in (T.CLAMPED, T.UNCLAMPED):
        for w in range(5):
            a, b = np.zeros(256, np.float32), np.zeros(256, np.float32)
            o.pe_or_float_table(cl, w, T.ptr(a))
            assert r.ref_get_float_table(cl, w, T.ptr(b)) == 0
            assert (a.view(np.uint32) == b.view(np.uint32)).all(), (cl, w)
        ty = _rgb_y(o, cl)
        got, exp = np.zeros_like(yuv), np.zeros_like(yuv)
        o.pe_or_yuv2rgb_float(0, cl, T.ptr(ty), T.ptr(yuv), T.ptr(got), None, n)
        r.ref_yuv2rgb_float_bulk(cl, T.ptr(yuv), T.ptr(exp), n)
        assert (got == exp).all(), (cl, "as written")
        sg, se = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        o.pe_or_yuv2rgb_float(1, cl, T.ptr(ty), T.ptr(yuv), T.ptr(got), T.ptr(sg), n)
        r.ref_yuv2rgb_floaty_bulk(cl, T.ptr(yuv), T.ptr(exp), T.ptr(se), n)
        assert (got == exp).all() and (sg.view(np.uint32) == se.view(np.uint32)).all(), (cl, "RGBf_Y form")


def test_clamped_float_luma_table_quirk():
    o = _o()
    t = np.zeros(256, np.float32)
    o.pe_or_float_table(T.CLAMPED, 0, T.ptr(t))
    assert t[234] > 253 and (t[235:] == 0).all() and (t[:17] == 0).all()


def test_product_float_tables_equal_the_oracle():
    lb = pytest.importorskip("lives_b200")
    o = _o()
    for cl in (T.CLAMPED, T.UNCLAMPED):
        for w in range(5):
            a = np.zeros(256, np.float32)
            o.pe_or_float_table(cl, w, T.ptr(a))
            assert (lb.float_yuv_table(cl, w).view(np.uint32) == a.view(np.uint32)).all(), (cl, w)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("cl", [T.CLAMPED, T.UNCLAMPED])
def test_cuda_float_path_is_bit_exact_every_triple(mode, cl):
    """all 2^24 (Y, U, V) triples as a 4096 x 4096 YUV888 frame: bytes equal to the oracle's, float sums 0 ULP apart"""
    lb = pytest.importorskip("lives_b200")
    o = _o()
    eng = lb.Engine()
    yuv = _all_triples()
    w = h = 4096
    src = np.zeros((h, T.rowstride(w, 3)), np.uint8)
    src[:, :w * 3] = yuv.reshape(h, w * 3)
    lay = lb.Layer.from_host(eng, T.PAL["YUV888"], w, h, [src], yuv_clamping=cl, yuv_subspace=T.SUB_BT709)
    ok, sums = lb.convert_yuv888_to_rgb_float(lay, T.PAL["RGB24"], mode, want_sums=True)
    assert ok and lay.palette == T.PAL["RGB24"]
    got = lay.to_host()[0][:, :w * 3].reshape(-1, 3)
    exp, se = np.zeros_like(yuv), np.zeros((len(yuv), 3), np.float32)
    ty = _rgb_y(o, cl)   # (kept alive across the call: T.ptr() of a temporary is a dangling pointer once the array is collected)
    o.pe_or_yuv2rgb_float(mode, cl, T.ptr(ty), T.ptr(yuv), T.ptr(exp), T.ptr(se), len(yuv))
    bad = np.flatnonzero((got != exp).any(axis=1))
    assert bad.size == 0, "%d triples differ, first (Y, U, V) = %s: got %s, oracle %s" % (bad.size, yuv[bad[0]], got[bad[0]], exp[bad[0]])
    ulp = np.abs(sums.reshape(-1, 3).view(np.int32).astype(np.int64) - se.view(np.int32).astype(np.int64))
    assert ulp.max() == 0, "float sums differ by up to %d ULP" % ulp.max()
    eng.close()


@pytest.mark.gpu
def test_cuda_float_path_layouts_alpha_and_refusals():
    lb = pytest.importorskip("lives_b200")
    o = _o()
    eng = lb.Engine()
    rng = np.random.default_rng(8)
    w, h = 61, 17
    src = T.make_packed(rng, w, h, 4)
    yuv = np.ascontiguousarray(src[:, :w * 4].reshape(-1, 4)[:, :3])
    exp = np.zeros_like(yuv)
    ty = _rgb_y(o, T.UNCLAMPED)
    o.pe_or_yuv2rgb_float(1, T.UNCLAMPED, T.ptr(ty), T.ptr(yuv), T.ptr(exp), None, len(yuv))
    for pal, order in ((T.PAL["RGBA32"], (0, 1, 2, 3)), (T.PAL["BGRA32"], (2, 1, 0, 3)), (T.PAL["ARGB32"], (1, 2, 3, 0)), (T.PAL["BGR24"], (2, 1, 0))):
        lay = lb.Layer.from_host(eng, T.PAL["YUVA8888"], w, h, [src], yuv_clamping=T.UNCLAMPED, yuv_subspace=T.SUB_BT709)
        assert lb.convert_yuv888_to_rgb_float(lay, pal, 1)
        ps = len(order)
        got = lay.to_host()[0][:, :w * ps].reshape(-1, ps)
        for c in range(3):
            assert (got[:, order[c]] == exp[:, c]).all(), (pal, c)
        if ps == 4:
            assert (got[:, order[3]] == src[:, :w * 4].reshape(-1, 4)[:, 3]).all()
    lay = lb.Layer.from_host(eng, T.PAL["YUVA8888"], w, h, [src], yuv_clamping=T.UNCLAMPED, yuv_subspace=T.SUB_YCBCR)
    assert not lb.convert_yuv888_to_rgb_float(lay, T.PAL["RGB24"], 1) and lay.palette == T.PAL["YUVA8888"]  # no float tables for YCbCr
    eng.close()
