"""GPU: softlight / triple split / multi_transitions on DEVICE frames (pe_fx_*) against the oracle (itself pinned to the compiled
reference plugins on the CPU, tests/test_fx2_oracle_vs_reference.py), at ragged sizes and at 4K; errors leave the out frame alone."""
import ctypes as C
import itertools
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pe_testlib as T  # noqa: E402

lb = pytest.importorskip("lives_b200")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = lb.Engine()
    yield e
    e.close()


def _o():
    o = T.oracle()
    o.pe_or_softlight.argtypes = [T.VP, T.I, T.VP, T.I, T.I, T.I, T.I]
    o.pe_or_triple_split.argtypes = [T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.I, T.I, T.D, T.I, T.D, T.I, T.D, T.VP]
    o.pe_or_dissolve_mask.argtypes = [C.c_int64, C.c_long, T.VP]
    o.pe_or_multi_transition.argtypes = [T.I, T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.I, T.I, T.D, T.VP]
    return o


@pytest.mark.parametrize("size", [(64, 16), (70, 9), (6, 4), (3840, 2160), (1282, 50)])
@pytest.mark.parametrize("pal,clamped", [(512, 1), (522, 0), (545, 0)])
def test_softlight_device_frames(eng, size, pal, clamped):
    o = _o()
    w, ht = size
    if pal == 512:
        ht &= ~1  # 4:2:0 frames are created with even sizes (colourspace.c:11603-11604)
    rng = np.random.default_rng(w + pal)
    ys = T.rowstride(w, 1)
    cw = w if pal == 545 else w >> 1
    chh = (ht + 1) >> 1 if pal == 512 else ht
    cs = ys if pal == 545 else ys >> 1
    planes = [T.make_packed(rng, w, ht, 1, ys)] + [T.make_packed(rng, cw, chh, 1, cs) for _ in range(3 if pal == 545 else 2)]
    clamping = 0 if clamped else 1
    src = lb.Layer.from_host(eng, pal, w, ht, planes, yuv_clamping=clamping)
    dst = lb.Layer.create(eng, pal, w, ht, yuv_clamping=clamping)
    lb.softlight(src, dst)
    got = dst.to_host()
    exp = np.zeros_like(planes[0])
    o.pe_or_softlight(T.ptr(planes[0]), ys, T.ptr(exp), ys, w, ht, clamped)
    assert (got[0][:, :w] == exp[:, :w]).all()
    ch_rows = ht >> 1 if pal == 512 else ht  # the plugin copies height >> 1 chroma rows of a 4:2:0 frame (:147)
    for a, b in zip(planes[1:], got[1:]):
        assert (a[:ch_rows, :cw] == b[:ch_rows, :cw]).all()
    with pytest.raises(lb.PixelEngineError):
        lb.softlight(src, src)


@pytest.mark.parametrize("size", [(64, 32), (61, 17), (3840, 2160)])
def test_triple_split_device_frames(eng, size):
    o = _o()
    w, ht = size
    rng = np.random.default_rng(w)
    for pal, (xs, sym, xe, vert, bw) in itertools.product((1, 2), [(0.666667, 1, 0.333333, 0, 0.02), (0.25, 0, 0.6, 1, 0.05), (0.3, 0, 0.8, 0, 0.0)]):
        s1, s2 = T.make_packed(rng, w, ht, 3), T.make_packed(rng, w, ht, 3)
        col = (200, 100, 50)
        a, b = lb.Layer.from_host(eng, pal, w, ht, [s1]), lb.Layer.from_host(eng, pal, w, ht, [s2])
        d = lb.Layer.create(eng, pal, w, ht)
        lb.triple_split(a, b, d, xs, sym, xe, vert, bw, col)
        exp = np.zeros_like(s1)
        o.pe_or_triple_split(T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0], w, ht, int(pal == 2), xs, sym, xe,
                             vert, bw, (C.c_int * 3)(*col))
        assert (d.to_host()[0][:, :w * 3] == exp[:, :w * 3]).all(), (pal, xs, sym, xe, vert, bw)
        lb.triple_split(a, b, a, xs, sym, xe, vert, bw, col)  # in place
        assert (a.to_host()[0][:, :w * 3] == exp[:, :w * 3]).all()


@pytest.mark.parametrize("ftype", [0, 1, 2, 3])
@pytest.mark.parametrize("size", [(64, 32), (61, 17), (3840, 2160)])
def test_multi_transition_device_frames(eng, ftype, size):
    o = _o()
    w, ht = size
    rng = np.random.default_rng(w * 7 + ftype)
    for (pal, ps), bf in itertools.product(((1, 3), (3, 4)), (0.0, 1.0, 0.37, 2 / 3)):
        s1, s2 = T.make_packed(rng, w, ht, ps), T.make_packed(rng, w, ht, ps)
        seed = 987654321 + w
        mask_h = np.zeros(w * ht, np.float32)
        o.pe_or_dissolve_mask(seed, w * ht, T.ptr(mask_h))
        mask = lb.DissolveMask(eng, w, ht, seed) if ftype == 3 else None
        a, b = lb.Layer.from_host(eng, pal, w, ht, [s1]), lb.Layer.from_host(eng, pal, w, ht, [s2])
        d = lb.Layer.create(eng, pal, w, ht)
        lb.multi_transition(ftype, a, b, d, bf, mask)
        exp = np.zeros_like(s1)
        o.pe_or_multi_transition(ftype, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0], w, ht, ps, bf, T.ptr(mask_h))
        assert (d.to_host()[0][:, :w * ps] == exp[:, :w * ps]).all(), (ftype, pal, bf)
        if ftype == 2:
            with pytest.raises(lb.PixelEngineError):
                lb.multi_transition(ftype, a, b, a, bf, mask)
        else:
            lb.multi_transition(ftype, a, b, a, bf, mask)
            assert (a.to_host()[0][:, :w * ps] == exp[:, :w * ps]).all()
        if mask is not None:
            mask.close()
    if ftype == 3:
        with pytest.raises(lb.PixelEngineError):
            lb.multi_transition(3, a, b, d, 0.5, None)


@pytest.mark.parametrize("size", [(64, 16), (70, 9), (1920, 1080)])
@pytest.mark.parametrize("clamping", [0, 1])
def test_yuv444p_yuv422p_planar_both_ways(eng, size, clamping):
    """4:4:4 planar <-> 4:2:2 planar (X row of the quirk table: the reference's dispatcher calls its VERTICAL resamplers there,
    colourspace.c:12948,13719).  Contract = the reference's own horizontal converters composed: combineplanes + yuv888_to_yuv422 one way,
    double_chroma_packed + splitplanes the other."""
    o = T.oracle()
    w, ht = size
    rng = np.random.default_rng(w + clamping)
    rs = T.rowstride(w, 1)
    p444 = [T.make_packed(rng, w, ht, 1, rs) for _ in range(3)]
    # ---- 4:4:4 -> 4:2:2
    lay = lb.Layer.from_host(eng, 544, w, ht, p444, yuv_clamping=clamping)
    assert lb.convert_layer_palette(lay, 522, clamping)
    assert lay.palette == 522 and lay.width == (w & ~1)
    got = lay.to_host()
    t888 = np.zeros((ht, T.rowstride(w, 3)), np.uint8)
    o.pe_or_combine_planes(T.planes_arg(*p444, p444[0]), rs, w, ht, T.ptr(t888), t888.strides[0], 0, 0)
    exp = [np.zeros_like(g) for g in got]
    o.pe_or_yuv888_subsample(2, T.ptr(t888), t888.strides[0], w & ~1, ht, 0, T.planes_arg(*exp), T.strides_arg(*exp), clamping)
    cw = (w & ~1) >> 1
    assert (got[0][:, :w & ~1] == exp[0][:, :w & ~1]).all()
    assert (got[1][:, :cw] == exp[1][:, :cw]).all() and (got[2][:, :cw] == exp[2][:, :cw]).all()
    # ---- 4:2:2 -> 4:4:4 (+ alpha)
    w2 = w & ~1
    y, u, v = T.make_yuv_planar(rng, w2, ht, True, clamping == 0)
    for outpl in (544, 545):
        lay = lb.Layer.from_host(eng, 522, w2, ht, [y, u, v], yuv_clamping=clamping)
        assert lb.convert_layer_palette(lay, outpl, clamping)
        got = lay.to_host()
        t888 = np.zeros((ht, T.rowstride(w2, 3)), np.uint8)
        o.pe_or_chroma_upsample_packed(0, T.planes_arg(y, u, v), T.strides_arg(y, u, v), w2, ht, T.ptr(t888), t888.strides[0], 0, 1, clamping)
        exp = [np.zeros_like(got[0]) for _ in range(4)]
        o.pe_or_split_planes(T.ptr(t888), t888.strides[0], w2, ht, T.planes_arg(*exp), T.strides_arg(*exp), 0, 0)
        for k in range(3):
            assert (got[k][:, :w2] == exp[k][:, :w2]).all(), (outpl, k)
        if outpl == 545:
            assert (got[3][:, :w2] == 255).all()
