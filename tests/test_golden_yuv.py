"""Golden vectors of the YUV <-> YUV family, frozen from the compiled reference (tests/golden/make_golden_yuv.py ->
ref_vectors_yuv.npz).  CPU: the oracle reproduces every vector; GPU: the CUDA path reproduces them through the C ABI."""
import itertools
import os

import numpy as np
import pytest

import pe_testlib as T

G = np.load(os.path.join(T.GOLDEN, "ref_vectors_yuv.npz"))
W, H = 48, 10
WM = W // 2


def _c(name):
    return np.ascontiguousarray(G[name])


def test_oracle_yuv_family_matches_golden():
    o = T.oracle()
    for w in range(4):
        t = np.zeros(256, np.uint8)
        o.pe_or_yy_table(w, T.ptr(t))
        assert (t == G["yy%d" % w]).all(), w
    pl = [_c("p444_%d" % k) for k in range(4)]
    for cl in (0, 1):
        for nm, order, ia, oa in (("rgb24", 0, 0, 0), ("rgba32", 0, 1, 1), ("bgra32", 1, 0, 1)):
            exp = G["yuv444p_to_%s_cl%d" % (nm, cl)]
            d = np.zeros_like(exp)
            o.pe_or_yuv444p_to_rgb(T.planes_arg(*pl), pl[0].strides[0], W, H, T.ptr(d), d.strides[0], order, ia, oa, cl, T.Q_HIGH)
            assert (d == exp).all(), (nm, cl)
    for oa in (0, 1):
        exp = G["combine_a%d" % oa]
        d = np.zeros_like(exp)
        o.pe_or_combine_planes(T.planes_arg(*pl), pl[0].strides[0], W, H, T.ptr(d), d.strides[0], 0, oa)
        assert (d == exp).all(), oa
    src = _c("yuv888_src")
    sp = [np.zeros_like(G["split_0"]) for _ in range(4)]
    o.pe_or_split_planes(T.ptr(src), src.strides[0], W, H, T.planes_arg(*sp), T.strides_arg(*sp), 0, 0)
    for k in range(3):
        assert (sp[k] == G["split_%d" % k]).all(), k
    c422 = [np.zeros_like(G["c422_u"]), _c("c422_u"), _c("c422_v")]
    c420 = [np.zeros_like(G["c420_u"]), _c("c420_u"), _c("c420_v")]
    for cl in (0, 1):
        d = [np.zeros_like(G["halve_cl0_u"]) for _ in range(3)]
        o.pe_or_halve_chroma(T.planes_arg(*c422), T.strides_arg(*c422), WM, H, T.planes_arg(*d), T.strides_arg(*d), cl)
        assert (d[1] == G["halve_cl%d_u" % cl]).all() and (d[2] == G["halve_cl%d_v" % cl]).all(), cl
        d = [np.zeros_like(G["double_cl0_u"]) for _ in range(3)]
        o.pe_or_double_chroma(T.planes_arg(*c420), T.strides_arg(*c420), WM, H // 2, T.planes_arg(*d), T.strides_arg(*d), cl)
        assert (d[1] == G["double_cl%d_u" % cl]).all() and (d[2] == G["double_cl%d_v" % cl]).all(), cl
    m = _c("mpx_src")
    for fmt, nm in ((0, "uyvy"), (1, "yuyv")):
        d = [np.zeros((H, W), np.uint8), np.zeros((H, WM), np.uint8), np.zeros((H, WM), np.uint8)]
        o.pe_or_packed422_to_yuv422p(fmt, T.ptr(m), m.strides[0], WM, H, T.planes_arg(*d), T.strides_arg(*d), 1)
        for k, pn in enumerate("yuv"):
            assert (d[k] == G["%s_to_yuv422p_%s" % (nm, pn)]).all(), (nm, pn)
        d = [np.zeros_like(G["uyvy_to_yuva4444p_y"]) for _ in range(4)]
        o.pe_or_packed422_to_yuv444p(fmt, T.ptr(m), m.strides[0], WM, H, T.planes_arg(*d), T.strides_arg(*d), 1)
        for k, pn in enumerate("yuva"):
            assert (d[k] == G["%s_to_yuva4444p_%s" % (nm, pn)]).all(), (nm, pn)
        for aa in (0, 1):
            exp = G["%s_to_yuv888_a%d" % (nm, aa)]
            d8 = np.zeros_like(exp)
            o.pe_or_packed422_to_yuv888(fmt, T.ptr(m), m.strides[0], WM, H, T.ptr(d8), d8.strides[0], aa)
            assert (d8 == exp).all(), (nm, aa)
    sw = m.copy()
    o.pe_or_swab(T.ptr(sw), sw.strides[0], WM, H)
    assert (sw == G["swab"]).all()
    s888 = _c("yuv888_dense")
    for cl in (0, 1):
        for mode in range(4):
            exp = [G["yuv888_sub_m%d_cl%d_%d" % (mode, cl, k)] for k in range(1 if mode <= 1 else 3)]
            d = [np.zeros_like(e) for e in exp]
            pa = d + [d[0]] * (3 - len(d))
            o.pe_or_yuv888_subsample(mode, T.ptr(s888), s888.strides[0], W, H, 0, T.planes_arg(*pa), T.strides_arg(*pa), cl)
            for k in range(len(d)):
                assert (d[k] == exp[k]).all(), (mode, cl, k)
    qs = [np.zeros_like(G["quad_src_u"]), _c("quad_src_u"), _c("quad_src_v")]
    for samp in (0, 1):
        for cl in (0, 1):
            d = [np.zeros((9, T.align_ceil(W + 1, 32)), np.uint8) for _ in range(4)]
            o.pe_or_quad_chroma(T.planes_arg(*qs), T.strides_arg(*qs), W, 9, T.planes_arg(*d), d[0].strides[0], 0, int(samp == 0), cl)
            assert (d[1][:, :W] == G["quad_s%d_cl%d_u" % (samp, cl)]).all() and (d[2][:, :W] == G["quad_s%d_cl%d_v" % (samp, cl)]).all(), (samp, cl)
    for cl in (0, 1):
        for fmt, nm in ((0, "uyvy"), (1, "yuyv")):
            exp = G["yuv444p_to_%s_cl%d" % (nm, cl)]
            d = np.zeros_like(exp)
            o.pe_or_yuv444p_to_packed422(fmt, T.planes_arg(*pl[:3]), pl[0].strides[0], W, H, T.ptr(d), d.strides[0], cl)
            assert (d == exp).all(), (nm, cl)
        eu, ev = G["yuv444p_to_yuv420p_cl%d_u" % cl], G["yuv444p_to_yuv420p_cl%d_v" % cl]
        d3 = [np.zeros((H, pl[0].strides[0]), np.uint8), np.zeros_like(eu), np.zeros_like(ev)]
        o.pe_or_yuv444p_to_yuv420p(T.planes_arg(*pl[:3]), T.strides_arg(*pl[:3]), W, H, T.planes_arg(*d3), T.strides_arg(*d3), cl)
        assert (d3[1] == eu).all() and (d3[2] == ev).all(), cl
    uy = _c("up_y")
    for is420, nm, hh in ((0, "422", H), (1, "420", 9)):
        cs = [_c("up%s_u" % nm), _c("up%s_v" % nm)]
        for samp, cl, aa in itertools.product((0, 1), (0, 1), (0, 1)):
            exp = G["up%s_s%d_cl%d_a%d" % (nm, samp, cl, aa)]
            d = np.zeros((hh, T.align_ceil(W * (4 if aa else 3) + 8, 32)), np.uint8)
            o.pe_or_chroma_upsample_packed(is420, T.planes_arg(uy, *cs), T.strides_arg(uy, *cs), W, hh, T.ptr(d), d.strides[0], aa,
                                           int(samp == 0), cl)
            assert (d[:exp.shape[0], :exp.shape[1]] == exp).all(), (nm, samp, cl, aa)


@pytest.mark.gpu
def test_cuda_yuv_family_matches_golden():
    lb = pytest.importorskip("lives_b200")
    eng = lb.Engine()
    pl = [_c("p444_%d" % k) for k in range(4)]
    for cl in (0, 1):
        for nm, ipal, opal, ps in (("rgb24", 544, 1, 3), ("rgba32", 545, 3, 4), ("bgra32", 544, 4, 4)):
            lay = lb.Layer.from_host(eng, ipal, W, H, pl[:4 if ipal == 545 else 3], yuv_clamping=cl)
            assert lb.convert_layer_palette(lay, opal, cl)
            assert (lay.to_host()[0][:, :W * ps] == G["yuv444p_to_%s_cl%d" % (nm, cl)][:, :W * ps]).all(), (nm, cl)
    for oa, opal in ((0, 588), (1, 589)):
        lay = lb.Layer.from_host(eng, 544, W, H, pl[:3])
        assert lb.convert_layer_palette(lay, opal, 0)
        n = W * (4 if oa else 3)
        assert (lay.to_host()[0][:, :n] == G["combine_a%d" % oa][:, :n]).all(), oa
    lay = lb.Layer.from_host(eng, 588, W, H, [_c("yuv888_src")])
    assert lb.convert_layer_palette(lay, 544, 0)
    for k, g in enumerate(lay.to_host()):
        assert (g[:, :W] == G["split_%d" % k][:, :W]).all(), k
    ydummy = np.zeros((H, T.rowstride(W, 1)), np.uint8)
    for cl in (0, 1):
        lay = lb.Layer.from_host(eng, 522, W, H, [ydummy, _c("c422_u"), _c("c422_v")], yuv_clamping=cl)
        assert lb.convert_layer_palette(lay, 512, cl)
        got = lay.to_host()
        assert (got[1][:, :WM] == G["halve_cl%d_u" % cl][:, :WM]).all() and (got[2][:, :WM] == G["halve_cl%d_v" % cl][:, :WM]).all(), cl
        lay = lb.Layer.from_host(eng, 512, W, H, [ydummy, _c("c420_u"), _c("c420_v")], yuv_clamping=cl)
        assert lb.convert_layer_palette(lay, 522, cl)
        got = lay.to_host()
        assert (got[1][:, :WM] == G["double_cl%d_u" % cl][:, :WM]).all() and (got[2][:, :WM] == G["double_cl%d_v" % cl][:, :WM]).all(), cl
    m = _c("mpx_src")
    for ipal, nm in ((564, "uyvy"), (565, "yuyv")):
        lay = lb.Layer.from_host(eng, ipal, W, H, [m])
        assert lb.convert_layer_palette(lay, 522, 0)  # ref_quirks on: the first macropixel everywhere, as the reference
        got = lay.to_host()
        for k, pn in enumerate("yuv"):
            exp = G["%s_to_yuv422p_%s" % (nm, pn)]
            assert (got[k][:, :exp.shape[1]] == exp).all(), (nm, pn)
        lay = lb.Layer.from_host(eng, ipal, W, H, [m])
        assert lb.convert_layer_palette(lay, 545, 0)
        for k, (g, pn) in enumerate(zip(lay.to_host(), "yuva")):
            assert (g[:, :W] == G["%s_to_yuva4444p_%s" % (nm, pn)][:, :W]).all(), (nm, pn)
        for aa, opal in ((0, 588), (1, 589)):
            lay = lb.Layer.from_host(eng, ipal, W, H, [m])
            assert lb.convert_layer_palette(lay, opal, 0)
            n = W * (4 if aa else 3)
            assert (lay.to_host()[0][:, :n] == G["%s_to_yuv888_a%d" % (nm, aa)][:, :n]).all(), (nm, aa)
    for cl in (0, 1):
        for opal, nm in ((564, "uyvy"), (565, "yuyv")):
            lay = lb.Layer.from_host(eng, 544, W, H, pl[:3], yuv_clamping=cl)
            assert lb.convert_layer_palette(lay, opal, cl)
            assert (lay.to_host()[0][:, :2 * W] == G["yuv444p_to_%s_cl%d" % (nm, cl)]).all(), (nm, cl)
        lay = lb.Layer.from_host(eng, 544, W, H, pl[:3], yuv_clamping=cl)
        assert lb.convert_layer_palette(lay, 512, cl)
        got = lay.to_host()
        assert (got[1][:, :WM] == G["yuv444p_to_yuv420p_cl%d_u" % cl][:, :WM]).all() and (got[2][:, :WM] == G["yuv444p_to_yuv420p_cl%d_v" % cl][:, :WM]).all()
    lay = lb.Layer.from_host(eng, 564, W, H, [m])
    assert lb.convert_layer_palette(lay, 565, 0)
    assert (lay.to_host()[0][:, :WM * 4] == G["swab"][:, :WM * 4]).all()
    yy = [np.ascontiguousarray(G["yy%d" % k]) for k in range(4)]
    lay = lb.Layer.from_host(eng, 564, W, H, [m], yuv_clamping=0, yuv_subspace=1)  # UYVY clamped -> unclamped in place
    assert lb.convert_layer_palette_full(lay, 564, 1, 0, 1, 0)
    got, px = lay.to_host()[0][:, :WM * 4], m[:, :WM * 4]
    assert (got[:, 1::2] == yy[0][px[:, 1::2]]).all() and (got[:, 0::2] == yy[1][px[:, 0::2]]).all()
    uy = _c("up_y")
    for ipal, nm, hh in ((522, "422", H), (512, "420", 9)):
        cs = [_c("up%s_u" % nm), _c("up%s_v" % nm)]
        for samp, cl, (aa, opal) in itertools.product((0, 1), (0, 1), ((0, 588), (1, 589))):
            exp = G["up%s_s%d_cl%d_a%d" % (nm, samp, cl, aa)]
            lay = lb.Layer.from_host(eng, ipal, W, hh, [np.ascontiguousarray(uy[:hh]), *cs], yuv_clamping=cl, yuv_sampling=samp)
            assert lb.convert_layer_palette(lay, opal, cl)
            assert (lay.to_host()[0][:exp.shape[0], :exp.shape[1]] == exp).all(), (nm, samp, cl, aa)
    eng.close()
