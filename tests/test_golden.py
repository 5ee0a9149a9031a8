"""Golden vectors frozen from the compiled reference (tests/golden/make_golden.py -> ref_vectors.npz).

CPU: the oracle reproduces every vector (this is what pins the oracle on a box without /root/reference).
GPU: the CUDA path reproduces them through the C ABI.
"""
import hashlib
import os

import ctypes as C

import numpy as np
import pytest

import pe_testlib as T

G = np.load(os.path.join(T.GOLDEN, "ref_vectors.npz"))
ORD = {"rgb24": (0, 0, 1), "rgba32": (0, 1, 3), "bgr24": (1, 0, 2), "bgra32": (1, 1, 4)}


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


# ------------------------------------------------------------------------------------------------ oracle (CPU)

def test_oracle_tables_match_golden():
    o = T.oracle()
    for cl in (0, 1):
        for sub in (1, 2):
            for w in range(14):
                a = np.zeros(256, np.int32)
                o.pe_or_conv_table(cl, sub, w, T.ptr(a))
                assert (a == G["conv_cl%d_sub%d" % (cl, sub)][w]).all(), (cl, sub, w)
    for w in range(6):
        t = np.zeros(65536, np.int32)
        o.pe_or_premult_table(w, T.ptr(t))
        assert (sha(t) == G["premult%d_sha256" % w]).all()
        assert (t[128 * 256:129 * 256] == G["premult%d_row128" % w]).all()
    for f in (-1, 1, 2, 1024):
        for t in (-1, 1, 2, 1024):
            if f == t:
                continue
            a, a16 = np.zeros(256, np.uint8), np.zeros(65536, np.uint16)
            assert o.pe_or_gamma_lut8(1.0, f, t, 1.4, T.ptr(a)) == 0 and o.pe_or_gamma_lut16(1.0, f, t, 1.4, T.ptr(a16)) == 0
            assert (a == G["lut8_%d_%d" % (f, t)]).all(), (f, t)
            assert (sha(a16) == G["lut16_%d_%d_sha256" % (f, t)]).all(), (f, t)


def _planar_cases():
    for is422 in (0, 1):
        for cl in (0, 1):
            key = "p%d_cl%d" % (422 if is422 else 420, cl)
            for name, (order, add_alpha, pal) in ORD.items():
                if key + "_" + name in G.files:
                    yield is422, cl, key, name, order, add_alpha, pal


def _rows(is422, h):
    return slice(0, h) if is422 else slice(1, h - 1)  # 4:2:0: rows 0 and h-1 are the X rows


def test_oracle_frames_match_golden():
    o = T.oracle()
    w, h = 64, 48
    for is422, cl, key, name, order, add_alpha, pal in _planar_cases():
        y, u, v = T._plane_like(G[key + "_y"]), T._plane_like(G[key + "_u"]), T._plane_like(G[key + "_v"])
        exp = G[key + "_" + name]
        ps = 4 if add_alpha else 3
        got = np.zeros_like(exp)
        o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(got), got.strides[0], order, add_alpha,
                               is422, cl, 1, T.Q_HIGH, 1, None)
        rs = _rows(is422, h)
        assert (got[rs, :w * ps] == exp[rs, :w * ps]).all(), (key, name)
    src = G["uyvy_src"]
    for fmt, nm in ((0, "uyvy"), (1, "yuyv")):
        got = np.zeros_like(G[nm + "_to_rgb24"])
        o.pe_or_packed422_to_rgb(fmt, T.ptr(src), src.strides[0], 48, 20, T.ptr(got), got.strides[0], 0, 0, 0, 1, T.Q_HIGH)
        assert (got == G[nm + "_to_rgb24"]).all()
    src = G["rgb_src"]
    got = np.zeros_like(G["rgb_to_yuv888_cl0"])
    o.pe_or_rgb_to_yuv888(T.ptr(src), src.strides[0], 50, 12, T.ptr(got), got.strides[0], 0, 0, 0, 0, T.Q_HIGH)
    assert (got == G["rgb_to_yuv888_cl0"]).all()
    got = np.zeros_like(G["yuv888_to_rgb_cl0"])
    o.pe_or_yuv888_to_rgb(T.ptr(src), src.strides[0], 50, 12, T.ptr(got), got.strides[0], 0, 0, 0, 0, 1, T.Q_HIGH)
    assert (got == G["yuv888_to_rgb_cl0"]).all()


def test_oracle_rgb_to_yuv_match_golden():
    """RGB -> UYVY / YUYV / YUV444P / YUV420P / YUV422P and the chroma averaging tables, frozen from the compiled reference"""
    import hashlib
    o = T.oracle()
    src = np.ascontiguousarray(G["rgb2_src"])
    for fmt, nm in ((0, "uyvy"), (1, "yuyv")):
        got = np.zeros_like(G["rgb_to_%s_cl0" % nm])
        o.pe_or_rgb_to_packed422(fmt, T.ptr(src), src.strides[0], 64, 12, T.ptr(got), got.strides[0], 0, 0, 0, T.Q_HIGH, None)
        assert (got == G["rgb_to_%s_cl0" % nm]).all(), nm
    pl = [np.zeros((12, 64), np.uint8) for _ in range(4)]
    o.pe_or_rgb_to_yuv444p(T.ptr(src), src.strides[0], 64, 12, T.planes_arg(*pl), 64, 0, 0, 0, 0, T.Q_HIGH)
    for k, pn in enumerate("yuv"):
        assert (pl[k] == G["rgb_to_yuv444p_cl0_" + pn]).all(), pn
    for is422, nm in ((0, "yuv420p"), (1, "yuv422p")):
        ch = 12 if is422 else 6
        pl = [np.zeros((12, 64), np.uint8), np.zeros((ch, 32), np.uint8), np.zeros((ch, 32), np.uint8)]
        o.pe_or_rgb_to_yuv420p(T.ptr(src), src.strides[0], 64, 12, T.planes_arg(*pl), (C.c_int * 3)(64, 32, 32), 0, 0, is422, 0, 1, T.Q_HIGH)
        for k, pn in enumerate("yuv"):
            assert (pl[k] == G["rgb_to_%s_cl0_%s" % (nm, pn)]).all(), (nm, pn)
    for which, nm in ((0, "cavgc"), (1, "cavgu")):
        t = np.zeros(65536, np.uint8)
        o.pe_or_avg_table(which, T.ptr(t))
        assert hashlib.sha256(t.tobytes()).digest() == G["avg_%s_sha256" % nm].tobytes(), nm
        assert (t[200 * 256:201 * 256] == G["avg_%s_row200" % nm]).all()


def test_oracle_effects_match_golden():
    o = T.oracle()
    s1, s2 = G["blend_s1"], G["blend_s2"]
    for typ in range(4):
        d = np.zeros_like(s1)
        o.pe_or_simple_blend(typ, 1, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d), d.strides[0], 61, 9, 100, s2.size)
        assert (d == G["simple_blend_t%d_bf100" % typ]).all(), typ
    for typ in range(7):
        d = np.zeros_like(s1)
        o.pe_or_multi_blend(typ, 1, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d), d.strides[0], 61, 9, 100)
        assert (d == G["multi_blend_t%d_bf100" % typ]).all(), typ
    a1, a2 = G["blend4_s1"], G["blend4_s2"]
    d = a1.copy()
    o.pe_or_simple_blend(0, 3, T.ptr(a1), a1.strides[0], T.ptr(a2), a2.strides[0], T.ptr(d), d.strides[0], 33, 5, 77, a2.size)
    assert (d == G["simple_blend_rgba_bf77"]).all()
    g = np.arange(256, dtype=np.uint8)
    d0, s0 = np.meshgrid(g, g, indexing="ij")
    for alpha in (0.1, 0.5, 1.0 / 3.0):
        dst = np.repeat(d0.reshape(-1, 1), 3, axis=1).astype(np.uint8).copy()
        srcp = np.repeat(s0.reshape(-1, 1), 3, axis=1).astype(np.uint8).copy()
        o.pe_or_alpha_over(T.ptr(dst), 65536 * 3, T.ptr(srcp), 65536 * 3, 1, 65536, 1, alpha)
        assert (dst[:, 0].reshape(256, 256) == G["paint_alpha_%.4f" % alpha]).all(), alpha


# ------------------------------------------------------------------------------------------------ CUDA path (GPU)

@pytest.mark.gpu
def test_cuda_matches_golden():
    lb = pytest.importorskip("lives_b200")
    eng = lb.Engine()
    w, h = 64, 48
    for f in (-1, 1, 2, 1024):
        for t in (-1, 1, 2, 1024):
            if f != t:
                assert (eng.gamma_lut8(1.0, f, t) == G["lut8_%d_%d" % (f, t)]).all(), (f, t)
    for is422, cl, key, name, order, add_alpha, pal in _planar_cases():
        y, u, v = T._plane_like(G[key + "_y"]), T._plane_like(G[key + "_u"]), T._plane_like(G[key + "_v"])
        lay = lb.Layer.from_host(eng, 522 if is422 else 512, w, h, [y, u, v], yuv_clamping=cl, yuv_subspace=1)
        assert lb.convert_layer_palette_full(lay, pal, cl, 0, 1, 0)
        got, exp = lay.to_host()[0], G[key + "_" + name]
        ps = 4 if add_alpha else 3
        rs = _rows(is422, h)
        assert (got[rs, :w * ps] == exp[rs, :w * ps]).all(), (key, name)
    src = np.ascontiguousarray(G["uyvy_src"])
    for pal, nm in ((564, "uyvy"), (565, "yuyv")):
        lay = lb.Layer.from_host(eng, pal, 96, 20, [src], yuv_clamping=0, yuv_subspace=1)
        assert lb.convert_layer_palette_full(lay, 1, 0, 0, 1, 0)
        assert (lay.to_host()[0][:, :288] == G[nm + "_to_rgb24"][:, :288]).all(), nm
    src = np.ascontiguousarray(G["rgb_src"])
    lay = lb.Layer.from_host(eng, 1, 50, 12, [src])
    assert lb.convert_layer_palette(lay, 588, 0)
    assert (lay.to_host()[0][:, :150] == G["rgb_to_yuv888_cl0"][:, :150]).all()
    lay = lb.Layer.from_host(eng, 588, 50, 12, [src], yuv_clamping=0, yuv_subspace=1)
    assert lb.convert_layer_palette(lay, 1, 0)
    assert (lay.to_host()[0][:, :150] == G["yuv888_to_rgb_cl0"][:, :150]).all()
    src = np.ascontiguousarray(G["rgb2_src"])
    for pal, nm in ((564, "uyvy"), (565, "yuyv")):
        lay = lb.Layer.from_host(eng, 1, 64, 12, [src])
        assert lb.convert_layer_palette(lay, pal, 0)
        assert (lay.to_host()[0][:, :128] == G["rgb_to_%s_cl0" % nm]).all(), nm
    for pal, nm in ((544, "yuv444p"), (512, "yuv420p"), (522, "yuv422p")):
        lay = lb.Layer.from_host(eng, 1, 64, 12, [src])
        assert lb.convert_layer_palette_full(lay, pal, 0, 0, 1, 0)
        got = lay.to_host()
        for k, pn in enumerate("yuv"):
            exp = G["rgb_to_%s_cl0_%s" % (nm, pn)]
            assert (got[k][:exp.shape[0], :exp.shape[1]] == exp).all(), (nm, pn)
    s1, s2 = np.ascontiguousarray(G["blend_s1"]), np.ascontiguousarray(G["blend_s2"])
    l1, l2 = lb.Layer.from_host(eng, 1, 61, 9, [s1]), lb.Layer.from_host(eng, 1, 61, 9, [s2])
    for typ in range(4):
        lo = lb.Layer.create(eng, 1, 61, 9)
        lb.simple_blend(typ, l1, l2, lo, 100)
        assert (lo.to_host()[0][:, :183] == G["simple_blend_t%d_bf100" % typ][:, :183]).all(), typ
    for typ in range(7):
        lo = lb.Layer.create(eng, 1, 61, 9)
        lb.multi_blend(typ, l1, l2, lo, 100)
        assert (lo.to_host()[0][:, :183] == G["multi_blend_t%d_bf100" % typ][:, :183]).all(), typ
    a1, a2 = np.ascontiguousarray(G["blend4_s1"]), np.ascontiguousarray(G["blend4_s2"])
    l1, l2 = lb.Layer.from_host(eng, 3, 33, 5, [a1]), lb.Layer.from_host(eng, 3, 33, 5, [a2])
    lb.simple_blend(0, l1, l2, l1, 77)
    assert (l1.to_host()[0][:, :132] == G["simple_blend_rgba_bf77"][:, :132]).all()
    # paint_pixel tables: a 256 x 256 RGB24 "frame" whose pixel (d, s) holds bg = d, fg = s
    g = np.arange(256, dtype=np.uint8)
    d0, s0 = np.meshgrid(g, g, indexing="ij")
    bg = np.repeat(d0[:, :, None], 3, axis=2).reshape(256, 768).copy()
    fg = np.repeat(s0[:, :, None], 3, axis=2).reshape(256, 768).copy()
    for alpha in (0.1, 0.5, 1.0 / 3.0):
        out = lb.Layer.create(eng, 1, 256, 256)
        lb.compositor(out, [lb.Layer.from_host(eng, 1, 256, 256, [fg]), lb.Layer.from_host(eng, 1, 256, 256, [bg])], [alpha, 1.0])
        assert (out.to_host()[0][:, 0::3] == G["paint_alpha_%.4f" % alpha]).all(), alpha
    eng.close()
