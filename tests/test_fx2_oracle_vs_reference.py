"""CPU: the oracle's restatements of softlight.c, layout_blends.c ("triple split") and multi_transitions.c against the reference's
own plugins, compiled unchanged (oracle/build_ref.py, with the plugins' -ffast-math as lives-plugins/weed-plugins/Makefile.am:49 sets
it) and run by the mini weed host through the reference's libweed (weed_setup(weed_bootstrap), init / process / deinit)."""
import ctypes as C
import itertools
import os

import numpy as np
import pytest

import pe_testlib as T

pytestmark = pytest.mark.skipif(not T.have_ref() or not os.path.exists(os.path.join(T.REF_DIR, "multi_transitions.so")),
                                reason="oracle/_ref (reference plugins + minihost) not built")

F = C.c_float


def _o():
    o = T.oracle()
    o.pe_or_softlight.argtypes = [T.VP, T.I, T.VP, T.I, T.I, T.I, T.I]
    o.pe_or_triple_split.argtypes = [T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.I, T.I, T.D, T.I, T.D, T.I, T.D, T.VP]
    o.pe_or_dissolve_mask.argtypes = [C.c_int64, C.c_long, T.VP]
    o.pe_or_multi_transition.argtypes = [T.I, T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.I, T.I, T.D, T.VP]
    return o


def _open(name):
    h = T.minihost().mh_open(os.path.join(T.REF_DIR, name + ".so").encode())
    assert h >= 0
    return h


@pytest.mark.parametrize("pal,clamped", [(512, 0), (512, 1), (522, 1), (544, 0), (545, 1), (513, 0)])
def test_softlight_luma_plane(pal, clamped):
    o, h = _o(), _open("softlight")
    rng = np.random.default_rng(pal + clamped)
    for w, ht in ((64, 16), (70, 9), (6, 4), (256, 33)):
        ys = T.rowstride(w, 1)
        cw = w if pal in (544, 545) else w >> 1
        chh = ht >> 1 if pal in (512, 513) else ht
        cs = ys if pal in (544, 545) else ys >> 1
        y = T.make_packed(rng, w, ht, 1, ys)
        planes = [y] + [T.make_packed(rng, cw, chh, 1, cs) for _ in range(3 if pal == 545 else 2)]
        outs = [np.full_like(p, 7) for p in planes]
        # WEED_YUV_CLAMPING_CLAMPED 0, UNCLAMPED 1 (the plugin tests == UNCLAMPED, softlight.c:100)
        assert T.mh_run(h, 0, [T.chan(pal, w, ht, planes, 0 if clamped else 1)], T.chan(pal, w, ht, outs, 0 if clamped else 1)) == 0
        exp = np.full_like(y, 7)
        o.pe_or_softlight(T.ptr(y), ys, T.ptr(exp), ys, w, ht, clamped)
        assert (outs[0][:, :w] == exp[:, :w]).all(), (pal, w, ht)
        for p, q in zip(planes[1:], outs[1:]):
            assert (p[:, :cw] == q[:, :cw]).all()  # chroma / alpha planes are copied (:150-154)


def test_triple_split_every_parameter_shape():
    o, h = _o(), _open("layout_blends")
    rng = np.random.default_rng(2)
    cases = [(0.666667, 1, 0.333333, 0, 0.0), (0.666667, 1, 0.333333, 0, 0.05), (0.2, 0, 0.7, 0, 0.03), (0.8, 0, 0.1, 0, 0.1),
             (0.5, 1, 0.5, 1, 0.0), (0.3, 0, 0.9, 1, 0.07), (0.0, 1, 1.0, 0, 0.5), (1.0, 1, 0.0, 1, 0.2), (0.4, 0, 0.4, 0, 0.01)]
    for (w, ht), pal, (xs, sym, xe, vert, bw) in itertools.product(((64, 32), (61, 17), (200, 50)), (1, 2), cases):
        s1, s2 = T.make_packed(rng, w, ht, 3), T.make_packed(rng, w, ht, 3)
        col = [200, 100, 50]
        d = np.full_like(s1, 9)
        rc = T.mh_run(h, 0, [T.chan(pal, w, ht, [s1]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [d]),
                      [[xs], [sym], [1 - sym], [xe], [vert], [bw], col])
        assert rc == 0
        exp = np.full_like(s1, 9)
        o.pe_or_triple_split(T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0], w, ht, int(pal == 2), xs, sym, xe,
                             vert, bw, (C.c_int * 3)(*col))
        assert (d == exp).all(), (w, ht, pal, xs, sym, xe, vert, bw)
        # in place (CAN_DO_INPLACE): the pixels that keep src1 are not written
        a, b = s1.copy(), s1.copy()
        T.mh_run(h, 0, [T.chan(pal, w, ht, [a]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [a]), [[xs], [sym], [1 - sym], [xe], [vert], [bw], col])
        o.pe_or_triple_split(T.ptr(b), b.strides[0], T.ptr(s2), s2.strides[0], T.ptr(b), b.strides[0], w, ht, int(pal == 2), xs, sym, xe, vert, bw,
                             (C.c_int * 3)(*col))
        assert (a == b).all()


@pytest.mark.parametrize("ftype", [0, 1, 2, 3])
def test_multi_transitions_against_the_compiled_plugin(ftype):
    o, h = _o(), _open("multi_transitions")
    rng = np.random.default_rng(10 + ftype)
    amounts = [0.0, 1.0, 0.5, 0.25, 1 / 3, 0.9, 0.013, 0.77, 0.999]
    for (pal, ps), (w, ht) in itertools.product(((1, 3), (3, 4), (565, 4), (588, 3)), ((64, 32), (61, 17), (37, 50), (130, 9), (8, 8))):
        s1, s2 = T.make_packed(rng, w, ht, ps), T.make_packed(rng, w, ht, ps)
        seed = int(rng.integers(1, 2 ** 62))
        mask = np.zeros(w * ht, np.float32)
        o.pe_or_dissolve_mask(seed, w * ht, T.ptr(mask))
        for bf in amounts + [float(x) for x in rng.random(6)]:
            d = np.full_like(s1, 9)
            assert T.mh_run(h, ftype, [T.chan(pal, w, ht, [s1]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [d]), [[bf]], seed=seed) == 0
            exp = np.full_like(s1, 9)
            o.pe_or_multi_transition(ftype, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0], w, ht, ps, bf, T.ptr(mask))
            assert (d == exp).all(), (ftype, pal, w, ht, bf, np.argwhere(d != exp)[:4])
            if ftype != 2:  # in place (the "4 way split" out channel is not CAN_DO_INPLACE, multi_transitions.c:268)
                a, b = s1.copy(), s1.copy()
                T.mh_run(h, ftype, [T.chan(pal, w, ht, [a]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [a]), [[bf]], seed=seed)
                o.pe_or_multi_transition(ftype, T.ptr(b), b.strides[0], T.ptr(s2), s2.strides[0], T.ptr(b), b.strides[0], w, ht, ps, bf, T.ptr(mask))
                assert (a == b).all(), (ftype, "inplace", pal, w, ht, bf)


def test_rand_replace_is_a_whole_frame_choice():
    """type 4: every frame is src1 or src2 as a whole (the plugin's own random stream decides; amount 0 -> always src1, 1 -> src2)"""
    h = _open("multi_transitions")
    rng = np.random.default_rng(5)
    s1, s2 = T.make_packed(rng, 32, 8, 3), T.make_packed(rng, 32, 8, 3)
    for bf, want in ((0.0, s1), (1.0, s2)):
        d = np.full_like(s1, 9)
        assert T.mh_run(h, 4, [T.chan(1, 32, 8, [s1]), T.chan(1, 32, 8, [s2])], T.chan(1, 32, 8, [d]), [[bf]]) == 0
        assert (d[:, :96] == want[:, :96]).all()
    d = np.full_like(s1, 9)
    T.mh_run(h, 4, [T.chan(1, 32, 8, [s1]), T.chan(1, 32, 8, [s2])], T.chan(1, 32, 8, [d]), [[0.5]])
    assert (d[:, :96] == s1[:, :96]).all() or (d[:, :96] == s2[:, :96]).all()
