"""Boundary B1: libpe_weed_plugin.so is a real libweed effect plugin.

It is loaded by tests/host/weed_minihost.c through the REFERENCE's own libweed (oracle/_ref/libweed*.so:
dlopen + dlsym("weed_setup") + setup(weed_bootstrap), as src/effects-weed.c:4468-4568 does) and must behave like the
reference's simple_blend.so / multi_blends.so loaded the same way.
"""
import ctypes as C
import itertools
import os

import numpy as np
import pytest

import pe_testlib as T

PLUGIN = os.path.join(T.REPO, "lives_b200", "libpe_weed_plugin.so")
pytestmark = pytest.mark.skipif(not T.have_ref() or not os.path.exists(os.path.join(T.REF_DIR, "libweed_minihost.so")),
                                reason="oracle/_ref (reference libweed + minihost) not built")

NAMES = ["chroma blend", "luma overlay", "luma underlay", "negative luma overlay", "blend_multiply", "blend_screen",
         "blend_darken", "blend_lighten", "blend_overlay", "blend_dodge", "blend_burn", "slide over", "compositor", "softlight",
         "triple split", "iris rectangle", "iris circle", "4 way split", "dissolve", "rand replace", "averaged luma overlay"]


def _minihost():
    mh = C.CDLL(os.path.join(T.REF_DIR, "libweed_minihost.so"))
    mh.mh_open.argtypes = [C.c_char_p]
    mh.mh_run2.argtypes = [T.I, T.I, T.I, T.I, T.I, T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.I]
    mh.mh_run2v.argtypes = [T.I, T.I, T.I, T.I, T.I, T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.VP, T.I]
    return mh


def _sover_params(transval, direction, mvlower, mvupper):
    """the 8 in-parameters of slide_over.c:164-172 that make sover_init :38-52 pick `direction` (1 .. 4)"""
    return (C.c_int * 8)(transval, 0, int(direction == 1), int(direction == 2), int(direction == 3), 0, mvlower, mvupper)


def _open_ours(mh):
    if not os.path.exists(PLUGIN):
        from lives_b200.build import build
        build()
    h = mh.mh_open(PLUGIN.encode())
    assert h >= 0, "weed_setup(weed_bootstrap) of libpe_weed_plugin.so failed: %d" % h
    return h


def test_plugin_bootstraps_through_the_reference_libweed():
    mh = _minihost()
    h = _open_ours(mh)
    assert mh.mh_num_filters(h) == len(NAMES)
    buf = C.create_string_buffer(64)
    for i, name in enumerate(NAMES):
        assert mh.mh_filter_name(h, i, buf, 64) == 0 and buf.value.decode() == name
        flags = mh.mh_filter_flags(h, i)
        assert not flags & (1 << 6), "a GPU plugin must not set WEED_FILTER_HINT_MAY_THREAD (one process call per frame)"
    # same names, same order as the reference plugins
    ref = mh.mh_open(os.path.join(T.REF_DIR, "simple_blend.so").encode())
    for i in range(4):
        mh.mh_filter_name(ref, i, buf, 64)
        assert buf.value.decode() == NAMES[i]
    mh.mh_filter_name(ref, 4, buf, 64)
    assert buf.value.decode() == NAMES[20]   # the file's fifth filter is registered last in ours (earlier indices kept)
    ref = mh.mh_open(os.path.join(T.REF_DIR, "multi_blends.so").encode())
    for i in range(7):
        mh.mh_filter_name(ref, i, buf, 64)
        assert buf.value.decode() == NAMES[4 + i]
    ref = mh.mh_open(os.path.join(T.REF_DIR, "slide_over.so").encode())
    mh.mh_filter_name(ref, 0, buf, 64)
    assert buf.value.decode() == NAMES[11]
    for so, first, cnt in (("softlight", 13, 1), ("layout_blends", 14, 1), ("multi_transitions", 15, 5)):
        ref = mh.mh_open(os.path.join(T.REF_DIR, so + ".so").encode())
        assert mh.mh_num_filters(ref) == cnt
        for i in range(cnt):
            mh.mh_filter_name(ref, i, buf, 64)
            assert buf.value.decode() == NAMES[first + i]
    # the 8 parameter templates of "slide over" are accepted with the reference's seed types (a wrong count / type is rc -4 / garbage)
    s = np.zeros((8, 32), np.uint8)
    import torch
    if not torch.cuda.is_available():
        d = s.copy()
        assert mh.mh_run2v(h, 11, 1, 8, 8, T.ptr(s), 32, T.ptr(s), 32, T.ptr(d), 32, 8, _sover_params(10, 1, 1, 0), 1) == 64


def test_plugin_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    mh = _minihost()
    h = _open_ours(mh)
    s = np.zeros((8, 32), np.uint8)
    d = np.zeros((8, 32), np.uint8)
    rc = mh.mh_run2(h, 0, 1, 8, 8, T.ptr(s), 32, T.ptr(s), 32, T.ptr(d), 32, 100, 1)
    assert rc == 64  # WEED_ERROR_PLUGIN_INVALID: no CPU fallback inside the plugin


@pytest.mark.gpu
def test_plugin_matches_reference_plugins_bit_for_bit():
    mh = _minihost()
    ours = _open_ours(mh)
    ref_s = mh.mh_open(os.path.join(T.REF_DIR, "simple_blend.so").encode())
    ref_m = mh.mh_open(os.path.join(T.REF_DIR, "multi_blends.so").encode())
    rng = np.random.default_rng(31)
    for pal, bf, (w, ht) in itertools.product((1, 2, 3, 4, 5), (0, 100, 255), ((64, 32), (61, 7), (640, 360))):
        ps = T.psize_of(pal)
        s1, s2 = T.make_packed(rng, w, ht, ps), T.make_packed(rng, w, ht, ps)
        if ps == 4:
            al = s2[:, 3::4] if pal != 5 else s2[:, 0::4]
            al[rng.random(al.shape) < 0.4] = 255
        for typ in range(5):
            d_ref, d_our = np.full_like(s1, 9), np.full_like(s1, 9)
            assert mh.mh_run2(ref_s, typ, pal, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_ref), d_ref.strides[0], bf, 1) == 0
            assert mh.mh_run2(ours, 20 if typ == 4 else typ, pal, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_our), d_our.strides[0], bf, 1) == 0
            if pal == 5 and s2.strides[0] == w * 4:
                d_ref[-1, w * 4 - 3:] = d_our[-1, w * 4 - 3:]  # the reference reads one byte past the frame there (UB)
            assert (d_ref == d_our).all(), (pal, bf, typ, w, ht)
        # in place (CAN_DO_INPLACE, effects-weed.c:2304-2314)
        a, b = s1.copy(), s1.copy()
        mh.mh_run2(ref_s, 0, pal, w, ht, T.ptr(a), a.strides[0], T.ptr(s2), s2.strides[0], T.ptr(a), a.strides[0], bf, 1)
        mh.mh_run2(ours, 0, pal, w, ht, T.ptr(b), b.strides[0], T.ptr(s2), s2.strides[0], T.ptr(b), b.strides[0], bf, 1)
        if pal == 5 and s2.strides[0] == w * 4:
            a[-1, w * 4 - 3:] = b[-1, w * 4 - 3:]
        assert (a == b).all(), (pal, bf, "inplace")
    for pal, bf, typ in itertools.product((1, 2), (0, 17, 128, 255), range(7)):
        w, ht = 53, 12
        s1, s2 = T.make_packed(rng, w, ht, 3), T.make_packed(rng, w, ht, 3)
        d_ref, d_our = np.full_like(s1, 9), np.full_like(s1, 9)
        assert mh.mh_run2(ref_m, typ, pal, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_ref), d_ref.strides[0], bf, 1) == 0
        assert mh.mh_run2(ours, 4 + typ, pal, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_our), d_our.strides[0], bf, 1) == 0
        assert (d_ref[:, :w * 3] == d_our[:, :w * 3]).all(), (pal, bf, typ)


@pytest.mark.gpu
def test_slide_over_plugin_matches_reference_plugin():
    """"slide over" of libpe_weed_plugin.so against slide_over.so, both run by the minihost with the same 8 parameters"""
    mh = _minihost()
    ours = _open_ours(mh)
    ref = mh.mh_open(os.path.join(T.REF_DIR, "slide_over.so").encode())
    rng = np.random.default_rng(32)
    for (pal, ps), (w, ht) in itertools.product(((1, 3), (4, 4), (589, 4), (565, 4)), ((64, 32), (61, 7))):
        s1, s2 = T.make_packed(rng, w, ht, ps), T.make_packed(rng, w, ht, ps)
        for direction, mvl, mvu, tv in itertools.product((1, 2, 3, 4), (0, 1), (0, 1), (0, 77, 128, 255)):
            d_ref, d_our = np.full_like(s1, 9), np.full_like(s1, 9)
            pr = _sover_params(tv, direction, mvl, mvu)
            assert mh.mh_run2v(ref, 0, pal, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_ref), d_ref.strides[0], 8, pr, 1) == 0
            assert mh.mh_run2v(ours, 11, pal, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_our), d_our.strides[0], 8, pr, 1) == 0
            assert (d_ref == d_our).all(), (pal, w, ht, direction, mvl, mvu, tv)


@pytest.mark.gpu
def test_compositor_filter_config3_through_the_plugin_boundary():
    """BASELINE config 3 reached the way weed_apply_instance reaches it: the "compositor" filter of libpe_weed_plugin.so (templates of
    gdk/compositor.c:300-340: one repeating in channel, per-layer double arrays) run by the minihost through the reference's libweed,
    against bgcol fill + paint_pixel (:120, :172-197) of the oracle (pinned to the compiled paint_pixel).  Layers are painted last
    first; "revz" reverses; a disabled channel is skipped; offsets / scales other than 0 / 1 fail loudly."""
    mh = _minihost()
    mh.mh_run_compositor.argtypes = [T.I] * 6 + [T.VP, T.VP, T.VP, T.I] + [T.VP] * 6 + [T.I]
    ours = _open_ours(mh)
    o = T.oracle()
    rng = np.random.default_rng(33)
    fidx = NAMES.index("compositor")
    D = C.c_double
    for (pal, ps), (w, ht), revz in itertools.product(((3, 4), (1, 3), (2, 3)), ((3840, 2160), (61, 7)), (0, 1)):
        if (w, ht) == (3840, 2160) and (pal != 3 or revz):
            continue
        layers = [T.make_packed(rng, w, ht, ps) for _ in range(3)]
        alphas = [0.5, 0.3, 1.0] if (w, ht) != (3840, 2160) else [0.5, 1.0, 1.0]
        enabled = [True, (w, ht) != (61, 7) or revz == 0, True]
        bg = (10, 200, 77)
        exp = np.zeros_like(layers[0])
        o.pe_or_fill(T.ptr(exp), exp.strides[0], pal, w, ht, bg[0], bg[1], bg[2])
        order = range(3) if revz else range(2, -1, -1)
        for z in order:
            if enabled[z]:
                o.pe_or_alpha_over(T.ptr(exp), exp.strides[0], T.ptr(layers[z]), layers[z].strides[0], pal, w, ht, alphas[z])
        if ps == 4:
            exp[:, 3:w * 4:4] = 255
        srcs = (C.c_void_p * 3)(*[layers[z].ctypes.data if enabled[z] else None for z in range(3)])
        rss = (C.c_int * 3)(*[a.strides[0] for a in layers])
        zero, one = (D * 3)(0, 0, 0), (D * 3)(1, 1, 1)
        got = np.full_like(layers[0], 9)
        rc = mh.mh_run_compositor(ours, fidx, pal, w, ht, 3, srcs, rss, T.ptr(got), got.strides[0], zero, zero, one, one, (D * 3)(*alphas),
                                  (C.c_int * 3)(*bg), revz)
        assert rc == 0, rc
        assert (got[:, :w * ps] == exp[:, :w * ps]).all(), (pal, w, ht, revz)
    # a scaled layer is refused, the out channel keeps its bytes
    got = np.full_like(layers[0], 9)
    half = (D * 3)(0.5, 1, 1)
    rc = mh.mh_run_compositor(ours, fidx, pal, w, ht, 3, srcs, rss, T.ptr(got), got.strides[0], zero, zero, half, one, (D * 3)(*alphas),
                              (C.c_int * 3)(*bg), 0)
    assert rc == 65 and (got == 9).all()  # WEED_ERROR_FILTER_INVALID


@pytest.mark.gpu
def test_softlight_triple_split_multi_transitions_match_the_reference_plugins():
    """"softlight", "triple split" and the five multi_transitions filters of libpe_weed_plugin.so against the
    reference's softlight.so / layout_blends.so / multi_transitions.so, both run by the minihost with the same channels and parameters"""
    mh = _minihost()
    ours = _open_ours(mh)
    T.minihost()  # binds mh_run_generic on the same library object
    ref_sl = mh.mh_open(os.path.join(T.REF_DIR, "softlight.so").encode())
    ref_ts = mh.mh_open(os.path.join(T.REF_DIR, "layout_blends.so").encode())
    ref_mt = mh.mh_open(os.path.join(T.REF_DIR, "multi_transitions.so").encode())
    rng = np.random.default_rng(34)
    # ---- softlight
    for pal, clamp, (w, ht) in itertools.product((512, 513, 522, 544, 545), (0, 1), ((64, 16), (70, 9), (1920, 1080))):
        if (w, ht) == (1920, 1080) and pal != 512:
            continue
        if pal in (512, 513):
            ht &= ~1  # the host creates 4:2:0 layers with even sizes (colourspace.c:11603-11604)
        ys = T.rowstride(w, 1)
        cw = w if pal in (544, 545) else w >> 1
        chh = ht >> 1 if pal in (512, 513) else ht
        cs = ys if pal in (544, 545) else ys >> 1
        planes = [T.make_packed(rng, w, ht, 1, ys)] + [T.make_packed(rng, cw, chh, 1, cs) for _ in range(3 if pal == 545 else 2)]
        o_ref, o_our = [np.full_like(p, 7) for p in planes], [np.full_like(p, 7) for p in planes]
        assert T.mh_run(ref_sl, 0, [T.chan(pal, w, ht, planes, clamp)], T.chan(pal, w, ht, o_ref, clamp)) == 0
        assert T.mh_run(ours, NAMES.index("softlight"), [T.chan(pal, w, ht, planes, clamp)], T.chan(pal, w, ht, o_our, clamp)) == 0
        assert (o_ref[0][:, :w] == o_our[0][:, :w]).all(), (pal, clamp, w, ht)
        for a, b in zip(o_ref[1:], o_our[1:]):
            assert (a[:, :cw] == b[:, :cw]).all(), (pal, "chroma")
    # ---- triple split
    cases = [(0.666667, 1, 0.333333, 0, 0.0), (0.666667, 1, 0.333333, 0, 0.05), (0.2, 0, 0.7, 0, 0.03), (0.5, 1, 0.5, 1, 0.0),
             (0.3, 0, 0.9, 1, 0.07), (1.0, 1, 0.0, 1, 0.2)]
    for (w, ht), pal, (xs, sym, xe, vert, bw) in itertools.product(((64, 32), (61, 17), (1920, 1080)), (1, 2), cases):
        s1, s2 = T.make_packed(rng, w, ht, 3), T.make_packed(rng, w, ht, 3)
        params = [[xs], [sym], [1 - sym], [xe], [vert], [bw], [200, 100, 50]]
        d_ref, d_our = np.full_like(s1, 9), np.full_like(s1, 9)
        assert T.mh_run(ref_ts, 0, [T.chan(pal, w, ht, [s1]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [d_ref]), params) == 0
        assert T.mh_run(ours, NAMES.index("triple split"), [T.chan(pal, w, ht, [s1]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [d_our]), params) == 0
        assert (d_ref[:, :w * 3] == d_our[:, :w * 3]).all(), (w, ht, pal, xs, sym, xe, vert, bw)
        a, b = s1.copy(), s1.copy()  # in place
        T.mh_run(ref_ts, 0, [T.chan(pal, w, ht, [a]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [a]), params)
        T.mh_run(ours, NAMES.index("triple split"), [T.chan(pal, w, ht, [b]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [b]), params)
        assert (a[:, :w * 3] == b[:, :w * 3]).all()
    # ---- multi_transitions
    first = NAMES.index("iris rectangle")
    for ftype, (pal, ps), (w, ht) in itertools.product(range(4), ((1, 3), (3, 4), (565, 4), (588, 3)), ((64, 32), (61, 17), (37, 50), (1280, 720))):
        if (w, ht) == (1280, 720) and pal not in (1, 3):
            continue
        s1, s2 = T.make_packed(rng, w, ht, ps), T.make_packed(rng, w, ht, ps)
        seed = int(rng.integers(1, 2 ** 62))
        for bf in [0.0, 1.0, 0.5, 1 / 3, 0.013, 0.999] + [float(x) for x in rng.random(3)]:
            d_ref, d_our = np.full_like(s1, 9), np.full_like(s1, 9)
            assert T.mh_run(ref_mt, ftype, [T.chan(pal, w, ht, [s1]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [d_ref]), [[bf]], seed=seed) == 0
            assert T.mh_run(ours, first + ftype, [T.chan(pal, w, ht, [s1]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [d_our]), [[bf]], seed=seed) == 0
            assert (d_ref[:, :w * ps] == d_our[:, :w * ps]).all(), (ftype, pal, w, ht, bf)
            if ftype != 2:
                a, b = s1.copy(), s1.copy()
                T.mh_run(ref_mt, ftype, [T.chan(pal, w, ht, [a]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [a]), [[bf]], seed=seed)
                T.mh_run(ours, first + ftype, [T.chan(pal, w, ht, [b]), T.chan(pal, w, ht, [s2])], T.chan(pal, w, ht, [b]), [[bf]], seed=seed)
                assert (a[:, :w * ps] == b[:, :w * ps]).all(), (ftype, "inplace", pal, w, ht, bf)
    # ---- rand replace: a whole-frame choice
    s1, s2 = T.make_packed(rng, 64, 8, 3), T.make_packed(rng, 64, 8, 3)
    for bf, want in ((0.0, s1), (1.0, s2)):
        d = np.full_like(s1, 9)
        assert T.mh_run(ours, NAMES.index("rand replace"), [T.chan(1, 64, 8, [s1]), T.chan(1, 64, 8, [s2])], T.chan(1, 64, 8, [d]), [[bf]]) == 0
        assert (d[:, :192] == want[:, :192]).all()
