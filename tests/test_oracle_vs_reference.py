"""Pin the CPU oracle (oracle/pe_oracle.c, our restatement) against the reference itself.

The reference has no golden vectors for this path, so the oracle is pinned
against binaries that oracle/build_ref.py compiles from the reference sources where they lie:
  oracle/_ref/libref_oracle.so   line-range slices of src/colourspace.c (all convert_*_frame loops)
  oracle/_ref/simple_blend.so / multi_blends.so  the unmodified effect plugins, driven through the
                                 real libweed bootstrap by tests/host/weed_minihost.c
  oracle/_ref/ref_paint_pixel.so compositor.c paint_pixel()
CPU only (no GPU, no product code).  Skipped when oracle/_ref is absent.
"""
import ctypes as C
import itertools
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pe_testlib import *  # noqa: E402,F401,F403
import pe_testlib as T  # noqa: E402

pytestmark = pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def test_conversion_tables_match_reference():
    o, r = T.oracle(), T.ref()
    for cl, sub, w in itertools.product((0, 1), (1, 2), range(14)):
        a = np.zeros(256, np.int32)
        b = np.zeros(256, np.int32)
        o.pe_or_conv_table(cl, sub, w, T.ptr(a))
        r.ref_get_conv_table(cl, sub, w, T.ptr(b))
        assert (a == b).all(), (cl, sub, w)


def test_premult_tables_match_reference():
    o, r = T.oracle(), T.ref()
    for w in range(6):
        a = np.zeros(65536, np.int32)
        b = np.zeros(65536, np.int32)
        o.pe_or_premult_table(w, T.ptr(a))
        r.ref_get_premult_table(w, T.ptr(b))
        assert (a == b).all(), w


_GAMMA_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import pe_testlib as T
f, t = int(sys.argv[1]), int(sys.argv[2])
o, r = T.oracle(), T.ref()
a = np.zeros(256, np.uint8); b = np.zeros(256, np.uint8)
ra = o.pe_or_gamma_lut8(1.0, f, t, 1.4, T.ptr(a)); rb = r.ref_gamma_lut8(1.0, f, t, T.ptr(b))
a16 = np.zeros(65536, np.uint16); b16 = np.zeros(65536, np.uint16)
o.pe_or_gamma_lut16(1.0, f, t, 1.4, T.ptr(a16)); r.ref_gamma_lut16(1.0, f, t, T.ptr(b16))
assert ra == rb == 0
assert (a == b).all() and (a16 == b16).all()
print("OK")
"""


@pytest.mark.parametrize("pair", [(f, t) for f in (-1, 1, 2, 1024) for t in (-1, 1, 2, 1024) if f != t])
def test_gamma_luts_match_reference(pair):
    # the reference caches LUTs under a key it mutates (colourspace.c:701,721-733): fresh process per pair
    code = _GAMMA_CHILD % os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-c", code, str(pair[0]), str(pair[1])], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr


def test_gamma_lut_known_quirks():
    """* -> LINEAR is the identity table; LINEAR -> sRGB is off the textbook curve"""
    o = T.oracle()
    a = np.zeros(256, np.uint8)
    assert o.pe_or_gamma_lut8(1.0, T.G_SRGB, T.G_LINEAR, 1.4, T.ptr(a)) == 0
    assert (a == np.arange(256)).all()
    assert o.pe_or_gamma_lut8(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(a)) == 0
    assert a[128] == 181 and a[255] == 246
    assert o.pe_or_gamma_lut8(1.0, T.G_SRGB, T.G_SRGB, 1.4, T.ptr(a)) == -1


@pytest.mark.parametrize("quality", [T.Q_HIGH, T.Q_MED])
def test_pixel_kernels_exhaustive(quality):
    """all 2^24 triples x {clamped,unclamped} x {YCbCr,BT.709}: rgb2yuv :2119, yuv2rgb :2345"""
    o, r = T.oracle(), T.ref()
    g = np.arange(256, dtype=np.uint8)
    trip = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).copy()
    n = len(trip)
    r.ref_set_prefs(1, quality, 1.4)
    for cl, sub in itertools.product((0, 1), (1, 2)):
        a = np.zeros_like(trip)
        b = np.zeros_like(trip)
        o.pe_or_yuv2rgb(cl, sub, quality, T.ptr(trip), T.ptr(a), n)
        r.ref_yuv2rgb_bulk(cl, sub, T.ptr(trip), T.ptr(b), n)
        assert (a == b).all()
        o.pe_or_rgb2yuv(cl, sub, quality, T.ptr(trip), T.ptr(a), n)
        r.ref_rgb2yuv_bulk(cl, sub, T.ptr(trip), T.ptr(b), n)
        assert (a == b).all()
        # HIGH (float32 divide) and MED (>>16) agree after the clamps
        o.pe_or_yuv2rgb(cl, sub, T.Q_HIGH, T.ptr(trip), T.ptr(a), n)
        o.pe_or_yuv2rgb(cl, sub, T.Q_MED, T.ptr(trip), T.ptr(b), n)
        assert (a == b).all()
    r.ref_set_prefs(1, T.Q_HIGH, 1.4)


def test_chroma_weight_integer_form():
    """(int)(n/3. + .5) == (2n+3)/6 for every reachable n (colourspace.c:3465), used by the CUDA kernels"""
    for n in range(0, 766):
        assert int(n / 3.0 + 0.5) == (2 * n + 3) // 6


def _run_planar(o, r, rng, w, h, order, add_alpha, is422, cl, sub, q, tgt_gamma=0):
    r.ref_set_prefs(1, q, 1.4)
    y, u, v = T.make_yuv_planar(rng, w, h, is422, cl == 0)
    ps = 4 if (add_alpha or order == 2) else 3
    ors = T.rowstride(w, ps)
    a = np.full((h, ors), 7, np.uint8)
    b = np.full((h + 16, ors), 7, np.uint8)
    lut = None
    if tgt_gamma:
        lut = np.zeros(65536, np.uint16)
        assert o.pe_or_gamma_lut16(1.0, T.G_SRGB, tgt_gamma, 1.4, T.ptr(lut)) == 0
    pl, st = T.planes_arg(y, u, v), T.strides_arg(y, u, v)
    o.pe_or_yuv420p_to_rgb(pl, st, w, h, T.ptr(a), ors, order, add_alpha, is422, cl, sub, q, 1, T.ptr(lut))
    bb = b[8:]  # slack rows: the reference has stray writes (colourspace.c:3584,3704)
    r.ref_yuv420p_to_rgb(pl, w, h, st, ors, T.ptr(bb), order, add_alpha, is422, 0, cl, sub,
                         T.G_SRGB if tgt_gamma else 0, tgt_gamma)
    r.ref_set_prefs(1, T.Q_HIGH, 1.4)
    return a, bb[:h], ps, b


@pytest.mark.parametrize("size", [(64, 48), (130, 34), (640, 360)])
@pytest.mark.parametrize("order", [0, 1, 2])
def test_planar_420_422_to_rgb_matches_reference(size, order):
    """convert_yuv420p_to_{rgb,bgr,argb}_frame colourspace.c:3260,3927,4527 with nfx_threads = 1.
    Compared on every row the reference defines: 4:2:2 -> all rows,
    4:2:0 -> interior rows 1..h-2 (rows 0 and h-1 are the X rows)."""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(2)
    w, h = size
    for add_alpha, is422, cl, sub, q in itertools.product((0, 1), (0, 1), (0, 1), (1, 2), (T.Q_HIGH, T.Q_LOW)):
        if order == 2 and is422:
            continue  # reference ARGB 4:2:2 branch never resets `or` (colourspace.c:4853) and runs off the frame
        a, b, ps, full = _run_planar(o, r, rng, w, h, order, add_alpha, is422, cl, sub, q)
        wb = w * ps
        if is422:
            assert (a[:, :wb] == b[:, :wb]).all(), (add_alpha, cl, sub, q)
        elif cl == 1 and order == 0:
            # unclamped RGB variant writes interior rows one byte early (`or = orowstride * i - y_delta`, :3704)
            ors = a.shape[1]
            fa, fb = a.reshape(-1), full.reshape(-1)[8 * ors:]
            for row in range(1, h - 1):
                lo = 1 if row == 1 else 0  # byte ors-1 of row 0 is later clobbered by the last-row slip (:3584)
                assert (fa[row * ors + lo:row * ors + wb] == fb[row * ors - 1 + lo:row * ors - 1 + wb]).all()
        else:
            assert (a[1:h - 1, :wb] == b[1:h - 1, :wb]).all(), (add_alpha, cl, sub, q)


def test_planar_420_inline_gamma_matches_reference():
    """xyuv2rgb_with_gamma colourspace.c:2386 (16-bit LUT inside the converter)"""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(3)
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import pe_testlib as T
import test_oracle_vs_reference as M
o, r = T.oracle(), T.ref()
rng = np.random.default_rng(3)
a, b, ps, _ = M._run_planar(o, r, rng, 64, 48, 0, 1, 0, 0, 1, T.Q_HIGH, tgt_gamma=T.G_LINEAR)
assert (a[1:47, :64 * ps] == b[1:47, :64 * ps]).all()
a, b, ps, _ = M._run_planar(o, r, rng, 64, 48, 0, 0, 0, 0, 1, T.Q_HIGH, tgt_gamma=T.G_BT709)
assert (a[1:47, :64 * ps] == b[1:47, :64 * ps]).all()
print("OK")
""" % os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr + out.stdout


@pytest.mark.parametrize("fmt", [0, 1])
def test_packed422_to_rgb_matches_reference(fmt):
    """convert_{uyvy,yuyv}_to_{rgb,bgr,argb}_frame colourspace.c:6616-7103 (single band: the threaded uyvy path
    dereferences a zeroed table set, SURVEY.md finding 1)"""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(4)
    wm, h = 48, 20
    src = T.make_packed(rng, wm, h, 4)
    for order, add_alpha, cl, sub in itertools.product((0, 1, 2), (0, 1), (0, 1), (1, 2)):
        ps = 4 if (add_alpha or order == 2) else 3
        ors = T.rowstride(wm * 2, ps)
        a = np.zeros((h, ors), np.uint8)
        b = np.zeros((h, ors), np.uint8)
        o.pe_or_packed422_to_rgb(fmt, T.ptr(src), src.strides[0], wm, h, T.ptr(a), ors, order, add_alpha, cl, sub, T.Q_HIGH)
        r.ref_packed422_to_rgb(fmt, T.ptr(src), wm, h, src.strides[0], ors, T.ptr(b), order, add_alpha, cl, sub)
        assert (a == b).all(), (order, add_alpha, cl, sub)


def test_yuv888_and_back_matches_reference():
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(6)
    w, h = 50, 12
    for order, in_alpha, out_alpha, cl, sub in itertools.product((0, 1, 2), (0, 1), (0, 1), (0, 1), (1, 2)):
        ips, ops = (4 if in_alpha else 3), (4 if (out_alpha or order == 2) else 3)
        src = T.make_packed(rng, w, h, ips)
        ors = T.rowstride(w, ops)
        a = np.zeros((h, ors), np.uint8)
        b = np.zeros((h, ors), np.uint8)
        if order == 2 and not out_alpha:
            continue
        o.pe_or_yuv888_to_rgb(T.ptr(src), src.strides[0], w, h, T.ptr(a), ors, order, in_alpha, out_alpha, cl, sub, T.Q_HIGH)
        r.ref_yuv888_to_rgb(T.ptr(src), w, h, src.strides[0], ors, T.ptr(b), order, in_alpha, out_alpha, cl, sub)
        if in_alpha and order == 2:
            continue  # reference yuva8888->argb leaves alpha position inconsistent; colour path covered above
        assert (a[:, :w * ops] == b[:, :w * ops]).all(), ("yuv888->rgb", order, in_alpha, out_alpha, cl, sub)
    for order, in_alpha, out_alpha, cl in itertools.product((0, 1), (0, 1), (0, 1), (0, 1)):
        ips, ops = (4 if in_alpha else 3), (4 if out_alpha else 3)
        src = T.make_packed(rng, w, h, ips)
        ors = T.rowstride(w, ops)
        a = np.zeros((h, ors), np.uint8)
        b = np.zeros((h, ors), np.uint8)
        o.pe_or_rgb_to_yuv888(T.ptr(src), src.strides[0], w, h, T.ptr(a), ors, order, in_alpha, out_alpha, cl, T.Q_HIGH)
        r.ref_rgb_to_yuv888(T.ptr(src), w, h, src.strides[0], ors, T.ptr(b), order, in_alpha, out_alpha, cl)
        assert (a == b).all(), ("rgb->yuv888", order, in_alpha, out_alpha, cl)


# (reference op code, in palette, out palette, width unit the *worker* loop expects, alpha_first flag)
_PERMS = [
    (0, "RGB24", "BGR24", "bytes", 0), (1, "BGRA32", "ARGB32", "bytes", 0), (1, "ARGB32", "BGRA32", "bytes", 1),
    (2, "RGB24", "BGRA32", "px", 0), (3, "BGR24", "ARGB32", "px", 0), (4, "BGRA32", "RGB24", "px", 0),
    (5, "ARGB32", "BGR24", "px", 0), (6, "RGB24", "ARGB32", "px", 0), (7, "RGB24", "RGBA32", "px", 0),
    (8, "ARGB32", "RGB24", "px", 0), (9, "RGBA32", "RGB24", "px", 0), (10, "BGRA32", "RGBA32", "px", 0),
]
# Broken in the reference snapshot: convert_delpre_frame without LUT copies the first
# pixel of every row (no `+ i`, colourspace.c:10262); _convert_swapprepost_frame indexes uint64 words with byte
# rowstrides when there is no LUT (:10458-10463, heap corruption) and emits A,G,B,R with one (:10467-10488).
# For those our contract is the evident intent (the palette's byte order) and only self-consistency is tested.
_PERMS_REF_BROKEN = {(8, False)}


@pytest.mark.parametrize("perm", _PERMS)
@pytest.mark.parametrize("use_lut", [False, True])
def test_rgb_permutations_match_reference(perm, use_lut):
    """RGB<->RGB loops colourspace.c:9259-10515, called as a band worker (thread_id 0) with the width unit the
    threaded dispatcher passes (e.g. hsize = width*3 for swap3, :9279) == the intended whole-row semantics"""
    op, ip, opal, unit, alpha_first = perm
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(7)
    w, h = 40, 9
    if (op, use_lut) in _PERMS_REF_BROKEN:
        pytest.skip("reference loop is broken for this variant (see comment above)")
    ipal, opl = T.PAL[ip], T.PAL[opal]
    ips, ops = T.psize_of(ipal), T.psize_of(opl)
    src = T.make_packed(rng, w, h, ips)
    ors = T.rowstride(w, ops)
    a = np.zeros((h, ors), np.uint8)
    b = np.zeros((h, ors), np.uint8)
    lut = None
    if use_lut:
        lut = np.zeros(256, np.uint8)
        o.pe_or_gamma_lut8(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut))
    assert o.pe_or_rgb_to_rgb(ipal, opl, T.ptr(src), src.strides[0], w, h, T.ptr(a), ors, T.ptr(lut)) == 0
    wref = w * ips if unit == "bytes" else w
    r.ref_rgb_permute(op, T.ptr(src), wref, h, src.strides[0], ors, T.ptr(b), T.ptr(lut), alpha_first, 0)
    if op == 0 and not use_lut:
        # no-LUT swap3 never writes the middle byte (colourspace.c:9313-9316): fine in place, compare R/B only
        m = np.ones(ors, bool)
        m[1::3] = False
        m[w * 3:] = False
        assert (a[:, m] == b[:, m]).all()
    else:
        assert (a[:, :w * ops] == b[:, :w * ops]).all()


def test_rgb_to_rgb_all_pairs_self_consistent():
    """every (in, out) pair of the five RGB palettes: out == the in pixel re-ordered, alpha kept or 255"""
    o = T.oracle()
    rng = np.random.default_rng(11)
    w, h = 23, 5
    lay = {1: (0, 1, 2, -1), 2: (2, 1, 0, -1), 3: (0, 1, 2, 3), 4: (2, 1, 0, 3), 5: (1, 2, 3, 0)}
    for ipal, opal in itertools.product(lay, lay):
        ips, ops = T.psize_of(ipal), T.psize_of(opal)
        src = T.make_packed(rng, w, h, ips)
        dst = np.zeros((h, T.rowstride(w, ops)), np.uint8)
        assert o.pe_or_rgb_to_rgb(ipal, opal, T.ptr(src), src.strides[0], w, h, T.ptr(dst), dst.strides[0], None) == 0
        s = src[:, :w * ips].reshape(h, w, ips)
        d = dst[:, :w * ops].reshape(h, w, ops)
        for k in range(3):
            assert (d[..., lay[opal][k]] == s[..., lay[ipal][k]]).all()
        if lay[opal][3] >= 0:
            exp = s[..., lay[ipal][3]] if lay[ipal][3] >= 0 else 255
            assert (d[..., lay[opal][3]] == exp).all()


def test_swap3_single_thread_width_bug_documented():
    """SURVEY.md finding 1: with nfx_threads == 1 convert_layer_palette_full passes width in PIXELS to a loop that
    steps in bytes (colourspace.c:9311 vs :12381): only the first third of each row is swapped.  Our contract is
    the threaded (whole-row) behaviour; this test documents the difference."""
    r = T.ref()
    rng = np.random.default_rng(1)
    w, h = 640, 4
    src = T.make_packed(rng, w, h, 3)
    work = src.copy()
    r.ref_set_prefs(1, T.Q_HIGH, 1.4)
    r.ref_rgb_permute(0, T.ptr(work), w, h, work.strides[0], work.strides[0], T.ptr(work), None, 0, -1)
    third = (w // 3) * 3  # bytes [0, ~w) processed: pixels whose first byte index < w
    npx = (w + 2) // 3
    sw = src[:, :w * 3].reshape(h, w, 3)
    ww = work[:, :w * 3].reshape(h, w, 3)
    assert (ww[:, :npx, 0] == sw[:, :npx, 2]).all() and (ww[:, :npx, 2] == sw[:, :npx, 0]).all()
    assert (ww[:, npx:] == sw[:, npx:]).all()


def test_premult_planar_yuva4444p_matches_reference():
    """alpha_premult on YUVA4444P ("special case - planar with alpha", colourspace.c:12001-12049): padded planes, both clampings and
    directions; the planar branch returns before the flag update of :12100-12104"""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(4444)
    for (w, h), cl, direction in itertools.product(((45, 9), (64, 3)), (T.CLAMPED, T.UNCLAMPED), (1, -1)):
        st = T.rowstride(w, 1)
        a = [np.zeros((h, st), np.uint8) for _ in range(4)]
        for p in a:
            p[:, :w] = rng.integers(0, 256, (h, w), dtype=np.uint8)
        b = [p.copy() for p in a]
        flags = C.c_int(0)
        o.pe_or_alpha_premult_planar(T.planes_arg(*a), T.strides_arg(*a), cl, w, h, direction)
        r.ref_alpha_premult_planar(T.planes_arg(*b), T.strides_arg(*b), w, h, cl, direction, C.byref(flags))
        for k in range(4):
            assert (a[k] == b[k]).all(), (w, h, cl, direction, k)
        assert flags.value == 0


def test_gamma_apply_and_premult_match_reference():
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(8)
    lut = np.zeros(256, np.uint8)
    o.pe_or_gamma_lut8(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut))
    w, h = 37, 11
    for pal in (1, 2, 3, 4, 5):
        ps = T.psize_of(pal)
        a = T.make_packed(rng, w, h, ps)
        b = a.copy()
        x, y, sw, sh = 3, 2, 20, 7
        o.pe_or_gamma_apply(T.ptr(a), a.strides[0], pal, x, y, sw, sh, T.ptr(lut))
        r.ref_gamma_apply(b[y:].ctypes.data, b.strides[0], ps, x, sw, sh, 1 if pal == 5 else 0, T.ptr(lut))
        assert (a == b).all(), pal
    for pal, cl, direction in itertools.product((3, 4, 5, 589), (0, 1), (1, -1)):
        a = T.make_packed(rng, w, h, 4)
        b = a.copy()
        flags = C.c_int(0)
        o.pe_or_alpha_premult(T.ptr(a), a.strides[0], pal, cl, w, h, direction)
        r.ref_alpha_premult(T.ptr(b), w, h, b.strides[0], pal, cl, direction, C.addressof(flags))
        assert (a == b).all(), (pal, cl, direction)
        assert flags.value == (1 if direction == 1 else 0)  # WEED_LAYER_ALPHA_PREMULT set on FORWARD (:12102)


# ------------------------------------------------------------------ effect plugins via the real bootstrap

def _minihost():
    mh = C.CDLL(os.path.join(T.REF_DIR, "libweed_minihost.so"))
    mh.mh_open.argtypes = [C.c_char_p]
    mh.mh_run2.argtypes = [T.I, T.I, T.I, T.I, T.I, T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.I]
    mh.mh_run2v.argtypes = [T.I, T.I, T.I, T.I, T.I, T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.VP, T.I]
    return mh


def slide_over_params(transval, direction, mvlower, mvupper):
    """the 8 in-parameters of slide_over.c:164-172 that make sover_init :38-52 pick `direction` (1 .. 4)"""
    vals = [transval, 0, int(direction == 1), int(direction == 2), int(direction == 3), 0, mvlower, mvupper]
    return (C.c_int * 8)(*vals)


def test_simple_blend_matches_real_plugin():
    """simple_blend.c loaded through dlopen + weed_setup(weed_bootstrap) exactly as LiVES does"""
    o, mh = T.oracle(), _minihost()
    h = mh.mh_open(os.path.join(T.REF_DIR, "simple_blend.so").encode())
    assert h >= 0 and mh.mh_num_filters(h) == 5
    buf = C.create_string_buffer(64)
    mh.mh_filter_name(h, 0, buf, 64)
    assert buf.value == b"chroma blend" and mh.mh_filter_flags(h, 0) == 0x4C
    rng = np.random.default_rng(5)
    for pal, bf, (w, ht) in itertools.product((1, 2, 3, 4, 5), (0, 1, 100, 128, 255), ((64, 32), (61, 7))):
        ps = T.psize_of(pal)
        s1 = T.make_packed(rng, w, ht, ps)
        s2 = T.make_packed(rng, w, ht, ps)
        if ps == 4:
            al = s2[:, 3::4]
            al[rng.random(al.shape) < 0.4] = 255
        for typ in (0, 1, 2, 3, 4):   # 4 "averaged luma overlay": its averaging branch is dead code (:153-169), == type 1
            d_ref = np.full_like(s1, 9)
            d_or = np.full_like(s1, 9)
            assert mh.mh_run2(h, typ, pal, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_ref),
                              d_ref.strides[0], bf, 1) == 0
            o.pe_or_simple_blend(typ, pal, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_or),
                                 d_or.strides[0], w, ht, bf, s2.size)
            if pal == 5 and typ == 0 and s2.strides[0] == w * 4:
                # ARGB: the plugin tests the NEXT pixel's alpha byte (start = 1, :80,:130); for the last pixel of
                # the frame that byte is out of bounds -> excluded
                d_ref[-1, w * 4 - 3:] = d_or[-1, w * 4 - 3:]
            assert (d_ref == d_or).all(), (pal, bf, typ, w, ht)
        # in-place (what LiVES does for CAN_DO_INPLACE when the filter can't thread, effects-weed.c:2304-2314)
        d_ref = s1.copy()
        d_or = s1.copy()
        mh.mh_run2(h, 0, pal, w, ht, T.ptr(d_ref), d_ref.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_ref),
                   d_ref.strides[0], bf, 1)
        o.pe_or_simple_blend(0, pal, T.ptr(d_or), d_or.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_or),
                             d_or.strides[0], w, ht, bf, s2.size)
        if pal == 5 and s2.strides[0] == w * 4:
            d_ref[-1, w * 4 - 3:] = d_or[-1, w * 4 - 3:]
        assert (d_ref == d_or).all()


def test_multi_blends_matches_real_plugin():
    o, mh = T.oracle(), _minihost()
    h = mh.mh_open(os.path.join(T.REF_DIR, "multi_blends.so").encode())
    assert h >= 0 and mh.mh_num_filters(h) == 7
    rng = np.random.default_rng(9)
    w, ht = 53, 12
    for pal, bf, typ in itertools.product((1, 2), (0, 17, 127, 128, 200, 255), range(7)):
        s1 = T.make_packed(rng, w, ht, 3)
        s2 = T.make_packed(rng, w, ht, 3)
        d_ref = np.full_like(s1, 9)
        d_or = np.full_like(s1, 9)
        assert mh.mh_run2(h, typ, pal, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_ref),
                          d_ref.strides[0], bf, 1) == 0
        o.pe_or_multi_blend(typ, pal, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d_or),
                            d_or.strides[0], w, ht, bf)
        assert (d_ref == d_or).all(), (pal, bf, typ)


def test_alpha_over_matches_paint_pixel():
    """compositor.c paint_pixel :120 -- every (dst, src) byte pair for a set of alphas"""
    o, p = T.oracle(), T.ref_paint()
    g = np.arange(256, dtype=np.uint8)
    d0, s0 = np.meshgrid(g, g, indexing="ij")
    n = 65536 // 1
    for alpha in (0.0, 0.1, 0.25, 1.0 / 3.0, 0.5, 0.7, 0.9, 0.999, 1.0):
        dst = np.repeat(d0.reshape(-1, 1), 3, axis=1).astype(np.uint8).copy()
        src = np.repeat(s0.reshape(-1, 1), 3, axis=1).astype(np.uint8).copy()
        a = dst.copy()
        b = dst.copy()
        o.pe_or_alpha_over(T.ptr(a), n * 3, T.ptr(src), n * 3, 1, n, 1, alpha)
        p.ref_paint_rows(T.ptr(b), T.ptr(src), n, 3, alpha)
        assert (a == b).all(), alpha


def test_rgb_to_packed422_and_planar444_match_reference():
    """convert_{rgb,bgr,argb}_to_{uyvy,yuyv}_frame :5129-5700 (rows unpadded: the reference's row advance only works then)
    and convert_{rgb,bgr,argb}_to_yuvp_frame :5786-6240"""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(12)
    w, h = 48, 10  # 48 px: 96-byte macropixel rows and 144 / 192-byte RGB rows are multiples of 32 -> no padding
    r.ref_set_prefs(1, T.Q_HIGH, 1.4)
    for fmt, order, in_alpha, cl in itertools.product((0, 1), (0, 1, 2), (0, 1), (0, 1)):
        if order == 2 and not in_alpha:
            continue
        ips = 4 if in_alpha else 3
        src = T.make_packed(rng, w, h, ips)
        a = np.zeros((h, w * 2), np.uint8)
        b = np.zeros((h, w * 2), np.uint8)
        o.pe_or_rgb_to_packed422(fmt, T.ptr(src), src.strides[0], w, h, T.ptr(a), a.strides[0], order, in_alpha, cl, T.Q_HIGH, None)
        r.ref_rgb_to_packed422(fmt, T.ptr(src), w, h, src.strides[0], b.strides[0], T.ptr(b), order, in_alpha, cl, 0, 0)
        assert (a == b).all(), ("packed422", fmt, order, in_alpha, cl)
    for order, in_alpha, out_alpha, cl in itertools.product((0, 1, 2), (0, 1), (0, 1), (0, 1)):
        if order == 2 and not in_alpha:
            continue
        ips = 4 if in_alpha else 3
        src = T.make_packed(rng, w, h, ips)
        ors = T.rowstride(w, 1)
        pa = [np.zeros((h, ors), np.uint8) for _ in range(4)]
        pb = [np.zeros((h, ors), np.uint8) for _ in range(4)]
        o.pe_or_rgb_to_yuv444p(T.ptr(src), src.strides[0], w, h, T.planes_arg(*pa), ors, order, in_alpha, out_alpha, cl, T.Q_HIGH)
        r.ref_rgb_to_yuv444p(T.ptr(src), w, h, src.strides[0], ors, T.planes_arg(*pb), order, in_alpha, out_alpha, cl)
        for k in range(4 if out_alpha else 3):
            if k == 3 and order == 2:
                continue  # the ARGB variant has no in_has_alpha argument and copies byte 3 (blue) as alpha (:6220): X
            assert (pa[k] == pb[k]).all(), ("yuv444p", order, in_alpha, out_alpha, cl, k)


def test_rgb_to_packed422_gamma_lut_variant():
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import pe_testlib as T
o, r = T.oracle(), T.ref()
rng = np.random.default_rng(13)
w, h = 48, 6
src = T.make_packed(rng, w, h, 3)
lut = np.zeros(65536, np.uint16)
assert o.pe_or_gamma_lut16(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut)) == 0
for fmt in (0, 1):
    a = np.zeros((h, w * 2), np.uint8); b = np.zeros((h, w * 2), np.uint8)
    o.pe_or_rgb_to_packed422(fmt, T.ptr(src), src.strides[0], w, h, T.ptr(a), a.strides[0], 0, 0, 0, T.Q_HIGH, T.ptr(lut))
    r.ref_rgb_to_packed422(fmt, T.ptr(src), w, h, src.strides[0], b.strides[0], T.ptr(b), 0, 0, 0, T.G_LINEAR, T.G_SRGB)
    assert (a == b).all(), fmt
print("OK")
""" % os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr + out.stdout


def test_avg_tables_and_rgb_to_planar420_422_match_reference():
    """init_average :190 (cavgc / cavgu) and convert_{rgb,bgr}_to_yuv420_frame :6250 / :6385, 4:2:0 and 4:2:2, on unpadded planes
    (the reference advances its plane pointers densely)"""
    o, r = T.oracle(), T.ref()
    for which in (0, 1):
        a, b = np.zeros(65536, np.uint8), np.zeros(65536, np.uint8)
        o.pe_or_avg_table(which, T.ptr(a))
        assert r.ref_get_avg_table(which, T.ptr(b)) == 0
        assert (a == b).all(), which
    rng = np.random.default_rng(14)
    r.ref_set_prefs(1, T.Q_HIGH, 1.4)
    for (w, h), order, in_alpha, is422, cl, sub in itertools.product(((64, 10), (32, 2), (96, 36)), (0, 1), (0, 1), (0, 1), (0, 1), (1, 2)):
        ips = 4 if in_alpha else 3
        src = T.make_packed(rng, w, h, ips)
        ch = h if is422 else h // 2
        pa = [np.full((h, w), 7, np.uint8), np.full((ch, w // 2), 7, np.uint8), np.full((ch, w // 2), 7, np.uint8)]
        pb = [p.copy() for p in pa]
        strides = (C.c_int * 3)(w, w // 2, w // 2)
        o.pe_or_rgb_to_yuv420p(T.ptr(src), src.strides[0], w, h, T.planes_arg(*pa), strides, order, in_alpha, is422, cl, sub, T.Q_HIGH)
        r.ref_rgb_to_yuv420(T.ptr(src), w, h, src.strides[0], strides, T.planes_arg(*pb), order, is422, in_alpha, sub, cl)
        for k in range(3):
            assert (pa[k] == pb[k]).all(), ("yuv420p", w, h, order, in_alpha, is422, cl, sub, k)


# ---- YUV <-> YUV family ------------------------------------------------------------------------------

def _planes444(rng, w, h, n, stride=None):
    stride = stride or T.rowstride(w, 1)
    out = []
    for _ in range(n):
        a = np.zeros((h, stride), np.uint8)
        a[:, :w] = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out.append(a)
    return out


def test_yuv444p_to_rgb_matches_reference():
    """convert_yuv_planar_to_{rgb,bgr}_frame: orders x alpha in / out x clamping.  Excluded (X): BGR24 -- the reference walks
    4 bytes per BGR24 pixel (colourspace.c:7313) and runs past its buffer; ARGB32 -- the row advance subtracts the width from
    the OUTPUT stride twice and never from the input stride (:7471-7472), rows after the first are read and written askew"""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(70)
    for (w, h), order, in_alpha, out_alpha, cl in itertools.product(((64, 6), (37, 5)), (0, 1, 2), (0, 1), (0, 1), (0, 1)):
        if order == 2 or (order == 1 and not out_alpha):
            continue  # X (see docstring)
        pl = _planes444(rng, w, h, 4)
        ps = 4 if out_alpha else 3
        ors = T.rowstride(w, ps)
        a, b = np.zeros((h, ors), np.uint8), np.zeros((h, ors), np.uint8)
        o.pe_or_yuv444p_to_rgb(T.planes_arg(*pl), pl[0].strides[0], w, h, T.ptr(a), ors, order, in_alpha, out_alpha, cl, T.Q_HIGH)
        r.ref_yuv444p_to_rgb(T.planes_arg(*pl), w, h, pl[0].strides[0], ors, T.ptr(b), order, in_alpha, out_alpha, cl)
        assert (a == b).all(), (w, h, order, in_alpha, out_alpha, cl)


def test_combine_and_split_planes_match_reference():
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(71)
    for (w, h, padded), ia, oa in itertools.product(((64, 5, False), (37, 4, True), (33, 3, False)), (0, 1), (0, 1)):
        pl = _planes444(rng, w, h, 4, stride=None if padded else w)
        ops = 4 if oa else 3
        ors = T.rowstride(w, ops) if padded else w * ops
        a, b = np.zeros((h, ors), np.uint8), np.zeros((h, ors), np.uint8)
        o.pe_or_combine_planes(T.planes_arg(*pl), pl[0].strides[0], w, h, T.ptr(a), ors, ia, oa)
        if padded and ia and oa:
            # X: on padded planes the reference never advances its alpha pointer by the row padding (colourspace.c:7625-7637)
            for k in range(4):
                assert (a[:, :w * 4].reshape(h, w, 4)[:, :, k] == pl[k][:, :w]).all()
        else:
            r.ref_combineplanes(T.planes_arg(*pl), w, h, pl[0].strides[0], ors, T.ptr(b), ia, oa)
            assert (a == b).all(), ("combine", w, h, padded, ia, oa)
        # and back: packed (ia = source alpha) -> planes (oa = destination alpha)
        ips = 4 if ia else 3
        irs = T.rowstride(w, ips) if padded else w * ips
        src = T.make_packed(rng, w, h, ips, stride=irs)
        st = T.rowstride(w, 1) if padded else w
        pa = [np.zeros((h, st), np.uint8) for _ in range(4)]
        pb = [np.zeros((h, st), np.uint8) for _ in range(4)]
        o.pe_or_split_planes(T.ptr(src), irs, w, h, T.planes_arg(*pa), T.strides_arg(*pa), ia, oa)
        if oa or ia:
            # X: with a destination alpha plane the reference advances that plane by (stride - width * ipsize) per row
            # (colourspace.c:9233-9236: the width has already been multiplied) and writes before the buffer; with a source
            # alpha and no destination alpha it never skips the source's 4th byte (:9238-9246).  The restatement is the
            # evident intent, checked against numpy here
            px = src[:, :w * ips].reshape(h, w, ips)
            for k in range(3):
                assert (pa[k][:, :w] == px[:, :, k]).all()
            if oa:
                assert (pa[3][:, :w] == (px[:, :, 3] if ia else 255)).all()
            continue
        r.ref_splitplanes(T.ptr(src), w, h, irs, T.strides_arg(*pb), T.planes_arg(*pb), ia, oa)
        for k in range(3):
            assert (pa[k] == pb[k]).all(), ("split", w, h, padded, ia, oa, k)


def test_halve_and_double_chroma_match_reference():
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(72)
    for (cw, ch), cl in itertools.product(((32, 8), (33, 7), (5, 2), (16, 1)), (0, 1)):
        st = T.align_ceil(cw, 16)
        src = [np.zeros((ch, st), np.uint8) for _ in range(3)]
        for p in src[1:]:
            p[:, :cw] = rng.integers(0, 256, (ch, cw), dtype=np.uint8)
        # halve: (ch + 1) // 2 rows out
        da = [np.zeros(((ch + 1) // 2, st), np.uint8) for _ in range(3)]
        db = [np.zeros(((ch + 1) // 2, st), np.uint8) for _ in range(3)]
        o.pe_or_halve_chroma(T.planes_arg(*src), T.strides_arg(*src), cw, ch, T.planes_arg(*da), T.strides_arg(*da), cl)
        r.ref_halve_chroma(T.planes_arg(*src), cw, ch, T.strides_arg(*src), T.strides_arg(*db), T.planes_arg(*db), cl)
        for k in (1, 2):
            assert (da[k] == db[k]).all(), ("halve", cw, ch, cl, k)
        # double: 2 ch rows out
        da = [np.zeros((2 * ch, st), np.uint8) for _ in range(3)]
        db = [np.zeros((2 * ch, st), np.uint8) for _ in range(3)]
        o.pe_or_double_chroma(T.planes_arg(*src), T.strides_arg(*src), cw, ch, T.planes_arg(*da), T.strides_arg(*da), cl)
        r.ref_double_chroma(T.planes_arg(*src), cw, ch, T.strides_arg(*src), T.strides_arg(*db), T.planes_arg(*db), cl)
        for k in (1, 2):
            assert (da[k] == db[k]).all(), ("double", cw, ch, cl, k)


def test_packed422_to_yuv_planar_and_yuv888_match_reference():
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(73)
    for (wm, h), fmt in itertools.product(((32, 4), (17, 3)), (0, 1)):
        irs = T.rowstride(wm, 4)
        src = T.make_packed(rng, wm, h, 4, stride=irs)
        # -> planar 4:2:2: the reference walks everything densely, so compare on unpadded buffers; its source pointer never
        # advances (colourspace.c:8103): every sample is the first macropixel's -- quirks = 1
        dense = np.ascontiguousarray(src[:, :wm * 4])
        pa = [np.zeros((h, 2 * wm), np.uint8), np.zeros((h, wm), np.uint8), np.zeros((h, wm), np.uint8)]
        pb = [np.zeros_like(p) for p in pa]
        o.pe_or_packed422_to_yuv422p(fmt, T.ptr(dense), wm * 4, wm, h, T.planes_arg(*pa), T.strides_arg(*pa), 1)
        r.ref_packed422_to_yuv422p(fmt, T.ptr(dense), wm, h, T.planes_arg(*pb))
        for k in range(3):
            assert (pa[k] == pb[k]).all(), ("422p", wm, h, fmt, k)
        # quirks = 0: the intended per-macropixel split
        o.pe_or_packed422_to_yuv422p(fmt, T.ptr(dense), wm * 4, wm, h, T.planes_arg(*pa), T.strides_arg(*pa), 0)
        yo, uo = (1, 0) if fmt == 0 else (0, 1)
        assert (pa[0] == dense[:, yo::2]).all() and (pa[1] == dense[:, uo::4]).all() and (pa[2] == dense[:, uo + 2::4]).all()
        # -> planar 4:4:4 (+ alpha)
        for aa in (0, 1):
            st = T.rowstride(2 * wm, 1)
            pa = [np.zeros((h, st), np.uint8) for _ in range(4)]
            pb = [np.zeros((h, st), np.uint8) for _ in range(4)]
            o.pe_or_packed422_to_yuv444p(fmt, T.ptr(src), irs, wm, h, T.planes_arg(*pa), T.strides_arg(*pa), aa)
            r.ref_packed422_to_yuv444p(fmt, T.ptr(src), wm, h, irs, T.strides_arg(*pb), T.planes_arg(*pb), aa)
            for k in range(4):
                assert (pa[k] == pb[k]).all(), ("444p", wm, h, fmt, aa, k)
            ors = T.rowstride(2 * wm, 4 if aa else 3)
            a, b = np.zeros((h, ors), np.uint8), np.zeros((h, ors), np.uint8)
            o.pe_or_packed422_to_yuv888(fmt, T.ptr(src), irs, wm, h, T.ptr(a), ors, aa)
            r.ref_packed422_to_yuv888(fmt, T.ptr(src), wm, h, irs, ors, T.ptr(b), aa)
            assert (a == b).all(), ("888", wm, h, fmt, aa)
        # UYVY <-> YUYV
        a, b = src.copy(), src.copy()
        o.pe_or_swab(T.ptr(a), irs, wm, h)
        r.ref_swab(T.ptr(b), wm, h, irs)
        assert (a == b).all(), ("swab", wm, h)


def test_clamping_tables_match_reference():
    """init_YUV_to_YUV_tables; the frame walk of switch_yuv_clamping_and_subspace is a LUT over every byte and is restated from
    the source (it needs a weed_layer_t, which the slices do not build)"""
    o, r = T.oracle(), T.ref()
    for w in range(4):
        a, b = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
        o.pe_or_yy_table(w, T.ptr(a))
        assert r.ref_get_yy_table(w, T.ptr(b)) == 0
        assert (a == b).all(), w
    rng = np.random.default_rng(74)
    ty, tc = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
    o.pe_or_yy_table(0, T.ptr(ty)); o.pe_or_yy_table(1, T.ptr(tc))
    buf = rng.integers(0, 256, 96, dtype=np.uint8)
    for kind, luma_mask in ((0, np.ones(96, bool)), (1, np.zeros(96, bool)), (2, np.arange(96) % 3 == 0), (4, np.arange(96) % 2 == 1),
                            (5, np.arange(96) % 2 == 0)):
        x = buf.copy()
        o.pe_or_switch_clamping_plane(T.ptr(x), 96, kind, 1)
        assert (x == np.where(luma_mask, ty[buf], tc[buf])).all(), kind
    x = buf.copy()
    o.pe_or_switch_clamping_plane(T.ptr(x), 96, 3, 1)
    i = np.arange(96)
    assert (x == np.where(i % 4 == 3, buf, np.where(i % 4 == 0, ty[buf], tc[buf]))).all()


def test_yuv444p_to_packed422_and_yuv420p_match_reference():
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(75)
    for (w, h), cl in itertools.product(((32, 6), (34, 5), (2, 2)), (0, 1)):
        # packed 4:2:2: the reference is only right on dense buffers (its strided branch runs `width` macropixels per row)
        pl = _planes444(rng, w, h, 3, stride=w)
        for fmt in (0, 1):
            a, b = np.zeros((h, 2 * w), np.uint8), np.zeros((h, 2 * w), np.uint8)
            o.pe_or_yuv444p_to_packed422(fmt, T.planes_arg(*pl), w, w, h, T.ptr(a), 2 * w, cl)
            r.ref_yuv444p_to_packed422(fmt, T.planes_arg(*pl), w, h, w, 2 * w, T.ptr(b), cl)
            assert (a == b).all(), ("packed422", w, h, fmt, cl)
        # planar 4:2:0, padded planes
        pl = _planes444(rng, w, h, 3)
        ys, cs = T.rowstride(w, 1), T.rowstride(w, 1) >> 1
        ch = (h + 1) >> 1
        da = [np.zeros((h, ys), np.uint8), np.zeros((ch, cs), np.uint8), np.zeros((ch, cs), np.uint8)]
        db = [np.zeros_like(p) for p in da]
        o.pe_or_yuv444p_to_yuv420p(T.planes_arg(*pl), T.strides_arg(*pl), w, h, T.planes_arg(*da), T.strides_arg(*da), cl)
        r.ref_yuv444p_to_yuv420p(T.planes_arg(*pl), w, h, T.strides_arg(*pl), T.strides_arg(*db), T.planes_arg(*db), cl)
        for k in range(3):
            assert (da[k][:, :w >> (k > 0)] == db[k][:, :w >> (k > 0)]).all(), ("420p", w, h, cl, k)


def test_planar42x_to_packed422_matches_reference():
    """4:2:0 -> UYVY / YUYV on unpadded chroma planes (where the reference's pointer rewind is right); 4:2:2 -> UYVY / YUYV on a
    one-row frame (the only row the reference gets right)"""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(76)
    for (w, h), fmt in itertools.product(((32, 6), (34, 4), (2, 2)), (0, 1)):
        # (the YUYV variant never skips the luma row padding, colourspace.c:7181-7195: unpadded luma for it)
        ys = T.rowstride(w, 1) if fmt == 0 else w
        y = np.zeros((h, ys), np.uint8); y[:, :w] = rng.integers(0, 256, (h, w), dtype=np.uint8)
        u = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
        v = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
        ors = T.rowstride(w // 2, 4)
        a, b = np.zeros((h, ors), np.uint8), np.zeros((h, ors), np.uint8)
        o.pe_or_yuv42xp_to_packed422(fmt, T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, 0, T.ptr(a), ors)
        r.ref_yuv420_to_packed422(fmt, T.planes_arg(y, u, v), w, h, T.strides_arg(y, u, v), ors, T.ptr(b), 0)
        assert (a == b).all(), ("420", w, h, fmt)
        y1 = rng.integers(0, 256, (1, w), dtype=np.uint8)
        u1, v1 = rng.integers(0, 256, (1, w // 2), dtype=np.uint8), rng.integers(0, 256, (1, w // 2), dtype=np.uint8)
        a, b = np.zeros((1, 2 * w), np.uint8), np.zeros((1, 2 * w), np.uint8)
        o.pe_or_yuv42xp_to_packed422(fmt, T.planes_arg(y1, u1, v1), T.strides_arg(y1, u1, v1), w, 1, 1, T.ptr(a), 2 * w)
        r.ref_yuv422p_to_packed422(fmt, T.planes_arg(y1, u1, v1), w // 2, 1, T.strides_arg(y1, u1, v1), 2 * w, T.ptr(b))
        assert (a == b).all(), ("422", w, fmt)


def test_quad_chroma_matches_reference():
    """4:2:0 -> 4:4:4 chroma (convert_quad_chroma) on padded planes, JPEG and MPEG sampling, even and odd heights; for an even
    height the reference never writes the last row (X: not compared)"""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(77)
    for (w, h), samp, cl, aa in itertools.product(((32, 8), (34, 7), (6, 2), (36, 9)), (0, 1), (0, 1), (0, 1)):
        cw, ch = w >> 1, (h + 1) >> 1
        cs, os_ = T.align_ceil(cw + 1, 16), T.align_ceil(w + 1, 32)
        src = [np.zeros((ch, cs), np.uint8) for _ in range(3)]
        for p in src[1:]:
            p[:, :cw] = rng.integers(0, 256, (ch, cw), dtype=np.uint8)
        da = [np.zeros((h, os_), np.uint8) for _ in range(4)]
        db = [np.zeros((h + 2, os_), np.uint8) for _ in range(4)]  # (slack rows: the reference's odd-row pass writes past `width`)
        o.pe_or_quad_chroma(T.planes_arg(*src), T.strides_arg(*src), w, h, T.planes_arg(*da), os_, aa, int(samp == 0), cl)
        r.ref_quad_chroma(T.planes_arg(*src), w, h, T.strides_arg(*src), os_, T.planes_arg(*db), aa, samp, cl)
        rows = h if (h & 1) else h - 1
        for k in (1, 2):
            assert (da[k][:rows, :w] == db[k][:rows, :w]).all(), ("quad", w, h, samp, cl, k)
        if aa:
            assert (da[3] == 255).all() and (db[3][:h] == 255).all()


def test_yuv888_subsample_matches_reference():
    """YUV888 / YUVA8888 -> UYVY / YUYV / YUV422P / YUV420P on unpadded buffers (the reference's strided branches are broken)"""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(78)
    for (w, h), sa, cl, mode in itertools.product(((32, 6), (34, 4), (2, 2)), (0, 1), (0, 1), (0, 1, 2, 3)):
        ips = 4 if sa else 3
        src = rng.integers(0, 256, (h, w * ips), dtype=np.uint8)
        if mode <= 1:
            da, db = [np.zeros((h, 2 * w), np.uint8)], [np.zeros((h, 2 * w), np.uint8)]
        else:
            ch = h if mode == 2 else h // 2
            da = [np.zeros((h, w), np.uint8), np.zeros((ch, w // 2), np.uint8), np.zeros((ch, w // 2), np.uint8)]
            db = [np.zeros_like(p) for p in da]
        pa, pb = da + [da[0]] * (3 - len(da)), db + [db[0]] * (3 - len(db))
        o.pe_or_yuv888_subsample(mode, T.ptr(src), src.strides[0], w, h, sa, T.planes_arg(*pa), T.strides_arg(*pa), cl)
        r.ref_yuv888_subsample(mode, T.ptr(src), w, h, src.strides[0], T.strides_arg(*pb), T.planes_arg(*pb), sa, cl)
        for k in range(len(da)):
            assert (da[k] == db[k]).all(), ("yuv888 subsample", w, h, sa, cl, mode, k)


def test_packed422_to_yuv420p_matches_reference():
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(79)
    for (wm, h), fmt, cl in itertools.product(((16, 6), (17, 4), (1, 2)), (0, 1), (0, 1)):
        src = rng.integers(0, 256, (h, wm * 4), dtype=np.uint8)
        da = [np.zeros((h, 2 * wm), np.uint8), np.zeros((h // 2, wm), np.uint8), np.zeros((h // 2, wm), np.uint8)]
        db = [np.zeros_like(p) for p in da]
        o.pe_or_packed422_to_yuv420p(fmt, T.ptr(src), src.strides[0], wm, h, T.planes_arg(*da), T.strides_arg(*da), cl)
        r.ref_packed422_to_yuv420p(fmt, T.ptr(src), wm, h, T.planes_arg(*db), cl)
        for k in range(3):
            assert (da[k] == db[k]).all(), ("packed422 -> 420p", wm, h, fmt, cl, k)


def test_chroma_upsample_packed_matches_reference():
    """4:2:2 / 4:2:0 planar -> YUV888 / YUVA8888 (convert_double_chroma_packed / convert_quad_chroma_packed) on padded planes; for
    4:2:0 the rows the reference leaves undefined (last row of an even height, last two of an odd height) are not compared"""
    o, r = T.oracle(), T.ref()
    rng = np.random.default_rng(80)
    for (w, h), is420, samp, cl, aa in itertools.product(((32, 8), (34, 7), (6, 2), (36, 9)), (0, 1), (0, 1), (0, 1), (0, 1)):
        cw, ch = w >> 1, ((h + 1) >> 1) if is420 else h
        ys, cs = T.align_ceil(w + 1, 32), T.align_ceil(cw + 1, 16)
        y = np.zeros((h, ys), np.uint8); y[:, :w] = rng.integers(0, 256, (h, w), dtype=np.uint8)
        u, v = np.zeros((ch, cs), np.uint8), np.zeros((ch, cs), np.uint8)
        u[:, :cw] = rng.integers(0, 256, (ch, cw), dtype=np.uint8); v[:, :cw] = rng.integers(0, 256, (ch, cw), dtype=np.uint8)
        ps = 4 if aa else 3
        ors = T.align_ceil(w * ps + 8, 32)
        a = np.zeros((h, ors), np.uint8)
        b = np.zeros((h + 2, ors), np.uint8)  # slack rows: the reference's post-loop touches the row past the frame
        o.pe_or_chroma_upsample_packed(is420, T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(a), ors, aa, int(samp == 0), cl)
        r.ref_chroma_upsample_packed(is420, T.planes_arg(y, u, v), w, h, T.strides_arg(y, u, v), ors, T.ptr(b), aa, samp, cl)
        rows = h if not is420 else (h - 2 if (h & 1) else h - 1)
        if aa and not is420:
            # X: convert_double_chroma_packed never writes the alpha byte of the second pixel of a pair (colourspace.c:10848-10863)
            assert (a[:, 3:w * 4:4] == 255).all()
            b[:h, 7:w * 4:8] = 255
        if aa and is420:
            # X: convert_quad_chroma_packed writes no alpha on odd rows (colourspace.c:10768-10784)
            assert (a[:, 3:w * 4:4] == 255).all()
            b[1:h:2, 3:w * 4:4] = 255
        assert (a[:rows, :w * ps] == b[:rows, :w * ps]).all(), ("upsample packed", w, h, is420, samp, cl, aa)
        # luma (and alpha) of every row
        assert (a[:, 0:w * ps:ps] == b[:h, 0:w * ps:ps]).all()


def test_slide_over_matches_real_plugin():
    """slide_over.c through the real bootstrap: every direction x moving clips, every transition value on small frames, all the
    packed pixel sizes; the dividing line for all 256 values at the benchmark sizes"""
    o, mh = T.oracle(), _minihost()
    h = mh.mh_open(os.path.join(T.REF_DIR, "slide_over.so").encode())
    assert h >= 0 and mh.mh_num_filters(h) == 1
    rng = np.random.default_rng(90)
    for (pal, ps), (w, ht) in itertools.product(((1, 3), (3, 4), (564, 4)), ((37, 19), (64, 32))):
        s1, s2 = T.make_packed(rng, w, ht, ps), T.make_packed(rng, w, ht, ps)
        for direction, mvl, mvu in itertools.product((1, 2, 3, 4), (0, 1), (0, 1)):
            for tv in (list(range(0, 256, 5)) + [1, 127, 128, 254, 255]):
                a, b = np.zeros_like(s1), np.zeros_like(s1)
                o.pe_or_slide_over(direction, tv, mvl, mvu, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(a), a.strides[0],
                                   w, ht, ps)
                assert mh.mh_run2v(h, 0, pal, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(b), b.strides[0], 8,
                                   slide_over_params(tv, direction, mvl, mvu), 1) == 0
                assert (a == b).all(), ("slide over", pal, w, ht, direction, mvl, mvu, tv)
    # the line itself at large sizes: rows / columns of the result that come from in1 (constant 1) and in2 (constant 2)
    for w, ht in ((3840, 2), (2, 2160), (1920, 1), (1, 1080), (1280, 1), (1, 720), (255, 1), (1, 255), (765, 1), (1, 510), (37, 1), (1, 333)):
        s1, s2 = np.full((ht, T.rowstride(w, 3)), 1, np.uint8), np.full((ht, T.rowstride(w, 3)), 2, np.uint8)
        for direction in ((1, 2) if ht <= 2 else (3, 4)):
            for tv in range(256):
                b = np.zeros_like(s1)
                assert mh.mh_run2v(h, 0, 1, w, ht, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(b), b.strides[0], 8,
                                   slide_over_params(tv, direction, 0, 0), 1) == 0
                line = b[0, 0:w * 3:3] if ht <= 2 else b[:, 0]
                first = 1 if direction in (1, 3) else 2
                assert int((line == first).sum()) == o.pe_or_slide_over_bound(direction, tv, w, ht), (w, ht, direction, tv)
