"""CPU checks of the integer identities the CUDA kernels rely on (each one is a claim in a kernel comment / DESIGN.md):
exhaustive over the value ranges that can occur, no GPU needed; the last test checks the (unpinned) resize contract of the
oracle against OpenCV."""
import numpy as np


def test_third_round_forms():
    """(int)(n / 3. + .5) of colourspace.c:3465 == (2n+3)/6 == ((2n+3) * 43691) >> 18 for every reachable chroma sum"""
    n = np.arange(0, 766, dtype=np.int64)
    ref = (n / 3.0 + 0.5).astype(np.int64)
    assert (ref == (2 * n + 3) // 6).all()
    assert (ref == ((2 * n + 3) * 43691) >> 18).all()          # pe_device.cuh third_round
    assert ref.max() == 255


def test_packed_multiply_high_index():
    """k_fused3 idx_hi / idx_lo: Q = 2n+3 packed in 16-bit halves; hi32(Q * 10923 * 128) & 0x7F80 == 128 * third_round(n) for the
    high half whatever the low half holds, and hi32((Q << 16) * ...) for the low half"""
    K = 10923 * 128
    n = np.arange(0, 766, dtype=np.uint64)
    q = 2 * n + 3
    want = 128 * ((2 * n + 3) // 6)
    for lo in (0, 3, 777, 1533):
        packed = (q << np.uint64(16)) | np.uint64(lo)
        got = ((packed * np.uint64(K)) >> np.uint64(32)) & np.uint64(0x7F80)
        assert (got == want).all(), lo
    for hi in (0, 3, 1533):
        packed = ((np.uint64(hi) << np.uint64(16)) | q) & np.uint64(0xFFFFFFFF)
        shifted = (packed << np.uint64(16)) & np.uint64(0xFFFFFFFF)
        got = ((shifted * np.uint64(K)) >> np.uint64(32)) & np.uint64(0x7F80)
        assert (got == want).all(), hi


def test_packed_chroma_sums_do_not_carry():
    """Q = 2 s1 + (s2 & ~1) + 3 with s <= 510 stays below 2^16: the two halves of a packed register never interact"""
    assert 2 * 510 + 510 + 3 < 1 << 16
    s1, s2 = np.meshgrid(np.arange(0, 511), np.arange(0, 511))
    n_up = s1 + (s2 >> 1)
    assert (2 * n_up + 3 == 2 * s1 + (s2 & ~1) + 3).all()
    n_lo = (s1 >> 1) + s2
    assert (2 * n_lo + 3 == (s1 & ~1) + 2 * s2 + 3).all()


def test_integer_alpha_over_is_the_double_expression():
    """compositor.c:120 dst = (u8)(dst * (1 - a) + src * a) in double == (bg * (256 - k) + fg * k) >> 8 for a = k / 256"""
    bg, fg = np.meshgrid(np.arange(256), np.arange(256))
    for k in (0, 1, 64, 96, 128, 192, 255, 256):
        a = k / 256.0
        want = (bg * (1.0 - a) + fg * a).astype(np.uint8)
        got = ((bg * (256 - k) + fg * k) >> 8).astype(np.uint8)
        assert (want == got).all(), k


def test_single_pass_vertical_filter_equals_two_pass_contract():
    """with an identity horizontal pass tmp = pix * 128 and out = clip((sum c12 * tmp + 2^18) >> 19)
    == (sum c12 * pix + 2^11) >> 12, the form k_fused2 / k_fused3 evaluate with DP2A"""
    rng = np.random.default_rng(0)
    pix = rng.integers(0, 256, (20000, 4))
    c = rng.integers(0, 4097, (20000, 3))
    c = np.concatenate([c, np.zeros((20000, 1), np.int64)], axis=1)
    c = (c * 4096 // np.maximum(c.sum(1, keepdims=True), 1))
    c[:, 3] = 4096 - c[:, :3].sum(1)                       # coefficients sum to 1 << 12
    tmp = np.minimum((pix * 16384) >> 7, 32767)
    two = np.clip(((c * tmp).sum(1) + (1 << 18)) >> 19, 0, 255)
    one = ((c * pix).sum(1) + (1 << 11)) >> 12
    assert (two == one).all()


def test_window_push_and_tap_order():
    """k_fused3 window: W = [B_k, A_k, B_k-1, A_k-1] (byte 0 = newest row); an output row whose taps start at source row f is
    emitted after step k = ceil((f + 3) / 2) with j0 = 2k - f - 3 in {0, 1}; byte i of the window is row f + 3 - i"""
    for f in range(-2, 40):
        k = (f + 4) >> 1
        j0 = 2 * k - f - 3
        assert j0 in (0, 1) and 2 * (k - 1) < f + 3 <= 2 * k
        hist = [2 * k - j for j in range(8)]               # rows held by [Wcur bytes 0..3, Wprev bytes 0..3] = steps k, k-1, (k-1, k-2)
        wcur, wprev = hist[0:4], [2 * (k - 1) - j for j in range(4)]
        win = wcur if j0 == 0 else [wcur[1], wcur[2], wcur[3], wprev[2]]   # PRMT 0x6321
        assert win == [f + 3, f + 2, f + 1, f]


def _prmt(x, y, sel):
    """__byte_perm(x, y, sel) on uint32 (selector nibbles 0-7 only: no sign replication)"""
    b = [(x >> (8 * k)) & 0xFF for k in range(4)] + [(y >> (8 * k)) & 0xFF for k in range(4)]
    return sum(b[(sel >> (4 * i)) & 0x7] << (8 * i) for i in range(4))


def test_byte_permutation_selectors_of_the_yuv_family_kernels():
    """pe_kernels_yuv3.cu: the PRMT selector constants do what their comments say (4 x 4 byte transposes of combine / split
    planes, 3-byte pixel packing / unpacking, macropixel shuffles), on words with all-distinct bytes"""
    yw, uw, vw, aw = 0x03020100, 0x13121110, 0x23222120, 0x33323130          # byte k of plane p = 0x(p)(k)
    # k_combine_planes: px[k] = y_k | u_k << 8 | v_k << 16 | a_k << 24
    yu_lo, yu_hi = _prmt(yw, uw, 0x5140), _prmt(yw, uw, 0x7362)
    va_lo, va_hi = _prmt(vw, aw, 0x5140), _prmt(vw, aw, 0x7362)
    px = [_prmt(yu_lo, va_lo, 0x5410), _prmt(yu_lo, va_lo, 0x7632), _prmt(yu_hi, va_hi, 0x5410), _prmt(yu_hi, va_hi, 0x7632)]
    assert px == [0x30201000 + 0x01010101 * k for k in range(4)]
    # k_split_planes: the inverse transpose
    a01, b01 = _prmt(px[0], px[1], 0x5140), _prmt(px[0], px[1], 0x7362)
    a23, b23 = _prmt(px[2], px[3], 0x5140), _prmt(px[2], px[3], 0x7362)
    assert (_prmt(a01, a23, 0x5410), _prmt(a01, a23, 0x7632), _prmt(b01, b23, 0x5410), _prmt(b01, b23, 0x7632)) == (yw, uw, vw, aw)
    # st_packed4 / ld_packed4: four 3-byte pixels <-> three words
    p = [0x00A2A1A0, 0x00B2B1B0, 0x00C2C1C0, 0x00D2D1D0]
    w = [_prmt(p[0], p[1], 0x4210), _prmt(p[1], p[2], 0x5421), _prmt(p[2], p[3], 0x6542)]
    stream = b"".join(int(x).to_bytes(4, "little") for x in w)
    assert stream == bytes([0xA0, 0xA1, 0xA2, 0xB0, 0xB1, 0xB2, 0xC0, 0xC1, 0xC2, 0xD0, 0xD1, 0xD2])
    back = [w[0], _prmt(w[0], w[1], 0x0543), _prmt(w[1], w[2], 0x0432), w[2] >> 8]
    assert [x & 0xFFFFFF for x in back] == p
    # macropixels: UYVY -> YUYV byte order, luma word of two macropixels, duplicated chroma, YUV(A)888 pixels
    uyvy0, uyvy1 = 0xB1C0B0A0, 0xB3C1B2A1          # u y0 v y1 (u = A*, y = B*, v = C*)
    m0, m1 = _prmt(uyvy0, 0, 0x2301), _prmt(uyvy1, 0, 0x2301)
    assert (m0, m1) == (0xC0B1A0B0, 0xC1B3A1B2)     # y0 u y1 v
    assert _prmt(m0, m1, 0x6420) == 0xB3B2B1B0      # y0 y1 y0' y1'
    assert _prmt(m0, m1, 0x5511) == 0xA1A1A0A0 and _prmt(m0, m1, 0x7733) == 0xC1C1C0C0
    ff = 0xFF000000
    assert _prmt(m0, ff, 0x7310) == 0xFFC0A0B0 and _prmt(m0, ff, 0x7312) == 0xFFC0A0B1
    assert _prmt(m0, m1, 0x4451) & 0xFFFF == 0xA1A0 and _prmt(m0, m1, 0x4473) & 0xFFFF == 0xC1C0
    # k_yuv_march seed lane: byte 0 of a chroma word replaced by byte k of the seed word
    sd, cw = 0x44332211, 0xDDCCBBAA
    assert [_prmt(cw, sd, 0x3214 + k) for k in range(4)] == [0xDDCCBB11, 0xDDCCBB22, 0xDDCCBB33, 0xDDCCBB44]
    # k_alpha_over / clamp LUT: (x >> 3) & 0x1FE0 == 32 * ((x >> 8) & 0xFF)
    x = np.arange(0, 1 << 16, dtype=np.int64)
    assert (((x >> 3) & 0x1FE0) == 32 * ((x >> 8) & 0xFF)).all()


def test_clamp_walk_phase():
    """k_clamp_lut on YUV888: byte i of a densely walked plane is luma iff i % 3 == 0; a 4-byte word starting at i0 sees phases
    (i0 % 3 + k) % 3; with a rowstride that is not a multiple of 3 the dense phase drifts from the per-row phase (the reference's
    behaviour, replicated under ref_quirks)"""
    for rs, rows in ((160, 4), (96, 3), (224, 5)):
        dense = np.arange(rs * rows) % 3 == 0
        per_row = (np.arange(rs * rows) % rs) % 3 == 0
        for i0 in range(0, rs * rows, 4):
            assert [((i0 % 3) + k) % 3 == 0 for k in range(4)] == list(dense[i0:i0 + 4])
            assert [(((i0 % rs) % 3) + k) % 3 == 0 for k in range(4)] == list(per_row[i0:i0 + 4])
        assert (dense == per_row).all() == (rs % 3 == 0)


def test_resize_contract_tracks_independent_resamplers():
    """The resize filter is OUR contract (the reference calls libswscale, which is neither in its tree nor in this image:
    parity unpinned).  Sanity against an independent implementation (OpenCV): a smooth image resized by the
    oracle's filter stays within a couple of LSB of cv2's INTER_AREA (downscale) / INTER_LINEAR (upscale) result."""
    cv2 = __import__("pytest").importorskip("cv2")
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import pe_testlib as T
    o = T.oracle()
    yy, xx = np.mgrid[0:270, 0:480]
    img = np.stack([(127 + 100 * np.sin(xx / 37.0) * np.cos(yy / 23.0)), (xx * 255 / 479), (yy * 255 / 269), np.full(xx.shape, 255)], -1)
    src = np.ascontiguousarray(img.astype(np.uint8))
    for (dw, dh), interp, tol_mean, tol_max in (((320, 180), cv2.INTER_AREA, 0.6, 4), ((720, 404), cv2.INTER_LINEAR, 0.6, 4),
                                                ((480, 202), cv2.INTER_AREA, 0.6, 4)):
        got = np.zeros((dh, dw * 4), np.uint8)
        o.pe_or_resize_packed(T.ptr(src), 480 * 4, 480, 270, T.ptr(got), dw * 4, dw, dh, 4)
        ref = cv2.resize(src, (dw, dh), interpolation=interp).reshape(dh, dw * 4)
        d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        assert d.mean() < tol_mean and d.max() <= tol_max, (dw, dh, float(d.mean()), int(d.max()))


def test_slide_over_dividing_line_of_the_shipped_library():
    """pe_fx_slide_over_bound (host arithmetic inside libpe_b200.so, no GPU involved) == the oracle's line for every transition
    value, direction and a spread of sizes; the oracle's line is pinned against the compiled plugin in test_oracle_vs_reference.py"""
    import lives_b200 as lb
    import pe_testlib as T
    o = T.oracle()
    for dim in list(range(1, 70)) + [255, 256, 333, 510, 720, 765, 1080, 1280, 1920, 2160, 3840, 4096, 7680]:
        for direction in (1, 2, 3, 4):
            w, h = (dim, 7) if direction <= 2 else (7, dim)
            for tv in range(256):
                got = lb.slide_over_bound(direction, tv, w, h)
                assert got == o.pe_or_slide_over_bound(direction, tv, w, h), (direction, tv, w, h)
                assert 0 <= got <= dim
    assert lb.slide_over_bound("dir_r2l", 255, 37, 1) == 0 and lb.slide_over_bound("dir_l2r", 255, 37, 1) == 36  # the fast-math build


def test_slide_over_realigning_loads():
    """k_slide_over (pe_kernels_rgb.cu): 16 bytes at ANY source address = the 4 or 5 aligned 32-bit words that hold them, re-aligned
    with __funnelshift_r(w[i], w[i + 1], 8 * (addr & 3)) = low 32 bits of ((w[i + 1] << 32 | w[i]) >> shift); and the words touched
    never leave the row: with a 4-byte aligned row start and a stride that is a multiple of 4 they end at round_up4(payload) <= stride"""
    rng = np.random.default_rng(7)
    buf = rng.integers(0, 256, 64, dtype=np.uint8)
    words = buf.view("<u4").astype(np.uint64)
    for addr in range(0, 40):
        w0, sh = addr >> 2, 8 * (addr & 3)
        n = 4 if sh == 0 else 5
        w = words[w0:w0 + n]
        if sh == 0:
            out = w.astype("<u4")
        else:
            out = np.array([((int(w[i + 1]) << 32 | int(w[i])) >> sh) & 0xFFFFFFFF for i in range(4)], dtype="<u4")
        assert (out.view(np.uint8) == buf[addr:addr + 16]).all(), addr
        assert 4 * (w0 + n) <= ((addr + 16 + 3) // 4) * 4  # last word touched = the word of the last byte needed
    for payload in range(1, 200):
        for align in (4, 32):
            stride = (payload + align - 1) // align * align
            assert (payload + 3) // 4 * 4 <= stride


def test_resize_coefficient_banks_of_the_shipped_library():
    """pe_resize_filter_host (the bank libpe_b200.so builds on the host; no GPU involved) == the oracle's, for the round-1 triangle
    contract (0) and for each of libswscale's recipes (1 bilinear -- the default --, 2 bicubic, 3 Lanczos, 4 / 5 fast bilinear vertical
    / horizontal), over the BASELINE geometries, odd sizes, extreme ratios; every bank sums to 1 << bits per output sample and the
    libswscale banks keep their taps inside the frame with monotonic positions"""
    import ctypes as C
    import lives_b200  # noqa: F401
    import pe_testlib as T
    from lives_b200 import _capi
    lib, o = _capi.lib(), T.oracle()
    geoms = [(1080, 720), (1920, 1280), (2160, 1608), (3840, 3840), (720, 1080), (640, 320), (300, 75), (37, 37), (5, 3), (3, 7),
             (4, 4), (1000, 33), (61, 64), (64, 61), (1081, 719), (2, 2), (360, 720), (240, 300)]
    o.pe_or_resize_filter_kind.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    for recipe in range(6):
        for (s, d), bits in [(g, b) for g in geoms for b in (12, 14)]:
            fa, ca = np.zeros(d, np.int32), np.zeros((d, 64), np.int16)
            fb, cb = np.zeros(d, np.int32), np.zeros((d, 64), np.int16)
            if recipe == 0:
                ta = o.pe_or_resize_filter(s, d, bits, T.ptr(fa), T.ptr(ca), 64)
            else:
                ta = o.pe_or_resize_filter_kind(recipe, s, d, bits, fa.ctypes.data, ca.ctypes.data, 64)
            tb = lib.pe_resize_filter_host(recipe, s, d, bits, C.c_void_p(fb.ctypes.data), C.c_void_p(cb.ctypes.data), 64)
            if recipe == 5 and bits != 14:
                continue  # the position walk only exists as the (14-bit) horizontal pass
            if ta == -1 and tb == -1 and recipe in (2, 3) and s > 15 * d:
                continue  # more than 64 taps: both refuse (the engine fails the call loudly)
            assert ta == tb and ta > 0, (recipe, s, d, bits, ta, tb)
            assert (fa == fb).all() and (ca == cb).all(), (recipe, s, d, bits)
            assert (cb.astype(np.int64).sum(axis=1) == (1 << bits)).all(), (recipe, s, d, bits)
            if recipe >= 1:
                assert fb.min() >= 0 and (fb + tb).max() <= max(s, tb) and (np.diff(fb) >= 0).all(), (recipe, s, d)
            if recipe in (2, 3) and s > 8 and d > 8 and s != d:
                assert cb.min() < 0, "bicubic / Lanczos banks have negative lobes"
    # the default recipe is libswscale's bilinear
    fa, ca = np.zeros(720, np.int32), np.zeros((720, 64), np.int16)
    fb, cb = np.zeros(720, np.int32), np.zeros((720, 64), np.int16)
    assert o.pe_or_resize_filter_sws(1080, 720, 12, T.ptr(fa), T.ptr(ca), 64) == o.pe_or_resize_filter_kind(1, 1080, 720, 12, fb.ctypes.data, cb.ctypes.data, 64)
    assert (fa == fb).all() and (ca == cb).all()


def test_avg_chroma_closed_form_equals_the_reference_tables():
    """the chroma up-sampling kernels compute avg_chroma(x, y) as clamp((((x + y) * A + B) * M) >> 32, lo, hi) instead of gathering
    from the 64 KB tables of init_average (colourspace.c:190-217): the constants the library ships reproduce the oracle's tables
    (which tests/test_oracle_vs_reference.py pins to the compiled reference) entry by entry, for both clampings"""
    import ctypes as C
    import lives_b200  # noqa: F401
    import pe_testlib as T
    from lives_b200 import _capi
    lib, o = _capi.lib(), T.oracle()
    o.pe_or_avg_table.argtypes = [C.c_int, C.c_void_p]
    for clamped in (1, 0):
        k = np.zeros(5, np.uint32)
        assert lib.pe_avg_closed_form(clamped, C.c_void_p(k.ctypes.data)) == 1
        A, B, M, lo, hi = (int(v) for v in k)
        tab = np.zeros(65536, np.uint8)
        o.pe_or_avg_table(0 if clamped else 1, C.c_void_p(tab.ctypes.data))
        s = np.add.outer(np.arange(256, dtype=np.int64), np.arange(256, dtype=np.int64))
        n = s * A + B
        assert n.max() < 2 ** 32
        f = np.clip((n * M) >> 32, lo, hi).astype(np.uint8)
        assert (f.reshape(-1) == tab).all(), clamped


def test_parallel_host_copy_pool():
    """pe_host_parallel_copy2d = the copy-thread pool behind the staging of pageable host planes (pe_hoststage.h): dense blocks cut into
    page-sized pieces per thread, strided rows cut by rows; every byte of the payload arrives, nothing outside it is written"""
    import ctypes as C
    import lives_b200  # noqa: F401
    from lives_b200 import _capi
    lib = _capi.lib()
    rng = np.random.default_rng(5)
    for threads in (1, 3, 8):
        for (rows, wbytes, ss, ds) in ((1, 5_000_003, 5_000_003, 5_000_003), (1080, 1920, 1920, 1920), (270, 1921, 2048, 1984), (7, 13, 64, 32),
                                      (1, 1, 1, 1), (2160, 3840 * 4, 3840 * 4, 3840 * 4 + 64)):
            src = rng.integers(0, 256, rows * ss, dtype=np.uint8)
            dst = np.full(rows * ds, 7, np.uint8)
            assert lib.pe_host_parallel_copy2d(C.c_void_p(dst.ctypes.data), ds, C.c_void_p(src.ctypes.data), ss, wbytes, rows, threads) == 0
            s2, d2 = src.reshape(rows, ss), dst.reshape(rows, ds)
            assert (d2[:, :wbytes] == s2[:, :wbytes]).all(), (threads, rows, wbytes)
            assert (d2[:, wbytes:] == 7).all(), (threads, rows, wbytes)
