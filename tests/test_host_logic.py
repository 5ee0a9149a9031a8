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
