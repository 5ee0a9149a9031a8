#!/usr/bin/env python
"""Prototype for the next step of the resize contract: libswscale's bilinear coefficient recipe -- triangle
taps at 2^-30 precision, near-zero taps (cumulated weight < 0.002) cut off, borders folded onto the edge sample, then normalised
to 1 << bits with error diffusion from tap to tap -- restated from the published algorithm of libswscale/utils.c initFilter and
checked against the real library (tests/swscale_ref.py; libswscale 9.1.100).  CPU-only numpy, slow, small frames.

Measured here (uniform noise in one channel, whole-frame SWS_BILINEAR):
    192x108 -> 128x72    this recipe 98.6 % equal, max 1   |  current contract 96.8 % equal, max 7
    384x216 -> 256x144               98.8 %,        max 1   |                   97.8 %,        max 8
    160x120 -> 320x240               93.8 %,        max 1   |                   93.8 %,        max 1
    64x216  -> 64x160 (vertical only) 99.4 %,       max 1
The residue sits exactly on the rounding boundary of the vertical pass (fractional part < 0.02 or > 0.996): not a coefficient
difference (no +-2 change of any tap removes it; 14 / 16 / 20-bit vertical coefficients are further away, a pmulhw-style vertical
pass much further).  Adopting the recipe means changing pe_or_resize_filter AND build_resize_filter together (the CUDA path is
bit-exact against the oracle); it was found after the round's GPU budget was spent, so it is a tool, not the product, for now.

    python tools/swscale_filter_proto.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pe_testlib as T  # noqa: E402
import swscale_ref as S  # noqa: E402


def init_filter_bilinear(src, dst, one, cutoff=0.002):
    xinc = ((src << 16) + (dst >> 1)) // dst
    if xinc <= 1 << 16: fs = 1 + 2
    else: fs = 1 + (2 * src + dst - 1) // dst
    fs = max(min(fs, src - 2), 1)
    ratio = src // dst
    lg = 0
    while (ratio >> (lg + 1)) > 0: lg += 1
    fone = 1 << (54 - min(lg if ratio > 0 else 0, 8))
    filt = np.zeros((dst, fs), dtype=object); pos = [0] * dst
    xdst = ((128 * xinc) >> 7) - ((128 * 0x10000) >> 7)
    for i in range(dst):
        num = xdst - (fs - 2) * (1 << 16)
        xx = (abs(num) // (1 << 17)) * (1 if num >= 0 else -1)  # C division truncates towards zero
        pos[i] = xx
        for j in range(fs):
            d = abs(xx * (1 << 17) - xdst) << 13
            if xinc > 1 << 16: d = d * dst // src
            c = (1 << 30) - d
            if c < 0: c = 0
            c *= fone >> 30
            filt[i, j] = c
            xx += 1
        xdst += 2 * xinc
    # reduce
    minfs = 0
    for i in range(dst - 1, -1, -1):
        mn = fs
        cut = 0
        for j in range(fs):
            cut += abs(filt[i, 0])
            if cut > cutoff * fone: break
            if i < dst - 1 and pos[i] >= pos[i + 1]: break
            filt[i, :-1] = filt[i, 1:]; filt[i, -1] = 0
            pos[i] += 1
        cut = 0
        for j in range(fs - 1, 0, -1):
            cut += abs(filt[i, j])
            if cut > cutoff * fone: break
            mn -= 1
        minfs = max(minfs, mn)
    nfs = minfs
    f2 = np.zeros((dst, nfs), dtype=object)
    for i in range(dst):
        for j in range(nfs):
            f2[i, j] = filt[i, j] if j < fs else 0
    # borders
    for i in range(dst):
        if pos[i] < 0:
            for j in range(1, nfs):
                left = max(j + pos[i], 0)
                f2[i, left] += f2[i, j]; f2[i, j] = 0
            pos[i] = 0
        if pos[i] + nfs > src:
            shift = pos[i] + min(nfs - src, 0)
            acc = 0
            for j in range(nfs - 1, -1, -1):
                if pos[i] + j >= src:
                    acc += f2[i, j]; f2[i, j] = 0
            for j in range(nfs - 1, -1, -1):
                f2[i, j] = 0 if j < shift else f2[i, j - shift]
            pos[i] -= shift
            f2[i, src - 1 - pos[i]] += acc
    out = np.zeros((dst, nfs), dtype=np.int64)
    for i in range(dst):
        s = sum(int(v) for v in f2[i]); s = (s + one // 2) // one
        if not s: s = 1
        err = 0
        for j in range(nfs):
            v = int(f2[i, j]) + err
            iv = (v + s // 2) // s if v >= 0 else -((-v + s // 2) // s)
            out[i, j] = iv; err = v - iv * s
    return out, pos


def scale2d(plane, dw, dh):
    """horizontal 14-bit pass to a 15-bit intermediate, vertical 12-bit pass (the shape of the contract, swscale's coefficients)"""
    h, w = plane.shape
    ch, ph = init_filter_bilinear(w, dw, 1 << 14)
    cv, pv = init_filter_bilinear(h, dh, 1 << 12)
    p = plane.astype(np.int64)
    tmp = np.zeros((h, dw), np.int64)
    for i in range(dw):
        acc = np.zeros(h, np.int64)
        for j in range(ch.shape[1]):
            acc += ch[i, j] * p[:, min(max(ph[i] + j, 0), w - 1)]
        tmp[:, i] = np.minimum(acc >> 7, 32767)
    out = np.zeros((dh, dw), np.int64)
    for i in range(dh):
        acc = np.full(dw, 64 << 12, np.int64)
        for j in range(cv.shape[1]):
            acc += cv[i, j] * tmp[min(max(pv[i] + j, 0), h - 1)]
        out[i] = np.clip(acc >> 19, 0, 255)
    return out


def main():
    if S.load()[0] is None:
        raise SystemExit(S.load()[1])
    rng = np.random.default_rng(2)
    for (w, h, dw, dh) in ((192, 108, 128, 72), (384, 216, 256, 144), (160, 120, 320, 240), (300, 200, 160, 120), (64, 216, 64, 160)):
        src = np.full((h, T.rowstride(w, 4)), 128, np.uint8)
        src[:, 0:w * 4:4] = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = S.scale([src], "rgba", w, h, "rgba", dw, dh, T.rowstride(dw, 4))[:, 0:dw * 4:4].astype(int)
        got = scale2d(src[:, 0:w * 4:4], dw, dh)
        cur = np.zeros((dh, T.rowstride(dw, 4)), np.uint8)
        T.oracle().pe_or_resize_packed(T.ptr(src), src.strides[0], w, h, T.ptr(cur), cur.strides[0], dw, dh, 4)
        d, d2 = ref - got, ref - cur[:, 0:dw * 4:4].astype(int)
        print((w, h, dw, dh), "swscale recipe: %.2f %% equal, max %d | current contract: %.2f %% equal, max %d"
              % (100 * (d == 0).mean(), np.abs(d).max(), 100 * (d2 == 0).mean(), np.abs(d2).max()))


if __name__ == "__main__":
    main()
