#!/usr/bin/env python
"""Generate tests/golden/ref_vectors_yuv411.npz -- YUV411 (IYU1) as a conversion source -- from the COMPILED REFERENCE
(oracle/_ref/libref_oracle.so, built by oracle/build_ref.py from /root/reference).  Run in the build container only; the GPU box
consumes the committed .npz.  Dense buffers (what convert_yuv411_to_*_frame walks), RGBA / BGRA destinations prefilled with 255 (the
alpha bytes the reference never writes, src/colourspace.c:8338-8370)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pe_testlib as T  # noqa: E402
from test_yuv411 import CASES, RGB_SOURCES, _src, _out_planes  # noqa: E402

WM, H = 12, 6


def main():
    assert T.have_ref(), "build oracle/_ref first (python oracle/build_ref.py)"
    r = T.ref()
    rng = np.random.default_rng(595)
    out = {}
    for cl in (T.CLAMPED, T.UNCLAMPED):
        src = _src(rng, WM, H, cl == T.CLAMPED)
        out["src_cl%d" % cl] = src
        for target, order, add_alpha, pal in CASES:
            exp = _out_planes(target, add_alpha, WM, H, dense=True)
            if target == 0 and add_alpha:
                exp[0][:] = 255
            pl = exp + [exp[0]] * (4 - len(exp))
            r.ref_yuv411_to(target, T.ptr(src), WM, H, exp[0].strides[0], T.planes_arg(*pl), order, add_alpha, cl)
            for k, p in enumerate(exp):
                out["%s_cl%d_p%d" % (pal, cl, k)] = p
    for order, has_alpha, pal in RGB_SOURCES:   # RGB sources -> YUV411 (convert_{rgb,bgr,argb}_to_yuv411_frame), 24 x 6 pixels
        rgb = T.make_packed(rng, 24, H, 4 if has_alpha else 3)
        out["rgbsrc_%s" % pal] = rgb
        for cl in (T.CLAMPED, T.UNCLAMPED):
            d = np.zeros((H, 6 * 6), np.uint8)
            r.ref_rgb_to_yuv411(T.ptr(rgb), 24, H, rgb.strides[0], T.ptr(d), order, has_alpha, cl)
            out["from_%s_cl%d" % (pal, cl)] = d
    np.savez_compressed(os.path.join(HERE, "ref_vectors_yuv411.npz"), **out)
    print("wrote %d arrays" % len(out))


if __name__ == "__main__":
    main()
