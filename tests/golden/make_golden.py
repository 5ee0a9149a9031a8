#!/usr/bin/env python
"""Generate tests/golden/ref_vectors.npz from the COMPILED REFERENCE (oracle/_ref/*.so, built by oracle/build_ref.py from
/root/reference).  Run in the build container only; the GPU box and CI consume the committed .npz.

Every array is an output of reference code (colourspace.c slices, the unmodified simple_blend.so / multi_blends.so driven
through the real weed_bootstrap, compositor.c paint_pixel), together with the seeded inputs that produced it.
"""
import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pe_testlib as T  # noqa: E402

_GAMMA_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import pe_testlib as T
r = T.ref()
f, t = int(sys.argv[1]), int(sys.argv[2])
a = np.zeros(256, np.uint8); a16 = np.zeros(65536, np.uint16)
assert r.ref_gamma_lut8(1.0, f, t, T.ptr(a)) == 0 and r.ref_gamma_lut16(1.0, f, t, T.ptr(a16)) == 0
sys.stdout.write(a.tobytes().hex() + " " + __import__("hashlib").sha256(a16.tobytes()).hexdigest())
"""


def main():
    assert T.have_ref(), "build oracle/_ref first (python oracle/build_ref.py)"
    r = T.ref()
    r.ref_set_prefs(1, T.Q_HIGH, 1.4)
    out = {}
    # conversion tables, all four (clamping, subspace) variants
    for cl in (0, 1):
        for sub in (1, 2):
            tabs = np.zeros((14, 256), np.int32)
            for w in range(14):
                r.ref_get_conv_table(cl, sub, w, T.ptr(tabs[w]))
            out["conv_cl%d_sub%d" % (cl, sub)] = tabs
    # premultiply tables: sha256 of the int32 tables + the alpha = 128 row of each
    for w in range(6):
        t = np.zeros(65536, np.int32)
        r.ref_get_premult_table(w, T.ptr(t))
        out["premult%d_sha256" % w] = np.frombuffer(hashlib.sha256(t.tobytes()).digest(), np.uint8)
        out["premult%d_row128" % w] = t[128 * 256:129 * 256].astype(np.uint8)
    # gamma LUTs: one fresh process per pair (the reference's LUT cache is keyed by a value it mutates)
    for f in (-1, 1, 2, 1024):
        for t in (-1, 1, 2, 1024):
            if f == t:
                continue
            res = subprocess.run([sys.executable, "-c", _GAMMA_CHILD % os.path.dirname(HERE), str(f), str(t)], capture_output=True, text=True)
            assert res.returncode == 0, res.stderr
            hx, sha = res.stdout.split()
            out["lut8_%d_%d" % (f, t)] = np.frombuffer(bytes.fromhex(hx), np.uint8)
            out["lut16_%d_%d_sha256" % (f, t)] = np.frombuffer(bytes.fromhex(sha), np.uint8)
    # planar 4:2:0 / 4:2:2 -> RGB(A): 64 x 48, seed 2; interior rows are the defined region for 4:2:0
    rng = np.random.default_rng(2)
    w, h = 64, 48
    for is422 in (0, 1):
        for cl in (0, 1):
            y, u, v = T.make_yuv_planar(rng, w, h, bool(is422), cl == 0)
            key = "p%d_cl%d" % (422 if is422 else 420, cl)
            out[key + "_y"], out[key + "_u"], out[key + "_v"] = y.copy(), u.copy(), v.copy()
            for order, add_alpha, name in ((0, 0, "rgb24"), (0, 1, "rgba32"), (1, 0, "bgr24"), (1, 1, "bgra32")):
                if cl == 1 and order == 0 and not is422:
                    continue  # the unclamped RGB-order loop writes one byte early (colourspace.c:3704): covered by the oracle test
                ps = 4 if add_alpha else 3
                ors = T.rowstride(w, ps)
                full = np.full((h + 16, ors), 7, np.uint8)
                dst = full[8:8 + h]
                r.ref_yuv420p_to_rgb(T.planes_arg(y, u, v), w, h, T.strides_arg(y, u, v), ors, T.ptr(dst), order, add_alpha, is422, 0, cl, 1, 0, 0)
                out[key + "_" + name] = dst.copy()
    # packed 4:2:2 and 4:4:4
    src = T.make_packed(rng, 48, 20, 4)
    out["uyvy_src"] = src
    for fmt, nm in ((0, "uyvy"), (1, "yuyv")):
        dst = np.zeros((20, T.rowstride(96, 3)), np.uint8)
        r.ref_packed422_to_rgb(fmt, T.ptr(src), 48, 20, src.strides[0], dst.strides[0], T.ptr(dst), 0, 0, 0, 1)
        out[nm + "_to_rgb24"] = dst
    src = T.make_packed(rng, 50, 12, 3)
    out["rgb_src"] = src
    dst = np.zeros((12, T.rowstride(50, 3)), np.uint8)
    r.ref_rgb_to_yuv888(T.ptr(src), 50, 12, src.strides[0], dst.strides[0], T.ptr(dst), 0, 0, 0, 0)
    out["rgb_to_yuv888_cl0"] = dst
    dst = np.zeros((12, T.rowstride(50, 3)), np.uint8)
    r.ref_yuv888_to_rgb(T.ptr(src), 50, 12, src.strides[0], dst.strides[0], T.ptr(dst), 0, 0, 0, 0, 1)
    out["yuv888_to_rgb_cl0"] = dst
    # RGB -> packed 4:2:2 / planar 4:4:4 / planar 4:2:0 / 4:2:2 (unpadded planes: the reference advances its pointers densely)
    src = T.make_packed(np.random.default_rng(2024), 64, 12, 3)  # (its own generator: the entries above and below keep their values)
    out["rgb2_src"] = src
    for fmt, nm in ((0, "uyvy"), (1, "yuyv")):
        dst = np.zeros((12, 128), np.uint8)
        r.ref_rgb_to_packed422(fmt, T.ptr(src), 64, 12, src.strides[0], dst.strides[0], T.ptr(dst), 0, 0, 0, 0, 0)
        out["rgb_to_%s_cl0" % nm] = dst
    pl = [np.zeros((12, 64), np.uint8) for _ in range(4)]
    r.ref_rgb_to_yuv444p(T.ptr(src), 64, 12, src.strides[0], 64, T.planes_arg(*pl), 0, 0, 0, 0)
    for k, nm in enumerate("yuv"):
        out["rgb_to_yuv444p_cl0_" + nm] = pl[k]
    for is422, nm in ((0, "yuv420p"), (1, "yuv422p")):
        ch = 12 if is422 else 6
        pl = [np.zeros((12, 64), np.uint8), np.zeros((ch, 32), np.uint8), np.zeros((ch, 32), np.uint8)]
        strides = (C.c_int * 3)(64, 32, 32)
        r.ref_rgb_to_yuv420(T.ptr(src), 64, 12, src.strides[0], strides, T.planes_arg(*pl), 0, is422, 0, 1, 0)
        for k, pn in enumerate("yuv"):
            out["rgb_to_%s_cl0_%s" % (nm, pn)] = pl[k]
    for which, nm in ((0, "cavgc"), (1, "cavgu")):
        t = np.zeros(65536, np.uint8)
        assert r.ref_get_avg_table(which, T.ptr(t)) == 0
        out["avg_%s_sha256" % nm] = np.frombuffer(hashlib.sha256(t.tobytes()).digest(), np.uint8)
        out["avg_%s_row200" % nm] = t[200 * 256:201 * 256].copy()
    # effect plugins through the real bootstrap
    mh = C.CDLL(os.path.join(T.REF_DIR, "libweed_minihost.so"))
    mh.mh_open.argtypes = [C.c_char_p]
    mh.mh_run2.argtypes = [T.I, T.I, T.I, T.I, T.I, T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.I]
    hs = mh.mh_open(os.path.join(T.REF_DIR, "simple_blend.so").encode())
    hm = mh.mh_open(os.path.join(T.REF_DIR, "multi_blends.so").encode())
    s1, s2 = T.make_packed(rng, 61, 9, 3), T.make_packed(rng, 61, 9, 3)
    out["blend_s1"], out["blend_s2"] = s1, s2
    for typ in range(4):
        d = np.zeros_like(s1)
        assert mh.mh_run2(hs, typ, 1, 61, 9, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d), d.strides[0], 100, 1) == 0
        out["simple_blend_t%d_bf100" % typ] = d
    for typ in range(7):
        d = np.zeros_like(s1)
        assert mh.mh_run2(hm, typ, 1, 61, 9, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d), d.strides[0], 100, 1) == 0
        out["multi_blend_t%d_bf100" % typ] = d
    a1, a2 = T.make_packed(rng, 33, 5, 4), T.make_packed(rng, 33, 5, 4)
    a2[:, 3::4][rng.random((5, 40))[:, :a2[:, 3::4].shape[1]] < 0.4] = 255
    out["blend4_s1"], out["blend4_s2"] = a1, a2
    d = a1.copy()
    assert mh.mh_run2(hs, 0, 3, 33, 5, T.ptr(a1), a1.strides[0], T.ptr(a2), a2.strides[0], T.ptr(d), d.strides[0], 77, 1) == 0
    out["simple_blend_rgba_bf77"] = d
    # paint_pixel: all (dst, src) byte pairs for three alphas
    p = T.ref_paint()
    g = np.arange(256, dtype=np.uint8)
    d0, s0 = np.meshgrid(g, g, indexing="ij")
    for alpha in (0.1, 0.5, 1.0 / 3.0):
        dst = np.repeat(d0.reshape(-1, 1), 3, axis=1).astype(np.uint8).copy()
        srcp = np.repeat(s0.reshape(-1, 1), 3, axis=1).astype(np.uint8).copy()
        p.ref_paint_rows(T.ptr(dst), T.ptr(srcp), 65536, 3, alpha)
        out["paint_alpha_%.4f" % alpha] = dst[:, 0].reshape(256, 256).copy()
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_vectors.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
