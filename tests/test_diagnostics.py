"""Per-frame diagnostics: is_all_black_ish (src/colourspace.c:2554-2594, exact and "ish" branches) and the row hashes of hash_cmp_layer
(:16044-16075 = minimd5, src/maths.c:575: the reference's own MD5 variant).  CPU: the oracle restatement against the compiled reference
(oracle/_ref/ref_diag.so, built by oracle/build_ref.py from the sources where they lie).  GPU: k_stats / k_row_hash against the oracle."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import pe_testlib as T


def _oracle():
    o = T.oracle()
    o.pe_or_is_all_black_ish.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    o.pe_or_minimd5.restype = C.c_uint64
    o.pe_or_minimd5.argtypes = [C.c_void_p, C.c_size_t]
    o.pe_or_row_hashes.restype = C.c_uint64
    o.pe_or_row_hashes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return o


def _ref():
    path = os.path.join(T.REF_DIR, "ref_diag.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/ref_diag.so not built")
    r = C.CDLL(path)
    r.ref_is_all_black_ish.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    r.ref_row_hashes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return r


def test_minimd5_is_not_rfc1321_and_the_oracle_equals_the_reference():
    o, r = _oracle(), _ref()
    rng = np.random.default_rng(0)
    differs = 0
    for n in list(range(0, 200)) + [1919, 1920, 3840, 5760, 15360]:
        a = rng.integers(0, 256, max(n, 1), dtype=np.uint8)
        ref = np.zeros(1, np.uint64)
        r.ref_row_hashes(a.ctypes.data, n, 1, max(n, 1), ref.ctypes.data)
        assert int(ref[0]) == o.pe_or_minimd5(a.ctypes.data, n), n
        u = np.frombuffer(hashlib.md5(a[:n].tobytes()).digest(), dtype="<u8")
        differs += int(u[0] ^ u[1]) != int(ref[0])
    assert differs > 190, "the reference's round 1 is its own (src/maths.h:48): a textbook MD5 would not be a parity check"


def test_is_all_black_ish_oracle_equals_the_reference_on_every_pixel_value():
    """all 2^24 (a, b, c) triples as one frame, each alone in its own row, both branches, with and without alpha"""
    o, r = _oracle(), _ref()
    v = np.arange(1 << 24, dtype=np.uint32)
    for has_alpha in (0, 1):
        ps = 4 if has_alpha else 3
        px = np.zeros((1 << 12, (1 << 12) * ps), np.uint8)
        flat = px.reshape(-1, ps)
        flat[:, 0], flat[:, 1], flat[:, 2] = v & 255, (v >> 8) & 255, v >> 16
        if has_alpha:
            flat[:, 3] = 255
        # per-row verdicts (a row = 4096 pixels sharing c and the high nibble of b): cheap to compare exhaustively row by row
        for exact in (0, 1):
            for row in range(0, 1 << 12, 97):
                a_ = o.pe_or_is_all_black_ish(1 << 12, 1, px.strides[0], has_alpha, px[row].ctypes.data, exact)
                b_ = r.ref_is_all_black_ish(1 << 12, 1, px.strides[0], has_alpha, px[row].ctypes.data, exact)
                assert a_ == b_, (has_alpha, exact, row)
        # and pixel by pixel on a random sample + the corner cases of the bit expression
        rng = np.random.default_rng(1)
        sample = np.concatenate([rng.integers(0, 1 << 24, 20000), [0, 0x202020, 0x201820, 0x201020, 0x200820, 0x1F1F1F, 0xE0E0E0, 0x20FF40]])
        for s in sample:
            p1 = flat[int(s):int(s) + 1]
            for exact in (0, 1):
                assert (o.pe_or_is_all_black_ish(1, 1, ps, has_alpha, p1.ctypes.data, exact)
                        == r.ref_is_all_black_ish(1, 1, ps, has_alpha, p1.ctypes.data, exact)), (hex(int(s)), exact)


@pytest.mark.gpu
def test_gpu_stats_and_row_hashes_against_the_oracle():
    lb = pytest.importorskip("lives_b200")
    o = _oracle()
    eng = lb.Engine()
    rng = np.random.default_rng(19)
    for pal, (w, h) in [(1, (333, 77)), (3, (333, 77)), (2, (1920, 1080)), (4, (3840, 2160)), (3, (1361, 9))]:
        ps = T.psize_of(pal)
        for kind in ("noise", "dark", "ish", "black"):
            if kind == "noise":
                src = T.make_packed(rng, w, h, ps, lo=3, hi=250)
            elif kind == "dark":
                src = T.make_packed(rng, w, h, ps, lo=0, hi=32)   # every byte < 32: black-ish, not black
            elif kind == "ish":
                src = T.make_packed(rng, w, h, ps, lo=0, hi=32)
                src[h // 2, (w // 3) * ps:(w // 3) * ps + 3] = (0x20, 0x18, 0x20)  # one pixel past the "ish" threshold
            else:
                src = np.zeros((h, T.rowstride(w, ps)), np.uint8)
                if ps == 4:
                    src[:, 3:w * 4:4] = 255  # opaque black
            lay = lb.Layer.from_host(eng, pal, w, h, [src])
            st = lay.stats()
            px = src[:, :w * ps].reshape(h, w, ps)
            a_off = 3 if ps == 4 else -1
            for k in range(ps):
                assert st["min"][k] == px[..., k].min() and st["max"][k] == px[..., k].max()
            col = [k for k in range(ps) if k != a_off]
            assert (st["hist"] == np.bincount(px[..., col].reshape(-1), minlength=256)).all()
            assert st["sum"] == int(px.astype(np.uint64).sum())
            assert st["all_black"] == o.pe_or_is_all_black_ish(w, h, src.strides[0], int(ps == 4), src.ctypes.data, 1), (pal, kind)
            assert st["all_black_ish"] == o.pe_or_is_all_black_ish(w, h, src.strides[0], int(ps == 4), src.ctypes.data, 0), (pal, kind)
            if kind == "ish":
                assert st["all_black_ish"] == 0 and st["all_black"] == 0
            if kind == "dark":
                assert st["all_black_ish"] == 1 and st["all_black"] == 0
            # row hashes: the reference's choice (`width` bytes) and the whole payload
            for nb in (0, w * ps, 55, 56, 64, 1):
                if nb > w * ps:
                    continue
                exp = np.zeros(h, np.uint64)
                par = o.pe_or_row_hashes(src.ctypes.data, nb if nb else w, h, src.strides[0], exp.ctypes.data)
                got, gpar = lay.row_hashes(nb)
                assert (got == exp).all() and gpar == par, (pal, kind, nb)
    # a planar frame: statistics of plane 0, no black verdicts
    y, u, v = T.make_yuv_planar(rng, 640, 360, False, True)
    st = lb.Layer.from_host(eng, 512, 640, 360, [y, u, v]).stats()
    assert st["min"][0] == y[:, :640].min() and st["max"][0] == y[:, :640].max() and st["all_black"] == -1 and st["all_black_ish"] == -1
    assert (st["hist"] == np.bincount(y[:, :640].reshape(-1), minlength=256)).all()
    eng.close()
