"""GPU: V-first planar output (WEED_PALETTE_YVU420P).

The reference's converters write Cb to dest[1] and Cr to dest[2] whatever the target and then, at conv_done, swap the chroma plane
pointers of every V-first palette (src/colourspace.c:13895, swap_chroma_planes :12108); a V-first SOURCE is swapped on the way in
(:12354).  YUV420P <-> YVU420P is an in-place relabel (pconv_can_inplace :12152-12155, no pixel work :13618-13623).
Checked here: every converter into YVU420P equals the same converter into YUV420P (itself pinned to the oracle by
tests/test_gpu_parity.py / test_golden_yuv.py) with planes 1 and 2 exchanged; round trips through either palette agree.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pe_testlib as T  # noqa: E402

pytestmark = pytest.mark.gpu
lb = pytest.importorskip("lives_b200")

YUV420P, YVU420P = 512, 513


@pytest.fixture(scope="module")
def eng():
    e = lb.Engine()
    yield e
    e.close()


def _make(eng, rng, pal, w, h):
    """a factory of identical source layers of palette `pal`"""
    if pal in (1, 2, 3, 4, 588, 589):
        a = T.make_packed(rng, w, h, T.psize_of(pal))
        return lambda: lb.Layer.from_host(eng, pal, w, h, [a])
    if pal in (564, 565):
        a = T.make_packed(rng, w // 2, h, 4)
        return lambda: lb.Layer.from_host(eng, pal, w, h, [a])
    if pal in (544, 545):
        n = 3 if pal == 544 else 4
        pl = [T.make_packed(rng, w, h, 1) for _ in range(n)]
        return lambda: lb.Layer.from_host(eng, pal, w, h, pl)
    raise AssertionError(pal)


@pytest.mark.parametrize("ipal", [1, 2, 3, 4, 544, 545, 588, 589, 564, 565])
@pytest.mark.parametrize("size", [(64, 48), (130, 38)])
def test_every_converter_into_yvu420p(eng, ipal, size):
    rng = np.random.default_rng(ipal)
    w, h = size
    mk = _make(eng, rng, ipal, w, h)
    a, b = mk(), mk()
    assert lb.convert_layer_palette(a, YUV420P, 0)
    assert lb.convert_layer_palette(b, YVU420P, 0)
    assert b.palette == YVU420P and (b.width, b.height) == (a.width, a.height)
    ya, ua, va = a.to_host()
    yb, p1, p2 = b.to_host()
    assert (ya == yb).all()
    assert (p1 == va).all() and (p2 == ua).all(), "plane 1 of a YVU420P layer is V (colourspace.c:13895)"
    assert ua.any() and not (ua == va).all()


@pytest.mark.parametrize("size", [(64, 48), (1920, 1080)])
def test_yvu420p_round_trip(eng, size):
    """RGB24 -> YVU420P -> RGB24 == RGB24 -> YUV420P -> RGB24 (a missing swap exchanges Cb and Cr)"""
    rng = np.random.default_rng(7)
    w, h = size
    src = T.make_packed(rng, w, h, 3)
    a = lb.Layer.from_host(eng, 1, w, h, [src])
    b = lb.Layer.from_host(eng, 1, w, h, [src])
    assert lb.convert_layer_palette(a, YUV420P, 0) and lb.convert_layer_palette(a, 1, 0)
    assert lb.convert_layer_palette(b, YVU420P, 0) and lb.convert_layer_palette(b, 1, 0)
    assert (a.to_host()[0] == b.to_host()[0]).all()


def test_yuv420p_yvu420p_relabel(eng):
    """in place, no kernel: the chroma planes change places (:12354 / :13895)"""
    rng = np.random.default_rng(8)
    w, h = 96, 64
    y, u, v = T.make_yuv_planar(rng, w, h, False, True)
    lay = lb.Layer.from_host(eng, YUV420P, w, h, [y, u, v])
    n0 = eng.launch_count
    assert lb.convert_layer_palette(lay, YVU420P, 0)
    assert eng.launch_count == n0 and lay.palette == YVU420P
    gy, g1, g2 = lay.to_host()
    assert (gy == y).all() and (g1 == v).all() and (g2 == u).all()
    assert lb.convert_layer_palette(lay, YUV420P, 0)
    assert eng.launch_count == n0 and lay.palette == YUV420P
    gy, g1, g2 = lay.to_host()
    assert (g1 == u).all() and (g2 == v).all()
    # a V-first source converts like its U-first twin
    a = lb.Layer.from_host(eng, YUV420P, w, h, [y, u, v], yuv_subspace=1)
    b = lb.Layer.from_host(eng, YVU420P, w, h, [y, v, u], yuv_subspace=1)
    assert lb.convert_layer_palette(a, 3, 0) and lb.convert_layer_palette(b, 3, 0)
    assert (a.to_host()[0] == b.to_host()[0]).all()
