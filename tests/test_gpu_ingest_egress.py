"""GPU: SURVEY 8f rank 1 -- frames that enter the chain in device memory (the decoder plugin's get_frame shape, src/plugins.h:442,
over a clip resident in HBM) and leave it as ONE packed frame (the render tail of src/events.c:4247-4263).  Same bytes as the
host-upload path; no host memory is touched between load() and render_out()."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pe_testlib as T  # noqa: E402

pytestmark = pytest.mark.gpu
lb = pytest.importorskip("lives_b200")


@pytest.fixture(scope="module")
def eng():
    e = lb.Engine()
    yield e
    e.close()


def test_ingest_from_device_clip_equals_host_upload(eng):
    rng = np.random.default_rng(5)
    w, h, n = 640, 360, 3
    frames = [T.make_yuv_planar(rng, w, h, False, True) for _ in range(n)]
    cache = lb.ClipCache(eng, lb.WEED_PALETTE_YUV420P, w, h, n, yuv_subspace=1)
    for i, f in enumerate(frames):
        cache.load(i, list(f))
    for i in (2, 0, 1, 4):  # frame numbers wrap around
        a = cache.frame(i)
        b = lb.Layer.from_host(eng, lb.WEED_PALETTE_YUV420P, w, h, list(frames[i % n]), yuv_subspace=1)
        assert a.palette == 512 and (a.width, a.height) == (w, h)
        assert lb.convert_layer_palette(a, 3, 0) and lb.convert_layer_palette(b, 3, 0)
        assert (a.to_host()[0] == b.to_host()[0]).all()
    # zero-copy borrow: read-only use as the fused chain's fg
    bg = T.make_packed(rng, w, h, 4)
    outs = []
    for fg_l in (cache.borrow(1), lb.Layer.from_host(eng, lb.WEED_PALETTE_YUV420P, w, h, list(frames[1]), yuv_subspace=1)):
        bg_l = lb.Layer.from_host(eng, 3, w, h, [bg], gamma_type=T.G_LINEAR)
        out_l = lb.Layer.create(eng, 3, w, h)
        lb.fused_convert_letterbox_over_gamma(fg_l, bg_l, out_l, w, 300, 0.5, T.G_LINEAR, T.G_SRGB)
        outs.append(out_l.to_host()[0])
    assert (outs[0] == outs[1]).all()
    cache.close()


def test_render_out_downloads_only_the_final_rgb24(eng):
    """convert_layer_palette(RGBA32 -> RGB24) + download == render_out; the async form overlaps with later work"""
    o = T.oracle()
    rng = np.random.default_rng(6)
    w, h = 1920, 1080
    src = T.make_packed(rng, w, h, 4)
    exp = np.zeros((h, T.rowstride(w, 3)), np.uint8)
    assert o.pe_or_rgb_to_rgb(3, 1, T.ptr(src), src.strides[0], w, h, T.ptr(exp), exp.strides[0], None) == 0
    lay = lb.Layer.from_host(eng, 3, w, h, [src])
    dst = np.zeros((h, w * 3), np.uint8)  # a tightly packed destination (a pixbuf's rowstride differs from the layer's)
    lb.render_out(lay, 1, dst)
    assert lay.palette == 1 and (dst == exp[:, :w * 3]).all()
    # four frames in flight on the four slots
    lays = [lb.Layer.from_host(eng, 3, w, h, [np.roll(src, k, axis=0)]) for k in range(4)]
    dsts = [np.zeros((h, w * 3), np.uint8) for _ in range(4)]
    for k in range(4):
        lb.render_out_begin(lays[k], 1, dsts[k], k)
    for k in range(4):
        lb.render_out_wait(eng, k)
        assert (dsts[k] == np.roll(exp, k, axis=0)[:, :w * 3]).all()
    # a planar frame is refused unless a packed palette is asked for
    y, u, v = T.make_yuv_planar(rng, 64, 48, False, True)
    pl = lb.Layer.from_host(eng, 512, 64, 48, [y, u, v], yuv_subspace=1)
    with pytest.raises(lb.PixelEngineError):
        lb.render_out(pl, 0, np.zeros((48, 64 * 3), np.uint8))
    lb.render_out(pl, 1, np.zeros((48, 64 * 3), np.uint8))


@pytest.mark.parametrize("geom", [(1280, 720, 536), (640, 360, 300)])
def test_convert_plan_descriptor_fused_equals_op_by_op(eng, geom):
    """get_op_order's result handed over as one descriptor (src/nodemodel.c:161, :1065-1282): the chain leaves as one
    fused launch, and the same descriptor run op by op (the reference's substeps) gives the same bytes"""
    rng = np.random.default_rng(9)
    w, h, ih = geom
    y, u, v = T.make_yuv_planar(rng, w, h, False, True)
    bg = T.make_packed(rng, w, h, 4)
    res = []
    for no_fuse in (False, True):
        # YUV420P -> RGBA32 (substep 1, done by the resize), letterbox to w x ih inside w x h (substep 2), no gamma before the effect
        plan = lb.convert_plan({lb.OP_RESIZE: 1, lb.OP_PCONV: 1, lb.OP_LETTERBOX: 2}, width=w, height=ih, lb_width=w, lb_height=h,
                               out_palette=3, no_fuse=no_fuse)
        fg_l = lb.Layer.from_host(eng, 512, w, h, [y, u, v], yuv_subspace=1)
        bg_l = lb.Layer.from_host(eng, 3, w, h, [bg], gamma_type=T.G_LINEAR)
        out_l = lb.Layer.create(eng, 3, w, h)
        if not no_fuse:  # once untimed: a cold engine builds its gamma / filter tables with a kernel of its own on first use
            warm = lb.Layer.from_host(eng, 512, w, h, [y, u, v], yuv_subspace=1)
            lb.run_convert_plan_over(warm, plan, bg_l, lb.Layer.create(eng, 3, w, h), 0.5, T.G_SRGB)
        n0 = eng.launch_count
        path = lb.run_convert_plan_over(fg_l, plan, bg_l, out_l, 0.5, T.G_SRGB)
        assert path == (0 if no_fuse else 1)
        if not no_fuse:
            assert eng.launch_count - n0 == 1, "the whole chain is one kernel launch"
        res.append(out_l.to_host()[0])
    assert (res[0] == res[1]).all()
    # the plain CONVERT step on its own layer: pconv then resize then gamma, in the plan's order
    src = T.make_packed(rng, 320, 240, 3)
    lay = lb.Layer.from_host(eng, 1, 320, 240, [src], gamma_type=T.G_SRGB)
    plan = lb.convert_plan({lb.OP_PCONV: 1, lb.OP_RESIZE: 2, lb.OP_GAMMA: 3}, width=160, height=120, out_palette=3, out_gamma=T.G_LINEAR)
    assert lb.run_convert_plan(lay, plan)
    ref = lb.Layer.from_host(eng, 1, 320, 240, [src], gamma_type=T.G_SRGB)
    assert lb.convert_layer_palette(ref, 3, 0) and lb.resize_layer(ref, 160, 120, 1, 3, 0) and lb.gamma_convert_layer(T.G_LINEAR, ref)
    assert lay.palette == 3 and (lay.width, lay.height) == (160, 120) and (lay.to_host()[0] == ref.to_host()[0]).all()
