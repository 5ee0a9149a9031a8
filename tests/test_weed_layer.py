"""Boundary B2 over real weed_layer_t plants: lives_b200/libpe_weed_layer.so exports the reference's frame ops with the exact
signatures of src/colourspace.h:387-415.  The layers are built by tests/host/weed_minihost.c through the REFERENCE's own libweed
(oracle/_ref/libweed*.so: weed_plant_new(WEED_PLANT_LAYER) + the leaves of src/layers.c:292-510), handed to the drop-ins, and read back
leaf by leaf.  Pixels are checked bit for bit against the oracle; a failing call must leave every leaf and byte as it was
(src/colourspace.c:13906-13927).
"""
import ctypes as C
import os

import numpy as np
import pytest

import pe_testlib as T

LAYERLIB = os.path.join(T.REPO, "lives_b200", "libpe_weed_layer.so")
pytestmark = pytest.mark.skipif(not T.have_ref() or not os.path.exists(os.path.join(T.REF_DIR, "libweed_minihost.so")),
                                reason="oracle/_ref (reference libweed + minihost) not built")

EXPORTS = ["alpha_premult", "gamma_convert_layer", "gamma_convert_sub_layer", "convert_layer_palette", "convert_layer_palette_full",
           "resize_layer_full", "resize_layer", "letterbox_layer", "pe_weed_layer_bind", "pe_weed_layer_set_allocator",
           "pe_weed_layer_set_pinning", "pe_weed_layer_engine"]


class Host:
    """the reference's libweed (RTLD_GLOBAL, so that the drop-ins find its weed_leaf_get ... variables) + the minihost + our library"""

    def __init__(self):
        # (the minihost is the "host program": it defines libweed's function-pointer variables, weed-host.h, and pulls libweed in)
        mh = C.CDLL(os.path.join(T.REF_DIR, "libweed_minihost.so"), mode=C.RTLD_GLOBAL)
        mh.mh_layer_new.restype = C.c_void_p
        mh.mh_layer_new.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int] * 5
        mh.mh_layer_has.argtypes = [C.c_void_p, C.c_char_p]
        mh.mh_layer_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        mh.mh_layer_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        mh.mh_layer_nplanes.argtypes = [C.c_void_p]
        mh.mh_layer_plane.restype = C.c_void_p
        mh.mh_layer_plane.argtypes = [C.c_void_p, C.c_int]
        mh.mh_layer_rowstride.argtypes = [C.c_void_p, C.c_int]
        mh.mh_layer_free.argtypes = [C.c_void_p]
        self.mh = mh
        if not os.path.exists(LAYERLIB):
            from lives_b200.build import build
            build()
        lib = C.CDLL(LAYERLIB)
        VP, I, D = C.c_void_p, C.c_int, C.c_double
        lib.convert_layer_palette_full.argtypes = [VP, I, I, I, I, I]
        lib.convert_layer_palette.argtypes = [VP, I, I]
        lib.resize_layer_full.argtypes = [VP, I, I, I, I, I, I, I, I]
        lib.resize_layer.argtypes = [VP, I, I, I, I, I]
        lib.letterbox_layer.argtypes = [VP, I, I, I, I, I, I, I]
        lib.gamma_convert_layer.argtypes = [I, VP]
        lib.gamma_convert_sub_layer.argtypes = [I, D, VP, I, I, I, I, I]
        lib.alpha_premult.argtypes = [VP, I]
        lib.alpha_premult.restype = None
        lib.pe_weed_layer_bind.argtypes = [VP]
        lib.pe_weed_layer_engine.restype = VP
        self.lib = lib

    def layer(self, palette, width_px, height, planes, clamping=0, sampling=0, subspace=0, gamma=0, yuv_leaves=None):
        n = len(planes)
        pp = (C.c_void_p * 4)(*([p.ctypes.data for p in planes] + [0] * (4 - n)))
        rs = (C.c_int * 4)(*([p.strides[0] for p in planes] + [0] * (4 - n)))
        rows = (C.c_int * 4)(*([p.shape[0] for p in planes] + [0] * (4 - n)))
        mpx = width_px // 2 if palette in (564, 565) else width_px
        if yuv_leaves is None:
            yuv_leaves = palette >= 512
        h = self.mh.mh_layer_new(palette, mpx, height, n, pp, rs, rows, clamping, sampling, subspace, gamma, int(yuv_leaves))
        assert h
        return h

    def snapshot(self, layer):
        """every leaf the ops may touch + the plane bytes"""
        mh = self.mh
        pal, h = mh.mh_layer_int(layer, b"current_palette", 0), mh.mh_layer_int(layer, b"height", 0)
        n = mh.mh_layer_nplanes(layer)
        planes = []
        for p in range(n):
            rs = mh.mh_layer_rowstride(layer, p)
            rows = h if p == 0 or pal not in (512, 513) else h // 2
            planes.append(np.ctypeslib.as_array(C.cast(mh.mh_layer_plane(layer, p), C.POINTER(C.c_uint8)), shape=(rows, rs)).copy())
        leaves = {k: (mh.mh_layer_has(layer, k), mh.mh_layer_int(layer, k, -99)) for k in
                  (b"current_palette", b"width", b"height", b"YUV_clamping", b"YUV_sampling", b"YUV_subspace", b"gamma_type", b"flags")}
        ptrs = [mh.mh_layer_plane(layer, p) for p in range(n)]
        return dict(leaves=leaves, planes=planes, ptrs=ptrs, rowstrides=[mh.mh_layer_rowstride(layer, p) for p in range(n)])


@pytest.fixture(scope="module")
def host():
    return Host()


def test_library_exports_the_reference_signatures(host):
    for name in EXPORTS:
        assert hasattr(host.lib, name), name
    # binds to the libweed that is loaded in this process (the function-pointer variables of libweed/weed.h:340-351, filled by the
    # host's weed_init(): before that there is nothing to bind to)
    host.mh.mh_layer_free(host.mh.mh_layer_new(1, 4, 4, 0, None, None, None, 0, 0, 0, 0, 0))
    assert host.lib.pe_weed_layer_bind(None) == 0


def test_without_a_gpu_every_op_fails_and_leaves_the_layer_untouched(host):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rng = np.random.default_rng(1)
    src = T.make_packed(rng, 64, 48, 3)
    lay = host.layer(1, 64, 48, [src])
    before = host.snapshot(lay)
    assert host.lib.convert_layer_palette(lay, 2, 0) == 0       # FALSE: no CPU fallback behind the drop-in
    assert host.lib.resize_layer(lay, 32, 24, 1, 1, 0) == 0
    assert host.lib.gamma_convert_layer(1, lay) == 0
    after = host.snapshot(lay)
    assert before["leaves"] == after["leaves"] and before["ptrs"] == after["ptrs"]
    assert all((a == b).all() for a, b in zip(before["planes"], after["planes"]))
    assert not host.lib.pe_weed_layer_engine()
    host.mh.mh_layer_free(lay)


def test_not_a_layer_is_refused(host):
    rng = np.random.default_rng(2)
    src = T.make_packed(rng, 16, 8, 3)
    lay = host.layer(1, 16, 8, [src])
    host.mh.mh_layer_set_int(lay, b"type", 4)  # a channel template, not a layer (WEED_IS_LAYER)
    assert host.lib.convert_layer_palette(lay, 2, 0) == 0
    assert host.lib.convert_layer_palette(None, 2, 0) == 0
    host.mh.mh_layer_set_int(lay, b"type", 128)
    host.mh.mh_layer_free(lay)


# ------------------------------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_config1_rgb24_to_bgr24_in_place_on_a_real_layer(host):
    """BASELINE config 1 through the drop-in: 640x480 RGB24 -> BGR24, same buffer, same rowstride (pconv_can_inplace :12148)"""
    o = T.oracle()
    rng = np.random.default_rng(1)
    src = T.make_packed(rng, 640, 480, 3)
    exp = np.zeros_like(src)
    assert o.pe_or_rgb_to_rgb(1, 2, T.ptr(src), src.strides[0], 640, 480, T.ptr(exp), exp.strides[0], None) == 0
    lay = host.layer(1, 640, 480, [src])
    before = host.snapshot(lay)
    assert host.lib.convert_layer_palette(lay, 2, 0) == 1
    after = host.snapshot(lay)
    assert after["leaves"][b"current_palette"] == (1, 2) and after["ptrs"] == before["ptrs"] and after["rowstrides"] == before["rowstrides"]
    assert (after["planes"][0][:, :1920] == exp[:, :1920]).all()
    host.mh.mh_layer_free(lay)


@pytest.mark.gpu
@pytest.mark.parametrize("pin", [0, 1])
def test_yuv420p_layer_to_rgba_and_resize(host, pin):
    """config 2 in its two-call form: convert_layer_palette(YUV420P -> RGBA32) then resize_layer(1280x720): new pixel buffers, the YUV
    leaves deleted (conv_done :13878-13881), width / height / rowstrides rewritten; with and without page-locking the buffers"""
    o = T.oracle()
    host.lib.pe_weed_layer_set_pinning(pin)
    try:
        rng = np.random.default_rng(2)
        w, h = 1920, 1080
        y, u, v = T.make_yuv_planar(rng, w, h, False, True)
        rgba = np.zeros((h, T.rowstride(w, 4)), np.uint8)
        o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(rgba), rgba.strides[0], 0, 1, 0, 0, 1, T.Q_HIGH, 1, None)
        lay = host.layer(512, w, h, [y, u, v], clamping=0, sampling=0, subspace=1)
        assert host.lib.convert_layer_palette(lay, 3, 0) == 1
        s = host.snapshot(lay)
        assert s["leaves"][b"current_palette"] == (1, 3) and s["leaves"][b"width"] == (1, w) and s["leaves"][b"height"] == (1, h)
        assert s["leaves"][b"YUV_clamping"][0] == 0 and s["leaves"][b"YUV_subspace"][0] == 0, "YUV leaves are deleted on an RGB layer"
        assert len(s["planes"]) == 1 and s["rowstrides"] == [T.rowstride(w, 4)]
        assert (s["planes"][0][:, :w * 4] == rgba[:, :w * 4]).all()
        exp = np.zeros((720, T.rowstride(1280, 4)), np.uint8)
        o.pe_or_resize_packed(T.ptr(rgba), rgba.strides[0], w, h, T.ptr(exp), exp.strides[0], 1280, 720, 4)
        assert host.lib.resize_layer(lay, 1280, 720, 1, 3, 0) == 1
        s = host.snapshot(lay)
        assert s["leaves"][b"width"] == (1, 1280) and s["leaves"][b"height"] == (1, 720)
        assert (s["planes"][0][:, :1280 * 4] == exp[:, :1280 * 4]).all()
        host.mh.mh_layer_free(lay)
    finally:
        host.lib.pe_weed_layer_set_pinning(0)


@pytest.mark.gpu
def test_letterbox_gamma_premult_on_a_real_layer(host):
    o = T.oracle()
    rng = np.random.default_rng(3)
    w, h = 320, 240
    src = T.make_packed(rng, w, h, 4)
    # letterbox 320x240 -> inner 320x180 inside 320x240
    inner = np.zeros((180, T.rowstride(w, 4)), np.uint8)
    o.pe_or_resize_packed(T.ptr(src), src.strides[0], w, h, T.ptr(inner), inner.strides[0], w, 180, 4)
    boxed = np.zeros((h, T.rowstride(w, 4)), np.uint8)
    o.pe_or_letterbox_packed(T.ptr(inner), inner.strides[0], w, 180, T.ptr(boxed), boxed.strides[0], w, h, 3)
    lay = host.layer(3, w, h, [src], gamma=T.G_SRGB)
    assert host.lib.letterbox_layer(lay, w, h, w, 180, 1, 3, 0) == 1
    s = host.snapshot(lay)
    assert s["leaves"][b"width"] == (1, w) and s["leaves"][b"height"] == (1, h)
    assert (s["planes"][0][:, :w * 4] == boxed[:, :w * 4]).all()
    # gamma sRGB -> linear in place: the gamma_type leaf follows
    lut = np.zeros(256, np.uint8)
    o.pe_or_gamma_lut8(1.0, T.G_SRGB, T.G_LINEAR, 1.4, T.ptr(lut))
    exp = boxed.copy()
    o.pe_or_gamma_apply(T.ptr(exp), exp.strides[0], 3, 0, 0, w, h, T.ptr(lut))
    ptrs = s["ptrs"]
    assert host.lib.gamma_convert_layer(T.G_LINEAR, lay) == 1
    s = host.snapshot(lay)
    assert s["leaves"][b"gamma_type"] == (1, T.G_LINEAR) and s["ptrs"] == ptrs
    assert (s["planes"][0][:, :w * 4] == exp[:, :w * 4]).all()
    # alpha_premult forward: flags leaf gets WEED_LAYER_ALPHA_PREMULT
    pm = exp.copy()
    o.pe_or_alpha_premult(T.ptr(pm), pm.strides[0], 3, 0, w, h, 1)
    host.lib.alpha_premult(lay, 1)
    s = host.snapshot(lay)
    assert (s["planes"][0][:, :w * 4] == pm[:, :w * 4]).all() and s["leaves"][b"flags"][1] & 1
    host.mh.mh_layer_free(lay)


@pytest.mark.gpu
def test_failed_op_leaves_the_layer_untouched(host):
    """a target this build refuses (WEED_PALETTE_RGBFLOAT = 64, libweed/weed-palettes.h:59: no converter in the reference either):
    FALSE, nothing changed"""
    rng = np.random.default_rng(4)
    src = T.make_packed(rng, 64, 48, 4)
    lay = host.layer(5, 64, 48, [src])
    before = host.snapshot(lay)
    assert host.lib.convert_layer_palette(lay, 64, 0) == 0
    after = host.snapshot(lay)
    assert before["leaves"] == after["leaves"] and before["ptrs"] == after["ptrs"] and (before["planes"][0] == after["planes"][0]).all()
    # a layer without pixel data: resize_layer_full records the target and returns FALSE (:14820-14832)
    empty = host.mh.mh_layer_new(3, 64, 48, 0, None, None, None, 0, 0, 0, 0, 0)
    assert host.lib.resize_layer(empty, 32, 24, 1, 3, 0) == 0
    assert host.mh.mh_layer_int(empty, b"width", 0) == 32 and host.mh.mh_layer_int(empty, b"height", 0) == 24
    host.mh.mh_layer_free(lay)
    host.mh.mh_layer_free(empty)
