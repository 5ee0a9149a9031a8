"""SURVEY 8f rank 4, display hand-off: lives_b200/libpe_vpp.so is a LiVES video playback plugin (videoplugin.h ABI; the host's view
src/plugins.h:153-215) whose screen is a ring of device surfaces.  Layers are built through the reference's libweed by the minihost and
handed to play_frame the way src/player.c:1508 does; the shown surface and the VPP_CAN_RETURN data are read back and compared."""
import ctypes as C
import os

import numpy as np
import pytest

import pe_testlib as T

VPPLIB = os.path.join(T.REPO, "lives_b200", "libpe_vpp.so")
pytestmark = pytest.mark.skipif(not T.have_ref() or not os.path.exists(os.path.join(T.REF_DIR, "libweed_minihost.so")),
                                reason="oracle/_ref (reference libweed + minihost) not built")

ABI = ["module_check_init", "get_description", "get_palette_list", "set_palette", "get_capabilities", "init_screen", "play_frame",
       "render_frame", "exit_screen", "module_unload"]  # videoplugin.h:62-150
EXT = ["pe_vpp_play_device_frame", "pe_vpp_acquire", "pe_vpp_read_surface", "pe_vpp_counters"]


def _load():
    mh = C.CDLL(os.path.join(T.REF_DIR, "libweed_minihost.so"), mode=C.RTLD_GLOBAL)
    mh.mh_layer_new.restype = C.c_void_p
    mh.mh_layer_new.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int] * 5
    mh.mh_layer_plane.restype = C.c_void_p
    mh.mh_layer_plane.argtypes = [C.c_void_p, C.c_int]
    mh.mh_layer_rowstride.argtypes = [C.c_void_p, C.c_int]
    mh.mh_layer_free.argtypes = [C.c_void_p]
    if not os.path.exists(VPPLIB):
        from lives_b200.build import build
        build()
    v = C.CDLL(VPPLIB)
    v.module_check_init.restype = C.c_char_p
    v.get_description.restype = C.c_char_p
    v.get_palette_list.restype = C.POINTER(C.c_int)
    v.get_capabilities.restype = C.c_uint64
    v.init_screen.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p]
    v.play_frame.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    v.render_frame.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    v.exit_screen.argtypes = [C.c_int16, C.c_int16]
    v.pe_vpp_read_surface.argtypes = [C.c_void_p, C.c_int]
    v.pe_vpp_play_device_frame.argtypes = [C.c_void_p, C.c_int64]
    v.pe_vpp_counters.argtypes = [C.POINTER(C.c_uint64)] * 3
    return mh, v


def _layer(mh, pal, w, h, arr):
    pp = (C.c_void_p * 4)(arr.ctypes.data, 0, 0, 0)
    rs = (C.c_int * 4)(arr.strides[0], 0, 0, 0)
    rows = (C.c_int * 4)(h, 0, 0, 0)
    lay = mh.mh_layer_new(pal, w, h, 1, pp, rs, rows, 0, 0, 0, 0, 0)
    assert lay
    return lay


def test_playback_plugin_abi_is_exported():
    _, v = _load()
    for name in ABI + EXT:
        assert hasattr(v, name), name
    pals = v.get_palette_list()
    got = []
    while pals[len(got)] != 0:  # WEED_PALETTE_END
        got.append(pals[len(got)])
    assert got == [3, 4, 1, 2]
    assert v.get_capabilities(3) == 3  # VPP_CAN_RESIZE | VPP_CAN_RETURN
    assert v.set_palette(3) and not v.set_palette(512)
    assert b"GPU memory" in v.get_description()


def test_playback_plugin_refuses_to_load_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _, v = _load()
    assert v.module_check_init() is not None  # an error string: the host drops the plugin (no CPU path)
    assert not v.init_screen(64, 32, 0, 0, 0, None)


@pytest.mark.gpu
def test_play_frame_keeps_the_frame_on_the_device_and_returns_it():
    mh, v = _load()
    assert v.module_check_init() is None
    rng = np.random.default_rng(50)
    for pal, ps in ((3, 4), (1, 3), (4, 4), (2, 3)):
        w, h = 640, 360
        assert v.set_palette(pal) and v.init_screen(w, h, 0, 0, 0, None)
        frames = [T.make_packed(rng, w, h, ps) for _ in range(5)]
        for k, fr in enumerate(frames):
            lay = _layer(mh, pal, w, h, fr)
            retbuf = np.zeros_like(fr)
            ret = _layer(mh, pal, w, h, retbuf) if k % 2 else None
            assert v.play_frame(lay, 1000 * k, ret)
            shown = np.zeros_like(fr)
            assert v.pe_vpp_read_surface(T.ptr(shown), shown.strides[0]) == 0
            assert (shown[:, :w * ps] == fr[:, :w * ps]).all(), (pal, k)
            if ret:
                rs = mh.mh_layer_rowstride(ret, 0)
                back = np.ctypeslib.as_array(C.cast(mh.mh_layer_plane(ret, 0), C.POINTER(C.c_uint8)), shape=(h, rs))
                assert (back[:, :w * ps] == fr[:, :w * ps]).all(), (pal, k, "return data")
                mh.mh_layer_free(ret)
            mh.mh_layer_free(lay)
        n, up, down = C.c_uint64(), C.c_uint64(), C.c_uint64()
        v.pe_vpp_counters(C.byref(n), C.byref(up), C.byref(down))
        assert n.value == 5 and up.value == 5 * w * h * ps and down.value == 2 * w * h * ps
        # a frame of another palette is refused (the host converts first, src/player.c:1359)
        other = _layer(mh, 1 if ps == 4 else 3, w, h, frames[0])
        assert not v.play_frame(other, 0, None)
        mh.mh_layer_free(other)
        v.exit_screen(0, 0)


@pytest.mark.gpu
def test_play_frame_resizes_to_the_screen_on_the_device():
    """VPP_CAN_RESIZE: a 1280 x 720 frame on a 640 x 360 screen = resize_layer's bilinear bank (the oracle's), return data unresized"""
    mh, v = _load()
    assert v.module_check_init() is None
    o = T.oracle()
    rng = np.random.default_rng(51)
    w, h, sw, sh = 1280, 720, 640, 360
    assert v.set_palette(3) and v.init_screen(sw, sh, 0, 0, 0, None)
    fr = T.make_packed(rng, w, h, 4)
    lay, retbuf = _layer(mh, 3, w, h, fr), np.zeros_like(fr)
    ret = _layer(mh, 3, w, h, retbuf)
    assert v.play_frame(lay, 7, ret)
    shown = np.zeros((sh, T.rowstride(sw, 4)), np.uint8)
    assert v.pe_vpp_read_surface(T.ptr(shown), shown.strides[0]) == 0
    exp = np.zeros_like(shown)
    o.pe_or_resize_packed(T.ptr(fr), fr.strides[0], w, h, T.ptr(exp), exp.strides[0], sw, sh, 4)
    assert (shown[:, :sw * 4] == exp[:, :sw * 4]).all()
    back = np.ctypeslib.as_array(C.cast(mh.mh_layer_plane(ret, 0), C.POINTER(C.c_uint8)), shape=(h, mh.mh_layer_rowstride(ret, 0)))
    assert (back[:, :w * 4] == fr[:, :w * 4]).all()
    mh.mh_layer_free(lay); mh.mh_layer_free(ret)
    # render_frame (the older entry point): packed rows without padding
    dense = np.ascontiguousarray(fr[:sh, :sw * 4])
    pd = (C.c_void_p * 1)(dense.ctypes.data)
    assert v.render_frame(sw, sh, 8, pd, None, None)
    got = np.zeros((sh, sw * 4), np.uint8)
    assert v.pe_vpp_read_surface(T.ptr(got), sw * 4) == 0 and (got == dense).all()
    v.exit_screen(0, 0)


@pytest.mark.gpu
def test_device_resident_frame_is_shown_without_crossing_pcie():
    """the output of the fused chain goes to the screen with zero host traffic: one device-to-device copy"""
    lb = pytest.importorskip("lives_b200")
    mh, v = _load()
    assert v.module_check_init() is None
    eng = lb.Engine.shared() if hasattr(lb.Engine, "shared") else None
    if eng is None:
        pytest.skip("no shared-engine handle in the Python mirror")
    rng = np.random.default_rng(52)
    w, h = 1920, 1080
    src = T.make_packed(rng, w, h, 4)
    lay = lb.Layer.from_host(eng, 3, w, h, [src])
    assert v.set_palette(3) and v.init_screen(w, h, 0, 0, 0, None)
    assert v.pe_vpp_play_device_frame(lay._h, 42)
    n, up, down = C.c_uint64(), C.c_uint64(), C.c_uint64()
    v.pe_vpp_counters(C.byref(n), C.byref(up), C.byref(down))
    assert n.value == 1 and up.value == 0 and down.value == 0
    shown = np.zeros_like(src)
    assert v.pe_vpp_read_surface(T.ptr(shown), shown.strides[0]) == 0 and (shown[:, :w * 4] == src[:, :w * 4]).all()
    v.exit_screen(0, 0)
