"""Host-side mirror of the reference's frame-op interface over the C ABI (include/pixel_engine.h).

Names, argument order and return conventions follow src/colourspace.h:387-415 of the reference:
    convert_layer_palette_full(layer, outpl, oclamping, osampling, osubspace, tgt_gamma) -> bool
    convert_layer_palette(layer, outpl, op_clamping) -> bool
    resize_layer_full / resize_layer / letterbox_layer / gamma_convert_layer / gamma_convert_sub_layer -> bool
    alpha_premult(layer, direction) -> None
Layers are mutated in place and left untouched on failure (colourspace.c:13906-13927).
A `Layer` is a device-resident frame (the weed_layer_t of src/layers.c); widths are in PIXELS.
This module is plumbing only: every pixel is computed by the sm_100a kernels behind libpe_b200.so.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import PixelEngineError, PixelEngineUnavailable  # noqa: F401

# ---- constants (libweed/weed-palettes.h:43-183, colourspace.h:26-30, preferences.h:100-104) -------------------
WEED_PALETTE_NONE = 0
WEED_PALETTE_RGB24, WEED_PALETTE_BGR24, WEED_PALETTE_RGBA32, WEED_PALETTE_BGRA32, WEED_PALETTE_ARGB32 = 1, 2, 3, 4, 5
WEED_PALETTE_YUV420P, WEED_PALETTE_YVU420P, WEED_PALETTE_YUV422P, WEED_PALETTE_YUV444P = 512, 513, 522, 544
WEED_PALETTE_YUVA4444P, WEED_PALETTE_UYVY, WEED_PALETTE_YUYV, WEED_PALETTE_YUV888, WEED_PALETTE_YUVA8888 = 545, 564, 565, 588, 589
WEED_PALETTE_YUV411 = 595
WEED_YUV_CLAMPING_CLAMPED, WEED_YUV_CLAMPING_UNCLAMPED = 0, 1
WEED_YUV_SAMPLING_DEFAULT, WEED_YUV_SAMPLING_MPEG = 0, 1
WEED_YUV_SUBSPACE_YUV, WEED_YUV_SUBSPACE_YCBCR, WEED_YUV_SUBSPACE_BT709 = 0, 1, 2
WEED_GAMMA_UNKNOWN, WEED_GAMMA_LINEAR, WEED_GAMMA_SRGB, WEED_GAMMA_BT709 = 0, -1, 1, 2
WEED_GAMMA_MONITOR, WEED_GAMMA_FILE, WEED_GAMMA_VARIANT = 1024, 1025, 2048
PB_QUALITY_LOW, PB_QUALITY_MED, PB_QUALITY_HIGH = 1, 2, 3
LIVES_INTERP_FAST, LIVES_INTERP_NORMAL, LIVES_INTERP_BEST = 0, 1, 2
LIVES_DIRECTION_REVERSE, LIVES_DIRECTION_FORWARD = -1, 1
WEED_LAYER_ALPHA_PREMULT = 1

_PLANAR = {512: 3, 513: 3, 522: 3, 544: 3, 545: 4}


def frame_layout(palette, width, height):
    """(nplanes, rowstrides, plane_heights, total_bytes) -- calc_rowstrides colourspace.c:11252"""
    n = C.c_int(0)
    rs = (C.c_int * 4)()
    ph = (C.c_int * 4)()
    total = capi.lib().pe_frame_layout(palette, width, height, C.byref(n), rs, ph)
    if not total:
        raise ValueError("bad frame geometry: palette %d %dx%d" % (palette, width, height))
    return n.value, list(rs)[:n.value], list(ph)[:n.value], total


def plane_row_bytes(palette, width, plane):
    if plane == 0:
        if palette in (WEED_PALETTE_UYVY, WEED_PALETTE_YUYV):
            return (width // 2) * 4
        if palette == WEED_PALETTE_YUV411:
            return (width // 4) * 6
        return width * {1: 3, 2: 3, 588: 3, 3: 4, 4: 4, 5: 4, 589: 4}.get(palette, 1)
    return width >> 1 if palette in (512, 513, 522) else width


class Engine:
    """One per (process, GPU): stream, conversion tables, LUT cache, device block pool (init_colour_engine :1973)."""

    def __init__(self, device=0, pb_quality=PB_QUALITY_HIGH, screen_gamma=1.4, apply_gamma=True, alpha_post=False,
                 ref_quirks=True, stream=None):
        lib = capi.lib()
        cfg = capi.pe_config_t()
        lib.pe_config_default(C.byref(cfg))
        cfg.device, cfg.pb_quality, cfg.screen_gamma = device, pb_quality, screen_gamma
        cfg.apply_gamma, cfg.alpha_post, cfg.ref_quirks = int(apply_gamma), int(alpha_post), int(ref_quirks)
        cfg.stream = stream
        h = C.c_void_p()
        capi.check(lib.pe_engine_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._lib = lib

    @classmethod
    def shared(cls):
        """the process-wide engine the weed plugin, the weed_layer_t drop-ins and the playback plugin run on (pe_engine_shared); never
        destroyed by this handle"""
        lib = capi.lib()
        h = lib.pe_engine_shared()
        if not h:
            raise capi.PixelEngineError(-1, lib.pe_last_error().decode())
        self = cls.__new__(cls)
        self._h, self._lib, self._borrowed = C.c_void_p(h), lib, True
        return self

    def close(self):
        if self._h and not getattr(self, "_borrowed", False):
            self._lib.pe_engine_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        capi.check(self._lib.pe_engine_sync(self._h))

    @property
    def stream(self):
        return self._lib.pe_engine_stream(self._h)

    @property
    def launch_count(self):
        return self._lib.pe_engine_launch_count(self._h)

    @property
    def sm_count(self):
        return self._lib.pe_sm_count(self._h)

    def set_resize_recipe(self, recipe):
        """1: libswscale's coefficient recipes, one bank per LiVESInterpType (default); 0: the round-1 triangle contract"""
        capi.check(self._lib.pe_engine_set_resize_recipe(self._h, recipe))

    def set_sm_limit(self, n):
        """persistent kernels use at most n SMs (0: all) -- leaves room for a collective that runs beside the engine"""
        capi.check(self._lib.pe_engine_set_sm_limit(self._h, int(n)))

    def mc_publish(self, mc_dst, src, nbytes, cuda_stream=None, max_ctas=0):
        """pe_mc_publish: `nbytes` from device address `src` through the multicast address `mc_dst` (see lives_b200/shard.py)"""
        capi.check(self._lib.pe_mc_publish(self._h, C.c_void_p(mc_dst), C.c_void_p(src), nbytes, C.c_void_p(cuda_stream or 0), max_ctas))

    def timer_start(self):
        capi.check(self._lib.pe_timer_start(self._h))

    def timer_stop_ms(self):
        ms = C.c_float()
        capi.check(self._lib.pe_timer_stop_ms(self._h, C.byref(ms)))
        return ms.value

    def gamma_lut8(self, fileg, gamma_from, gamma_to):
        out = np.zeros(256, np.uint8)
        rc = self._lib.pe_gamma_lut8(self._h, fileg, gamma_from, gamma_to, out.ctypes.data)
        return out if rc == capi.PE_OK else None


class Layer:
    """A device-resident frame.  Create with Layer.create / Layer.from_host; read back with to_host()."""

    def __init__(self, engine, handle):
        self.engine = engine
        self._h = handle

    @classmethod
    def create(cls, engine, palette, width, height, yuv_clamping=0, yuv_sampling=0, yuv_subspace=0, gamma_type=0,
               black_fill=False):
        h = C.c_void_p()
        capi.check(engine._lib.pe_frame_create(engine._h, palette, width, height, yuv_clamping, yuv_sampling, yuv_subspace,
                                               gamma_type, int(black_fill), C.byref(h)))
        return cls(engine, h)

    @classmethod
    def from_host(cls, engine, palette, width, height, planes, yuv_clamping=0, yuv_sampling=0, yuv_subspace=0,
                  gamma_type=0, flags=0):
        """planes: list of 2-D uint8 numpy arrays (rows x rowstride), one per plane"""
        layer = cls.create(engine, palette, width, height, yuv_clamping, yuv_sampling, yuv_subspace, gamma_type)
        if flags:
            layer.flags = flags
        layer.upload(planes)
        return layer

    @classmethod
    def wrap_device(cls, engine, palette, width, height, ptrs, rowstrides, yuv_clamping=0, yuv_sampling=0, yuv_subspace=0,
                    gamma_type=0, flags=0):
        """wrap caller-owned DEVICE memory (e.g. torch tensors' data_ptr()); never freed by the engine"""
        d = capi.pe_frame_desc_t()
        d.palette, d.width, d.height, d.nplanes = palette, width, height, len(ptrs)
        for i, (p, rs) in enumerate(zip(ptrs, rowstrides)):
            d.planes[i] = p
            d.rowstrides[i] = rs
        d.yuv_clamping, d.yuv_sampling, d.yuv_subspace, d.gamma_type, d.flags = yuv_clamping, yuv_sampling, yuv_subspace, gamma_type, flags
        h = C.c_void_p()
        capi.check(engine._lib.pe_frame_wrap(engine._h, C.byref(d), C.byref(h)))
        return cls(engine, h)

    def free(self):
        if self._h:
            self.engine._lib.pe_frame_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.engine._h:
                self.free()
        except Exception:
            pass

    # ---- metadata (weed_layer_get_* src/layers.c:292-510)
    @property
    def desc(self):
        d = capi.pe_frame_desc_t()
        capi.check(self.engine._lib.pe_frame_get_desc(self._h, C.byref(d)))
        return d

    palette = property(lambda self: self.desc.palette)
    width = property(lambda self: self.desc.width)
    height = property(lambda self: self.desc.height)
    rowstrides = property(lambda self: list(self.desc.rowstrides)[:self.desc.nplanes])
    plane_ptrs = property(lambda self: list(self.desc.planes)[:self.desc.nplanes])
    yuv_clamping = property(lambda self: self.desc.yuv_clamping)
    yuv_sampling = property(lambda self: self.desc.yuv_sampling)
    yuv_subspace = property(lambda self: self.desc.yuv_subspace)

    @property
    def gamma_type(self):
        return self.desc.gamma_type

    @gamma_type.setter
    def gamma_type(self, v):
        capi.check(self.engine._lib.pe_frame_set_gamma(self._h, v))

    @property
    def flags(self):
        return self.desc.flags

    @flags.setter
    def flags(self, v):
        capi.check(self.engine._lib.pe_frame_set_flags(self._h, v))

    # ---- host <-> device
    def upload(self, planes):
        d = self.desc
        if len(planes) != d.nplanes:
            raise ValueError("expected %d planes" % d.nplanes)
        ptrs = (C.c_void_p * 4)()
        rs = (C.c_int * 4)()
        for i, p in enumerate(planes):
            if p.dtype != np.uint8 or p.ndim != 2 or not p.flags.c_contiguous:
                raise ValueError("planes must be C-contiguous 2-D uint8 arrays (rows x rowstride)")
            ptrs[i] = p.ctypes.data
            rs[i] = p.strides[0]
        capi.check(self.engine._lib.pe_frame_upload(self.engine._h, self._h, ptrs, rs))
        self.engine.sync()  # `planes` may be pageable and die after this call

    def to_host(self, rowstrides=None):
        """list of (rows x rowstride) uint8 arrays; padding bytes are zero"""
        d = self.desc
        _, _, ph, _ = frame_layout(d.palette, d.width, d.height)
        out, ptrs, rs = [], (C.c_void_p * 4)(), (C.c_int * 4)()
        for i in range(d.nplanes):
            stride = rowstrides[i] if rowstrides else d.rowstrides[i]
            a = np.zeros((ph[i], stride), np.uint8)
            out.append(a)
            ptrs[i] = a.ctypes.data
            rs[i] = stride
        capi.check(self.engine._lib.pe_frame_download(self.engine._h, self._h, ptrs, rs))
        return out

    def copy(self):
        """weed_layer_copy (src/layers.c:840), deep, on the device"""
        h = C.c_void_p()
        capi.check(self.engine._lib.pe_frame_copy(self.engine._h, self._h, C.byref(h)))
        return Layer(self.engine, h)

    def stats(self):
        s = capi.pe_frame_stats_t()
        capi.check(self.engine._lib.pe_frame_stats(self.engine._h, self._h, C.byref(s)))
        return dict(min=list(s.min), max=list(s.max), hist=np.array(list(s.hist), np.uint32), sum=int(s.sum),
                    all_black_ish=s.all_black_ish, all_black=s.all_black)

    def row_hashes(self, nbytes=0):
        """hash_cmp_layer (colourspace.c:16044): (per-row minimd5 hashes as uint64 array, parity)"""
        h = np.zeros(self.height, np.uint64)
        par = C.c_uint64(0)
        capi.check(self.engine._lib.pe_frame_row_hashes(self.engine._h, self._h, nbytes, h.ctypes.data, C.byref(par)))
        return h, int(par.value)


# ---- boundary B2: the reference's frame ops --------------------------------------------------------------------

def convert_layer_palette_full(layer, outpl, oclamping, osampling, osubspace, tgt_gamma):
    """colourspace.h:395 / colourspace.c:12190"""
    e = layer.engine
    return bool(e._lib.pe_convert_layer_palette_full(e._h, layer._h, outpl, oclamping, osampling, osubspace, tgt_gamma))


def convert_layer_palette(layer, outpl, op_clamping):
    """colourspace.h:393 / colourspace.c:13931"""
    e = layer.engine
    return bool(e._lib.pe_convert_layer_palette(e._h, layer._h, outpl, op_clamping))


def resize_layer_full(layer, width, height, interp, opal_hint, oclamp_hint, osamp_hint, osubs_hint, tgt_gamma):
    """colourspace.h:409 / colourspace.c:14759"""
    e = layer.engine
    return bool(e._lib.pe_resize_layer_full(e._h, layer._h, width, height, interp, opal_hint, oclamp_hint, osamp_hint,
                                            osubs_hint, tgt_gamma))


def resize_layer(layer, width, height, interp, opal_hint, oclamp_hint):
    """colourspace.h:413 / colourspace.c:15331"""
    e = layer.engine
    return bool(e._lib.pe_resize_layer(e._h, layer._h, width, height, interp, opal_hint, oclamp_hint))


def resize_layer_batch(layers, width, height, interp, opal_hint, oclamp_hint):
    """resize_layer over a batch of independent layers (the render loop of src/events.c:4239-4253): one call, launches back to
    back; returns how many layers were resized"""
    e = layers[0].engine
    return int(e._lib.pe_resize_layer_batch(e._h, len(layers), _arr(layers), width, height, interp, opal_hint, oclamp_hint))


def convert_layer_palette_batch(layers, outpl, op_clamping):
    """convert_layer_palette over a batch of independent layers; returns how many were converted"""
    e = layers[0].engine
    return int(e._lib.pe_convert_layer_palette_batch(e._h, len(layers), _arr(layers), outpl, op_clamping))


def convert_yuv888_to_rgb_float(layer, outpl, mode=1, want_sums=False):
    """the reference's float ("experimental") YUV -> RGB path (colourspace.c:2367 yuv2rgb_float; BT.709 only).  mode 0: as written,
    1: the RGBf_Y form.  Returns TRUE / FALSE, or (ok, sums[h, w, 3] float32) with want_sums"""
    e = layer.engine
    sums = np.zeros((layer.height, layer.width, 3), np.float32) if want_sums else None
    ok = bool(e._lib.pe_convert_yuv888_to_rgb_float(e._h, layer._h, outpl, mode, sums.ctypes.data if want_sums else None))
    return (ok, sums) if want_sums else ok


def float_yuv_table(clamping, which):
    t = np.zeros(256, np.float32)
    capi.check(capi.lib().pe_float_yuv_table(clamping, which, t.ctypes.data))
    return t


def letterbox_layer(layer, nwidth, nheight, width, height, interp, tpal, tclamp):
    """colourspace.h:415 / colourspace.c:15343"""
    e = layer.engine
    return bool(e._lib.pe_letterbox_layer(e._h, layer._h, nwidth, nheight, width, height, interp, tpal, tclamp))


def gamma_convert_layer(gamma_type, layer):
    """colourspace.h:389 / colourspace.c:14146"""
    e = layer.engine
    return bool(e._lib.pe_gamma_convert_layer(e._h, gamma_type, layer._h))


def gamma_convert_sub_layer(gamma_type, fileg, layer, x, y, width, height, may_thread=True):
    """colourspace.h:391 / colourspace.c:14069"""
    e = layer.engine
    return bool(e._lib.pe_gamma_convert_sub_layer(e._h, gamma_type, fileg, layer._h, x, y, width, height, int(may_thread)))


def alpha_premult(layer, direction):
    """colourspace.h:387 / colourspace.c:11968"""
    e = layer.engine
    e._lib.pe_alpha_premult(e._h, layer._h, direction)


# ---- boundary B1 arithmetic: effect process functions -------------------------------------------------------------

SIMPLE_BLEND_TYPES = {"chroma blend": 0, "luma overlay": 1, "luma underlay": 2, "negative luma overlay": 3, "averaged luma overlay": 4}
MULTI_BLEND_TYPES = {"blend_multiply": 0, "blend_screen": 1, "blend_darken": 2, "blend_lighten": 3, "blend_overlay": 4,
                     "blend_dodge": 5, "blend_burn": 6}


def _arr(layers):
    a = (C.c_void_p * len(layers))()
    for i, l in enumerate(layers):
        a[i] = l._h
    return a


def simple_blend(filter_type, in1, in2, out, blend_factor):
    """simple_blend.c common_process :58 (weed process_func); out may be in1 (in place)"""
    t = SIMPLE_BLEND_TYPES.get(filter_type, filter_type)
    e = in1.engine
    capi.check(e._lib.pe_fx_simple_blend(e._h, t, in1._h, in2._h, out._h, blend_factor))


def simple_blend_batch(filter_type, in1, in2, out, blend_factor):
    """a batch of independent frames (render-to-disk / multitrack) in one launch"""
    t = SIMPLE_BLEND_TYPES.get(filter_type, filter_type)
    e = in1[0].engine
    capi.check(e._lib.pe_fx_simple_blend_batch(e._h, t, len(in1), _arr(in1), _arr(in2), _arr(out), blend_factor))


def multi_blend(filter_type, in1, in2, out, blend_factor):
    """multi_blends.c common_process :26"""
    t = MULTI_BLEND_TYPES.get(filter_type, filter_type)
    e = in1.engine
    capi.check(e._lib.pe_fx_multi_blend(e._h, t, in1._h, in2._h, out._h, blend_factor))


SLIDE_DIRECTIONS = {"dir_r2l": 1, "dir_l2r": 2, "dir_b2t": 3, "dir_t2b": 4}  # parameter name -> plugin_direction (sover_init :38-52)


def slide_over(in1, in2, out, transval, direction, mvlower=True, mvupper=False):
    """slide_over.c sover_process :55; direction 1 .. 4 or the name of the radio parameter that selects it; the defaults of
    mvlower / mvupper are the plugin's (:170-171)"""
    d = SLIDE_DIRECTIONS.get(direction, direction)
    e = in1.engine
    capi.check(e._lib.pe_fx_slide_over(e._h, in1._h, in2._h, out._h, transval, d, int(bool(mvlower)), int(bool(mvupper))))


def slide_over_bound(direction, transval, width, height):
    """the dividing line (rows / macropixels) as the reference's build computes it; host arithmetic only"""
    return capi.lib().pe_fx_slide_over_bound(SLIDE_DIRECTIONS.get(direction, direction), transval, width, height)


def softlight(in_layer, out):
    """softlight.c softlight_process :62 (planar YUV: the luma plane is filtered, the other planes are copied)"""
    e = in_layer.engine
    capi.check(e._lib.pe_fx_softlight(e._h, in_layer._h, out._h))


def triple_split(in1, in2, out, start=0.666667, sym=True, end=0.333333, vert=False, borderw=0., bordercol=(0, 0, 0)):
    """layout_blends.c common_process :19 ("triple split"); the defaults are the plugin's parameter templates :137-144"""
    e = in1.engine
    capi.check(e._lib.pe_fx_triple_split(e._h, in1._h, in2._h, out._h, start, int(bool(sym)), end, int(bool(vert)), borderw,
                                         (C.c_int * 3)(*bordercol)))


MULTI_TRANSITION_TYPES = {"iris rectangle": 0, "iris circle": 1, "4 way split": 2, "dissolve": 3}


class DissolveMask:
    """the per-instance mask of multi_transitions.c dissolve_init :42, drawn from the host's random seed"""

    def __init__(self, engine, width, height, random_seed):
        self.engine = engine
        self._h = C.c_void_p()
        capi.check(engine._lib.pe_fx_dissolve_mask_create(engine._h, width, height, random_seed, C.byref(self._h)))

    def close(self):
        if self._h:
            self.engine._lib.pe_fx_dissolve_mask_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def multi_transition(filter_type, in1, in2, out, amount, mask=None):
    """multi_transitions.c common_process :85: "iris rectangle", "iris circle", "4 way split", "dissolve" (needs a DissolveMask)"""
    t = MULTI_TRANSITION_TYPES.get(filter_type, filter_type)
    e = in1.engine
    capi.check(e._lib.pe_fx_multi_transition(e._h, t, in1._h, in2._h, out._h, amount, mask._h if mask is not None else None))


def compositor(out, layers, alphas, bgcol=(0, 0, 0)):
    """gdk/compositor.c compositor_process :127 at scale 1 / offset 0"""
    e = out.engine
    al = (C.c_double * max(len(alphas), 1))(*alphas)
    bg = (C.c_int * 3)(*bgcol)
    capi.check(e._lib.pe_fx_compositor(e._h, out._h, _arr(layers), al, len(layers), bg))


def compositor_gamma(out, layers, alphas, gamma_to, bgcol=(0, 0, 0)):
    """compositor() followed by gamma_convert_layer(gamma_to, out), the LUT folded into the last paint"""
    e = out.engine
    al = (C.c_double * max(len(alphas), 1))(*alphas)
    bg = (C.c_int * 3)(*bgcol)
    capi.check(e._lib.pe_fx_compositor_gamma(e._h, out._h, _arr(layers), al, len(layers), bg, gamma_to))


def compositor_gamma_batch(outs, layers_per_frame, alphas, gamma_to, bgcol=(0, 0, 0)):
    """compositor_gamma() for a batch of independent output frames: layers_per_frame[i] = the layers of frame i (same count and
    per-layer alphas for every frame); returns how many frames were composited"""
    e = outs[0].engine
    nl = len(alphas)
    flat = [l for ls in layers_per_frame for l in ls]
    assert len(flat) == nl * len(outs)
    al = (C.c_double * max(nl, 1))(*alphas)
    bg = (C.c_int * 3)(*bgcol)
    return int(e._lib.pe_fx_compositor_gamma_batch(e._h, len(outs), _arr(outs), _arr(flat), al, nl, bg, gamma_to))


def fused_convert_letterbox_over_gamma(fg, bg, out, inner_w, inner_h, alpha, gamma_from, gamma_to):
    """convert_layer_palette(fg -> RGBA32); letterbox_layer; compositor over bg; gamma_convert_layer -- one kernel"""
    e = fg.engine
    capi.check(e._lib.pe_fused_convert_letterbox_over_gamma(e._h, fg._h, bg._h, out._h, inner_w, inner_h, alpha, gamma_from,
                                                            gamma_to))


def convert_crossfade(clip, operand, outpl, op_clamping, blend_factor):
    """convert_layer_palette(clip, outpl) + 'chroma blend' (in1 = converted clip, in2 = operand) -> clip, one kernel"""
    e = clip.engine
    capi.check(e._lib.pe_fx_convert_crossfade(e._h, clip._h, operand._h, outpl, op_clamping, blend_factor))


def convert_crossfade_batch(clips, operand, outpl, op_clamping, blend_factor):
    """convert_crossfade for the clips of a multitrack stack that fade against ONE shared operand (BASELINE config 5): same-shaped
    clips leave as one kernel launch per 32; returns how many clips were converted"""
    e = operand.engine
    return int(e._lib.pe_fx_convert_crossfade_batch(e._h, len(clips), _arr(clips), operand._h, outpl, op_clamping, blend_factor))


def convert_crossfade_batchv(clips, operands, outpl, op_clamping, blend_factor):
    """convert_crossfade with one operand per clip (operands[i] for clips[i]); one kernel launch per 32 same-shaped clips"""
    e = clips[0].engine
    assert len(clips) == len(operands)
    return int(e._lib.pe_fx_convert_crossfade_batchv(e._h, len(clips), _arr(clips), _arr(operands), outpl, op_clamping, blend_factor))


def fused_convert_letterbox_over_gamma_batch(fg, bg, out, inner_w, inner_h, alpha, gamma_from, gamma_to):
    e = fg[0].engine
    capi.check(e._lib.pe_fused_convert_letterbox_over_gamma_batch(e._h, len(fg), _arr(fg), _arr(bg), _arr(out), inner_w,
                                                                  inner_h, alpha, gamma_from, gamma_to))


# ---- host-frame drop-ins (H2D -> op -> D2H); `HostLayer` is the plain-memory weed_layer_t --------------------------

class HostLayer:
    """A frame in HOST memory described the way the reference's weed_layer_t leaves describe it."""

    def __init__(self, palette, width, height, planes, yuv_clamping=0, yuv_sampling=0, yuv_subspace=0, gamma_type=0, flags=0):
        self.planes = list(planes)
        self.d = capi.pe_frame_desc_t()
        d = self.d
        d.palette, d.width, d.height, d.nplanes = palette, width, height, len(planes)
        d.yuv_clamping, d.yuv_sampling, d.yuv_subspace, d.gamma_type, d.flags = yuv_clamping, yuv_sampling, yuv_subspace, gamma_type, flags
        self._sync_ptrs()

    def _sync_ptrs(self):
        for i in range(4):
            if i < len(self.planes):
                self.d.planes[i] = self.planes[i].ctypes.data
                self.d.rowstrides[i] = self.planes[i].strides[0]
            else:
                self.d.planes[i] = None
                self.d.rowstrides[i] = 0
        self.d.nplanes = len(self.planes)


class _NumpyAllocator:
    """pe_host_allocator_t backed by numpy: new pixel buffers handed back by the drop-ins stay owned by Python"""

    def __init__(self):
        self.blocks = {}
        self._alloc = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)(self.alloc)
        self._free = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)(self.free)
        self.c = capi.pe_host_allocator_t(C.cast(self._alloc, C.c_void_p), C.cast(self._free, C.c_void_p), None)

    def alloc(self, nbytes, _user):
        a = np.zeros(nbytes, np.uint8)
        self.blocks[a.ctypes.data] = a
        return a.ctypes.data

    def free(self, ptr, _user):
        self.blocks.pop(ptr, None)  # buffers we did not hand out (the caller's numpy arrays) are simply dropped


def _rebind(host_layer, allocator):
    """after a host drop-in replaced the pixel buffers: rebuild numpy views over the new block"""
    d = host_layer.d
    base = d.planes[0]
    if base in allocator.blocks:
        blk = allocator.blocks[base]
        _, rs, ph, _ = frame_layout(d.palette, d.width, d.height)
        planes, off = [], 0
        for i in range(d.nplanes):
            planes.append(blk[off:off + rs[i] * ph[i]].reshape(ph[i], rs[i]))
            off += rs[i] * ph[i]
        host_layer.planes = planes
        host_layer._keep = blk


def host_convert_layer_palette_full(engine, host_layer, outpl, oclamping, osampling, osubspace, tgt_gamma):
    al = _NumpyAllocator()
    ok = bool(engine._lib.pe_host_convert_layer_palette_full(engine._h, C.byref(host_layer.d), outpl, oclamping, osampling,
                                                            osubspace, tgt_gamma, C.byref(al.c)))
    if ok:
        _rebind(host_layer, al)
    return ok


def host_resize_layer(engine, host_layer, width, height, interp, opal_hint, oclamp_hint):
    al = _NumpyAllocator()
    ok = bool(engine._lib.pe_host_resize_layer(engine._h, C.byref(host_layer.d), width, height, interp, opal_hint,
                                              oclamp_hint, C.byref(al.c)))
    if ok:
        _rebind(host_layer, al)
    return ok


def host_letterbox_layer(engine, host_layer, nwidth, nheight, width, height, interp, tpal, tclamp):
    al = _NumpyAllocator()
    ok = bool(engine._lib.pe_host_letterbox_layer(engine._h, C.byref(host_layer.d), nwidth, nheight, width, height, interp,
                                                 tpal, tclamp, C.byref(al.c)))
    if ok:
        _rebind(host_layer, al)
    return ok


def host_gamma_convert_layer(engine, gamma_type, host_layer):
    return bool(engine._lib.pe_host_gamma_convert_layer(engine._h, gamma_type, C.byref(host_layer.d)))


def host_simple_blend(engine, filter_type, in1, in2, out, blend_factor):
    t = SIMPLE_BLEND_TYPES.get(filter_type, filter_type)
    capi.check(engine._lib.pe_host_simple_blend(engine._h, t, C.byref(in1.d), C.byref(in2.d), C.byref(out.d), blend_factor))


def host_multi_blend(engine, filter_type, in1, in2, out, blend_factor):
    t = MULTI_BLEND_TYPES.get(filter_type, filter_type)
    capi.check(engine._lib.pe_host_multi_blend(engine._h, t, C.byref(in1.d), C.byref(in2.d), C.byref(out.d), blend_factor))


def host_fused_convert_letterbox_over_gamma(engine, fg, bg, out, inner_w, inner_h, alpha, gamma_from, gamma_to):
    capi.check(engine._lib.pe_host_fused_convert_letterbox_over_gamma(engine._h, C.byref(fg.d), C.byref(bg.d), C.byref(out.d),
                                                                      inner_w, inner_h, alpha, gamma_from, gamma_to))


def host_fused_convert_letterbox_over_gamma_batch(engine, fgs, bgs, outs, inner_w, inner_h, alpha, gamma_from, gamma_to):
    """n independent host frames; H2D, kernel and D2H of consecutive frames overlap (three streams)"""
    n = len(fgs)

    def arr(ls):
        a = (capi.PDESC * n)()
        for i, l in enumerate(ls):
            a[i] = C.pointer(l.d)
        return a
    capi.check(engine._lib.pe_host_fused_convert_letterbox_over_gamma_batch(engine._h, n, arr(fgs), arr(bgs), arr(outs), inner_w,
                                                                            inner_h, alpha, gamma_from, gamma_to))


# ---- ingest / egress in device memory ------------------------------------------------------------------

class pe_clip_source_t(C.Structure):
    _fields_ = [("clip_data", C.c_void_p), ("get_frame", C.c_void_p), ("palette", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("yuv_clamping", C.c_int), ("yuv_sampling", C.c_int), ("yuv_subspace", C.c_int), ("gamma_type", C.c_int)]


class ClipCache:
    """A clip resident in HBM (pe_clip_cache_t): the built-in device source behind the decoder plugin's get_frame shape
    (src/plugins.h:442).  load() is the one-time fill; frame() / borrow() never touch host memory."""

    def __init__(self, engine, palette, width, height, nframes, yuv_clamping=0, yuv_sampling=0, yuv_subspace=0, gamma_type=0):
        self.engine = engine
        h = C.c_void_p()
        capi.check(engine._lib.pe_clip_cache_create(engine._h, palette, width, height, nframes, yuv_clamping, yuv_sampling, yuv_subspace,
                                                    gamma_type, C.byref(h)))
        self._h = h
        self.nframes = nframes
        self._src = pe_clip_source_t()
        capi.check(engine._lib.pe_clip_cache_source(self._h, C.byref(self._src)))

    def load(self, frame, planes):
        ptrs, rs = (C.c_void_p * 4)(), (C.c_int * 4)()
        for i, p in enumerate(planes):
            ptrs[i], rs[i] = p.ctypes.data, p.strides[0]
        capi.check(self.engine._lib.pe_clip_cache_load(self._h, frame, ptrs, rs))

    def frame(self, n):
        """pull_frame on the device: a new Layer filled by the source's get_frame (device-to-device)"""
        h = C.c_void_p()
        capi.check(self.engine._lib.pe_ingest_frame(self.engine._h, C.byref(self._src), n, C.byref(h)))
        return Layer(self.engine, h)

    def borrow(self, n):
        """zero copy: the cached frame itself as a read-only Layer"""
        h = C.c_void_p()
        capi.check(self.engine._lib.pe_clip_cache_borrow(self._h, n, C.byref(h)))
        return Layer(self.engine, h)

    def close(self):
        if self._h:
            self.engine._lib.pe_clip_cache_destroy(self._h)
            self._h = None


def render_out(layer, out_palette, host_array):
    """the render tail (src/events.c:4247-4263): convert to out_palette, download only that packed frame"""
    e = layer.engine
    capi.check(e._lib.pe_render_out(e._h, layer._h, out_palette, host_array.ctypes.data, host_array.strides[0]))


def render_out_begin(layer, out_palette, host_array, slot):
    e = layer.engine
    capi.check(e._lib.pe_render_out_begin(e._h, layer._h, out_palette, host_array.ctypes.data, host_array.strides[0], slot))


def render_out_wait(engine, slot):
    capi.check(engine._lib.pe_render_out_wait(engine._h, slot))


# ---- the node model's CONVERT step as one descriptor -------------------------------------------------------

OP_RESIZE, OP_PCONV, OP_GAMMA, OP_LETTERBOX, N_OP_TYPES = 0, 1, 2, 3, 6  # src/nodemodel.h:717-722


class pe_convert_plan_t(C.Structure):
    _fields_ = [("op_order", C.c_int * 6), ("width", C.c_int), ("height", C.c_int), ("lb_width", C.c_int), ("lb_height", C.c_int),
                ("interp", C.c_int), ("out_palette", C.c_int), ("out_clamping", C.c_int), ("out_sampling", C.c_int),
                ("out_subspace", C.c_int), ("out_gamma", C.c_int), ("no_fuse", C.c_int)]


def convert_plan(op_order, width=0, height=0, lb_width=0, lb_height=0, interp=LIVES_INTERP_NORMAL, out_palette=0, out_clamping=0,
                 out_sampling=0, out_subspace=0, out_gamma=0, no_fuse=False):
    """op_order: dict {OP_*: substep} as get_op_order (src/nodemodel.c:161) fills its array"""
    p = pe_convert_plan_t()
    for k, v in op_order.items():
        p.op_order[k] = v
    p.width, p.height, p.lb_width, p.lb_height, p.interp = width, height, lb_width, lb_height, interp
    p.out_palette, p.out_clamping, p.out_sampling, p.out_subspace, p.out_gamma, p.no_fuse = (out_palette, out_clamping, out_sampling,
                                                                                             out_subspace, out_gamma, int(no_fuse))
    return p


def run_convert_plan(layer, plan):
    return bool(layer.engine._lib.pe_run_convert_plan(layer.engine._h, layer._h, C.byref(plan)))


def run_convert_plan_over(fg, plan, bg, out, alpha, gamma_to):
    """returns 1 when the chain left as one fused launch, 0 when it ran op by op"""
    e = fg.engine
    capi.check(e._lib.pe_run_convert_plan_over(e._h, fg._h, C.byref(plan), bg._h, out._h, alpha, gamma_to))
    return e._lib.pe_last_plan_path()
