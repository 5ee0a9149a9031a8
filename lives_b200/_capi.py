"""ctypes binding of include/pixel_engine.h (libpe_b200.so).

The library is loaded lazily and loudly: if it is missing or cannot be loaded there is no fallback --
`PixelEngineUnavailable` is raised (the product path never routes through oracle/ or a CPU loop).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpe_b200.so")

PE_MAXPLANES = 4
PE_TRUE, PE_FALSE = 1, 0
PE_OK, PE_ERR_CUDA, PE_ERR_ARG, PE_ERR_PALETTE, PE_ERR_MEMORY, PE_ERR_SIZE = 0, 1, 2, 3, 4, 5


class PixelEngineUnavailable(RuntimeError):
    pass


class PixelEngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pixel engine error %d: %s" % (code, msg))
        self.code = code


class pe_config_t(C.Structure):
    _fields_ = [("device", C.c_int), ("pb_quality", C.c_int), ("screen_gamma", C.c_double), ("apply_gamma", C.c_int),
                ("alpha_post", C.c_int), ("ref_quirks", C.c_int), ("stream", C.c_void_p)]


class pe_frame_desc_t(C.Structure):
    _fields_ = [("palette", C.c_int), ("width", C.c_int), ("height", C.c_int), ("nplanes", C.c_int),
                ("rowstrides", C.c_int * PE_MAXPLANES), ("planes", C.c_void_p * PE_MAXPLANES),
                ("yuv_clamping", C.c_int), ("yuv_sampling", C.c_int), ("yuv_subspace", C.c_int),
                ("gamma_type", C.c_int), ("flags", C.c_int)]


class pe_frame_stats_t(C.Structure):
    _fields_ = [("min", C.c_uint8 * 4), ("max", C.c_uint8 * 4), ("hist", C.c_uint32 * 256), ("sum", C.c_uint64),
                ("all_black_ish", C.c_int), ("all_black", C.c_int)]


class pe_host_allocator_t(C.Structure):
    _fields_ = [("alloc", C.c_void_p), ("free", C.c_void_p), ("user", C.c_void_p)]


I, D, VP, L, SZ = C.c_int, C.c_double, C.c_void_p, C.c_long, C.c_size_t
PVP = C.POINTER(C.c_void_p)
PI = C.POINTER(C.c_int)
PDESC = C.POINTER(pe_frame_desc_t)

# name -> (restype, argtypes); every symbol include/pixel_engine.h declares
PROTOTYPES = {
    "pe_config_default": (None, [C.POINTER(pe_config_t)]),
    "pe_engine_create": (I, [C.POINTER(pe_config_t), PVP]),
    "pe_engine_destroy": (None, [VP]),
    "pe_engine_sync": (I, [VP]),
    "pe_engine_stream": (VP, [VP]),
    "pe_last_error": (C.c_char_p, []),
    "pe_engine_launch_count": (L, [VP]),
    "pe_engine_shared_configure": (I, [C.c_void_p]),
    "pe_engine_shared": (VP, []),
    "pe_engine_set_prefs": (I, [VP, I, C.c_double, I, I]),
    "pe_engine_get_config": (I, [VP, C.c_void_p]),
    "pe_host_register": (I, [VP, SZ]),
    "pe_run_convert_plan": (I, [VP, VP, C.c_void_p]),
    "pe_run_convert_plan_over": (I, [VP, VP, C.c_void_p, VP, VP, C.c_double, I]),
    "pe_last_plan_path": (I, []),
    "pe_ingest_frame": (I, [VP, C.c_void_p, C.c_int64, PVP]),
    "pe_clip_cache_create": (I, [VP, I, I, I, I, I, I, I, I, PVP]),
    "pe_clip_cache_destroy": (None, [VP]),
    "pe_clip_cache_load": (I, [VP, C.c_int64, C.c_void_p, C.c_void_p]),
    "pe_clip_cache_frame_desc": (I, [VP, C.c_int64, PDESC]),
    "pe_clip_cache_source": (I, [VP, C.c_void_p]),
    "pe_clip_cache_borrow": (I, [VP, C.c_int64, PVP]),
    "pe_render_out_begin": (I, [VP, VP, I, VP, I, I]),
    "pe_render_out_wait": (I, [VP, I]),
    "pe_render_out": (I, [VP, VP, I, VP, I]),
    "pe_host_compositor": (I, [VP, C.c_void_p, C.c_void_p, C.c_void_p, I, C.c_void_p]),
    "pe_host_unregister": (I, [VP]),
    "pe_timer_start": (I, [VP]),
    "pe_timer_stop_ms": (I, [VP, C.POINTER(C.c_float)]),
    "pe_sm_count": (I, [VP]),
    "pe_engine_set_sm_limit": (I, [VP, I]),
    "pe_engine_set_resize_recipe": (I, [VP, I]),
    "pe_resize_filter_host": (I, [I, I, I, I, VP, VP, I]),
    "pe_avg_closed_form": (I, [I, VP]),
    "pe_host_parallel_copy2d": (I, [VP, C.c_size_t, VP, C.c_size_t, C.c_size_t, C.c_size_t, I]),
    "pe_frame_layout": (SZ, [I, I, I, PI, PI, PI]),
    "pe_frame_create": (I, [VP, I, I, I, I, I, I, I, I, PVP]),
    "pe_frame_wrap": (I, [VP, PDESC, PVP]),
    "pe_frame_destroy": (None, [VP]),
    "pe_frame_get_desc": (I, [VP, PDESC]),
    "pe_frame_set_gamma": (I, [VP, I]),
    "pe_frame_set_flags": (I, [VP, I]),
    "pe_frame_upload": (I, [VP, VP, PVP, PI]),
    "pe_frame_download": (I, [VP, VP, PVP, PI]),
    "pe_frame_copy": (I, [VP, VP, PVP]),
    "pe_host_alloc": (VP, [SZ]),
    "pe_host_free": (None, [VP]),
    "pe_convert_layer_palette_full": (I, [VP, VP, I, I, I, I, I]),
    "pe_convert_layer_palette": (I, [VP, VP, I, I]),
    "pe_resize_layer_full": (I, [VP, VP, I, I, I, I, I, I, I, I]),
    "pe_resize_layer": (I, [VP, VP, I, I, I, I, I]),
    "pe_resize_layer_batch": (I, [VP, I, VP, I, I, I, I, I]),
    "pe_fx_convert_crossfade": (I, [VP, VP, VP, I, I, I]),
    "pe_fx_convert_crossfade_batch": (I, [VP, I, VP, VP, I, I, I]),
    "pe_fx_convert_crossfade_batchv": (I, [VP, I, VP, VP, I, I, I]),
    "pe_mc_publish": (I, [VP, VP, VP, C.c_size_t, VP, I]),
    "pe_convert_layer_palette_batch": (I, [VP, I, VP, I, I]),
    "pe_letterbox_layer": (I, [VP, VP, I, I, I, I, I, I, I]),
    "pe_gamma_convert_layer": (I, [VP, I, VP]),
    "pe_gamma_convert_sub_layer": (I, [VP, I, D, VP, I, I, I, I, I]),
    "pe_alpha_premult": (None, [VP, VP, I]),
    "pe_gamma_lut8": (I, [VP, D, I, I, VP]),
    "pe_fx_simple_blend": (I, [VP, I, VP, VP, VP, I]),
    "pe_fx_multi_blend": (I, [VP, I, VP, VP, VP, I]),
    "pe_fx_slide_over": (I, [VP, VP, VP, VP, I, I, I, I]),
    "pe_fx_slide_over_bound": (I, [I, I, I, I]),
    "pe_convert_yuv888_to_rgb_float": (I, [VP, VP, I, I, VP]),
    "pe_float_yuv_table": (I, [I, I, VP]),
    "pe_fx_softlight": (I, [VP, VP, VP]),
    "pe_fx_triple_split": (I, [VP, VP, VP, VP, D, I, D, I, D, PI]),
    "pe_fx_triple_split_classes": (None, [I, I, D, I, D, I, D, VP, VP]),
    "pe_fx_dissolve_mask_create": (I, [VP, I, I, C.c_int64, C.POINTER(VP)]),
    "pe_fx_dissolve_mask_destroy": (None, [VP]),
    "pe_fx_multi_transition": (I, [VP, I, VP, VP, VP, D, VP]),
    "pe_host_softlight": (I, [VP, PDESC, PDESC]),
    "pe_host_triple_split": (I, [VP, PDESC, PDESC, PDESC, D, I, D, I, D, PI]),
    "pe_host_multi_transition": (I, [VP, I, PDESC, PDESC, PDESC, D, VP]),
    "pe_fx_compositor": (I, [VP, VP, PVP, C.POINTER(D), I, PI]),
    "pe_fx_compositor_gamma": (I, [VP, VP, PVP, C.POINTER(D), I, PI, I]),
    "pe_fx_compositor_gamma_batch": (I, [VP, I, PVP, PVP, C.POINTER(D), I, PI, I]),
    "pe_fx_simple_blend_batch": (I, [VP, I, I, PVP, PVP, PVP, I]),
    "pe_fused_convert_letterbox_over_gamma": (I, [VP, VP, VP, VP, I, I, D, I, I]),
    "pe_fused_convert_letterbox_over_gamma_batch": (I, [VP, I, PVP, PVP, PVP, I, I, D, I, I]),
    "pe_frame_stats": (I, [VP, VP, C.POINTER(pe_frame_stats_t)]),
    "pe_frame_row_hashes": (I, [VP, VP, I, C.c_void_p, C.c_void_p]),
    "pe_host_convert_layer_palette_full": (I, [VP, PDESC, I, I, I, I, I, VP]),
    "pe_host_resize_layer": (I, [VP, PDESC, I, I, I, I, I, VP]),
    "pe_host_letterbox_layer": (I, [VP, PDESC, I, I, I, I, I, I, I, VP]),
    "pe_host_gamma_convert_layer": (I, [VP, I, PDESC]),
    "pe_host_simple_blend": (I, [VP, I, PDESC, PDESC, PDESC, I]),
    "pe_host_multi_blend": (I, [VP, I, PDESC, PDESC, PDESC, I]),
    "pe_host_slide_over": (I, [VP, PDESC, PDESC, PDESC, I, I, I, I]),
    "pe_host_fused_convert_letterbox_over_gamma": (I, [VP, PDESC, PDESC, PDESC, I, I, D, I, I]),
    "pe_host_fused_convert_letterbox_over_gamma_batch": (I, [VP, I, C.POINTER(PDESC), C.POINTER(PDESC), C.POINTER(PDESC), I, I, D, I, I]),
}

_lib = None


def lib():
    """the loaded C-ABI library; raises PixelEngineUnavailable when it cannot be loaded"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PixelEngineUnavailable(
                "%s is missing: build it with `python -m lives_b200.build` (needs nvcc). There is no CPU fallback." % LIB_PATH)
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as exc:
            raise PixelEngineUnavailable("cannot load %s: %s" % (LIB_PATH, exc)) from exc
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError = header / library mismatch: loud
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    return lib().pe_last_error().decode("utf-8", "replace")


def check(rc):
    if rc != PE_OK:
        raise PixelEngineError(rc, last_error())
