"""ctypes loaders + frame helpers shared by the tests.

The oracle (oracle/libpe_oracle.so) and the compiled reference (oracle/_ref/*.so)
are CHECKERS: they are loaded only here, never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
GOLDEN = os.path.join(REPO, "tests", "golden")

PAL = dict(RGB24=1, BGR24=2, RGBA32=3, BGRA32=4, ARGB32=5, YUV420P=512, YVU420P=513, YUV422P=522,
           YUV444P=544, YUVA4444P=545, UYVY=564, YUYV=565, YUV888=588, YUVA8888=589, YUV411=595)
CLAMPED, UNCLAMPED = 0, 1
SUB_YUV, SUB_YCBCR, SUB_BT709 = 0, 1, 2
G_UNKNOWN, G_LINEAR, G_SRGB, G_BT709, G_MONITOR = 0, -1, 1, 2, 1024
Q_LOW, Q_MED, Q_HIGH = 1, 2, 3

u8p = C.POINTER(C.c_uint8)


def p8(a):
    return a.ctypes.data_as(u8p)


def align_ceil(a, b):
    return ((a + b - 1) // b) * b


def rowstride(width, psize):
    """colourspace.c:11299-11342 default rowstride = ALIGN_CEIL(width*psize, 32)"""
    return align_ceil(width * psize, 32)


def psize_of(pal):
    return {1: 3, 2: 3, 3: 4, 4: 4, 5: 4, 588: 3, 589: 4, 564: 4, 565: 4, 595: 6}.get(pal, 1)


I, L, D, VP = C.c_int, C.c_long, C.c_double, C.c_void_p

_ORACLE_PROTOS = {
    "pe_or_conv_table": [I, I, I, VP],
    "pe_or_premult_table": [I, VP],
    "pe_or_gamma_lut8": [D, I, I, D, VP],
    "pe_or_gamma_lut16": [D, I, I, D, VP],
    "pe_or_rgb2yuv": [I, I, I, VP, VP, L],
    "pe_or_yuv2rgb": [I, I, I, VP, VP, L],
    "pe_or_yuv420p_to_rgb": [VP, VP, I, I, VP, I, I, I, I, I, I, I, I, VP],
    "pe_or_packed422_to_rgb": [I, VP, I, I, I, VP, I, I, I, I, I, I],
    "pe_or_yuv888_to_rgb": [VP, I, I, I, VP, I, I, I, I, I, I, I],
    "pe_or_rgb_to_yuv888": [VP, I, I, I, VP, I, I, I, I, I, I],
    "pe_or_rgb_to_rgb": [I, I, VP, I, I, I, VP, I, VP],
    "pe_or_rgb_to_packed422": [I, VP, I, I, I, VP, I, I, I, I, I, VP],
    "pe_or_rgb_to_yuv444p": [VP, I, I, I, VP, I, I, I, I, I, I],
    "pe_or_rgb_to_yuv420p": [VP, I, I, I, VP, VP, I, I, I, I, I, I],
    "pe_or_avg_table": [I, VP],
    "pe_or_gamma_apply": [VP, I, I, I, I, I, I, VP],
    "pe_or_alpha_premult": [VP, I, I, I, I, I, I],
    "pe_or_simple_blend": [I, I, VP, I, VP, I, VP, I, I, I, I, L],
    "pe_or_multi_blend": [I, I, VP, I, VP, I, VP, I, I, I, I],
    "pe_or_alpha_over": [VP, I, VP, I, I, I, I, D],
    "pe_or_fill": [VP, I, I, I, I, I, I, I],
    "pe_or_resize_packed": [VP, I, I, I, VP, I, I, I, I],
    "pe_or_resize_filter": [I, I, I, VP, VP, I],
    "pe_or_resize_filter_sws": [I, I, I, VP, VP, I],
    "pe_or_set_resize_recipe": [I],
    "pe_or_letterbox_packed": [VP, I, I, I, VP, I, I, I, I],
    "pe_or_yuv444p_to_rgb": [VP, I, I, I, VP, I, I, I, I, I, I],
    "pe_or_combine_planes": [VP, I, I, I, VP, I, I, I],
    "pe_or_split_planes": [VP, I, I, I, VP, VP, I, I],
    "pe_or_halve_chroma": [VP, VP, I, I, VP, VP, I],
    "pe_or_double_chroma": [VP, VP, I, I, VP, VP, I],
    "pe_or_packed422_to_yuv422p": [I, VP, I, I, I, VP, VP, I],
    "pe_or_packed422_to_yuv444p": [I, VP, I, I, I, VP, VP, I],
    "pe_or_packed422_to_yuv888": [I, VP, I, I, I, VP, I, I],
    "pe_or_alpha_premult_planar": [VP, VP, I, I, I, I],
    "pe_or_to_yuv411": [I, VP, VP, I, I, VP, I, I],
    "pe_or_rgb_to_yuv411": [VP, I, I, I, VP, I, I, I, I],
    "pe_or_yuv411_to": [I, VP, I, I, I, VP, VP, I, I, I, I, I],
    "pe_or_swab": [VP, I, I, I],
    "pe_or_slide_over_bound": [I, I, I, I],
    "pe_or_slide_over": [I, I, I, I, VP, I, VP, I, VP, I, I, I, I],
    "pe_or_chroma_upsample_packed": [I, VP, VP, I, I, VP, I, I, I, I],
    "pe_or_packed422_to_yuv420p": [I, VP, I, I, I, VP, VP, I],
    "pe_or_yuv888_subsample": [I, VP, I, I, I, I, VP, VP, I],
    "pe_or_quad_chroma": [VP, VP, I, I, VP, I, I, I, I],
    "pe_or_yuv42xp_to_packed422": [I, VP, VP, I, I, I, VP, I],
    "pe_or_yuv444p_to_packed422": [I, VP, I, I, I, VP, I, I],
    "pe_or_yuv444p_to_yuv420p": [VP, VP, I, I, VP, VP, I],
    "pe_or_yy_table": [I, VP],
    "pe_or_switch_clamping_plane": [VP, L, I, I],
}

_REF_PROTOS = {
    "ref_init": [],
    "ref_set_prefs": [I, I, D],
    "ref_get_conv_table": [I, I, I, VP],
    "ref_get_premult_table": [I, VP],
    "ref_get_avg_table": [I, VP],
    "ref_get_yy_table": [I, VP],
    "ref_gamma_lut8": [D, I, I, VP],
    "ref_gamma_lut16": [D, I, I, VP],
    "ref_gamma_consts": [VP],
    "ref_rgb2yuv_bulk": [I, I, VP, VP, L],
    "ref_yuv2rgb_bulk": [I, I, VP, VP, L],
    "ref_yuv420p_to_rgb": [VP, I, I, VP, I, VP, I, I, I, I, I, I, I, I],
    "ref_packed422_to_rgb": [I, VP, I, I, I, I, VP, I, I, I, I],
    "ref_yuv888_to_rgb": [VP, I, I, I, I, VP, I, I, I, I, I],
    "ref_yuv444p_to_rgb": [VP, I, I, I, I, VP, I, I, I, I],
    "ref_rgb_permute": [I, VP, I, I, I, I, VP, VP, I, I],
    "ref_rgb_to_yuv888": [VP, I, I, I, I, VP, I, I, I, I],
    "ref_rgb_to_yuv444p": [VP, I, I, I, I, VP, I, I, I, I],
    "ref_rgb_to_yuv420": [VP, I, I, I, VP, VP, I, I, I, I, I],
    "ref_rgb_to_packed422": [I, VP, I, I, I, I, VP, I, I, I, I, I],
    "ref_alpha_premult": [VP, I, I, I, I, I, I, VP],
    "ref_gamma_apply": [VP, I, I, I, I, I, I, VP],
    "ref_combineplanes": [VP, I, I, I, I, VP, I, I],
    "ref_splitplanes": [VP, I, I, I, VP, VP, I, I],
    "ref_halve_chroma": [VP, I, I, VP, VP, VP, I],
    "ref_double_chroma": [VP, I, I, VP, VP, VP, I],
    "ref_packed422_to_yuv422p": [I, VP, I, I, VP],
    "ref_packed422_to_yuv444p": [I, VP, I, I, I, VP, VP, I],
    "ref_packed422_to_yuv888": [I, VP, I, I, I, I, VP, I],
    "ref_alpha_premult_planar": [VP, VP, I, I, I, I, VP],
    "ref_to_yuv411": [I, VP, I, I, I, VP, I],
    "ref_rgb_to_yuv411": [VP, I, I, I, VP, I, I, I],
    "ref_yuv411_to": [I, VP, I, I, I, VP, I, I, I],
    "ref_swab": [VP, I, I, I],
    "ref_chroma_upsample_packed": [I, VP, I, I, VP, I, VP, I, I, I],
    "ref_packed422_to_yuv420p": [I, VP, I, I, VP, I],
    "ref_yuv888_subsample": [I, VP, I, I, I, VP, VP, I, I],
    "ref_quad_chroma": [VP, I, I, VP, I, VP, I, I, I],
    "ref_yuv420_to_packed422": [I, VP, I, I, VP, I, VP, I],
    "ref_yuv422p_to_packed422": [I, VP, I, I, VP, I, VP],
    "ref_yuv444p_to_packed422": [I, VP, I, I, I, I, VP, I],
    "ref_yuv444p_to_yuv420p": [VP, I, I, VP, VP, VP, I],
}


def _bind(lib, protos):
    for name, args in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    return lib


def ptr(a):
    """raw address of a numpy array (or None)"""
    return None if a is None else a.ctypes.data


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "libpe_oracle.so")
        src = os.path.join(ORACLE_DIR, "pe_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libpe_oracle.so"], stdout=subprocess.DEVNULL)
        _oracle = _bind(C.CDLL(so), _ORACLE_PROTOS)
    return _oracle


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_oracle.so"))


_ref = None


def ref():
    """compiled reference slices (oracle/_ref/libref_oracle.so)"""
    global _ref
    if _ref is None:
        _ref = _bind(C.CDLL(os.path.join(REF_DIR, "libref_oracle.so")), _REF_PROTOS)
        _ref.ref_init()
    return _ref


_paint = None


def ref_paint():
    global _paint
    if _paint is None:
        _paint = C.CDLL(os.path.join(REF_DIR, "ref_paint_pixel.so"))
        _paint.ref_paint_rows.argtypes = [VP, VP, L, I, D]
    return _paint


# ---------------------------------------------------------------- frame makers

def make_packed(rng, width, height, psize, stride=None, lo=0, hi=256):
    stride = stride or rowstride(width, psize)
    a = np.zeros((height, stride), np.uint8)
    a[:, :width * psize] = rng.integers(lo, hi, (height, width * psize), dtype=np.uint8)
    return a


def _plane(rng, rows, cols, stride, lo, hi):
    """one plane + a guard byte right after it holding the replicated edge sample, so that the
    reference's one-past-row chroma read (colourspace.c:3508) sees what our contract defines"""
    buf = np.zeros(rows * stride + 16, np.uint8)
    pl = buf[:rows * stride].reshape(rows, stride)
    pl[:, :cols] = rng.integers(lo, hi, (rows, cols), dtype=np.uint8)
    if stride == cols:
        buf[rows * stride] = pl[rows - 1, cols - 1]
    return pl


def _plane_like(arr):
    """copy of a stored plane with the guard byte of _plane() behind it"""
    rows, stride = arr.shape
    buf = np.zeros(rows * stride + 16, np.uint8)
    pl = buf[:rows * stride].reshape(rows, stride)
    pl[...] = arr
    buf[rows * stride] = pl[rows - 1, stride - 1]
    return pl


def make_yuv_planar(rng, width, height, is_422=False, clamped=True):
    """Planar frame with the reference strides (colourspace.c:11344-11357)."""
    ys = rowstride(width, 1)
    cs = ys >> 1
    cw, ch = width >> 1, (height if is_422 else (height + 1) >> 1)
    if clamped:
        y = _plane(rng, height, width, ys, 16, 236)
        u = _plane(rng, ch, cw, cs, 16, 241)
        v = _plane(rng, ch, cw, cs, 16, 241)
    else:
        y = _plane(rng, height, width, ys, 0, 256)
        u = _plane(rng, ch, cw, cs, 0, 256)
        v = _plane(rng, ch, cw, cs, 0, 256)
    return y, u, v


def planes_arg(*planes):
    """uint8_t *planes[n] (keeps nothing alive: callers hold the arrays)"""
    return (C.c_void_p * len(planes))(*[p.ctypes.data for p in planes])


def strides_arg(*planes):
    return (C.c_int * len(planes))(*[p.strides[0] for p in planes])


# ---------------------------------------------------------------- the mini weed host (tests/host/weed_minihost.c)

class ChanDesc(C.Structure):
    """mh_chan_desc: one channel of mh_run_generic"""
    _fields_ = [("palette", I), ("width", I), ("height", I), ("nplanes", I), ("yuv_clamping", I), ("planes", VP * 4), ("rowstrides", I * 4)]


def chan(palette, width, height, planes, clamping=-1):
    d = ChanDesc()
    d.palette, d.width, d.height, d.nplanes, d.yuv_clamping = palette, width, height, len(planes), clamping
    for i, p in enumerate(planes):
        d.planes[i] = p.ctypes.data
        d.rowstrides[i] = p.strides[0]
    return d


_mh = None


def minihost():
    global _mh
    if _mh is None:
        _mh = C.CDLL(os.path.join(REF_DIR, "libweed_minihost.so"))
        _mh.mh_open.argtypes = [C.c_char_p]
        _mh.mh_run_generic.argtypes = [I, I, I, C.POINTER(ChanDesc), C.POINTER(ChanDesc), I, C.POINTER(D), C.POINTER(I), C.c_longlong, I]
    return _mh


def mh_run(h, fidx, ins, out, params=(), seed=1, nframes=1):
    """run filter fidx of plugin handle h: ins / out = ChanDesc, params = one list of numbers per in-parameter template"""
    mh = minihost()
    flat = [float(v) for p in params for v in p]
    counts = [len(p) for p in params]
    return mh.mh_run_generic(h, fidx, len(ins), (ChanDesc * len(ins))(*ins), C.byref(out), len(params), (D * max(1, len(flat)))(*flat),
                             (I * max(1, len(counts)))(*counts), seed, nframes)
