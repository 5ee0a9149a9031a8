"""TEST INFRASTRUCTURE: the real libswscale through ctypes, when one can be found.

The reference hands resizing to libswscale (src/colourspace.c:14711 sws_scale, context :15059, flags :14991-14997, colourspace
details :15079) and neither ships nor pins it (configure.ac:562).  This image has no system FFmpeg, but the opencv-python-headless
wheel bundles one (libswscale 9.1.100 = FFmpeg 8.0 at the time of writing); importing cv2 maps its dependencies, after which the
library loads by path.  Used by tests/test_resize_vs_swscale.py and tools/swscale_distance.py to MEASURE how far the published
resize contract is from what sws_scale produces -- not to pin it bit for bit."""
import ctypes as C
import glob
import os

import numpy as np

SWS_FAST_BILINEAR, SWS_BILINEAR, SWS_BICUBIC = 1, 2, 4
SWS_CS_ITU709, SWS_CS_ITU601 = 1, 5
PIX_FMT = {"yuv420p": 0, "rgb24": 2, "bgr24": 3, "argb": 25, "rgba": 26, "bgra": 28}  # checked against av_get_pix_fmt in load()

_lib = None


def load():
    """(libswscale, version string) or (None, reason)"""
    global _lib
    if _lib is not None:
        return _lib
    try:
        import cv2  # noqa: F401  (maps libavutil & co. of the wheel)
        d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
        av = C.CDLL(glob.glob(os.path.join(d, "libavutil-*"))[0])
        sw = C.CDLL(glob.glob(os.path.join(d, "libswscale-*"))[0])
    except Exception as exc:  # no cv2 / no bundled FFmpeg
        _lib = (None, "no loadable libswscale: %s" % exc)
        return _lib
    av.av_get_pix_fmt.restype, av.av_get_pix_fmt.argtypes = C.c_int, [C.c_char_p]
    for name, val in PIX_FMT.items():
        if av.av_get_pix_fmt(name.encode()) != val:
            _lib = (None, "pixel format numbering differs (%s)" % name)
            return _lib
    sw.swscale_version.restype = C.c_uint
    sw.sws_getContext.restype = C.c_void_p
    sw.sws_getContext.argtypes = [C.c_int] * 7 + [C.c_void_p] * 3
    sw.sws_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    sw.sws_freeContext.argtypes = [C.c_void_p]
    sw.sws_getCoefficients.restype, sw.sws_getCoefficients.argtypes = C.c_void_p, [C.c_int]
    sw.sws_setColorspaceDetails.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    v = sw.swscale_version()
    _lib = (sw, "%d.%d.%d" % (v >> 16, (v >> 8) & 255, v & 255))
    return _lib


def scale(planes, src_fmt, w, h, dst_fmt, dw, dh, dst_rowbytes, flags=SWS_BILINEAR, yuv=None):
    """one whole-frame sws_scale call, as the reference issues it with one thread (colourspace.c:15059, :15228);
    yuv = (bt709, src_unclamped, dst_unclamped) adds the sws_setColorspaceDetails of :15079"""
    sw, _ = load()
    ctx = sw.sws_getContext(w, h, PIX_FMT[src_fmt], dw, dh, PIX_FMT[dst_fmt], flags, None, None, None)
    assert ctx, "sws_getContext failed"
    if yuv is not None:
        co = sw.sws_getCoefficients(SWS_CS_ITU709 if yuv[0] else SWS_CS_ITU601)
        sw.sws_setColorspaceDetails(ctx, co, int(yuv[1]), co, int(yuv[2]), 0, 65536, 65536)
    dst = np.zeros((dh, dst_rowbytes), np.uint8)
    sp = (C.c_void_p * 4)(*([p.ctypes.data for p in planes] + [0] * (4 - len(planes))))
    ss = (C.c_int * 4)(*([p.strides[0] for p in planes] + [0] * (4 - len(planes))))
    dp = (C.c_void_p * 4)(dst.ctypes.data, 0, 0, 0)
    ds = (C.c_int * 4)(dst.strides[0], 0, 0, 0)
    rows = sw.sws_scale(ctx, sp, ss, 0, h, dp, ds)
    sw.sws_freeContext(ctx)
    assert rows == dh, rows
    return dst


class Scaler:
    """a cached SwsContext (the reference keeps its contexts too: sws_getCachedContext, colourspace.c:15059) for repeated calls on
    frames of one geometry -- what bench.py's CPU arm times"""

    def __init__(self, src_fmt, w, h, dst_fmt, dw, dh, flags=SWS_BILINEAR, yuv=None):
        sw, _ = load()
        self.sw, self.h, self.dh = sw, h, dh
        self.ctx = sw.sws_getContext(w, h, PIX_FMT[src_fmt], dw, dh, PIX_FMT[dst_fmt], flags, None, None, None)
        assert self.ctx, "sws_getContext failed"
        if yuv is not None:
            co = sw.sws_getCoefficients(SWS_CS_ITU709 if yuv[0] else SWS_CS_ITU601)
            sw.sws_setColorspaceDetails(self.ctx, co, int(yuv[1]), co, int(yuv[2]), 0, 65536, 65536)

    def run(self, planes, dst):
        sp = (C.c_void_p * 4)(*([p.ctypes.data for p in planes] + [0] * (4 - len(planes))))
        ss = (C.c_int * 4)(*([p.strides[0] for p in planes] + [0] * (4 - len(planes))))
        dp = (C.c_void_p * 4)(dst.ctypes.data, 0, 0, 0)
        ds = (C.c_int * 4)(dst.strides[0], 0, 0, 0)
        rows = self.sw.sws_scale(self.ctx, sp, ss, 0, self.h, dp, ds)
        assert rows == self.dh, rows

    def close(self):
        if self.ctx:
            self.sw.sws_freeContext(self.ctx)
            self.ctx = None
