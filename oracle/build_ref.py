#!/usr/bin/env python3
"""Build the REFERENCE-side checker binaries into oracle/_ref/ (git-ignored).

TEST INFRASTRUCTURE ONLY.  Nothing built here is linked into, imported by, or
called from the product (lives_b200/).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may load these files.

What it does (needs /root/reference; on the GPU box the prebuilt files travel
with the repo snapshot and this script is a no-op):

  1. libref_oracle.so  -- line-range slices of the reference's
     src/colourspace.c + colourspace.h (tables, gamma LUT builders, per-pixel
     macros and every convert_*_frame loop), compiled where they lie with
     oracle/ref_shim.h supplying the LiVES names they expect, plus
     oracle/ref_wrappers.c (ours) exporting a flat C ABI for ctypes.
     The slices are written to oracle/_ref/src/ only; no reference text is
     ever copied into tracked files.
  2. libweed.so / libweed-utils.so / libweed-host-utils.so -- the reference's
     libweed/*.c compiled unchanged.
  3. simple_blend.so / multi_blends.so / slide_over.so / softlight.so / layout_blends.so / multi_transitions.so -- the reference's effect plugins
     compiled unchanged against those libs (flags per
     lives-plugins/weed-plugins/Makefile.am).
  4. ref_paint_pixel.so -- compositor.c's paint_pixel() (gdk is not available,
     so only that function is sliced out) + our 3-line exporter.
  5. weed_minihost -- tests/host/weed_minihost.c (ours) linked against the real
     libweed: dlopen()s any weed plugin and runs process_func on raw frames.
"""
import os
import subprocess
import sys

REF = os.environ.get("PE_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_ref")
SRC = os.path.join(OUT, "src")

# (first, last) inclusive 1-based line ranges, see SURVEY.md appendix B
CS_H_SLICES = [(12, 30), (43, 63), (65, 131), (145, 258)]
CS_C_SLICES = [
    (52, 420),      # static tables, init_average, set_conversion_arrays, accessors
    (575, 575),     # unal_inited
    (592, 1160),    # gamma LUT builders, spc_rnd, table inits, init_unal
    (1985, 2478),   # thread-fn forward decls, per-pixel kernels
    (2750, 10875),  # every convert_*_frame loop
    (11968, 12106), # alpha_premult
    (14034, 14062), # gamma_convert_layer_thread
]
MATHS_LINES = [88, 101, 104, 118]


def sh(cmd, **kw):
    print("+", " ".join(cmd))
    subprocess.check_call(cmd, **kw)


def slice_file(path, ranges):
    with open(path, "r", errors="replace") as f:
        lines = f.readlines()
    out = []
    for a, b in ranges:
        out.append("/* ---- %s:%d-%d ---- */\n" % (os.path.basename(path), a, b))
        out.extend(lines[a - 1:b])
        out.append("\n")
    return "".join(out)


def build_colourspace():
    cs_c = os.path.join(REF, "src", "colourspace.c")
    cs_h = os.path.join(REF, "src", "colourspace.h")
    maths = os.path.join(REF, "src", "maths.h")
    with open(maths) as f:
        ml = f.readlines()
    with open(os.path.join(SRC, "ref_maths_slice.h"), "w") as f:
        for n in MATHS_LINES:
            f.write(ml[n - 1])
    tu = ['#include "ref_shim.h"\n',
          slice_file(cs_h, CS_H_SLICES),
          "static __thread struct _conv_array ref_tls_conv_arrays;\n",
          "ref_prefs_t ref_prefs_obj = {1, PB_QUALITY_HIGH, 1.4, TRUE, FALSE};\n",
          "ref_prefs_t *prefs = &ref_prefs_obj, *future_prefs = &ref_prefs_obj;\n",
          "ref_mainw_t ref_mainw_obj; ref_mainw_t *mainw = &ref_mainw_obj;\n",
          "#define USE_THREADS 1\n",
          "double weed_palette_get_bytes_per_macropixel(int pal);\n",
          "#define pixel_size(pal) ((int)weed_palette_get_bytes_per_macropixel(pal))\n",
          slice_file(cs_c, CS_C_SLICES),
          '#include "ref_wrappers.c"\n']
    with open(os.path.join(SRC, "ref_colourspace_tu.c"), "w") as f:
        f.write("".join(tu))
    so = os.path.join(OUT, "libref_oracle.so")
    sh(["gcc", "-O3", "-fPIC", "-shared", "-w", "-pthread",
        "-I", HERE, "-I", SRC, "-I", os.path.join(REF, "libweed"),
        os.path.join(SRC, "ref_colourspace_tu.c"), "-o", so, "-lm"])
    return so


def build_libweed():
    lw = os.path.join(REF, "libweed")
    inc = os.path.join(SRC, "inc")
    wdir = os.path.join(inc, "weed")
    os.makedirs(os.path.join(wdir, "weed-plugin-utils"), exist_ok=True)
    for h in os.listdir(lw):
        if h.endswith(".h"):
            dst = os.path.join(wdir, h)
            if not os.path.lexists(dst):
                os.symlink(os.path.join(lw, h), dst)
    dst = os.path.join(wdir, "weed-plugin-utils", "weed-plugin-utils.c")
    if not os.path.lexists(dst):
        os.symlink(os.path.join(lw, "weed-plugin-utils.c"), dst)
    common = ["gcc", "-O2", "-fPIC", "-shared", "-w", "-pthread", "-D_BUILD_LOCAL_", "-I", lw, "-I", inc]
    sh(common + [os.path.join(lw, "weed.c"), "-o", os.path.join(OUT, "libweed.so")])
    sh(common + [os.path.join(lw, "weed-utils.c"), "-o", os.path.join(OUT, "libweed-utils.so"),
                 "-L", OUT, "-lweed", "-Wl,-rpath,$ORIGIN"])
    sh(common + [os.path.join(lw, "weed-host-utils.c"), "-o", os.path.join(OUT, "libweed-host-utils.so"),
                 "-L", OUT, "-lweed", "-lweed-utils", "-Wl,-rpath,$ORIGIN"])
    return inc


def build_plugins(inc):
    pdir = os.path.join(REF, "lives-plugins", "weed-plugins")
    for name in ("simple_blend", "multi_blends", "slide_over", "softlight", "layout_blends", "multi_transitions"):
        sh(["gcc", "-O3", "-fPIC", "-shared", "-w", "-ffast-math", "-fno-math-errno",
            "-I", inc, os.path.join(pdir, name + ".c"),
            "-o", os.path.join(OUT, name + ".so"),
            "-L", OUT, "-lweed-utils", "-lweed", "-lm", "-Wl,-rpath,$ORIGIN"])
    # paint_pixel: compositor.c needs gdk-pixbuf (absent) -> slice the one function
    comp = os.path.join(pdir, "gdk", "compositor.c")
    with open(os.path.join(SRC, "ref_paint_pixel_tu.c"), "w") as f:
        f.write(slice_file(comp, [(120, 125)]))
        f.write("void ref_paint_pixel(unsigned char *dst, int dof, unsigned char *src, int sof, double alpha)"
                " { paint_pixel(dst, dof, src, sof, alpha); }\n"
                "void ref_paint_rows(unsigned char *dst, unsigned char *src, long npix, int psize, double alpha)"
                " { for (long i = 0; i < npix; i++) paint_pixel(dst, (int)(i * psize), src, (int)(i * psize), alpha); }\n")
    sh(["gcc", "-O3", "-fPIC", "-shared", "-w", "-ffast-math", "-fno-math-errno",
        os.path.join(SRC, "ref_paint_pixel_tu.c"), "-o", os.path.join(OUT, "ref_paint_pixel.so")])


def build_diag():
    """per-frame diagnostics: is_all_black_ish (colourspace.c:2554-2594, both branches) and the row hash of hash_cmp_layer (:16044:
    minimd5 of src/maths.c:473-585 -- an MD5 variant of the reference's own, see its BX macro, src/maths.h:42-57)"""
    cs_c = os.path.join(REF, "src", "colourspace.c")
    mc, mh = os.path.join(REF, "src", "maths.c"), os.path.join(REF, "src", "maths.h")
    tu = ["#include <stdint.h>\n#include <stddef.h>\n#include <string.h>\n#include <stdlib.h>\n",
          "typedef int boolean;\n#define TRUE 1\n#define FALSE 0\n#define LIVES_LOCAL_INLINE static inline\n#define LIVES_HOT\n"
          "#define lives_memcpy memcpy\n#define lives_calloc calloc\n",
          slice_file(mh, [(42, 57)]), slice_file(mc, [(473, 546), (575, 585)]), slice_file(cs_c, [(2554, 2594)]),
          "int ref_is_all_black_ish(int width, int height, int rstride, int has_alpha, const uint8_t *pdata, int exact)"
          " { return is_all_black_ish(width, height, rstride, has_alpha, pdata, exact); }\n"
          "void ref_row_hashes(const uint8_t *pd, int nbytes, int height, int rowstride, uint64_t *out)"
          " { for (int i = 0; i < height; i++) out[i] = minimd5((void *)&pd[(size_t)rowstride * i], nbytes); }\n"]
    with open(os.path.join(SRC, "ref_diag_tu.c"), "w") as f:
        f.write("".join(tu))
    sh(["gcc", "-O2", "-fPIC", "-shared", "-w", os.path.join(SRC, "ref_diag_tu.c"), "-o", os.path.join(OUT, "ref_diag.so")])


def build_minihost(inc):
    host = os.path.join(REPO, "tests", "host", "weed_minihost.c")
    if not os.path.exists(host):
        return
    sh(["gcc", "-O2", "-fPIC", "-shared", "-w", "-pthread", "-I", inc, host,
        "-o", os.path.join(OUT, "libweed_minihost.so"),
        "-L", OUT, "-lweed-host-utils", "-lweed-utils", "-lweed", "-ldl", "-lm", "-Wl,-rpath,$ORIGIN"])


def main():
    if not os.path.isdir(REF):
        print("reference tree %s not present: keeping prebuilt oracle/_ref as is" % REF)
        return 0
    os.makedirs(SRC, exist_ok=True)
    build_colourspace()
    inc = build_libweed()
    build_plugins(inc)
    build_diag()
    build_minihost(inc)
    print("oracle/_ref built")
    return 0


if __name__ == "__main__":
    sys.exit(main())
