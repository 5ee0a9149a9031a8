// pe_kernels.h -- launchers of the sm_100a kernels (definitions in pe_kernels_*.cu).
// Plain structs only; no CUDA types leak past this header except cudaStream_t.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pe {

struct Launch {
  cudaStream_t stream;
  int sm_count;
  long *launch_counter;  // host counter, incremented once per kernel launch
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: the launchers keep their opt-in state per CUDA ordinal (the ABI
// allows one engine per GPU in one process), never in a process-global flag
struct PerDevice {
  size_t v[64] = {};
  size_t &cur() { int d = 0; cudaGetDevice(&d); return v[d & 63]; }
};

// packed image view
struct Img {
  uint8_t *p;
  int rs;  // rowstride in bytes
};
struct CImg {
  const uint8_t *p;
  int rs;
};
// planar YUV view
struct Planes {
  const uint8_t *y, *u, *v;
  int rs_y, rs_u, rs_v;
  int cw, ch;  // chroma plane width / height in samples
};

// device-resident tables of one (clamping, subspace) variant: 14 x int[256] (ConvTab order)
struct DevConv {
  const int32_t *t;    // [14][256]
  int min_y, max_y, min_uv, max_uv;
  // YUV -> RGB tables for the planar converters: RGB_Y[256] then R_Cr, G_Cb, G_Cr, B_Cb indexed by the UN-divided chroma
  // sum n = u1 + (u2 >> 1) <= 765 (kExtN entries each):  ext[t][n] = table_t[clamp(third_round(n), lo, hi)], which folds the
  // (int)(n / 3. + .5) rounding and CLAMP16_240 / CLAMP0_255 of colourspace.c:3465-3469 into the lookup.  A plain chroma
  // sample m is looked up at n = 3 * m.
  const int32_t *ext;  // [256 + 4 * kExtN]
};
constexpr int kExtN = 768;

// byte order of an RGB palette: offsets of R,G,B,A inside a pixel (A = -1 when absent)
struct RgbLayout {
  int r, g, b, a, psize;
};

// ---- RGB <-> RGB (colourspace.c:9259-10515) ---------------------------------------------------
cudaError_t launch_rgb_to_rgb(const Launch &L, CImg src, Img dst, int width, int height, RgbLayout in, RgbLayout out,
                              const uint8_t *lut8_dev);
cudaError_t launch_rgb_to_rgb_batch(const Launch &L, const uint8_t *const *srcs, int irow, uint8_t *const *dsts, int orow, int n, int width,
                                    int height, RgbLayout in, RgbLayout out, const uint8_t *lut8_dev);
// ---- gamma LUT on a rectangle (colourspace.c:14034) ---------------------------------------------
cudaError_t launch_lut8_rect(const Launch &L, Img img, RgbLayout lay, int x, int y, int width, int height,
                             const uint8_t *lut8_dev);
// ---- premultiply (colourspace.c:11968) ------------------------------------------------------------
cudaError_t launch_premult_planar(const Launch &L, uint8_t *const planes[4], const int rowstrides[4], int width, int height,
                                  const uint8_t *tab_y, const uint8_t *tab_c);   // YUVA4444P, colourspace.c:12001-12049
cudaError_t launch_premult(const Launch &L, Img img, int width, int height, int coffs, int ncol, int aoffs,
                           const uint8_t *tab0, const uint8_t *tab1, const uint8_t *tab2, int yuva_fwd_quirk);
// ---- planar 4:2:0 / 4:2:2 -> RGB (colourspace.c:3260-5127) -----------------------------------------
struct YuvToRgbArgs {
  Planes src;
  Img dst;
  int width, height;
  RgbLayout out;      // out.a >= 0 -> alpha byte written as 255
  int is_422;
  int clamped;        // chroma clamp range 16..240 vs 0..255
  int low_quality;    // PB_QUALITY_LOW chroma shortcut (RGB order only, colourspace.c:3470)
  int quirks;         // replicate colourspace.c:3461,3544,3600
  DevConv conv;
  const uint16_t *lut16;  // optional inline gamma (xyuv2rgb_with_gamma :2386)
  // optional fused crossfade (simple_blend.c 'chroma blend' with the converted frame as in1): 3-byte output palettes only
  const uint8_t *blend2 = nullptr;  // in2 pixels (same palette and size as dst), nullptr: off
  int blend2_rs = 0, blend_bf = 0;
};
cudaError_t launch_yuv_planar_to_rgb(const Launch &L, const YuvToRgbArgs &a);
// the fast variant (pe_kernels_yuv2.cu: bank-replicated tables, packed chroma sums): width % 4 == 0, aligned planes, no
// PB_QUALITY_LOW, no inline gamma LUT; host_tables = the ConvTables a.conv was uploaded from
struct ConvTables;
bool yuv_planar_fast_ok(const YuvToRgbArgs &a, const ConvTables *host_tables);
cudaError_t launch_yuv_planar_to_rgb_fast(const Launch &L, const YuvToRgbArgs &a);
// n frames that differ only in their pointers (yuv_planar_same_shape), each passing yuv_planar_fast_ok: one launch per 32 frames
bool yuv_planar_same_shape(const YuvToRgbArgs &a, const YuvToRgbArgs &b);
cudaError_t launch_yuv_planar_to_rgb_batch(const Launch &L, const YuvToRgbArgs *frames, int n);
// ---- packed 4:2:2 / 4:4:4 (colourspace.c:6616-7103, :2750-3258, :5700-6239) ---------------------------
cudaError_t launch_packed422_to_rgb(const Launch &L, int fmt, CImg src, Img dst, int width_mpx, int height,
                                    RgbLayout out, DevConv conv);
cudaError_t launch_yuv888_to_rgb(const Launch &L, CImg src, Img dst, int width, int height, int in_alpha,
                                 RgbLayout out, DevConv conv);
cudaError_t launch_rgb_to_yuv888(const Launch &L, CImg src, Img dst, int width, int height, RgbLayout in,
                                 int out_alpha, DevConv conv);
// RGB(A) -> UYVY (fmt 0) / YUYV (fmt 1), colourspace.c:5129-5700; width in pixels (odd last column dropped); lut16 optional
cudaError_t launch_rgb_to_packed422(const Launch &L, int fmt, CImg src, Img dst, int width_px, int height, RgbLayout in, DevConv conv,
                                    const uint16_t *lut16_dev);
// RGB(A) -> YUV444P / YUVA4444P (planes[3] = alpha plane or nullptr), colourspace.c:5786-6240
cudaError_t launch_rgb_to_yuv444p(const Launch &L, CImg src, uint8_t *const planes[4], int orow, int width, int height, RgbLayout in,
                                  DevConv conv);
// RGB(A) -> YUV420P / YUV422P (colourspace.c:6250 / :6385); cavg_dev: the 64 KB averaging table of the output clamping
cudaError_t launch_rgb_to_yuv420p(const Launch &L, CImg src, uint8_t *const planes[3], const int rowstrides[3], int width, int height,
                                  RgbLayout in, int is_422, DevConv conv, const uint8_t *cavg_dev);
// the reference's float ("experimental") YUV -> RGB arithmetic (colourspace.c:2367 yuv2rgb_float; pe_kernels_float.cu): ftab_dev = the five
// float tables [RGBf_Y, Rf_Cr, Gf_Cb, Gf_Cr, Bf_Cb][256], rgb_y_dev = the integer RGB_Y table mode 0 adds them to; sums_dev optional
cudaError_t launch_yuv888_to_rgb_float(const Launch &L, int mode, CImg src, Img dst, int width, int height, int in_alpha, RgbLayout out,
                                       const float *ftab_dev, const int32_t *rgb_y_dev, float *sums_dev);
// YUV411 (IYU1) -> RGB(A) / packed 4:4:4 / planar 4:4:4 / UYVY / YUYV (convert_yuv411_to_*_frame, colourspace.c:8305-8910); target: 0 RGB,
// 1 YUV888 / YUVA8888, 2 YUV444P / YUVA4444P, 3 UYVY, 4 YUYV, 5 YUV422P, 6 YUV420P; cavg_dev: the 64 KB averaging table of the frame's clamping
// YUV -> YUV411: mode 0 UYVY, 1 YUYV, 2 YUV420P, 3 YUV422P, 4 YUV888, 5 YUVA8888, 6 planar 4:4:4 (colourspace.c:7755-8302, :9148)
cudaError_t launch_to_yuv411(const Launch &L, int mode, const uint8_t *const src[3], const int irow[3], int width_mpx, int height, Img dst,
                             const uint8_t *cavg_dev);
// RGB(A) / BGR(A) / ARGB -> YUV411 (colourspace.c:6499-6614): whole macropixels
cudaError_t launch_rgb_to_yuv411(const Launch &L, CImg src, Img dst, int width_mpx, int height, RgbLayout in, DevConv conv);
cudaError_t launch_yuv411_to(const Launch &L, CImg src, int width_mpx, int height, uint8_t *const dst[4], const int orow[4], int target,
                             int alpha, RgbLayout out, int bgr_quirk, DevConv conv, const uint8_t *cavg_dev);
// the owner's side of the multitrack operand exchange (pe_kernels_mc.cu): bytes from local memory through an NVSwitch multicast address
cudaError_t launch_mc_publish(const Launch &L, const void *src, void *mc_dst, size_t bytes, int max_ctas);
// ---- effects ---------------------------------------------------------------------------------------
// ---- YUV <-> YUV family + planar 4:4:4 -> RGB (pe_kernels_yuv3.cu) ---------------------------------------
cudaError_t launch_yuv444p_to_rgb(const Launch &L, const uint8_t *const planes[4], int irow, Img dst, int width, int height, int in_alpha,
                                  RgbLayout out, DevConv conv);
cudaError_t launch_combine_planes(const Launch &L, const uint8_t *const planes[4], int irow, Img dst, int width, int height, int in_alpha,
                                  int out_alpha);
cudaError_t launch_split_planes(const Launch &L, CImg src, uint8_t *const planes[4], const int orows[4], int width, int height,
                                int src_alpha, int dest_alpha);
// dbl 0: convert_halve_chroma (cw x ch source chroma -> (ch + 1) / 2 rows), 1: convert_double_chroma (-> 2 ch rows)
cudaError_t launch_resample_chroma_v(const Launch &L, int dbl, const uint8_t *su, const uint8_t *sv, int irs_u, int irs_v, uint8_t *du,
                                     uint8_t *dv, int ors_u, int ors_v, int cw, int ch, const uint8_t *cavg_dev);
// fmt 0 UYVY 1 YUYV; mode 0 -> planar 4:2:2, 1 -> planar 4:4:4 (+ alpha), 2 -> YUV888 / YUVA8888 (planes[0])
cudaError_t launch_packed422_unpack(const Launch &L, int fmt, int mode, CImg src, uint8_t *const planes[4], const int orows[4], int width_mpx,
                                    int height, int add_alpha, int first_only);
cudaError_t launch_yuv444p_to_packed422(const Launch &L, int fmt, const uint8_t *const planes[3], int irow, Img dst, int width_mpx, int height,
                                        const uint8_t *cavg_dev);
cudaError_t launch_yuv444p_to_chroma420(const Launch &L, const uint8_t *su, const uint8_t *sv, int irs, uint8_t *du, uint8_t *dv, int ors_u,
                                        int ors_v, int cw, int height, const uint8_t *cavg_dev);
cudaError_t launch_planar42x_to_packed422(const Launch &L, int fmt, int is_422, const uint8_t *const planes[3], const int irows[3], Img dst,
                                          int width_mpx, int height);
cudaError_t launch_quad_chroma(const Launch &L, const uint8_t *su, const uint8_t *sv, int irs_u, int irs_v, int ch, uint8_t *du, uint8_t *dv,
                               int ors, int width, int height, int jpeg, int clamped /* which avg_chroma table: closed form, pe_tables.h AvgForm */);
// mode 0 UYVY 1 YUYV (planes[0] = the packed frame) 2 planar 4:2:2 3 planar 4:2:0
cudaError_t launch_yuv888_subsample(const Launch &L, int mode, CImg src, int src_alpha, uint8_t *const planes[3], const int orows[3], int width,
                                    int height, const uint8_t *cavg_dev);
cudaError_t launch_packed422_to_yuv420p(const Launch &L, int fmt, CImg src, uint8_t *const planes[3], const int orows[3], int width_mpx,
                                        int height, const uint8_t *cavg_dev);
cudaError_t launch_chroma_upsample_packed(const Launch &L, int is_420, const uint8_t *const planes[3], const int irows[3], int ch, Img dst,
                                          int width, int height, int add_alpha, int jpeg, int clamped);
cudaError_t launch_swab(const Launch &L, Img img, int width_mpx, int height);
// kind 0 luma plane, 1 chroma plane, 2 YUV888, 3 YUVA8888, 4 UYVY, 5 YUYV; row_phase_stride (YUV888): 0 = the reference's dense
// walk across the row padding, else the rowstride (the Y U V phase restarts with every row)
cudaError_t launch_clamp_lut(const Launch &L, uint8_t *plane, long long nbytes, int kind, int row_phase_stride, const uint8_t *ty_dev,
                             const uint8_t *tc_dev);

struct BlendFrame {
  const uint8_t *s1, *s2;
  uint8_t *d;
  int rs1, rs2, rsd;
  long long s2_bytes;  // size of the src2 buffer (bounds the ARGB next-pixel alpha read, simple_blend.c:130)
};
cudaError_t launch_simple_blend(const Launch &L, int type, const BlendFrame *frames_dev, int nframes, int width,
                                int height, RgbLayout lay, int bf, const int32_t *luma_tabs_dev);
cudaError_t launch_multi_blend(const Launch &L, int type, BlendFrame f, int width, int height, int bgr, int bf,
                               const int32_t *luma_tabs_dev);
// slide_over.c:55: dst byte (j, x) = first[rs_first * j + off_first + x] before the line (j < bound when along_y, else x < bound,
// bound in rows / bytes), second[rs_second * j + off_second + x] behind it
struct SlideArgs {
  const uint8_t *first, *second;
  uint8_t *d;
  long long off_first, off_second;
  int rs_first, rs_second, rsd, row_bytes, height, along_y, bound;
  int word_safe;  // set by launch_slide_over: source strides and bases are multiples of 4 (aligned word loads stay inside the rows)
};
cudaError_t launch_slide_over(const Launch &L, const SlideArgs &a);
// softlight.c softlight_process :62 on one luma plane (rows / columns at the frame edge are copied)
cudaError_t launch_softlight(const Launch &L, CImg src, Img dst, int width, int height, int ymin, int ymax);
// the per-pixel selectors of layout_blends.c ("triple split", mode 0) and multi_transitions.c ("iris rectangle" 1, "iris circle" 2,
// "4 way split" 3, "dissolve" 4): dst pixel = in1 pixel, in2 pixel or the constant colour.  The host fills the fields its mode reads
// (the float fields hold the values the reference's -ffast-math build computes once per frame; pe_engine.cu pe_fx_multi_transition).
struct SelectArgs {
  const uint8_t *s1, *s2;
  uint8_t *d;
  int rs1, rs2, rsd, width, height, psize;
  int row_bytes;                        // set by launch_select
  const uint8_t *colclass, *rowclass;   // mode 0: bit 0 = "outside" test, bit 1 = "inside" test of the column / row (device memory)
  int colour[3];                        // mode 0: the border colour in the frame's byte order
  uint32_t colour_words[3];             // set by launch_select
  int xx, yy;                           // mode 1: inset in bytes / rows; mode 3: displacement in rows / bytes
  int ihwidth, ihheight;                // (width * psize) >> 1, height >> 1
  float bf, inv_psize, inv_maxradsq, hheight, hwidth, inv_hh, inv_hw;
  const float *mask;                    // mode 4: [height][width] (device memory)
};
cudaError_t launch_select(const Launch &L, int mode, const SelectArgs &a);
// dst = trunc(bg * (1 - alpha) + fg * alpha) in double (compositor.c:120), optional lut8 afterwards:
// launch_over_table builds the 64 KB [bg][fg] result table for one (alpha, lut8), launch_alpha_over applies it
cudaError_t launch_over_table(const Launch &L, double alpha, const uint8_t *lut8_dev, uint8_t *table_dev);
cudaError_t launch_alpha_over(const Launch &L, CImg bg, CImg fg, Img dst, int width, int height, int psize,
                              const uint8_t *over_table_dev, int force_opaque);
// alpha = k256 / 256 exactly: integer blend, no table; optional gamma LUT applied to the result
cudaError_t launch_alpha_over_arith(const Launch &L, CImg bg, CImg fg, Img dst, int width, int height, int psize, int k256,
                                    const uint8_t *lut8_dev, int force_opaque);
cudaError_t launch_alpha_over_arith_batch(const Launch &L, const uint8_t *const *bgs, const uint8_t *const *fgs, uint8_t *const *dsts, int n,
                                          int rs_bg, int rs_fg, int rs_d, int width, int height, int psize, int k256,
                                          const uint8_t *lut8_dev, int force_opaque);
cudaError_t launch_fill(const Launch &L, Img dst, int width, int height, int psize, uint32_t pixel);
// ---- resize (our contract) + letterbox (colourspace.c:15343) -------------------------------------------
struct DevFilter {
  const int32_t *first;
  const int16_t *coef;
  int taps;
  int nonneg;  // no negative coefficient (bilinear banks): the <= 4-tap kernels (unsigned DP2A / packed halves) may take it
};
cudaError_t launch_resize_h(const Launch &L, CImg src, int sw, int sh, int16_t *tmp, int dw, int psize, DevFilter fx);
cudaError_t launch_resize_v(const Launch &L, const int16_t *tmp, int sh, Img dst, int dw, int dh, int psize, DevFilter fy);
// both passes in one kernel (intermediate in shared memory); cudaErrorInvalidConfiguration when the scale factor is too large
cudaError_t launch_resize_tile(const Launch &L, CImg src, int sw, int sh, Img dst, int dw, int dh, int psize, DevFilter fx,
                               DevFilter fy, const int32_t *hx_first, const int32_t *hy_first);
// n frames of the same geometry and strides in one launch per 32 (4-byte pixels, <= 4 taps); cudaErrorInvalidConfiguration otherwise
cudaError_t launch_resize_tile_batch(const Launch &L, const uint8_t *const *srcs, int srs, int sw, int sh, uint8_t *const *dsts, int drs,
                                     int dw, int dh, int psize, DevFilter fx, DevFilter fy, const int32_t *hx_first,
                                     const int32_t *hy_first, int n);
cudaError_t launch_letterbox(const Launch &L, CImg inner, int iw, int ih, Img outer, int ow, int oh, int psize,
                             uint32_t black_pixel);
cudaError_t launch_copy2d(const Launch &L, const uint8_t *src, int srs, uint8_t *dst, int drs, int row_bytes, int rows,
                          int fill, uint8_t fill_value);
// ---- conversion + resize in one kernel (pe_kernels_fused4.cu): planar 4:2:0 -> RGBA32 / BGRA32 scaled on both axes, <= 4 non-negative taps
struct ResizeFilter;
bool cvt_resize_supported(const YuvToRgbArgs &a, int dw, int dh, int drs, const uint8_t *dst, const ResizeFilter &hx, const ResizeFilter &hy);
// frames: n same-shaped conversions (yuv_planar_same_shape); dsts[i]: the dw x dh destination of frames[i] (rowstride drs);
// cudaErrorInvalidConfiguration when a tile's source rectangle does not fit shared memory (the caller runs the unfused pair)
// px_dev / py_dev: the banks as int4 per output sample {c0 | c1 << 16, c2 | c3 << 16, first, aux} (pe_engine.cu get_pack4)
cudaError_t launch_cvt_resize(const Launch &L, const YuvToRgbArgs *frames, uint8_t *const *dsts, int n, int dw, int dh, int drs, const void *px_dev,
                              const void *py_dev, const ResizeFilter &hx, const ResizeFilter &hy);
// ---- fused chain -----------------------------------------------------------------------------------------
struct FusedArgs {
  Planes fg;
  int fw, fh;           // fg luma size
  int is_422, clamped, low_quality, quirks;
  DevConv conv;
  CImg bg;
  Img out;
  int ow, oh;           // outer (= bg = out) size
  int iw, ih, ox, oy;   // inner rect of the letterbox
  DevFilter fx, fy;     // fw -> iw (14 bit), fh -> ih (12 bit)
  const uint8_t *over_table;  // 64 KB [bg][fg] -> composited (+ gamma) byte, see launch_over_table
};
// frames_dev: FusedArgs[nframes] in DEVICE memory; max_src_rows / max_src_cols: the largest source extent one output tile
// touches (computed by the engine from the host copy of the filter banks)
cudaError_t launch_fused_dev(const Launch &L, const FusedArgs *frames_dev, int nframes, int ow, int oh, int max_src_rows,
                             int max_src_cols);
int fused_tile_w();
int fused_tile_h();
// fast path (pe_kernels_fused2.cu): no horizontal scaling, <= 4 vertical taps, 4-byte aligned planes
bool fused2_supported(const FusedArgs &a, int fy_taps, int unused);
int fused2_max_virtual_rows(int is422);
int fused2_max_tile_h();
// frames_host: FusedArgs[nframes] in HOST memory (they travel as kernel parameters).
// blend_a in 0..256: integer blend (alpha = blend_a / 256) + optional lut8; blend_a < 0: FusedArgs::over_table
cudaError_t launch_fused2(const Launch &L, const FusedArgs *frames_host, int nframes, int ow, int oh, int tile_h, int blend_a,
                          const uint8_t *lut8_dev);
// register-resident path (pe_kernels_fused3.cu): 4:2:0, full-width letterbox, alpha = k / 256, one conversion variant per batch
bool fused3_tables_ok(const ConvTables &t);
bool fused3_supported(const FusedArgs *frames_host, int nframes, int fy_taps);
// rows4_dev: int4 per inner output row {first, c3 | c2 << 16, c1 | c0 << 16, 0}; coef16: the coefficients in it are scaled by 16
cudaError_t launch_fused3(const Launch &L, const FusedArgs *frames_host, int nframes, int blend_a, const uint8_t *lut8_dev,
                          const void *rows4_dev, int coef16, unsigned int *sched_dev /* [2], zero; reset by the kernel */);
// ---- diagnostics -------------------------------------------------------------------------------------------
struct DevStats {
  unsigned int minv[4], maxv[4];
  unsigned int hist[256];
  unsigned long long sum;
  unsigned int not_black;      // some colour byte of bytes 0..2 is non-zero (is_all_black_ish exact branch fails)
  unsigned int not_black_ish;  // the reference's "ish" expression is non-zero for some pixel (:2583-2587)
};
cudaError_t launch_stats(const Launch &L, CImg img, int width, int height, int psize, int a_off, DevStats *out_dev);
// minimd5 (src/maths.c:575) of the first nbytes bytes of every row: the row hashes of hash_cmp_layer (colourspace.c:16044)
cudaError_t launch_row_hash(const Launch &L, CImg img, int nbytes, int height, unsigned long long *out_dev);

}  // namespace pe
