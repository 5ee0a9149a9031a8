/* pe_oracle.h -- CPU restatement of the LiVES per-frame pixel path.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (lives_b200/) never links, imports or calls it.
 *
 * PARITY STATUS: the reference holds no golden vectors for this path
 * (SURVEY.md section 4), so the oracle is pinned against the reference itself:
 * tests/test_oracle_vs_reference.py compares every function below with
 * oracle/_ref/libref_oracle.so / simple_blend.so / ref_paint_pixel.so, which
 * oracle/build_ref.py compiles from the reference sources where they lie, and
 * tests/golden/ freezes the outputs of that compiled reference.
 * Resize (pe_or_resize_*) is "parity unpinned": the reference delegates it to
 * libswscale, which is neither in the tree nor installed.
 *
 * All citations are file:line in /root/reference.
 */
#ifndef PE_ORACLE_H
#define PE_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* palette / enum values: libweed/weed-palettes.h:43-183 */
enum {
  OR_PAL_RGB24 = 1, OR_PAL_BGR24 = 2, OR_PAL_RGBA32 = 3, OR_PAL_BGRA32 = 4, OR_PAL_ARGB32 = 5,
  OR_PAL_YUV420P = 512, OR_PAL_YVU420P = 513, OR_PAL_YUV422P = 522, OR_PAL_YUV444P = 544,
  OR_PAL_YUVA4444P = 545, OR_PAL_UYVY = 564, OR_PAL_YUYV = 565, OR_PAL_YUV888 = 588, OR_PAL_YUVA8888 = 589
};
enum { OR_CLAMPED = 0, OR_UNCLAMPED = 1 };
enum { OR_SUBSPACE_YUV = 0, OR_SUBSPACE_YCBCR = 1, OR_SUBSPACE_BT709 = 2 };
enum { OR_GAMMA_UNKNOWN = 0, OR_GAMMA_LINEAR = -1, OR_GAMMA_SRGB = 1, OR_GAMMA_BT709 = 2, OR_GAMMA_MONITOR = 1024 };
enum { OR_QUALITY_LOW = 1, OR_QUALITY_MED = 2, OR_QUALITY_HIGH = 3 };
enum { OR_ORDER_RGB = 0, OR_ORDER_BGR = 1, OR_ORDER_ARGB = 2 };

/* which: 0..8 Y_R Y_G Y_B Cb_R Cb_G Cb_B Cr_R Cr_G Cr_B, 9..13 RGB_Y R_Cr G_Cb G_Cr B_Cb */
void pe_or_conv_table(int clamping, int subspace, int which, int32_t out[256]);
/* which: 0 unal 1 al 2 unalcy 3 alcy 4 unalcuv 5 alcuv (colourspace.c:1141) */
void pe_or_premult_table(int which, int32_t out[65536]);
int pe_or_gamma_lut8(double fileg, int gamma_from, int gamma_to, double screen_gamma, uint8_t out[256]);
int pe_or_gamma_lut16(double fileg, int gamma_from, int gamma_to, double screen_gamma, uint16_t out[65536]);

void pe_or_rgb2yuv(int clamping, int subspace, int quality, const uint8_t *rgb, uint8_t *yuv, long n);
void pe_or_yuv2rgb(int clamping, int subspace, int quality, const uint8_t *yuv, uint8_t *rgb, long n);

/* planar 4:2:0 / 4:2:2 -> packed RGB(A); quirks != 0 replicates the deterministic
 * copy-paste slips of colourspace.c:3461,3544,3600; plane_sizes bound the
 * one-past-row chroma read (see DESIGN.md "edge read"). lut16 may be NULL. */
void pe_or_yuv420p_to_rgb(const uint8_t *const src[3], const int istrides[3], int width, int height,
                          uint8_t *dest, int orowstride, int order, int add_alpha, int is_422,
                          int clamping, int subspace, int quality, int quirks, const uint16_t *lut16);
/* fmt 0 UYVY, 1 YUYV; width in macropixels (colourspace.c:6616,6862) */
void pe_or_packed422_to_rgb(int fmt, const uint8_t *src, int irow, int width_mpx, int height,
                            uint8_t *dest, int orowstride, int order, int add_alpha, int clamping, int subspace,
                            int quality);
void pe_or_yuv888_to_rgb(const uint8_t *src, int irow, int width, int height, uint8_t *dest, int orow,
                         int order, int in_alpha, int out_alpha, int clamping, int subspace, int quality);
void pe_or_rgb_to_yuv888(const uint8_t *src, int irow, int width, int height, uint8_t *dest, int orow,
                         int order, int in_alpha, int out_alpha, int clamping, int quality);

/* fmt 0 UYVY, 1 YUYV; width in PIXELS (rounded down to even); lut16 may be NULL (colourspace.c:5129-5700) */
void pe_or_rgb_to_packed422(int fmt, const uint8_t *src, int irow, int width, int height, uint8_t *dest, int orow, int order,
                            int in_alpha, int clamping, int quality, const uint16_t *lut16);
/* planar 4:4:4 (+ alpha plane), colourspace.c:5786,5971,6154 */
/* init_average :190 (which 0 cavgc, 1 cavgu) and convert_{rgb,bgr}_to_yuv420_frame :6250 / :6385 (4:2:0 and 4:2:2 planar) */
void pe_or_avg_table(int which, uint8_t out[65536]);
void pe_or_rgb_to_yuv420p(const uint8_t *src, int irow, int width, int height, uint8_t *const dest[3], const int ostrides[3],
                          int order, int in_alpha, int is_422, int clamping, int subspace, int quality);
void pe_or_rgb_to_yuv444p(const uint8_t *src, int irow, int width, int height, uint8_t *const dest[4], int orow, int order,
                          int in_alpha, int out_alpha, int clamping, int quality);

/* RGB<->RGB: any of the 5 RGB palettes to any other, optional lut8 on colour bytes
 * (colourspace.c:12370-12556 + :9259-10515, intended whole-row semantics) */
int pe_or_rgb_to_rgb(int inpal, int outpal, const uint8_t *src, int irow, int width, int height,
                     uint8_t *dest, int orow, const uint8_t *lut8);

/* gamma_convert_layer_thread colourspace.c:14034 on an explicit rectangle */
void pe_or_gamma_apply(uint8_t *pixels, int rowstride, int palette, int x, int y, int width, int height,
                       const uint8_t lut8[256]);
/* alpha_premult colourspace.c:11968 for the packed 4-byte palettes; direction +1 forward, -1 reverse */
void pe_or_alpha_premult(uint8_t *pixels, int rowstride, int palette, int clamping, int width, int height,
                         int direction);

/* simple_blend.c common_process :58.  type 0 chroma blend, 1 luma overlay, 2 luma underlay, 3 neg luma overlay */
void pe_or_simple_blend(int type, int palette, const uint8_t *src1, int irow1, const uint8_t *src2, int irow2,
                        uint8_t *dst, int orow, int width, int height, int bf, long src2_bytes);
/* multi_blends.c common_process :26. type 0 multiply 1 screen 2 darken 3 lighten 4 overlay 5 dodge 6 burn */
void pe_or_multi_blend(int type, int palette, const uint8_t *src1, int irow1, const uint8_t *src2, int irow2,
                       uint8_t *dst, int orow, int width, int height, int bf);
/* compositor.c paint_pixel :120 over a same-size layer at offset 0, scalar alpha */
void pe_or_alpha_over(uint8_t *dst, int orow, const uint8_t *src, int irow, int palette, int width, int height,
                      double alpha);
/* compositor.c:178-186 background fill */
void pe_or_fill(uint8_t *dst, int orow, int palette, int width, int height, int r, int g, int b);

/* bilinear / triangle resize on packed 3- or 4-byte pixels -- OUR published contract
 * (parity unpinned, see header comment and DESIGN.md) */
void pe_or_resize_packed(const uint8_t *src, int irow, int sw, int sh, uint8_t *dst, int orow, int dw, int dh,
                         int psize);
int pe_or_resize_filter(int src_n, int dst_n, int shift_bits, int32_t *first, int16_t *coefs, int max_taps);
/* the same interface with libswscale's bilinear coefficient recipe (opt-in; pe_or_set_resize_recipe(1) makes pe_or_resize_packed use
 * it; 0 = the published contract the product ships, the default) */
int pe_or_resize_filter_sws(int src_n, int dst_n, int shift_bits, int32_t *first, int16_t *coefs, int max_taps);
void pe_or_set_resize_recipe(int recipe);
/* libswscale's recipe per flag (kind 1 bilinear, 2 bicubic, 3 Lanczos, 4 / 5 fast bilinear vertical / horizontal) and the resize with
 * a LiVESInterpType (0 FAST, 1 NORMAL, 2 BEST: flags of src/colourspace.c:14991-14997) */
int pe_or_resize_filter_kind(int kind, int src_n, int dst_n, int shift_bits, int32_t *first, int16_t *coefs, int max_taps);
void pe_or_resize_packed_interp(const uint8_t *src, int irow, int sw, int sh, uint8_t *dst, int orow, int dw, int dh, int psize,
                                int interp);
/* letterbox_layer colourspace.c:15343: centre an inner packed frame in a black outer one */
void pe_or_letterbox_packed(const uint8_t *inner, int irow, int iw, int ih, uint8_t *outer, int orow, int ow, int oh,
                            int palette);

/* ---- YUV <-> YUV family --------------------------------------------------------------------------- */
/* planar 4:4:4 (+ alpha plane) -> packed RGB(A), convert_yuv_planar_to_{rgb,bgr,argb}_frame colourspace.c:7200,7304,7405:
 * always the YCbCr tables; order 0 RGB 1 BGR 2 ARGB */
void pe_or_yuv444p_to_rgb(const uint8_t *const src[4], int irow, int width, int height, uint8_t *dest, int orow, int order,
                          int in_alpha, int out_alpha, int clamping, int quality);
/* convert_combineplanes_frame :7593 (4:4:4 planar -> YUV888 / YUVA8888) and convert_splitplanes_frame :9198 (the reverse) */
void pe_or_combine_planes(const uint8_t *const src[4], int irow, int width, int height, uint8_t *dest, int orow, int in_alpha,
                          int out_alpha);
void pe_or_split_planes(const uint8_t *src, int irow, int width, int height, uint8_t *const dest[4], const int orows[4],
                        int src_alpha, int dest_alpha);
/* convert_halve_chroma :10578 (4:2:2 -> 4:2:0 chroma planes) / convert_double_chroma :10612 (4:2:0 -> 4:2:2); cwidth x cheight =
 * the SOURCE chroma plane; planes 1 and 2 only (the caller copies luma, :13706,:13593) */
void pe_or_halve_chroma(const uint8_t *const src[3], const int istrides[3], int cwidth, int cheight, uint8_t *const dest[3],
                        const int ostrides[3], int clamping);
void pe_or_double_chroma(const uint8_t *const src[3], const int istrides[3], int cwidth, int cheight, uint8_t *const dest[3],
                         const int ostrides[3], int clamping);
/* packed 4:2:2 (fmt 0 UYVY, 1 YUYV; width in macropixels) -> planar 4:2:2 (convert_{uyvy,yuyv}_to_yuv422_frame :8093,:8111;
 * quirks != 0: the source pointer is never advanced, every sample is the frame's FIRST macropixel), planar 4:4:4 (+ alpha)
 * (convert_{uyvy,yuyv}_to_yuvp_frame :7800,:7823) and YUV888 / YUVA8888 (convert_{uyvy,yuyv}_to_yuv888_frame :7845,:7866) */
void pe_or_packed422_to_yuv422p(int fmt, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *const dest[3],
                                const int orows[3], int quirks);
void pe_or_packed422_to_yuv444p(int fmt, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *const dest[4],
                                const int orows[4], int add_alpha);
/* YUV411 (IYU1) -> RGB(A) / YUV888 / YUVA8888 / YUV444P / YUVA4444P / UYVY / YUYV, colourspace.c:8305-8910 (see pe_oracle.c) */
void pe_or_yuv411_to(int target, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *const dest[4], const int orow[4],
                     int order, int add_alpha, int clamping, int quality, int quirks);
void pe_or_to_yuv411(int mode, const uint8_t *const src[3], const int irow[3], int width, int height, uint8_t *dest, int orow,
                     int clamping);
void pe_or_alpha_premult_planar(uint8_t *const planes[4], const int rows[4], int clamping, int width, int height, int direction);
void pe_or_rgb_to_yuv411(const uint8_t *src, int irow, int width, int height, uint8_t *dest, int orow, int order, int in_alpha,
                         int clamping);
void pe_or_packed422_to_yuv888(int fmt, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *dest, int orow,
                               int add_alpha);
/* planar 4:4:4 -> packed 4:2:2 (convert_yuv_planar_to_{uyvy,yuyv}_frame :7500,:7548): chroma = avg_chroma of the pixel pair;
 * -> planar 4:2:0 (convert_yuvp_to_yuv420_frame :7690): avg_chroma(avg_chroma(row 2k pair), avg_chroma(row 2k+1 pair)) */
void pe_or_yuv444p_to_packed422(int fmt, const uint8_t *const src[3], int irow, int width, int height, uint8_t *dest, int orow,
                                int clamping);
void pe_or_yuv444p_to_yuv420p(const uint8_t *const src[3], const int irows[3], int width, int height, uint8_t *const dest[3],
                              const int orows[3], int clamping);
/* planar 4:2:0 / 4:2:2 -> packed 4:2:2 (convert_yuv420_to_{uyvy,yuyv}_frame :7104,:7152; convert_yuv422p_to_{uyvy,yuyv}_frame
 * :6442,:6470): luma and chroma interleaved, 4:2:0 chroma rows used twice without interpolation */
void pe_or_yuv42xp_to_packed422(int fmt, const uint8_t *const src[3], const int irows[3], int width, int height, int is_422,
                                uint8_t *dest, int orow);
/* convert_quad_chroma :10642 (4:2:0 -> 4:4:4 chroma planes 1 and 2, + alpha plane = 255): even rows interpolated horizontally
 * from chroma row i / 2 (plain average for JPEG sampling, 3:1 / 1:3 weights otherwise, U and V mirrored), odd rows = avg_chroma of
 * the rows below and above; width x height = the destination plane */
void pe_or_quad_chroma(const uint8_t *const src[3], const int istrides[3], int width, int height, uint8_t *const dest[4], int ostride,
                       int add_alpha, int sampling_jpeg, int clamping);
/* YUV888 / YUVA8888 -> UYVY (mode 0) / YUYV (1) / planar 4:2:2 (2) / planar 4:2:0 (3): convert_yuv888_to_{uyvy,yuyv,yuv422,yuv420}_frame
 * :8184,:8228,:8129,:8035.  Chroma of a pixel pair = avg_chroma(first, second); 4:2:0 additionally avg_chroma(row 2k, row 2k+1).
 * dest[0] is the packed frame for modes 0 / 1 */
void pe_or_yuv888_subsample(int mode, const uint8_t *src, int irow, int width, int height, int src_alpha, uint8_t *const dest[3],
                            const int orows[3], int clamping);
/* packed 4:2:2 -> planar 4:2:0 (convert_{uyvy,yuyv}_to_yuv420_frame :7887,:7930): luma split, chroma row k = avg_chroma(row 2k,
 * row 2k+1) */
void pe_or_packed422_to_yuv420p(int fmt, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *const dest[3],
                                const int orows[3], int clamping);
/* planar 4:2:0 / 4:2:2 -> YUV888 / YUVA8888 with the chroma up-sampled on the fly: convert_quad_chroma_packed :10715 (is_420 = 1)
 * and convert_double_chroma_packed :10811 (is_420 = 0) */
void pe_or_chroma_upsample_packed(int is_420, const uint8_t *const src[3], const int istrides[3], int width, int height, uint8_t *dest,
                                  int orow, int add_alpha, int sampling_jpeg, int clamping);
/* convert_swab_frame :10517 (UYVY <-> YUYV in place) */
void pe_or_swab(uint8_t *pixels, int irow, int width_mpx, int height);
/* init_YUV_to_YUV_tables :1108; which 0 Yc->Yu 1 UVc->UVu 2 Yu->Yc 3 UVu->UVc */
void pe_or_yy_table(int which, uint8_t out[256]);
/* switch_yuv_clamping_and_subspace :10929 on one plane walked densely (padding included, as the reference walks it):
 * kind 0 all luma, 1 all chroma, 2 YUV888, 3 YUVA8888, 4 UYVY, 5 YUYV; to_unclamped selects the table pair */
void pe_or_switch_clamping_plane(uint8_t *plane, long nbytes, int kind, int to_unclamped);
/* slide_over.c sover_process :55-145: a straight cut between in1 ("upper") and in2 ("lower") whose dividing line moves with
 * transval 0 .. 255.  direction 1 .. 4 = the plugin's "plugin_direction" (sover_init :38-52): 1 / 2 the line runs along x, 3 / 4
 * along y; mvlower / mvupper: the clip slides with the line instead of being uncovered in place.  width in macropixels of psize
 * bytes.  pe_or_slide_over_bound: the dividing line (rows or macropixels) */
int pe_or_slide_over_bound(int direction, int transval, int width, int height);
void pe_or_slide_over(int direction, int transval, int mvlower, int mvupper, const uint8_t *src1, int irow1, const uint8_t *src2,
                      int irow2, uint8_t *dest, int orow, int width, int height, int psize);

/* ---- per-frame diagnostics ------------------------------------------------------------------------------------------- */
/* is_all_black_ish src/colourspace.c:2554-2594, both branches (exact = 0: the bit expression of :2583-2587) */
int pe_or_is_all_black_ish(int width, int height, int rowstride, int has_alpha, const uint8_t *pixels, int exact);
/* minimd5 src/maths.c:575 = U[0] ^ U[1] of the reference's own MD5 variant (src/maths.h:42-56: RFC 1321 except round 1) */
uint64_t pe_or_minimd5(const uint8_t *data, size_t n);
/* hash_cmp_layer src/colourspace.c:16044-16075: per-row hashes of the first nbytes bytes + their XOR */
uint64_t pe_or_row_hashes(const uint8_t *pixels, int nbytes, int height, int rowstride, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif

/* ---- SURVEY 8f rank 3, second batch (lives-plugins/weed-plugins/softlight.c, layout_blends.c, multi_transitions.c) ---- */
void pe_or_softlight(const uint8_t *src, int irow, uint8_t *dst, int orow, int width, int height, int clamped);
void pe_or_triple_split_classes(int width, int height, double xstart, int sym, double xend, int vert, double bw, uint8_t *colclass,
                                uint8_t *rowclass);
void pe_or_triple_split(const uint8_t *src1, int irow1, const uint8_t *src2, int irow2, uint8_t *dst, int orow, int width, int height,
                        int bgr, double xstart, int sym, double xend, int vert, double bw, const int bordercol[3]);
void pe_or_dissolve_mask(int64_t seed, long n, float *mask);
void pe_or_multi_transition(int type, const uint8_t *src1, int irow1, const uint8_t *src2, int irow2, uint8_t *dst, int orow, int width,
                            int height, int psize, double bfd, const float *mask);

/* ---- the reference's float YUV -> RGB path (colourspace.c:101-172, :592, :1040-1104, :2367) ---- */
void pe_or_float_table(int clamping, int which, float out[256]);
void pe_or_yuv2rgb_float(int mode, int clamping, const int32_t *rgb_y_int, const uint8_t *yuv, uint8_t *rgb, float *sums, long n);
