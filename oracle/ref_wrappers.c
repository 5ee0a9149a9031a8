/* ref_wrappers.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat C ABI around the *static* functions of the reference's
 * src/colourspace.c.  This file is #included at the end of the translation
 * unit that oracle/build_ref.py assembles from reference line-range slices,
 * so it can see the static tables and converters.  It contains no reference
 * code: each wrapper only forwards its arguments.
 *
 * Conventions: thread_id == -1 lets the reference do its own (pthread)
 * row-band fan-out according to prefs->nfx_threads; ref_set_prefs(1, ...)
 * gives the deterministic single-band result used for parity.
 */

double weed_palette_get_bytes_per_macropixel(int pal) {
  switch (pal) {
  case WEED_PALETTE_RGB24: case WEED_PALETTE_BGR24: case WEED_PALETTE_YUV888: return 3.;
  case WEED_PALETTE_RGBA32: case WEED_PALETTE_BGRA32: case WEED_PALETTE_ARGB32:
  case WEED_PALETTE_YUVA8888: case WEED_PALETTE_UYVY: case WEED_PALETTE_YUYV: return 4.;
  case WEED_PALETTE_YUV411: return 6.;
  default: return 1.;
  }
}

static int ref_inited = 0;

void ref_init(void) {
  if (ref_inited) return;
  /* init_colour_engine (colourspace.c:1973) minus init_advanced_palettes */
  init_RGB_to_YUV_tables();
  init_YUV_to_RGB_tables();
  init_YUV_to_YUV_tables();
  init_average();
  init_unal();
  init_gamma_tx();
  avg_chromaf = avg_chromaf_fast;
  ref_inited = 1;
}

void ref_set_prefs(int nfx_threads, int pb_quality, double screen_gamma) {
  prefs->nfx_threads = nfx_threads;
  prefs->pb_quality = (short)pb_quality;
  prefs->screen_gamma = screen_gamma;
}

/* ---- table dumps -------------------------------------------------------- */

/* which: 0..8 = Y_R,Y_G,Y_B,Cb_R,Cb_G,Cb_B,Cr_R,Cr_G,Cr_B ; 9..13 = RGB_Y,R_Cr,G_Cb,G_Cr,B_Cb */
int ref_get_conv_table(int clamping, int subspace, int which, int *out) {
  ref_init();
  set_conversion_arrays(clamping, subspace);
  struct _conv_array *ca = &THREADVAR(conv_arrays);
  int *t[14] = {ca->Yx_R, ca->Yx_G, ca->Yx_B, ca->Cbx_R, ca->Cbx_G, ca->Cbx_B,
                ca->Crx_R, ca->Crx_G, ca->Crx_B,
                ca->RGBx_Y, ca->Rx_Cr, ca->Gx_Cb, ca->Gx_Cr, ca->Bx_Cb};
  if (which < 0 || which > 13) return -1;
  memcpy(out, t[which], 256 * sizeof(int));
  return 0;
}

/* which: 0 unal 1 al 2 unalcy 3 alcy 4 unalcuv 5 alcuv */
int ref_get_premult_table(int which, int *out) {
  ref_init();
  int *t[6] = {&unal[0][0], &al[0][0], &unalcy[0][0], &alcy[0][0], &unalcuv[0][0], &alcuv[0][0]};
  if (which < 0 || which > 5) return -1;
  memcpy(out, t[which], 65536 * sizeof(int));
  return 0;
}

/* which: 0 cavgc 1 cavgu 2 cavgrgb */
int ref_get_avg_table(int which, uint8_t *out) {
  ref_init();
  uint8_t *t[3] = {&cavgc[0][0], &cavgu[0][0], &cavgrgb[0][0]};
  if (which < 0 || which > 2) return -1;
  memcpy(out, t[which], 65536);
  return 0;
}

/* which: 0 Yc->Yu 1 UVc->UVu 2 Yu->Yc 3 UVu->UVc */
int ref_get_yy_table(int which, uint8_t *out) {
  ref_init();
  uint8_t *t[4] = {Yclamped_to_Yunclamped, UVclamped_to_UVunclamped, Yunclamped_to_Yclamped, UVunclamped_to_UVclamped};
  if (which < 0 || which > 3) return -1;
  memcpy(out, t[which], 256);
  return 0;
}

/* NOTE: the reference caches LUTs process-wide under a key that it mutates
 * (colourspace.c:701,721-733) -- dump each (from,to) pair in a fresh process. */
int ref_gamma_lut8(double fileg, int gamma_from, int gamma_to, uint8_t *out) {
  ref_init();
  uint8_t *l = create_gamma_lut8(fileg, gamma_from, gamma_to);
  if (!l) return -1;
  memcpy(out, l, 256);
  return 0;
}

int ref_gamma_lut16(double fileg, int gamma_from, int gamma_to, uint16_t *out) {
  ref_init();
  uint16_t *l = create_gamma_lut(fileg, gamma_from, gamma_to);
  if (!l) return -1;
  memcpy(out, l, 65536 * sizeof(uint16_t));
  return 0;
}

void ref_gamma_consts(float *out8) {
  ref_init();
  for (int i = 0; i < N_GAMMA_TYPES; i++) {
    out8[i * 4] = gamma_tx[i].offs; out8[i * 4 + 1] = gamma_tx[i].lin;
    out8[i * 4 + 2] = gamma_tx[i].thresh; out8[i * 4 + 3] = gamma_tx[i].pf;
  }
}

/* ---- single-pixel kernels (exhaustive tests) ---------------------------- */

void ref_rgb2yuv_bulk(int clamping, int subspace, const uint8_t *rgb, uint8_t *yuv, long n) {
  ref_init();
  set_conversion_arrays(clamping, subspace);
  for (long i = 0; i < n; i++) rgb2yuv(rgb[i * 3], rgb[i * 3 + 1], rgb[i * 3 + 2], &yuv[i * 3], &yuv[i * 3 + 1], &yuv[i * 3 + 2]);
}

void ref_yuv2rgb_bulk(int clamping, int subspace, const uint8_t *yuv, uint8_t *rgb, long n) {
  ref_init();
  set_conversion_arrays(clamping, subspace);
  for (long i = 0; i < n; i++) yuv2rgb(yuv[i * 3], yuv[i * 3 + 1], yuv[i * 3 + 2], &rgb[i * 3], &rgb[i * 3 + 1], &rgb[i * 3 + 2]);
}

/* ---- frame converters --------------------------------------------------- */

/* order: 0 = RGB(A), 1 = BGR(A), 2 = ARGB */
void ref_yuv420p_to_rgb(uint8_t **src, int width, int height, int *istrides, int orowstride, uint8_t *dest,
                        int order, int add_alpha, int is_422, int sampling, int clamping, int subspace,
                        int gamma, int tgt_gamma) {
  ref_init();
  if (order == 0)
    convert_yuv420p_to_rgb_frame(src, width, height, 0, istrides, orowstride, dest, add_alpha, is_422, sampling,
                                 clamping, subspace, gamma, tgt_gamma, NULL, -USE_THREADS);
  else if (order == 1)
    convert_yuv420p_to_bgr_frame(src, width, height, 0, istrides, orowstride, dest, add_alpha, is_422, sampling,
                                 clamping, subspace, gamma, tgt_gamma, NULL, -USE_THREADS);
  else
    convert_yuv420p_to_argb_frame(src, width, height, 0, istrides, orowstride, dest, is_422, sampling,
                                  clamping, subspace, gamma, tgt_gamma, NULL, -USE_THREADS);
}

/* fmt: 0 = UYVY, 1 = YUYV; width in macropixels */
void ref_packed422_to_rgb(int fmt, void *src, int width, int height, int irow, int orowstride, uint8_t *dest,
                          int order, int add_alpha, int clamping, int subspace) {
  ref_init();
  /* only convert_uyvy_to_rgb_frame selects tables itself with a subspace; the others
     call set_conversion_arrays(clamping, WEED_YUV_SUBSPACE_YCBCR) */
  if (fmt == 0) {
    if (order == 0) convert_uyvy_to_rgb_frame((uyvy_macropixel *)src, width, height, irow, orowstride, dest, add_alpha, clamping, subspace, -USE_THREADS);
    else if (order == 1) convert_uyvy_to_bgr_frame((uyvy_macropixel *)src, width, height, irow, orowstride, dest, add_alpha, clamping, -USE_THREADS);
    else convert_uyvy_to_argb_frame((uyvy_macropixel *)src, width, height, irow, orowstride, dest, clamping, -USE_THREADS);
  } else {
    if (order == 0) convert_yuyv_to_rgb_frame((yuyv_macropixel *)src, width, height, irow, orowstride, dest, add_alpha, clamping, -USE_THREADS);
    else if (order == 1) convert_yuyv_to_bgr_frame((yuyv_macropixel *)src, width, height, irow, orowstride, dest, add_alpha, clamping, -USE_THREADS);
    else convert_yuyv_to_argb_frame((yuyv_macropixel *)src, width, height, irow, orowstride, dest, clamping, -USE_THREADS);
  }
}

void ref_yuv888_to_rgb(uint8_t *src, int width, int height, int irow, int orow, uint8_t *dest, int order,
                       int in_alpha, int out_alpha, int clamping, int subspace) {
  ref_init();
  if (!in_alpha) {
    if (order == 0) convert_yuv888_to_rgb_frame(src, width, height, irow, orow, dest, out_alpha, clamping, subspace, -USE_THREADS);
    else if (order == 1) convert_yuv888_to_bgr_frame(src, width, height, irow, orow, dest, out_alpha, clamping, subspace, -USE_THREADS);
    else convert_yuv888_to_argb_frame(src, width, height, irow, orow, dest, clamping, subspace, -USE_THREADS);
  } else {
    if (order == 0) convert_yuva8888_to_rgba_frame(src, width, height, irow, orow, dest, !out_alpha, clamping, subspace, -USE_THREADS);
    else if (order == 1) convert_yuva8888_to_bgra_frame(src, width, height, irow, orow, dest, !out_alpha, clamping, subspace, -USE_THREADS);
    else convert_yuva8888_to_argb_frame(src, width, height, irow, orow, dest, clamping, subspace, -USE_THREADS);
  }
}

void ref_yuv444p_to_rgb(uint8_t **src, int width, int height, int irow, int orow, uint8_t *dest, int order,
                        int in_alpha, int out_alpha, int clamping) {
  ref_init();
  if (order == 0) convert_yuv_planar_to_rgb_frame(src, width, height, irow, orow, dest, in_alpha, out_alpha, clamping, -USE_THREADS);
  else if (order == 1) convert_yuv_planar_to_bgr_frame(src, width, height, irow, orow, dest, in_alpha, out_alpha, clamping, -USE_THREADS);
  else convert_yuv_planar_to_argb_frame(src, width, height, irow, orow, dest, in_alpha, clamping, -USE_THREADS);
}

/* RGB <-> RGB byte permutations; op codes follow the forward-declaration list
 * colourspace.c:2024-2036.  width is passed through exactly as given, so callers
 * choose pixels (what convert_layer_palette_full passes) or bytes. */
enum { REF_SWAP3 = 0, REF_SWAP4, REF_SWAP3ADDPOST, REF_SWAP3ADDPRE, REF_SWAP3DELPOST, REF_SWAP3DELPRE,
       REF_ADDPRE, REF_ADDPOST, REF_DELPRE, REF_DELPOST, REF_SWAP3POSTALPHA, REF_SWAP3PREALPHA, REF_SWAPPREPOST };

void ref_rgb_permute(int op, uint8_t *src, int width, int height, int irow, int orow, uint8_t *dest,
                     uint8_t *lut8, int alpha_first, int thread_id) {
  ref_init();
  switch (op) {
  case REF_SWAP3: convert_swap3_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_SWAP4: convert_swap4_frame(src, width, height, irow, orow, dest, lut8, alpha_first, thread_id); break;
  case REF_SWAP3ADDPOST: convert_swap3addpost_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_SWAP3ADDPRE: convert_swap3addpre_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_SWAP3DELPOST: convert_swap3delpost_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_SWAP3DELPRE: convert_swap3delpre_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_ADDPRE: convert_addpre_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_ADDPOST: convert_addpost_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_DELPRE: convert_delpre_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_DELPOST: convert_delpost_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_SWAP3POSTALPHA: convert_swap3postalpha_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_SWAP3PREALPHA: convert_swap3prealpha_frame(src, width, height, irow, orow, dest, lut8, thread_id); break;
  case REF_SWAPPREPOST: convert_swapprepost_frame(src, width, height, irow, orow, dest, lut8, alpha_first, thread_id); break;
  }
}

/* RGB -> YUV family.  order: 0 rgb 1 bgr 2 argb */
void ref_rgb_to_yuv888(uint8_t *rgb, int width, int height, int irow, int orow, uint8_t *dest, int order,
                       int in_alpha, int out_alpha, int clamping) {
  ref_init();
  if (order == 0) convert_rgb_to_yuv_frame(rgb, width, height, irow, orow, dest, in_alpha, out_alpha, clamping, -USE_THREADS);
  else if (order == 1) convert_bgr_to_yuv_frame(rgb, width, height, irow, orow, dest, in_alpha, out_alpha, clamping, -USE_THREADS);
  else convert_argb_to_yuv_frame(rgb, width, height, irow, orow, dest, out_alpha, clamping, -USE_THREADS);
}

void ref_rgb_to_yuv444p(uint8_t *rgb, int width, int height, int irow, int orow, uint8_t **dest, int order,
                        int in_alpha, int out_alpha, int clamping) {
  ref_init();
  if (order == 0) convert_rgb_to_yuvp_frame(rgb, width, height, irow, orow, dest, in_alpha, out_alpha, clamping, -USE_THREADS);
  else if (order == 1) convert_bgr_to_yuvp_frame(rgb, width, height, irow, orow, dest, in_alpha, out_alpha, clamping, -USE_THREADS);
  else convert_argb_to_yuvp_frame(rgb, width, height, irow, orow, dest, out_alpha, clamping, -USE_THREADS);
}

void ref_rgb_to_yuv420(uint8_t *rgb, int width, int height, int irow, int *ostrides, uint8_t **dest, int order,
                       int is_422, int has_alpha, int subspace, int clamping) {
  ref_init();
  if (order == 0) convert_rgb_to_yuv420_frame(rgb, width, height, irow, ostrides, dest, is_422, has_alpha, subspace, clamping);
  else if (order == 1) convert_bgr_to_yuv420_frame(rgb, width, height, irow, ostrides, dest, is_422, has_alpha, subspace, clamping);
  else convert_argb_to_yuv420_frame(rgb, width, height, irow, ostrides, dest, is_422, subspace, clamping);
}

/* fmt 0 uyvy 1 yuyv; order 0 rgb 1 bgr 2 argb; gamma LUT variant when gamma != tgt_gamma (both != 0) */
void ref_rgb_to_packed422(int fmt, uint8_t *rgb, int width, int height, int irow, int orow, void *dest, int order, int has_alpha,
                          int clamping, int gamma, int tgt_gamma) {
  ref_init();
  uint16_t *lut = (gamma != tgt_gamma) ? create_gamma_lut(1.0, gamma, tgt_gamma) : NULL;
  if (fmt == 0) {
    if (order == 0) convert_rgb_to_uyvy_frame(rgb, width, height, irow, orow, (uyvy_macropixel *)dest, has_alpha, clamping, lut, -USE_THREADS);
    else if (order == 1) convert_bgr_to_uyvy_frame(rgb, width, height, irow, orow, (uyvy_macropixel *)dest, has_alpha, clamping, lut, -USE_THREADS);
    else convert_argb_to_uyvy_frame(rgb, width, height, irow, orow, (uyvy_macropixel *)dest, clamping, lut, -USE_THREADS);
  } else {
    if (order == 0) convert_rgb_to_yuyv_frame(rgb, width, height, irow, orow, (yuyv_macropixel *)dest, has_alpha, clamping, lut, -USE_THREADS);
    else if (order == 1) convert_bgr_to_yuyv_frame(rgb, width, height, irow, orow, (yuyv_macropixel *)dest, has_alpha, clamping, lut, -USE_THREADS);
    else convert_argb_to_yuyv_frame(rgb, width, height, irow, orow, (yuyv_macropixel *)dest, clamping, lut, -USE_THREADS);
  }
}

/* ---- in-place layer ops -------------------------------------------------- */

void ref_alpha_premult(uint8_t *pixels, int width, int height, int rowstride, int palette, int clamping,
                       int direction, int *flags_inout) {
  ref_init();
  weed_layer_t l;
  memset(&l, 0, sizeof(l));
  l.width = width; l.height = height; l.palette = palette; l.clamping = clamping;
  l.nplanes = 1; l.rowstrides[0] = rowstride; l.pixel_data[0] = pixels;
  l.flags = flags_inout ? *flags_inout : 0;
  alpha_premult(&l, direction);
  if (flags_inout) *flags_inout = l.flags;
}

void ref_alpha_premult_planar(uint8_t **planes, int *rowstrides, int width, int height, int clamping, int direction, int *flags_inout) {
  ref_init();
  weed_layer_t l;
  memset(&l, 0, sizeof(l));
  l.width = width; l.height = height; l.palette = WEED_PALETTE_YUVA4444P; l.clamping = clamping;
  l.nplanes = 4;
  for (int p = 0; p < 4; p++) { l.rowstrides[p] = rowstrides[p]; l.pixel_data[p] = planes[p]; }
  l.flags = flags_inout ? *flags_inout : 0;
  alpha_premult(&l, direction);
  if (flags_inout) *flags_inout = l.flags;
}

/* the per-band worker of gamma_convert_sub_layer (colourspace.c:14034) on an
 * explicit rectangle; the banding arithmetic of the caller (:14096-14110) is
 * thread-count dependent and is not part of the contract */
void ref_gamma_apply(uint8_t *pixels, int rowstride, int psize, int x, int width, int height, int alpha_first,
                     uint8_t *lut8) {
  lives_cc_params cc;
  memset(&cc, 0, sizeof(cc));
  cc.src = pixels; cc.orowstrides[0] = rowstride; cc.psize = psize; cc.vsize = height; cc.hsize = width;
  cc.xoffset = (size_t)x * psize; cc.alpha_first = alpha_first; cc.lut8 = lut8;
  gamma_convert_layer_thread(&cc);
}

/* ---- YUV <-> YUV family -------------------------------------------------- */

void ref_combineplanes(uint8_t **src, int width, int height, int irow, int orow, uint8_t *dest, int in_alpha, int out_alpha) {
  convert_combineplanes_frame(src, width, height, irow, orow, dest, in_alpha, out_alpha);
}

void ref_splitplanes(uint8_t *src, int width, int height, int irow, int *orows, uint8_t **dest, int src_alpha, int dest_alpha) {
  int o[4] = {orows[0], orows[1], orows[2], orows[3]}; /* the reference subtracts the width from its argument array */
  convert_splitplanes_frame(src, width, height, irow, o, dest, src_alpha, dest_alpha);
}

/* cwidth x cheight = the SOURCE chroma plane */
void ref_halve_chroma(uint8_t **src, int cwidth, int cheight, int *istrides, int *ostrides, uint8_t **dest, int clamping) {
  ref_init();
  convert_halve_chroma(src, cwidth, cheight, istrides, ostrides, dest, clamping);
}

void ref_double_chroma(uint8_t **src, int cwidth, int cheight, int *istrides, int *ostrides, uint8_t **dest, int clamping) {
  ref_init();
  convert_double_chroma(src, cwidth, cheight, istrides, ostrides, dest, clamping);
}

/* fmt 0 uyvy 1 yuyv; width in macropixels */
void ref_packed422_to_yuv422p(int fmt, void *src, int width, int height, uint8_t **dest) {
  if (fmt == 0) convert_uyvy_to_yuv422_frame((uyvy_macropixel *)src, width, height, dest);
  else convert_yuyv_to_yuv422_frame((yuyv_macropixel *)src, width, height, dest);
}

void ref_packed422_to_yuv444p(int fmt, void *src, int width, int height, int irow, int *orows, uint8_t **dest, int add_alpha) {
  if (fmt == 0) convert_uyvy_to_yuvp_frame((uyvy_macropixel *)src, width, height, irow, orows, dest, add_alpha);
  else convert_yuyv_to_yuvp_frame((yuyv_macropixel *)src, width, height, irow, orows, dest, add_alpha);
}

void ref_packed422_to_yuv888(int fmt, void *src, int width, int height, int irow, int orow, uint8_t *dest, int add_alpha) {
  if (fmt == 0) convert_uyvy_to_yuv888_frame((uyvy_macropixel *)src, width, height, irow, orow, dest, add_alpha);
  else convert_yuyv_to_yuv888_frame((yuyv_macropixel *)src, width, height, irow, orow, dest, add_alpha);
}

void ref_swab(uint8_t *src, int width, int height, int irow) {
  convert_swab_frame(src, width, height, irow, -USE_THREADS);
}

/* fmt 0 uyvy 1 yuyv; only the dense branch (irow == width, orow == 2 * width) of the reference is usable */
void ref_yuv444p_to_packed422(int fmt, uint8_t **src, int width, int height, int irow, int orow, void *dest, int clamping) {
  ref_init();
  if (fmt == 0) convert_yuv_planar_to_uyvy_frame(src, width, height, irow, orow, (uyvy_macropixel *)dest, clamping);
  else convert_yuv_planar_to_yuyv_frame(src, width, height, irow, orow, (yuyv_macropixel *)dest, clamping);
}

void ref_yuv444p_to_yuv420p(uint8_t **src, int width, int height, int *irows, int *orows, uint8_t **dest, int clamping) {
  ref_init();
  convert_yuvp_to_yuv420_frame(src, width, height, irows, orows, dest, clamping);
}

/* planar 4:2:0 -> packed 4:2:2; width in pixels (what the dispatcher passes, colourspace.c:13566) */
void ref_yuv420_to_packed422(int fmt, uint8_t **src, int width, int height, int *irows, int orow, void *dest, int clamping) {
  int ir[3] = {irows[0], irows[1], irows[2]}; /* the reference modifies its argument */
  ref_init();
  if (fmt == 0) convert_yuv420_to_uyvy_frame(src, width, height, ir, orow, (uyvy_macropixel *)dest, clamping);
  else convert_yuv420_to_yuyv_frame(src, width, height, ir, orow, (yuyv_macropixel *)dest, clamping);
}

/* planar 4:2:2 -> packed 4:2:2; width = macropixels per row as the function's loop counts them */
void ref_yuv422p_to_packed422(int fmt, uint8_t **src, int width, int height, int *irows, int orow, uint8_t *dest) {
  int ir[3] = {irows[0], irows[1], irows[2]};
  if (fmt == 0) convert_yuv422p_to_uyvy_frame(src, width, height, ir, orow, dest);
  else convert_yuv422p_to_yuyv_frame(src, width, height, ir, orow, dest);
}

/* width x height = the destination (4:4:4) plane */
void ref_quad_chroma(uint8_t **src, int width, int height, int *istrides, int ostride, uint8_t **dest, int add_alpha, int sampling,
                     int clamping) {
  ref_init();
  convert_quad_chroma(src, width, height, istrides, ostride, dest, add_alpha, sampling, clamping);
}

/* mode 0 uyvy 1 yuyv 2 planar 4:2:2 3 planar 4:2:0; width in pixels */
void ref_yuv888_subsample(int mode, uint8_t *src, int width, int height, int irow, int *orows, uint8_t **dest, int src_alpha, int clamping) {
  int o[3] = {orows[0], orows[1], orows[2]};
  ref_init();
  if (mode == 0) convert_yuv888_to_uyvy_frame(src, width, height, irow, o[0], (uyvy_macropixel *)dest[0], src_alpha, clamping);
  else if (mode == 1) convert_yuv888_to_yuyv_frame(src, width, height, irow, o[0], (yuyv_macropixel *)dest[0], src_alpha, clamping);
  else if (mode == 2) convert_yuv888_to_yuv422_frame(src, width, height, irow, o, dest, src_alpha, clamping);
  else convert_yuv888_to_yuv420_frame(src, width, height, irow, o, dest, src_alpha, clamping);
}

/* fmt 0 uyvy 1 yuyv; width in macropixels; dense buffers */
void ref_packed422_to_yuv420p(int fmt, void *src, int width, int height, uint8_t **dest, int clamping) {
  ref_init();
  if (fmt == 0) convert_uyvy_to_yuv420_frame((uyvy_macropixel *)src, width, height, dest, clamping);
  else convert_yuyv_to_yuv420_frame((yuyv_macropixel *)src, width, height, dest, clamping);
}

/* is_420 1: convert_quad_chroma_packed, 0: convert_double_chroma_packed; width x height = the destination frame in pixels */
void ref_chroma_upsample_packed(int is_420, uint8_t **src, int width, int height, int *istrides, int ostride, uint8_t *dest, int add_alpha,
                                int sampling, int clamping) {
  ref_init();
  if (is_420) convert_quad_chroma_packed(src, width, height, istrides, ostride, dest, add_alpha, sampling, clamping);
  else convert_double_chroma_packed(src, width, height, istrides, ostride, dest, add_alpha, sampling, clamping);
}

/* ---- the reference's "float - experimental" YUV -> RGB path (colourspace.c:101-172 tables, :592 clamp0255f, :2367 yuv2rgb_float);
 *      float tables exist for the BT.709 subspace only (:279-310, :316-355) ---- */
/* which: 0 RGBf_Y 1 Rf_Cr 2 Gf_Cb 3 Gf_Cr 4 Bf_Cb */
/* YUV411 as a source; target as pe_or_yuv411_to; width in macropixels.  Only convert_yuv411_to_{rgb,bgr,argb}_frame take an output
 * rowstride; the others write densely */
void ref_yuv411_to(int target, void *src, int width, int height, int orow, uint8_t **dest, int order, int add_alpha, int clamping) {
  ref_init();
  yuv411_macropixel *s = (yuv411_macropixel *)src;
  if (target == 0) {
    if (order == 0) convert_yuv411_to_rgb_frame(s, width, height, orow, dest[0], add_alpha, clamping);
    else if (order == 1) convert_yuv411_to_bgr_frame(s, width, height, orow, dest[0], add_alpha, clamping);
    else convert_yuv411_to_argb_frame(s, width, height, orow, dest[0], clamping);
  } else if (target == 1) convert_yuv411_to_yuv888_frame(s, width, height, dest[0], add_alpha, clamping);
  else if (target == 2) convert_yuv411_to_yuvp_frame(s, width, height, dest, add_alpha, clamping);
  else if (target == 3) convert_yuv411_to_uyvy_frame(s, width, height, (uyvy_macropixel *)dest[0], clamping);
  else if (target == 4) convert_yuv411_to_yuyv_frame(s, width, height, (yuyv_macropixel *)dest[0], clamping);
  else if (target == 5) convert_yuv411_to_yuv422_frame(s, width, height, dest, clamping);
  else convert_yuv411_to_yuv420_frame(s, width, height, dest, FALSE, clamping);
}

/* mode as pe_or_to_yuv411; width as the reference's dispatcher passes it (UYVY / YUYV: macropixels; otherwise pixels); dense walks */
void ref_to_yuv411(int mode, uint8_t **src, int width, int height, int irow, void *dest, int clamping) {
  ref_init();
  yuv411_macropixel *d = (yuv411_macropixel *)dest;
  if (mode == 0) convert_uyvy_to_yuv411_frame((uyvy_macropixel *)src[0], width, height, d, clamping);
  else if (mode == 1) convert_yuyv_to_yuv411_frame((yuyv_macropixel *)src[0], width, height, d, clamping);
  else if (mode <= 3) convert_yuv420_to_yuv411_frame(src, width, height, d, mode == 3, clamping);
  else if (mode <= 5) convert_yuv888_to_yuv411_frame(src[0], width, height, irow, d, mode == 5);
  else convert_yuvp_to_yuv411_frame(src, width, height, irow, d, clamping);
}

/* order 0 RGB(A), 1 BGR(A), 2 ARGB; width in pixels; dense output */
void ref_rgb_to_yuv411(uint8_t *src, int width, int height, int irow, void *dest, int order, int has_alpha, int clamping) {
  ref_init();
  if (order == 0) convert_rgb_to_yuv411_frame(src, width, height, irow, (yuv411_macropixel *)dest, has_alpha, clamping);
  else if (order == 1) convert_bgr_to_yuv411_frame(src, width, height, irow, (yuv411_macropixel *)dest, has_alpha, clamping);
  else convert_argb_to_yuv411_frame(src, width, height, irow, (yuv411_macropixel *)dest, clamping);
}

int ref_get_float_table(int clamping, int which, float *out) {
  ref_init();
  set_conversion_arrays(clamping, WEED_YUV_SUBSPACE_BT709);
  float *t[5] = {RGBf_Y, Rf_Cr, Gf_Cb, Gf_Cr, Bf_Cb}; /* the reference's own accessor macros (:411-415) */
  if (which < 0 || which > 4 || !t[which]) return -1;
  memcpy(out, t[which], 256 * sizeof(float));
  return 0;
}

/* yuv2rgb_float exactly as the reference writes it (:2367-2372: the INTEGER RGB_Y table plus the float chroma tables) */
void ref_yuv2rgb_float_bulk(int clamping, const uint8_t *yuv, uint8_t *rgb, long n) {
  ref_init();
  set_conversion_arrays(clamping, WEED_YUV_SUBSPACE_BT709);
  for (long i = 0; i < n; i++) yuv2rgb_float(yuv[i * 3], yuv[i * 3 + 1], yuv[i * 3 + 2], &rgb[i * 3], &rgb[i * 3 + 1], &rgb[i * 3 + 2]);
}

/* the form the commented-out variant at :2398-2400 spells (RGBf_Y[y] + ...): the reference's float tables and clamp0255f, our loop;
 * sums (optional): the three float sums per pixel before the clamp */
void ref_yuv2rgb_floaty_bulk(int clamping, const uint8_t *yuv, uint8_t *rgb, float *sums, long n) {
  ref_init();
  set_conversion_arrays(clamping, WEED_YUV_SUBSPACE_BT709);
  for (long i = 0; i < n; i++) {
    const uint8_t y = yuv[i * 3], u = yuv[i * 3 + 1], v = yuv[i * 3 + 2];
    const float r = RGBf_Y[y] + Rf_Cr[v], g = RGBf_Y[y] + Gf_Cb[u] + Gf_Cr[v], b = RGBf_Y[y] + Bf_Cb[u];
    rgb[i * 3] = clamp0255f(r); rgb[i * 3 + 1] = clamp0255f(g); rgb[i * 3 + 2] = clamp0255f(b);
    if (sums) { sums[i * 3] = r; sums[i * 3 + 1] = g; sums[i * 3 + 2] = b; }
  }
}
