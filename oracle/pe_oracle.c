/* pe_oracle.c -- CPU restatement of the LiVES per-frame pixel path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY -- see pe_oracle.h.  Each function cites the
 * reference file:line it follows; deviations from the reference are listed in
 * DESIGN.md ("quirk table") and are limited to undefined / thread-count
 * dependent behaviour.
 */
#include "pe_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- scalar helpers ------------------------------------------------------ */

/* src/maths.h:118 */
static int or_myround(double n) { return n >= 0. ? (int)(n + 0.5) : (int)(n - 0.5); }

/* src/maths.h:88 */
static uint8_t or_clamp0255f(double a) { /* macro in the reference: evaluated in the argument's own type */
  return a >= 254.5 ? (uint8_t)255 : a < -0.5 ? (uint8_t)0 : (uint8_t)(a + .5);
}

/* src/colourspace.h:19-23 written as plain range clamps */
static int or_clamp16_240(int n) { return n < 16 ? 16 : n > 240 ? 240 : n; }
static int or_clamp0_255(int n) { return n < 0 ? 0 : n > 255 ? 255 : n; }

/* src/colourspace.c:832-835 */
static int32_t or_spc_rnd(int32_t val, int quality) {
  if (quality != OR_QUALITY_HIGH) return val >> 16;
  return (int32_t)((float)val / 65536.);
}

#define OR_SF 65793. /* SCALE_FACTOR src/colourspace.h:60 */

/* ---- conversion tables  src/colourspace.c:851-1105 ------------------------ */

typedef struct {
  int32_t t[14][256];
  int min_y, max_y, min_uv, max_uv;
} or_conv_t;

static void or_build_conv(int clamping, int subspace, or_conv_t *c) {
  const int bt709 = (subspace == OR_SUBSPACE_BT709);
  const double kr = bt709 ? 0.2126 : 0.299, kb = bt709 ? 0.0722 : 0.114; /* colourspace.h:84-91 */
  const double cfy = (235. - 16.) / 255., cfuv = (240. - 16.) / 255.;     /* colourspace.h:120-121 */
  int i;
  if (clamping == OR_CLAMPED) {
    /* colourspace.c:878-894 / :919-947 */
    for (i = 0; i < 256; i++) {
      double fac;
      c->t[0][i] = or_myround(kr * (double)i * cfy * OR_SF);
      c->t[1][i] = or_myround((1. - kr - kb) * (double)i * cfy * OR_SF);
      c->t[2][i] = or_myround((kb * (double)i * cfy + 16.) * OR_SF);
      fac = .5 / (1. - kb);
      c->t[3][i] = or_myround(-fac * kr * (double)i * cfuv * OR_SF);
      c->t[4][i] = or_myround(-fac * (1. - kb - kr) * (double)i * cfuv * OR_SF);
      c->t[5][i] = or_myround((0.5 * (double)i * cfuv + 128.) * OR_SF);
      fac = .5 / (1. - kr);
      c->t[6][i] = or_myround((0.5 * (double)i * cfuv + 128.) * OR_SF);
      c->t[7][i] = or_myround(-fac * (1. - kb - kr) * (double)i * cfuv * OR_SF);
      c->t[8][i] = or_myround(-fac * kb * (double)i * cfuv * OR_SF);
    }
    c->min_y = c->min_uv = 16; c->max_y = 235; c->max_uv = 240; /* :362-366 */
  } else {
    /* colourspace.c:896-912 / :949-977 */
    for (i = 0; i < 256; i++) {
      double fac;
      c->t[0][i] = or_myround(kr * (double)i * OR_SF);
      c->t[1][i] = or_myround((1. - kr - kb) * (double)i * OR_SF);
      c->t[2][i] = or_myround(kb * (double)i * OR_SF);
      fac = .5 / (1. - kb);
      c->t[3][i] = or_myround(-fac * kr * (double)i * OR_SF);
      c->t[4][i] = or_myround(-fac * (1. - kb - kr) * (double)i * OR_SF);
      c->t[5][i] = or_myround((0.5 * (double)i + 128.) * OR_SF);
      fac = .5 / (1. - kr);
      c->t[6][i] = or_myround((0.5 * (double)i + 128.) * OR_SF);
      c->t[7][i] = or_myround(-fac * (1. - kb - kr) * (double)i * OR_SF);
      c->t[8][i] = or_myround(-fac * kb * (double)i * OR_SF);
    }
    c->min_y = c->min_uv = 0; c->max_y = c->max_uv = 255; /* :367-370 */
  }
  /* YUV -> RGB, colourspace.c:984-1105.  G_Cb uses -.5/(1+Kb+Kr) for YCbCr (:1005) and
   * -.5/(1+Kb+Kb) for BT.709 (:1062) -- replicated as written. */
  {
    const double gcb = bt709 ? -.5 / (1. + kb + kb) : -.5 / (1. + kb + kr);
    if (clamping == OR_CLAMPED) {
      for (i = 0; i <= 16; i++) c->t[9][i] = 0;
      for (; i < 235; i++) c->t[9][i] = or_myround(((double)i - 16.) / (235. - 16.) * 255. * OR_SF);
      for (; i < 256; i++) c->t[9][i] = (int)(255 * OR_SF);
      for (i = 0; i <= 16; i++) c->t[10][i] = c->t[11][i] = c->t[12][i] = c->t[13][i] = 0;
      for (; i < 240; i++) {
        double cc = (((double)i - 16.) / (240. - 16.) * 255.) - 128.;
        c->t[10][i] = or_myround(2. * (1. - kr) * cc * OR_SF);
        c->t[11][i] = or_myround(gcb * cc * OR_SF);
        c->t[12][i] = or_myround(-.5 / (1. - kr) * cc * OR_SF);
        c->t[13][i] = or_myround(2. * (1. - kb) * cc * OR_SF);
      }
      for (; i < 256; i++) {
        /* saturating value: YCbCr uses c(240) (:1015-1024), BT.709 uses 255-128 (:1078-1081) */
        double cc = bt709 ? (255. - 128.) : (((240. - 16.) / (240. - 16.) * 255.) - 128.);
        c->t[10][i] = or_myround(2. * (1. - kr) * cc * OR_SF);
        c->t[11][i] = or_myround(gcb * cc * OR_SF);
        c->t[12][i] = or_myround(-.5 / (1. - kr) * cc * OR_SF);
        c->t[13][i] = or_myround(2. * (1. - kb) * cc * OR_SF);
      }
    } else {
      for (i = 0; i < 256; i++) {
        double cc = (double)i - 128.;
        c->t[9][i] = (int)(i * OR_SF);
        c->t[10][i] = or_myround(2. * (1. - kr) * cc * OR_SF);
        c->t[11][i] = or_myround(gcb * cc * OR_SF);
        c->t[12][i] = or_myround(-.5 / (1. - kr) * cc * OR_SF);
        c->t[13][i] = or_myround(2. * (1. - kb) * cc * OR_SF);
      }
    }
  }
}

/* cache of the four (clamping, subspace) variants */
static or_conv_t or_cache[2][2];
static int or_cache_ok[2][2];
static const or_conv_t *or_conv(int clamping, int subspace) {
  int a = clamping == OR_CLAMPED ? 0 : 1, b = subspace == OR_SUBSPACE_BT709 ? 1 : 0;
  if (!or_cache_ok[a][b]) { or_build_conv(clamping, subspace, &or_cache[a][b]); or_cache_ok[a][b] = 1; }
  return &or_cache[a][b];
}

void pe_or_conv_table(int clamping, int subspace, int which, int32_t out[256]) {
  memcpy(out, or_conv(clamping, subspace)->t[which], 256 * sizeof(int32_t));
}

/* ---- premultiply tables  src/colourspace.c:1141-1160 ---------------------- */

void pe_or_premult_table(int which, int32_t out[65536]) {
  for (int i = 0; i < 256; i++) {
    float alpha = (float)255. / (float)i;
    for (int j = 0; j < 256; j++) {
      int32_t v;
      switch (which) {
      case 0: v = or_clamp0255f((float)j / alpha); break;
      case 1: v = or_clamp0255f((float)j * alpha); break;
      case 2: v = (int)((float)j / alpha + .5) > (235. - 16.) ? 235
                  : (int)((float)(j - 16.) / alpha + 16. + .5); break;
      case 3: v = (int)((float)j / alpha + .5) > (240. - 16.) ? 240
                  : (int)((float)(j - 16.) / alpha + 16. + .5); break;
      case 4: v = or_clamp0255f((float)(j - 16.) * alpha + 16.); break;
      default: v = or_clamp0255f((float)(j - 128.) * alpha + 128.); break;
      }
      out[i * 256 + j] = v;
    }
  }
}

/* ---- gamma LUTs  src/colourspace.c:655-819, colourspace.h:152-185 ---------- */

typedef struct { float offs, lin, thresh, pf; } or_gamma_const;

static void or_gamma_consts(or_gamma_const g[2]) {
  /* INIT_GAMMA colourspace.h:157-161; sRGB (12.92, 0.04045, 2.4), BT709 (4.5, 0.018, 1/.45) :168-169 */
  g[0].offs = 0.; g[0].lin = 12.92; g[0].thresh = 0.04045; g[0].pf = 2.4;
  g[1].offs = 0.; g[1].lin = 4.5; g[1].thresh = 0.018; g[1].pf = 1. / .45;
  for (int k = 0; k < 2; k++) {
    g[k].offs = (powf((g[k].thresh / g[k].lin), (1. / g[k].pf)) - g[k].thresh)
                / (1. - (powf((g[k].thresh / g[k].lin), (1. / g[k].pf))));
  }
}

static int or_gamma_idx(int gamma_type) { return gamma_type == OR_GAMMA_BT709 ? 1 : 0; } /* :625-628 */

/* one LUT entry; *gamma_from is mutated across calls exactly as the by-value parameter is
 * mutated across loop iterations in the reference (:697-713) */
static float or_gamma_entry(float a0, double fileg, int *gamma_from, int gamma_to, double screen_gamma,
                            float inv_gamma, const or_gamma_const g[2]) {
  float a = a0, x = a0;
  int idx;
  if (fileg != 1.0) x = powf(a, fileg);
  if (*gamma_from == OR_GAMMA_MONITOR) {
    x = powf(a, screen_gamma);
    *gamma_from = OR_GAMMA_SRGB;
  }
  if (*gamma_from != OR_GAMMA_LINEAR && !(*gamma_from == OR_GAMMA_SRGB && gamma_to == OR_GAMMA_MONITOR)) {
    idx = or_gamma_idx(*gamma_from);
    a = (a < g[idx].thresh) ? a / g[idx].lin : powf((a + g[idx].offs) / (1. + g[idx].offs), g[idx].pf);
    *gamma_from = OR_GAMMA_LINEAR;
  }
  if (gamma_to != OR_GAMMA_LINEAR) {
    idx = (gamma_to == OR_GAMMA_MONITOR) ? or_gamma_idx(OR_GAMMA_SRGB) : or_gamma_idx(gamma_to);
    x = (a < (g[idx].thresh) / g[idx].lin) ? a * g[idx].lin
        : powf((1. + g[idx].offs) * a, 1. / g[idx].pf) - g[idx].offs;
  }
  if (gamma_to == OR_GAMMA_MONITOR) x = powf(a, inv_gamma);
  return x;
}

int pe_or_gamma_lut8(double fileg, int gamma_from, int gamma_to, double screen_gamma, uint8_t out[256]) {
  or_gamma_const g[2];
  float inv_gamma = 0.;
  if (fileg == 1.0 && (gamma_to == gamma_from || gamma_to == OR_GAMMA_UNKNOWN || gamma_from == OR_GAMMA_UNKNOWN))
    return -1; /* :662-663 returns NULL */
  or_gamma_consts(g);
  if (gamma_to == OR_GAMMA_MONITOR) inv_gamma = 1. / (float)screen_gamma;
  out[0] = 0;
  for (int i = 1; i < 256; ++i) {
    float x = or_gamma_entry((float)i / 255., fileg, &gamma_from, gamma_to, screen_gamma, inv_gamma, g);
    out[i] = (uint8_t)or_clamp0_255((int)(x * 255.)); /* CLAMP0_255i :716 */
  }
  return 0;
}

int pe_or_gamma_lut16(double fileg, int gamma_from, int gamma_to, double screen_gamma, uint16_t out[65536]) {
  or_gamma_const g[2];
  float inv_gamma = 0.;
  if (fileg == 1.0 && (gamma_to == gamma_from || gamma_to == OR_GAMMA_UNKNOWN || gamma_from == OR_GAMMA_UNKNOWN))
    return -1;
  or_gamma_consts(g);
  if (gamma_to == OR_GAMMA_MONITOR) inv_gamma = 1. / (float)screen_gamma;
  out[0] = 0;
  for (int i = 1; i < 65536; ++i) {
    float x = or_gamma_entry((float)i / 65536., fileg, &gamma_from, gamma_to, screen_gamma, inv_gamma, g);
    /* CLAMP16bit colourspace.h:16 */
    out[i] = (x) >= 0.99999 ? 65535 : x < 0.00001 ? 0 : (uint16_t)(x * 65535.9999);
  }
  return 0;
}

/* ---- per-pixel kernels  src/colourspace.c:2119-2127, :2345-2356 ------------ */

static inline void or_px_rgb2yuv(const or_conv_t *c, int q, uint8_t r, uint8_t g, uint8_t b, uint8_t *y, uint8_t *u,
                                 uint8_t *v) {
  short a;
  if ((a = or_spc_rnd(c->t[0][r] + c->t[1][g] + c->t[2][b], q)) > c->max_y) *y = c->max_y;
  else *y = a < c->min_y ? c->min_y : a;
  if ((a = or_spc_rnd(c->t[3][r] + c->t[4][g] + c->t[5][b], q)) > c->max_uv) *u = c->max_uv;
  else *u = a < c->min_uv ? c->min_uv : a;
  if ((a = or_spc_rnd(c->t[6][r] + c->t[7][g] + c->t[8][b], q)) > c->max_uv) *v = c->max_uv;
  else *v = a < c->min_uv ? c->min_uv : a;
}

static inline void or_px_yuv2rgb(const or_conv_t *c, int q, const uint16_t *lut16, uint8_t y, uint8_t u, uint8_t v,
                                 uint8_t *r, uint8_t *g, uint8_t *b) {
  int yy = c->t[9][y];
  if (!lut16) {
    *r = or_clamp0255f(or_spc_rnd(yy + c->t[10][v], q));
    *g = or_clamp0255f(or_spc_rnd(yy + c->t[11][u] + c->t[12][v], q));
    *b = or_clamp0255f(or_spc_rnd(yy + c->t[13][u], q));
  } else {
    /* xyuv2rgb_with_gamma :2386-2390 */
    int t;
    t = (yy + c->t[10][v]) >> 8; t = t > 65535 ? 65535 : t < 0 ? 0 : t; *r = lut16[t] >> 8;
    t = (yy + c->t[11][u] + c->t[12][v]) >> 8; t = t > 65535 ? 65535 : t < 0 ? 0 : t; *g = lut16[t] >> 8;
    t = (yy + c->t[13][u]) >> 8; t = t > 65535 ? 65535 : t < 0 ? 0 : t; *b = lut16[t] >> 8;
  }
}

void pe_or_rgb2yuv(int clamping, int subspace, int quality, const uint8_t *rgb, uint8_t *yuv, long n) {
  const or_conv_t *c = or_conv(clamping, subspace);
  for (long i = 0; i < n; i++)
    or_px_rgb2yuv(c, quality, rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], &yuv[3 * i], &yuv[3 * i + 1], &yuv[3 * i + 2]);
}

void pe_or_yuv2rgb(int clamping, int subspace, int quality, const uint8_t *yuv, uint8_t *rgb, long n) {
  const or_conv_t *c = or_conv(clamping, subspace);
  for (long i = 0; i < n; i++)
    or_px_yuv2rgb(c, quality, NULL, yuv[3 * i], yuv[3 * i + 1], yuv[3 * i + 2], &rgb[3 * i], &rgb[3 * i + 1], &rgb[3 * i + 2]);
}

/* byte offsets of R,G,B,(A) inside an output pixel for each order */
static void or_order_offsets(int order, int add_alpha, int *ro, int *go, int *bo, int *ao, int *psize) {
  if (order == OR_ORDER_ARGB) { *ao = 0; *ro = 1; *go = 2; *bo = 3; *psize = 4; return; }
  *psize = add_alpha ? 4 : 3; *ao = add_alpha ? 3 : -1; *go = 1;
  if (order == OR_ORDER_RGB) { *ro = 0; *bo = 2; } else { *ro = 2; *bo = 0; }
}

/* ---- planar 4:2:0 / 4:2:2 -> RGB  src/colourspace.c:3260-3904 --------------- */

/* chroma sample one past the end of a chroma row: the reference reads plane[stride*r + cw]
 * (:3508-3512, :3613).  That byte is row padding, or the first sample of row r+1; only on the
 * last chroma row of a plane whose stride equals its width does it lie outside the plane, and
 * there we define it as the replicated edge sample. */
static inline int or_chroma_at(const uint8_t *p, int stride, int r, int c, int cw, int ch) {
  if (c < cw) return p[(long)stride * r + c];
  if (cw < stride || r + 1 < ch) return p[(long)stride * r + cw];
  return p[(long)stride * r + cw - 1];
}

void pe_or_yuv420p_to_rgb(const uint8_t *const src[3], const int istrides[3], int width, int height,
                          uint8_t *dest, int orowstride, int order, int add_alpha, int is_422,
                          int clamping, int subspace, int quality, int quirks, const uint16_t *lut16) {
  const or_conv_t *c = or_conv(clamping, subspace);
  const uint8_t *s_y = src[0], *s_u = src[1], *s_v = src[2];
  const int cw = width >> 1, ch = is_422 ? height : (height + 1) >> 1;
  int ro, go, bo, ao, ps;
  int (*clampf)(int) = clamping == OR_CLAMPED ? or_clamp16_240 : or_clamp0_255;
  or_order_offsets(order, add_alpha, &ro, &go, &bo, &ao, &ps);

#define OR_EMIT(row, col, yv, uv, vv) do { \
    uint8_t *d_ = dest + (long)orowstride * (row) + (long)(col) * ps; \
    or_px_yuv2rgb(c, quality, lut16, (yv), (uint8_t)(uv), (uint8_t)(vv), &d_[ro], &d_[go], &d_[bo]); \
    if (ao >= 0) d_[ao] = 255; } while (0)

  if (is_422) {
    /* :3598-3642.  quirk (R): the running "last/this" pair is seeded from chroma row i>>1 */
    for (int i = 0; i < height; i++) {
      int seed_row = quirks ? (i >> 1) : i;
      int last_u = s_u[(long)istrides[1] * seed_row], last_v = s_v[(long)istrides[2] * seed_row];
      int this_u = last_u, this_v = last_v;
      for (int j = 0; j < width; j += 2) {
        int jc = j >> 1;
        int u1 = clampf((this_u + last_u) >> 1), v1 = clampf((this_v + last_v) >> 1);
        int next_u = or_chroma_at(s_u, istrides[1], i, jc + 1, cw, ch);
        int next_v = or_chroma_at(s_v, istrides[2], i, jc + 1, cw, ch);
        int u2 = clampf((this_u + next_u) >> 1), v2 = clampf((this_v + next_v) >> 1);
        last_u = this_u; last_v = this_v; this_u = next_u; this_v = next_v;
        OR_EMIT(i, j, s_y[(long)istrides[0] * i + j], u1, v1);
        OR_EMIT(i, j + 1, s_y[(long)istrides[0] * i + j + 1], u2, v2);
      }
    }
    return;
  }

  /* rows 0 and (even height) height-1: single chroma row, horizontal average only.
   * X rows: the reference indexes the tables with an unshifted sum (:3421-3428) and reads
   * luma row 0 / writes row 0 for the last row (:3561,:3584); we define the evident intent. */
  for (int pass = 0; pass < 2; pass++) {
    int i = pass == 0 ? 0 : height - 1;
    if (pass == 1 && (height < 2 || (height & 1))) break;
    int r2 = i >> 1;
    int last_u = s_u[(long)istrides[1] * r2], last_v = s_v[(long)istrides[2] * r2];
    int this_u = last_u, this_v = last_v;
    for (int j = 0; j < width; j += 2) {
      int jc = j >> 1;
      int u1 = clampf((this_u + last_u) >> 1), v1 = clampf((this_v + last_v) >> 1);
      int next_u = or_chroma_at(s_u, istrides[1], r2, jc + 1, cw, ch);
      int next_v = or_chroma_at(s_v, istrides[2], r2, jc + 1, cw, ch);
      int u2 = clampf((this_u + next_u) >> 1), v2 = clampf((this_v + next_v) >> 1);
      OR_EMIT(i, j, s_y[(long)istrides[0] * i + j], u1, v1);
      OR_EMIT(i, j + 1, s_y[(long)istrides[0] * i + j + 1], u2, v2);
      last_u = this_u; last_v = this_v; this_u = next_u; this_v = next_v;
    }
  }

  /* interior row pairs (i, i+1), i odd: chroma rows r2 = i>>1 and r2+1, weights 2/3-1/3 (:3440-3549) */
  for (int i = 1; i < height - 1; i += 2) {
    int r2 = i >> 1;
    int last_u1 = s_u[(long)istrides[1] * r2], last_v1 = s_v[(long)istrides[2] * r2];
    int last_u2 = s_u[(long)istrides[1] * (r2 + 1)], last_v2 = s_v[(long)istrides[2] * (r2 + 1)];
    int this_u1 = last_u1, this_v1 = last_v1, this_u2 = last_u2, this_v2 = last_v2;
    for (int j = 0; j < width; j += 2) {
      int jc = j >> 1, u1, u2, v1, v2, u3, u4, v3, v4;
      int next_u1, next_v1, next_u2, next_v2;
      /* left pixel */
      u1 = this_u1 + last_u1;
      v1 = this_v1 + last_v1;
      u2 = quirks ? this_u1 + last_u1 : this_u2 + last_u2; /* :3461 */
      v2 = this_v2 + last_v2;
      if (quality != OR_QUALITY_LOW || order != OR_ORDER_RGB) { /* bgr/argb variants have no LOW branch (:4090) */
        u3 = clampf((int)((u1 + (u2 >> 1)) / 3. + .5));
        u4 = clampf((int)(((u1 >> 1) + u2) / 3. + .5));
        v3 = clampf((int)((v1 + (v2 >> 1)) / 3. + .5));
        v4 = clampf((int)(((v1 >> 1) + v2) / 3. + .5));
      } else {
        u3 = clampf(u1 >> 1); u4 = clampf(u2 >> 1); v3 = clampf(v1 >> 1); v4 = clampf(v2 >> 1);
      }
      OR_EMIT(i, j, s_y[(long)istrides[0] * i + j], u3, v3);
      OR_EMIT(i + 1, j, s_y[(long)istrides[0] * (i + 1) + j], u4, v4);
      /* right pixel */
      next_u1 = or_chroma_at(s_u, istrides[1], r2, jc + 1, cw, ch);
      next_v1 = or_chroma_at(s_v, istrides[2], r2, jc + 1, cw, ch);
      next_u2 = or_chroma_at(s_u, istrides[1], r2 + 1, jc + 1, cw, ch);
      next_v2 = or_chroma_at(s_v, istrides[2], r2 + 1, jc + 1, cw, ch);
      u1 = this_u1 + next_u1; v1 = this_v1 + next_v1;
      u2 = this_u2 + next_u2; v2 = this_v2 + next_v2;
      if (quality != OR_QUALITY_LOW || order != OR_ORDER_RGB) {
        u3 = clampf((int)((u1 + (u2 >> 1)) / 3. + .5));
        u4 = clampf((int)(((u1 >> 1) + u2) / 3. + .5));
        v3 = clampf((int)((v1 + (v2 >> 1)) / 3. + .5));
        v4 = clampf((int)(((v1 >> 1) + v2) / 3. + .5));
      } else {
        u3 = clampf(u1 >> 1); u4 = clampf(u2 >> 1); v3 = clampf(v1 >> 1); v4 = clampf(v2 >> 1);
      }
      OR_EMIT(i, j + 1, s_y[(long)istrides[0] * i + j + 1], u3, v3);
      OR_EMIT(i + 1, j + 1, s_y[(long)istrides[0] * (i + 1) + j + 1], u4, v4);
      /* :3538-3546 -- with quirks, last_v1 takes this_v2 and last_v2 is never advanced */
      last_u1 = this_u1; this_u1 = next_u1;
      last_u2 = this_u2; this_u2 = next_u2;
      if (quirks) { last_v1 = this_v2; }
      else { last_v1 = this_v1; last_v2 = this_v2; }
      this_v1 = next_v1; this_v2 = next_v2;
    }
  }
#undef OR_EMIT
}

/* ---- packed 4:2:2 -> RGB  src/colourspace.c:6616-7103, uyvy2rgb :2410 ------- */

void pe_or_packed422_to_rgb(int fmt, const uint8_t *src, int irow, int width_mpx, int height,
                            uint8_t *dest, int orowstride, int order, int add_alpha, int clamping, int subspace,
                            int quality) {
  /* only convert_uyvy_to_rgb_frame honours the subspace (:6624); the other five select YCbCr */
  const or_conv_t *c = or_conv(clamping, (fmt == 0 && order == OR_ORDER_RGB) ? subspace : OR_SUBSPACE_YCBCR);
  int ro, go, bo, ao, ps;
  or_order_offsets(order, add_alpha, &ro, &go, &bo, &ao, &ps);
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    uint8_t *d = dest + (long)orowstride * i;
    for (int j = 0; j < width_mpx; j++, s += 4, d += 2 * ps) {
      uint8_t y0, y1, u, v;
      if (fmt == 0) { u = s[0]; y0 = s[1]; v = s[2]; y1 = s[3]; }
      else { y0 = s[0]; u = s[1]; y1 = s[2]; v = s[3]; }
      or_px_yuv2rgb(c, quality, NULL, y0, u, v, &d[ro], &d[go], &d[bo]);
      or_px_yuv2rgb(c, quality, NULL, y1, u, v, &d[ps + ro], &d[ps + go], &d[ps + bo]);
      if (ao >= 0) d[ao] = d[ps + ao] = 255;
    }
  }
}

/* ---- packed 4:4:4  src/colourspace.c:2750-3258 / :5700-6239 ---------------- */

void pe_or_yuv888_to_rgb(const uint8_t *src, int irow, int width, int height, uint8_t *dest, int orow,
                         int order, int in_alpha, int out_alpha, int clamping, int subspace, int quality) {
  const or_conv_t *c = or_conv(clamping, subspace);
  int ro, go, bo, ao, ps, ips = in_alpha ? 4 : 3;
  or_order_offsets(order, out_alpha, &ro, &go, &bo, &ao, &ps);
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    uint8_t *d = dest + (long)orow * i;
    for (int j = 0; j < width; j++, s += ips, d += ps) {
      or_px_yuv2rgb(c, quality, NULL, s[0], s[1], s[2], &d[ro], &d[go], &d[bo]);
      if (ao >= 0) d[ao] = in_alpha ? s[3] : 255;
    }
  }
}

void pe_or_rgb_to_yuv888(const uint8_t *src, int irow, int width, int height, uint8_t *dest, int orow,
                         int order, int in_alpha, int out_alpha, int clamping, int quality) {
  /* always YCbCr tables (:5710); width is rounded down to even (:5750) */
  const or_conv_t *c = or_conv(clamping, OR_SUBSPACE_YCBCR);
  int ro, go, bo, ao, ips, ops = out_alpha ? 4 : 3;
  or_order_offsets(order, in_alpha, &ro, &go, &bo, &ao, &ips);
  width = (width >> 1) << 1;
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    uint8_t *d = dest + (long)orow * i;
    for (int j = 0; j < width; j++, s += ips, d += ops) {
      if (out_alpha) d[3] = ao >= 0 ? s[ao] : 255;
      or_px_rgb2yuv(c, quality, s[ro], s[go], s[bo], &d[0], &d[1], &d[2]);
    }
  }
}

/* ---- RGB -> packed 4:2:2  src/colourspace.c:5129-5700 (rgb2uyvy :2162, rgb2yuyv :2176) ---------------------
 * One macropixel from two pixels: Cb from the FIRST pixel, Cr from the SECOND (no chroma averaging), always the YCbCr tables
 * (:5146).  YUYV: the max clamp of U and V is overwritten by the following statement (missing `else`, :2183-2184,2189-2190),
 * so values above max_UV pass through -- replicated.  Rows are written with the output rowstride (the reference's own row
 * advance `orowstride / 2 - hsize` in macropixel units only works for unpadded rows, :5206). */
void pe_or_rgb_to_packed422(int fmt, const uint8_t *src, int irow, int width, int height, uint8_t *dest, int orow, int order,
                            int in_alpha, int clamping, int quality, const uint16_t *lut16) {
  const or_conv_t *c = or_conv(clamping, OR_SUBSPACE_YCBCR);
  int ro, go, bo, ao, ips;
  or_order_offsets(order, in_alpha, &ro, &go, &bo, &ao, &ips);
  width = (width >> 1) << 1;
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    uint8_t *d = dest + (long)orow * i;
    for (int j = 0; j < width; j += 2, s += 2 * ips, d += 4) {
      const uint8_t r0 = s[ro], g0 = s[go], b0 = s[bo], r1 = s[ips + ro], g1 = s[ips + go], b1 = s[ips + bo];
      short au, ay0, av, ay1;
      if (lut16) { /* rgb2uyvy_with_gamma :2146 */
        au = lut16[(c->t[3][r0] + c->t[4][g0] + c->t[5][b0]) >> 8] >> 8;
        ay0 = lut16[(c->t[0][r0] + c->t[1][g0] + c->t[2][b0]) >> 8] >> 8;
        av = lut16[(c->t[6][r1] + c->t[7][g1] + c->t[8][b1]) >> 8] >> 8;
        ay1 = lut16[(c->t[0][r1] + c->t[1][g1] + c->t[2][b1]) >> 8] >> 8;
      } else {
        au = or_spc_rnd(c->t[3][r0] + c->t[4][g0] + c->t[5][b0], quality);
        ay0 = or_spc_rnd(c->t[0][r0] + c->t[1][g0] + c->t[2][b0], quality);
        av = or_spc_rnd(c->t[6][r1] + c->t[7][g1] + c->t[8][b1], quality);
        ay1 = or_spc_rnd(c->t[0][r1] + c->t[1][g1] + c->t[2][b1], quality);
      }
      {
        const uint8_t y0 = ay0 > c->max_y ? c->max_y : ay0 < c->min_y ? c->min_y : ay0;
        const uint8_t y1 = ay1 > c->max_y ? c->max_y : ay1 < c->min_y ? c->min_y : ay1;
        uint8_t u, v;
        if (fmt == 0) {
          u = au > c->max_uv ? c->max_uv : au < c->min_uv ? c->min_uv : au;
          v = av > c->max_uv ? c->max_uv : av < c->min_uv ? c->min_uv : av;
          d[0] = u; d[1] = y0; d[2] = v; d[3] = y1;
        } else {
          u = au < c->min_uv ? c->min_uv : (uint8_t)au; /* missing else: no max clamp */
          v = av < c->min_uv ? c->min_uv : (uint8_t)av;
          d[0] = y0; d[1] = u; d[2] = y1; d[3] = v;
        }
      }
    }
  }
}

/* ---- RGB -> planar 4:4:4  src/colourspace.c:5786-5880 (rgb), :5971 (bgr), :6154 (argb) ----------------------- */
void pe_or_rgb_to_yuv444p(const uint8_t *src, int irow, int width, int height, uint8_t *const dest[4], int orow, int order,
                          int in_alpha, int out_alpha, int clamping, int quality) {
  const or_conv_t *c = or_conv(clamping, OR_SUBSPACE_YCBCR);
  int ro, go, bo, ao, ips;
  or_order_offsets(order, in_alpha, &ro, &go, &bo, &ao, &ips);
  width = (width >> 1) << 1;
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    for (int j = 0; j < width; j++, s += ips) {
      const long o = (long)orow * i + j;
      if (out_alpha) dest[3][o] = ao >= 0 ? s[ao] : 255;
      or_px_rgb2yuv(c, quality, s[ro], s[go], s[bo], &dest[0][o], &dest[1][o], &dest[2][o]);
    }
  }
}

/* ---- chroma averaging tables  src/colourspace.c:190-217 (init_average, !MULT_AVG) ---------------------------
 * which 0: cavgc (clamped: float maths, result clamped to 16..240)   1: cavgu (unclamped: ((x-128)+(y-128) >> 1) + 128) */
void pe_or_avg_table(int which, uint8_t out[65536]) {
  for (int x = 0; x < 256; x++) {
    float fa = (float)(x - 128.) * 255. / 244.;
    short sa = (short)(x - 128);
    for (int y = 0; y < 256; y++) {
      float fb = (float)(y - 128.) * 255. / 244.;
      short sb = (short)(y - 128);
      float fc = (fa + fb) * 224. / 512. + 128.;
      short c = ((sa + sb) >> 1) + 128;
      out[x * 256 + y] = which == 0 ? (uint8_t)(fc > 240. ? 240 : fc < 16. ? 16 : fc) : (uint8_t)(c > 255 ? 255 : c < 0 ? 0 : c);
    }
  }
}

/* ---- RGB -> planar 4:2:0 / 4:2:2  src/colourspace.c:6250-6320 (rgb), :6385-6440 (bgr) ---------------------------
 * One rgb2uyvy macropixel (:2162) per pixel pair: Y of both pixels, Cb of the FIRST, Cr of the SECOND; tables of the
 * (clamping, subspace) handed in -- the dispatcher passes osubspace for 4:2:0 and WEED_YUV_SAMPLING_DEFAULT (= YCbCr) in
 * that slot for 4:2:2 (:12681-12690).  Width and height are cut to even.
 * 4:2:0 chroma (the pointer dance of :6291-6306): luma row 0 writes chroma row 0, row 1 overwrites it; from then on every
 * even row 2c+2 writes chroma row c+1 and folds itself into row c, avg_chroma(new, old); the following odd row overwrites
 * row c+1.  Net effect:  C[c] = cavg[C(2c+2)][C(2c+1)] for c < h/2 - 1,  C[h/2 - 1] = C(h - 1)  (luma row 0 contributes
 * nothing).  The reference advances its Y / Cb / Cr pointers densely, so it is only right for unpadded planes
 * (ostride == width); this restatement uses the plane strides and equals it there.
 * ARGB (:6323): reads G1 / B1 of the second pixel two / one bytes too far (:6357-6358, past the row on the last pair) -- X,
 * not restated. */
void pe_or_rgb_to_yuv420p(const uint8_t *src, int irow, int width, int height, uint8_t *const dest[3], const int ostrides[3],
                          int order, int in_alpha, int is_422, int clamping, int subspace, int quality) {
  const or_conv_t *c = or_conv(clamping, subspace);
  static uint8_t avg[2][65536];
  static int avg_ok = 0;
  int ro, go, bo, ao, ips;
  if (!avg_ok) { pe_or_avg_table(0, avg[0]); pe_or_avg_table(1, avg[1]); avg_ok = 1; }
  const uint8_t *cavg = avg[clamping == OR_CLAMPED ? 0 : 1];
  or_order_offsets(order, in_alpha, &ro, &go, &bo, &ao, &ips);
  width = (width >> 1) << 1;
  height = (height >> 1) << 1;
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    uint8_t *y = dest[0] + (long)ostrides[0] * i;
    /* chroma row this luma row lands in (rows 2c and 2c+1 both write row c, the odd one last), and whether it also folds
     * into the row above */
    const int crow = is_422 ? i : i >> 1;
    const int fold = !is_422 && i > 0 && !(i & 1);           /* even rows > 0 average into row crow - 1 */
    uint8_t *cb = dest[1] + (long)ostrides[1] * crow;
    uint8_t *cr = dest[2] + (long)ostrides[2] * crow;
    for (int j = 0; j < width; j += 2, s += 2 * ips) {
      const uint8_t r0 = s[ro], g0 = s[go], b0 = s[bo], r1 = s[ips + ro], g1 = s[ips + go], b1 = s[ips + bo];
      short a;
      uint8_t u, v, y0, y1;
      a = or_spc_rnd(c->t[3][r0] + c->t[4][g0] + c->t[5][b0], quality);
      u = a > c->max_uv ? c->max_uv : a < c->min_uv ? c->min_uv : a;
      a = or_spc_rnd(c->t[0][r0] + c->t[1][g0] + c->t[2][b0], quality);
      y0 = a > c->max_y ? c->max_y : a < c->min_y ? c->min_y : a;
      a = or_spc_rnd(c->t[6][r1] + c->t[7][g1] + c->t[8][b1], quality);
      v = a > c->max_uv ? c->max_uv : a < c->min_uv ? c->min_uv : a;
      a = or_spc_rnd(c->t[0][r1] + c->t[1][g1] + c->t[2][b1], quality);
      y1 = a > c->max_y ? c->max_y : a < c->min_y ? c->min_y : a;
      y[j] = y0; y[j + 1] = y1;
      if (fold) {
        uint8_t *pb = dest[1] + (long)ostrides[1] * (crow - 1) + (j >> 1), *pr = dest[2] + (long)ostrides[2] * (crow - 1) + (j >> 1);
        *pb = cavg[((int)u << 8) + *pb];
        *pr = cavg[((int)v << 8) + *pr];
      }
      cb[j >> 1] = u; cr[j >> 1] = v;
    }
  }
}

/* ---- RGB <-> RGB  src/colourspace.c:12370-12556 dispatch, loops :9259-10515 -- */

static int or_rgb_layout(int pal, int *ro, int *go, int *bo, int *ao, int *ps) {
  switch (pal) {
  case OR_PAL_RGB24: *ro = 0; *go = 1; *bo = 2; *ao = -1; *ps = 3; return 0;
  case OR_PAL_BGR24: *ro = 2; *go = 1; *bo = 0; *ao = -1; *ps = 3; return 0;
  case OR_PAL_RGBA32: *ro = 0; *go = 1; *bo = 2; *ao = 3; *ps = 4; return 0;
  case OR_PAL_BGRA32: *ro = 2; *go = 1; *bo = 0; *ao = 3; *ps = 4; return 0;
  case OR_PAL_ARGB32: *ro = 1; *go = 2; *bo = 3; *ao = 0; *ps = 4; return 0;
  }
  return -1;
}

int pe_or_rgb_to_rgb(int inpal, int outpal, const uint8_t *src, int irow, int width, int height,
                     uint8_t *dest, int orow, const uint8_t *lut8) {
  int iro, igo, ibo, iao, ips, oro, ogo, obo, oao, ops;
  if (or_rgb_layout(inpal, &iro, &igo, &ibo, &iao, &ips) || or_rgb_layout(outpal, &oro, &ogo, &obo, &oao, &ops)) return -1;
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    uint8_t *d = dest + (long)orow * i;
    for (int j = 0; j < width; j++, s += ips, d += ops) {
      uint8_t r = s[iro], g = s[igo], b = s[ibo], a = iao >= 0 ? s[iao] : 255;
      if (lut8) { r = lut8[r]; g = lut8[g]; b = lut8[b]; }
      d[oro] = r; d[ogo] = g; d[obo] = b;
      if (oao >= 0) d[oao] = a;
    }
  }
  return 0;
}

/* ---- gamma apply  src/colourspace.c:14034-14062 ---------------------------- */

void pe_or_gamma_apply(uint8_t *pixels, int rowstride, int palette, int x, int y, int width, int height,
                       const uint8_t lut8[256]) {
  int ro, go, bo, ao, ps;
  if (or_rgb_layout(palette, &ro, &go, &bo, &ao, &ps)) return;
  {
    const int px = ps < 3 ? ps : 3;
    const int start = x * ps + (palette == OR_PAL_ARGB32 ? 1 : 0);
    const int end = start + width * ps;
    for (int i = 0; i < height; i++) {
      uint8_t *row = pixels + (long)rowstride * (y + i);
      for (int j = start; j < end; j += ps)
        for (int k = 0; k < px; k++) row[j + k] = lut8[row[j + k]];
    }
  }
}

/* ---- alpha premultiply  src/colourspace.c:11968-12106 ---------------------- */

void pe_or_alpha_premult(uint8_t *pixels, int rowstride, int palette, int clamping, int width, int height,
                         int direction) {
  static int32_t *tabs[6];
  int psize = 4, psizel, coffs, aoffs;
  if (!tabs[0]) for (int k = 0; k < 6; k++) { tabs[k] = (int32_t *)malloc(65536 * sizeof(int32_t)); pe_or_premult_table(k, tabs[k]); }
  switch (palette) {
  case OR_PAL_RGBA32: case OR_PAL_BGRA32: case OR_PAL_YUVA8888: psizel = 3; coffs = 0; aoffs = 3; break;
  case OR_PAL_ARGB32: psizel = 4; coffs = 1; aoffs = 0; break;
  default: return;
  }
  if (palette != OR_PAL_YUVA8888 || clamping != OR_CLAMPED) {
    /* REVERSE indexes unal, FORWARD indexes al (:12058-12074) */
    const int32_t *t = direction < 0 ? tabs[0] : tabs[1];
    for (int i = 0; i < height; i++) {
      uint8_t *ptr = pixels + (long)i * rowstride;
      for (int j = 0; j < width * psize; j += psize) {
        int alpha = ptr[j + aoffs];
        for (int p = coffs; p < psizel; p++) ptr[j + p] = (uint8_t)t[alpha * 256 + ptr[j + p]];
      }
    }
  } else {
    /* clamped YUVA8888 (:12076-12098); forward path reads ptr[j] for U and V as written (:12093-12094) */
    for (int i = 0; i < height; i++) {
      uint8_t *ptr = pixels + (long)i * rowstride;
      for (int j = 0; j < width * psize; j += psize) {
        int alpha = ptr[j + 3];
        if (direction < 0) {
          ptr[j] = (uint8_t)tabs[2][alpha * 256 + ptr[j]];
          ptr[j + 1] = (uint8_t)tabs[4][alpha * 256 + ptr[j + 1]];
          ptr[j + 2] = (uint8_t)tabs[4][alpha * 256 + ptr[j + 2]];
        } else {
          ptr[j] = (uint8_t)tabs[3][alpha * 256 + ptr[j]];
          ptr[j + 1] = (uint8_t)tabs[5][alpha * 256 + ptr[j]];
          ptr[j + 2] = (uint8_t)tabs[5][alpha * 256 + ptr[j]];
        }
      }
    }
  }
}

/* YUVA4444P (:12001-12049): every plane through its own table with the sample's alpha -- unal / al when unclamped, (un)alcy for the
 * luma plane and (un)alcuv for both chroma planes when clamped; each plane reads its ORIGINAL value (no rewritten-Y slip here) */
void pe_or_alpha_premult_planar(uint8_t *const planes[4], const int rows[4], int clamping, int width, int height, int direction) {
  static int32_t *tabs[6];
  if (!tabs[0]) for (int k = 0; k < 6; k++) { tabs[k] = (int32_t *)malloc(65536 * sizeof(int32_t)); pe_or_premult_table(k, tabs[k]); }
  const int32_t *ty, *tc;
  if (clamping != OR_CLAMPED) ty = tc = direction < 0 ? tabs[0] : tabs[1];
  else if (direction < 0) { ty = tabs[2]; tc = tabs[4]; }
  else { ty = tabs[3]; tc = tabs[5]; }
  for (int i = 0; i < height; i++)
    for (int j = 0; j < width; j++) {
      const int alpha = planes[3][(long)rows[3] * i + j];
      uint8_t *y = planes[0] + (long)rows[0] * i + j, *u = planes[1] + (long)rows[1] * i + j, *v = planes[2] + (long)rows[2] * i + j;
      *y = (uint8_t)ty[alpha * 256 + *y]; *u = (uint8_t)tc[alpha * 256 + *u]; *v = (uint8_t)tc[alpha * 256 + *v];
    }
}

/* ---- effects --------------------------------------------------------------- */

/* calc_luma libweed/weed-plugin-utils.c:924-934 with its own 16.16 tables (:881-886, SCALE_FACTOR 65536) */
static int32_t or_lY_R[256], or_lY_G[256], or_lY_B[256];
static int or_luma_ok;
static uint8_t or_calc_luma(const uint8_t *px, int palette) {
  if (!or_luma_ok) {
    for (int i = 0; i < 256; i++) {
      or_lY_R[i] = or_myround(0.299 * (double)i * 65536.);
      or_lY_G[i] = or_myround((1. - 0.299 - 0.114) * (double)i * 65536.);
      or_lY_B[i] = or_myround(0.114 * (double)i * 65536.);
    }
    or_luma_ok = 1;
  }
  switch (palette) {
  case OR_PAL_RGB24: case OR_PAL_RGBA32: return (or_lY_R[px[0]] + or_lY_G[px[1]] + or_lY_B[px[2]]) >> 16;
  case OR_PAL_BGR24: case OR_PAL_BGRA32: return (or_lY_R[px[2]] + or_lY_G[px[1]] + or_lY_B[px[0]]) >> 16;
  case OR_PAL_ARGB32: return (or_lY_R[px[1]] + or_lY_G[px[2]] + or_lY_B[px[3]]) >> 16;
  }
  return 0;
}

/* simple_blend.c:58-200.  src2_bytes bounds the ARGB "next pixel alpha" read (:130, start = 1) */
void pe_or_simple_blend(int type, int palette, const uint8_t *src1, int irow1, const uint8_t *src2, int irow2,
                        uint8_t *dst, int orow, int width, int height, int bf, long src2_bytes) {
  const int psize = (palette == OR_PAL_RGB24 || palette == OR_PAL_BGR24) ? 3 : 4;
  const int widthx = width * psize;
  const int start = palette == OR_PAL_ARGB32 ? 1 : 0;
  const uint8_t blend_factor = (uint8_t)bf, blendneg = 0xFF - blend_factor;
  const int inplace = (src1 == dst);
  if (type == 0) {
#define OR_BL(i_, j_) ((uint8_t)((blend_factor * (i_) + blendneg * (j_)) >> 8)) /* make_blend_table :31-35 */
    for (int i = 0; i < height; i++) {
      const long o = (long)orow * i, r1 = (long)irow1 * i, r2 = (long)irow2 * i;
      if (psize == 3) {
        for (int j = start; j < widthx; j++) dst[o + j] = OR_BL(src2[r2 + j], src1[r1 + j]);
      } else {
        for (int j = start; j < widthx; j += 4) {
          /* with start == 1 (ARGB) byte j+3 is the alpha of the NEXT pixel; past the end of the
           * buffer we take 255 (the reference reads out of bounds there) */
          int a2 = (r2 + j + 3 < src2_bytes) ? src2[r2 + j + 3] : 255;
          if (a2 == 255) {
            for (int k = 0; k < 3; k++) dst[o + j + k] = OR_BL(src2[r2 + j + k], src1[r1 + j + k]);
          } else {
            const float alpha = (float)a2 / 255., inv_alpha = 1. - alpha;
            for (int k = 0; k < 3; k++)
              dst[o + j + k] = OR_BL((uint8_t)((float)src2[r2 + j + k] * alpha), (uint8_t)((float)src1[r1 + j + k] * inv_alpha));
          }
        }
      }
    }
#undef OR_BL
    return;
  }
  /* luma overlay / underlay / negative overlay / averaged luma overlay :150-197.
   * Type 4 ("averaged luma overlay" :153-169): the 3 x 3 luma average runs only `if (j > start && j < width - 1 && row > 0 &&
   * row < height - 1)`; `row` starts at 0 (:73) and its only `row++` (:167) sits INSIDE that guard, so the guard never holds and
   * every pixel falls through into `case 1`: type 4 computes exactly the luma overlay (== the compiled plugin, tests).
   * ARGB (start == 1): calc_luma is handed the pixel pointer + 1, so it weighs G, B and the NEXT pixel's alpha byte
   * (libweed/weed-plugin-utils.c:924-934 reads px[1..3]) -- replicated.  For the last pixel of the buffer that byte
   * lies outside it (the reference reads out of bounds); there we take 255, as for the chroma blend above.  Both
   * inputs are assumed to have src2's geometry. */
  for (int i = 0; i < height; i++) {
    const long o = (long)orow * i, r1 = (long)irow1 * i, r2 = (long)irow2 * i;
    for (int j = start; j < widthx; j += psize) {
      int take2;
      uint8_t p1[4], p2[4];
      memcpy(p1, &src1[r1 + j], 3); memcpy(p2, &src2[r2 + j], 3);
      p1[3] = (psize == 4 && r1 + j + 3 < src2_bytes) ? src1[r1 + j + 3] : 255;
      p2[3] = (psize == 4 && r2 + j + 3 < src2_bytes) ? src2[r2 + j + 3] : 255;
      if (type == 1 || type == 4) take2 = or_calc_luma(p1, palette) < blend_factor;
      else if (type == 2) take2 = or_calc_luma(p2, palette) > blendneg;
      else take2 = or_calc_luma(p1, palette) > blendneg;
      if (take2) memcpy(&dst[o + j], &src2[r2 + j], 3);
      else if (!inplace) memcpy(&dst[o + j], &src1[r1 + j], 3);
    }
  }
}

/* multi_blends.c:26-165; RGB24/BGR24 only (width*3 at :33) */
void pe_or_multi_blend(int type, int palette, const uint8_t *src1, int irow1, const uint8_t *src2, int irow2,
                       uint8_t *dst, int orow, int width, int height, int bf) {
  const uint8_t blend_factor = (uint8_t)bf;
  const uint8_t blend1 = blend_factor * 2, blendneg1 = 255 - blend_factor * 2;
  const uint8_t blend2 = (255 - blend_factor) * 2, blendneg2 = (blend_factor - 128) * 2;
  for (int i = 0; i < height; i++) {
    const uint8_t *s1 = src1 + (long)irow1 * i, *s2 = src2 + (long)irow2 * i;
    uint8_t *d = dst + (long)orow * i;
    for (int j = 0; j < width * 3; j += 3) {
      uint8_t pixel[3];
      int intval, k, mpy = 0, scr = 0;
      uint8_t luma1, luma2;
      switch (type) {
      case 0: mpy = 1; break;
      case 1: scr = 1; break;
      case 2:
        luma1 = or_calc_luma(&s1[j], palette); luma2 = or_calc_luma(&s2[j], palette);
        memcpy(pixel, luma1 <= luma2 ? &s1[j] : &s2[j], 3); break;
      case 3:
        luma1 = or_calc_luma(&s1[j], palette); luma2 = or_calc_luma(&s2[j], palette);
        memcpy(pixel, luma1 >= luma2 ? &s1[j] : &s2[j], 3); break;
      case 4:
        luma1 = or_calc_luma(&s1[j], palette);
        if (luma1 < 128) mpy = 1; else scr = 1;
        break;
      case 5:
        for (k = 0; k < 3; k++) {
          if (s2[j + k] == 255) pixel[k] = 255;
          else { intval = ((int)(s1[j + k]) << 8) / (int)(255 - s2[j + k]); pixel[k] = intval > 255 ? 255 : (uint8_t)intval; }
        }
        break;
      default:
        for (k = 0; k < 3; k++) {
          if (s2[j + k] == 0) pixel[k] = 0;
          else { intval = 255 - (255 - ((int)(s1[j + k]) << 8)) / (int)(s2[j + k]); pixel[k] = intval < 0 ? 0 : (uint8_t)intval; }
        }
        break;
      }
      if (mpy) for (k = 0; k < 3; k++) pixel[k] = (uint8_t)((s2[j + k] * s1[j + k]) >> 8);
      if (scr) for (k = 0; k < 3; k++) pixel[k] = (uint8_t)(255 - (((255 - s2[j + k]) * (255 - s1[j + k])) >> 8));
      if (blend_factor < 128) for (k = 0; k < 3; k++) d[j + k] = (blend1 * pixel[k] + blendneg1 * s1[j + k]) >> 8;
      else for (k = 0; k < 3; k++) d[j + k] = (blend2 * pixel[k] + blendneg2 * s2[j + k]) >> 8;
    }
  }
}

/* gdk/compositor.c paint_pixel :120-125 (double arithmetic, truncation on store) */
void pe_or_alpha_over(uint8_t *dst, int orow, const uint8_t *src, int irow, int palette, int width, int height,
                      double alpha) {
  const int psize = (palette == OR_PAL_RGB24 || palette == OR_PAL_BGR24) ? 3 : 4;
  for (int y = 0; y < height; y++) {
    uint8_t *d = dst + (long)orow * y;
    const uint8_t *s = src + (long)irow * y;
    for (int x = 0; x < width; x++, d += psize, s += psize) {
      double invalpha = 1. - alpha;
      d[0] = d[0] * invalpha + s[0] * alpha;
      d[1] = d[1] * invalpha + s[1] * alpha;
      d[2] = d[2] * invalpha + s[2] * alpha;
    }
  }
}

/* gdk/compositor.c:172-186 */
void pe_or_fill(uint8_t *dst, int orow, int palette, int width, int height, int r, int g, int b) {
  const int psize = (palette == OR_PAL_RGB24 || palette == OR_PAL_BGR24) ? 3 : 4;
  const int swap = (palette == OR_PAL_BGR24 || palette == OR_PAL_BGRA32);
  for (int y = 0; y < height; y++) {
    uint8_t *d = dst + (long)orow * y;
    for (int x = 0; x < width; x++, d += psize) {
      d[0] = swap ? b : r; d[1] = g; d[2] = swap ? r : b;
      if (psize == 4) d[3] = 0xFF;
    }
  }
}

/* ---- resize: OUR contract (reference = libswscale, unavailable) ------------- */
/* Separable, swscale-shaped data path:
 *   per axis a filter bank of `taps` coefficients per output sample, centre aligned:
 *     centre(i) = (i + 0.5) * src_n / dst_n - 0.5
 *     scale >= 1 (upscale or equal): 2 taps, triangle of half-width 1
 *     scale <  1 (downscale): triangle of half-width 1/scale, taps = ceil(2/scale) + 1
 *   weights are computed in double, normalised so they sum to exactly 1 << shift_bits
 *   (residual added to the largest tap), source indices clamped to the edge;
 *   horizontal pass: shift_bits = 14, tmp = min(sum(coef * pix) >> 7, 32767)      (15-bit)
 *   vertical pass:   shift_bits = 12, out = clip_u8((sum(coef * tmp) + (1 << 18)) >> 19)
 * At scale 1 both passes are exact identities. */
int pe_or_resize_filter(int src_n, int dst_n, int shift_bits, int32_t *first, int16_t *coefs, int max_taps) {
  const double ratio = (double)src_n / (double)dst_n; /* source samples per destination sample */
  const double support = ratio > 1. ? ratio : 1.;
  int taps = ratio > 1. ? (int)ceil(2. * ratio) + 1 : 2;
  const int one = 1 << shift_bits;
  if (taps > max_taps) return -1;
  for (int i = 0; i < dst_n; i++) {
    const double centre = ((double)i + 0.5) * ratio - 0.5;
    int left = (int)floor(centre - support) + 1;
    double w[64], sum = 0.;
    int acc = 0, big = 0;
    if (ratio <= 1.) left = (int)floor(centre);
    for (int k = 0; k < taps; k++) {
      double d = fabs((double)(left + k) - centre) / support;
      w[k] = d < 1. ? 1. - d : 0.;
      sum += w[k];
    }
    for (int k = 0; k < taps; k++) {
      int q = (int)floor(w[k] / sum * one + 0.5);
      coefs[(long)i * max_taps + k] = (int16_t)q;
      acc += q;
      if (w[k] > w[big]) big = k;
    }
    coefs[(long)i * max_taps + big] += (int16_t)(one - acc);
    for (int k = taps; k < max_taps; k++) coefs[(long)i * max_taps + k] = 0;
    first[i] = left;
  }
  return taps;
}

/* The bilinear coefficient recipe of libswscale (third-party, not in the reference tree and not pinned by it: configure.ac:562;
 * restated from the published algorithm of libswscale/utils.c initFilter and checked against libswscale 9.1.100 in
 * tests/test_resize_vs_swscale.py): triangle taps at 2^-30 precision around a centre-aligned position, near-zero taps (cumulated
 * weight below 0.002) dropped from either end, taps past the frame folded onto the edge sample, then normalised to 1 << shift_bits
 * with the rounding error carried from tap to tap.  OPT-IN second recipe beside pe_or_resize_filter. */
static int64_t or_rounded_div(int64_t a, int64_t b) { return a >= 0 ? (a + (b >> 1)) / b : (a - (b >> 1)) / b; }

/* kind: 1 SWS_BILINEAR (LIVES_INTERP_NORMAL), 2 SWS_BICUBIC (BEST, shrinking), 3 SWS_LANCZOS (BEST, growing),
 * 4 SWS_FAST_BILINEAR as a filter bank (its vertical pass: two taps whatever the scale factor), 5 SWS_FAST_BILINEAR's horizontal pass
 * (not a bank in libswscale but a 16.16 position walk from the left edge, `(s[xx] << 7) + (s[xx + 1] - s[xx]) * xalpha` with a 7-bit
 * xalpha: the same numbers as two 14-bit taps) -- the flags resize_layer_full picks, src/colourspace.c:14991-14997 */
int pe_or_resize_filter_kind(int kind, int src_n, int dst_n, int shift_bits, int32_t *first, int16_t *coefs, int max_taps) {
  const int64_t xinc = (((int64_t)src_n << 16) + (dst_n >> 1)) / dst_n, one = (int64_t)1 << shift_bits;
  const int size_factor = kind == 2 ? 4 : kind == 3 ? 6 : 2;
  int fs = kind == 4 ? 2 : xinc <= (1 << 16) ? 1 + size_factor : 1 + (size_factor * src_n + dst_n - 1) / dst_n, lg = 0, min_fs = 0;
  int64_t *f, fone, xdst;
  if (kind < 1 || kind > 5 || src_n < 1 || dst_n < 1 || max_taps < 2) return -1;
  if (kind == 5) { /* the position walk */
    uint32_t xpos = 0;
    for (int i = 0; i < dst_n; i++, xpos += (uint32_t)xinc) {
      int xx = (int)(xpos >> 16), xalpha = (int)((xpos & 0xFFFF) >> 9);
      if (xx >= src_n - 1) { xx = src_n - 2; xalpha = 128; } /* "dst[i] = src[srcW - 1] * 128" for the tail */
      if (xx < 0) { xx = 0; xalpha = 0; }                    /* one-sample-wide source */
      first[i] = xx;
      coefs[(long)i * max_taps] = (int16_t)((128 - xalpha) << 7);
      coefs[(long)i * max_taps + 1] = (int16_t)(xalpha << 7);
      for (int j = 2; j < max_taps; j++) coefs[(long)i * max_taps + j] = 0;
    }
    return 2;
  }
  if (kind != 4) {
    if (fs > src_n - 2) fs = src_n - 2;
    if (fs < 1) fs = 1;
  }
  if (fs > 64) return -1;
  for (int r = src_n / dst_n; r > 1; r >>= 1) lg++;
  fone = (int64_t)1 << (54 - (lg < 8 ? lg : 8));
  f = (int64_t *)calloc((size_t)dst_n * fs, sizeof(int64_t));
  if (kind == 4) {
    xdst = (xinc >> 1) - 0x8000; /* ((128 * xinc) >> 8) - ((128 * 0x8000) >> 7) */
    for (int i = 0; i < dst_n; i++, xdst += xinc) {
      int xx = (int)(xdst >> 16); /* arithmetic shift: floor */
      first[i] = xx;
      for (int j = 0; j < 2; j++, xx++) {
        const int64_t c = fone - llabs(((int64_t)xx << 16) - xdst) * (fone >> 16);
        f[(long)i * fs + j] = c < 0 ? 0 : c;
      }
    }
  } else {
    xdst = xinc - 65536; /* ((128 * xinc) >> 7) - ((128 * 65536) >> 7): both grids sampled at pixel centres */
    for (int i = 0; i < dst_n; i++, xdst += 2 * xinc) {
      int xx = (int)((xdst - (int64_t)(fs - 2) * 65536) / (1 << 17)); /* C division: towards zero */
      first[i] = xx;
      for (int j = 0; j < fs; j++, xx++) {
        int64_t d = llabs((int64_t)xx * (1 << 17) - xdst) << 13, c;
        if (xinc > (1 << 16)) d = d * dst_n / src_n;
        if (kind == 2) { /* Mitchell-Netravali family with libswscale's defaults B = 0, C = 0.6, in 24-bit fixed point */
          const int64_t B = 0, Cc = (int64_t)(0.6 * (1 << 24));
          if (d >= (int64_t)1 << 31) c = 0;
          else {
            const int64_t dd = (d * d) >> 30, ddd = (dd * d) >> 30;
            if (d < (int64_t)1 << 30)
              c = (12 * (1 << 24) - 9 * B - 6 * Cc) * ddd + (-18 * (1 << 24) + 12 * B + 6 * Cc) * dd + (6 * (1 << 24) - 2 * B) * ((int64_t)1 << 30);
            else
              c = (-B - 6 * Cc) * ddd + (6 * B + 30 * Cc) * dd + (-12 * B - 48 * Cc) * d + (8 * B + 24 * Cc) * ((int64_t)1 << 30);
          }
          c /= ((int64_t)1 << 54) / fone;
        } else if (kind == 3) { /* Lanczos, 3 lobes, in double */
          const double fd = (double)d * (1.0 / (1 << 30)), p = 3.0;
          double v = fd == 0.0 ? 1.0 : sin(fd * M_PI) * sin(fd * M_PI / p) / (fd * fd * M_PI * M_PI / p);
          if (fd > p) v = 0;
          c = (int64_t)(v * (double)fone);
        } else {
          c = ((int64_t)1 << 30) - d;
          c = c < 0 ? 0 : c * (fone >> 30);
        }
        f[(long)i * fs + j] = c;
      }
    }
  }
  for (int i = dst_n - 1; i >= 0; i--) { /* shrink: drop near-zero taps on the left (shifting), count them on the right */
    int64_t *r = f + (long)i * fs, cut = 0;
    int mn = fs;
    for (int j = 0; j < fs; j++) {
      cut += llabs(r[0]);
      if ((double)cut > 0.002 * (double)fone) break;
      if (i < dst_n - 1 && first[i] >= first[i + 1]) break; /* positions stay monotonic */
      memmove(r, r + 1, sizeof(int64_t) * (fs - 1));
      r[fs - 1] = 0;
      first[i]++;
    }
    cut = 0;
    for (int j = fs - 1; j > 0; j--) {
      cut += llabs(r[j]);
      if ((double)cut > 0.002 * (double)fone) break;
      mn--;
    }
    if (mn > min_fs) min_fs = mn;
  }
  if (min_fs > max_taps) { free(f); return -1; }
  for (int i = 0; i < dst_n; i++) {
    int64_t t[64], sum = 0, err = 0;
    for (int j = 0; j < min_fs; j++) t[j] = f[(long)i * fs + j];
    if (first[i] < 0) { /* taps left of the frame land on sample 0 */
      for (int j = 1; j < min_fs; j++) {
        const int left = j + first[i] > 0 ? j + first[i] : 0;
        t[left] += t[j];
        t[j] = 0;
      }
      first[i] = 0;
    }
    if (first[i] + min_fs > src_n) { /* taps right of the frame land on the last sample, the window moves back inside */
      const int shift = first[i] + (min_fs - src_n < 0 ? min_fs - src_n : 0);
      int64_t acc = 0;
      for (int j = min_fs - 1; j >= 0; j--)
        if (first[i] + j >= src_n) { acc += t[j]; t[j] = 0; }
      for (int j = min_fs - 1; j >= 0; j--) t[j] = j < shift ? 0 : t[j - shift];
      first[i] -= shift;
      t[src_n - 1 - first[i]] += acc;
    }
    for (int j = 0; j < min_fs; j++) sum += t[j];
    sum = (sum + one / 2) / one;
    if (!sum) sum = 1;
    for (int j = 0; j < min_fs; j++) {
      const int64_t v = t[j] + err, q = or_rounded_div(v, sum);
      coefs[(long)i * max_taps + j] = (int16_t)q;
      err = v - q * sum;
    }
    for (int j = min_fs; j < max_taps; j++) coefs[(long)i * max_taps + j] = 0;
  }
  free(f);
  return min_fs;
}

int pe_or_resize_filter_sws(int src_n, int dst_n, int shift_bits, int32_t *first, int16_t *coefs, int max_taps) {
  return pe_or_resize_filter_kind(1, src_n, dst_n, shift_bits, first, coefs, max_taps);
}

static int or_resize_recipe = 1; /* 1: libswscale's recipes (default, what the product ships), 0: the round-1 triangle contract */
void pe_or_set_resize_recipe(int recipe) { or_resize_recipe = recipe; }
/* which bank an axis takes: interp 0 FAST, 1 NORMAL, 2 BEST (LiVESInterpType); `up` = the frame grows in either direction (:14991) */
static int or_resize_filter(int interp, int up, int horizontal, int src_n, int dst_n, int shift_bits, int32_t *first, int16_t *coefs,
                            int max_taps) {
  int kind = 1;
  if (!or_resize_recipe) return pe_or_resize_filter(src_n, dst_n, shift_bits, first, coefs, max_taps);
  if (interp == 2) kind = up ? 3 : 2;
  else if (interp == 0) kind = horizontal ? 5 : 4;
  if (src_n == dst_n) kind = 1; /* an unscaled axis is the identity under every flag */
  return pe_or_resize_filter_kind(kind, src_n, dst_n, shift_bits, first, coefs, max_taps);
}

void pe_or_resize_packed_interp(const uint8_t *src, int irow, int sw, int sh, uint8_t *dst, int orow, int dw, int dh, int psize,
                                int interp);
void pe_or_resize_packed(const uint8_t *src, int irow, int sw, int sh, uint8_t *dst, int orow, int dw, int dh,
                         int psize) {
  pe_or_resize_packed_interp(src, irow, sw, sh, dst, orow, dw, dh, psize, 1);
}

void pe_or_resize_packed_interp(const uint8_t *src, int irow, int sw, int sh, uint8_t *dst, int orow, int dw, int dh, int psize,
                                int interp) {
  enum { MT = 64 };
  const int up = dw > sw || dh > sh;
  int32_t *fx = (int32_t *)malloc(sizeof(int32_t) * dw), *fy = (int32_t *)malloc(sizeof(int32_t) * dh);
  int16_t *cx = (int16_t *)malloc(sizeof(int16_t) * MT * dw), *cy = (int16_t *)malloc(sizeof(int16_t) * MT * dh);
  int tx = or_resize_filter(interp, up, 1, sw, dw, 14, fx, cx, MT), ty = or_resize_filter(interp, up, 0, sh, dh, 12, fy, cy, MT);
  int16_t *tmp = (int16_t *)malloc(sizeof(int16_t) * (size_t)sh * dw * psize);
  if (tx > 0 && ty > 0) {
    for (int y = 0; y < sh; y++) {
      const uint8_t *s = src + (long)irow * y;
      for (int x = 0; x < dw; x++)
        for (int ch = 0; ch < psize; ch++) {
          int acc = 0;
          for (int k = 0; k < tx; k++) {
            int sx = fx[x] + k; sx = sx < 0 ? 0 : sx >= sw ? sw - 1 : sx;
            acc += cx[(long)x * MT + k] * s[sx * psize + ch];
          }
          acc >>= 7;
          tmp[((long)y * dw + x) * psize + ch] = (int16_t)(acc > 32767 ? 32767 : acc);
        }
    }
    for (int y = 0; y < dh; y++) {
      uint8_t *d = dst + (long)orow * y;
      for (int x = 0; x < dw * psize; x++) {
        int acc = 1 << 18;
        for (int k = 0; k < ty; k++) {
          int sy = fy[y] + k; sy = sy < 0 ? 0 : sy >= sh ? sh - 1 : sy;
          acc += cy[(long)y * MT + k] * tmp[(long)sy * dw * psize + x];
        }
        acc >>= 19;
        d[x] = (uint8_t)(acc < 0 ? 0 : acc > 255 ? 255 : acc);
      }
    }
  }
  free(fx); free(fy); free(cx); free(cy); free(tmp);
}

/* letterbox_layer src/colourspace.c:15343: offsets ((outer - inner + 1) >> 1) (:15522-15523),
 * border = black with opaque alpha (blank_pixel via :15489) */
void pe_or_letterbox_packed(const uint8_t *inner, int irow, int iw, int ih, uint8_t *outer, int orow, int ow, int oh,
                            int palette) {
  int ro, go, bo, ao, ps;
  if (or_rgb_layout(palette, &ro, &go, &bo, &ao, &ps)) return;
  {
    const int ox = (ow - iw + 1) >> 1, oy = (oh - ih + 1) >> 1;
    for (int y = 0; y < oh; y++) {
      uint8_t *d = outer + (long)orow * y;
      for (int x = 0; x < ow; x++, d += ps) { d[ro] = d[go] = d[bo] = 0; if (ao >= 0) d[ao] = 255; }
    }
    for (int y = 0; y < ih; y++)
      memcpy(outer + (long)orow * (y + oy) + (long)ox * ps, inner + (long)irow * y, (size_t)iw * ps);
  }
}

/* ==== YUV <-> YUV family ======================================================================================= */

/* convert_yuv_planar_to_rgb_frame src/colourspace.c:7200-7290 (bgr :7304, argb :7405): yuv2rgb per sample, the YCbCr tables of
 * the clamping (set_conversion_arrays(clamping, WEED_YUV_SUBSPACE_YCBCR)), alpha from the 4th plane or 255.
 * X: the BGR variant starts with opstep = 4 (:7313) and only ever sets it to 4, so BGR24 output walks 4 bytes per pixel and runs
 * past its rows / buffer -- restated with the palette's own 3-byte step.  X: the ARGB variant subtracts the width from the output
 * stride twice and never from the input stride (:7471-7472) and never selects its tables -- restated as the RGB variant with
 * the alpha byte first. */
void pe_or_yuv444p_to_rgb(const uint8_t *const src[4], int irow, int width, int height, uint8_t *dest, int orow, int order,
                          int in_alpha, int out_alpha, int clamping, int quality) {
  const or_conv_t *c = or_conv(clamping, OR_SUBSPACE_YCBCR);
  int ro, go, bo, ao, ps;
  or_order_offsets(order, out_alpha, &ro, &go, &bo, &ao, &ps);
  for (int i = 0; i < height; i++) {
    uint8_t *d = dest + (long)orow * i;
    const long o = (long)irow * i;
    for (int j = 0; j < width; j++, d += ps) {
      or_px_yuv2rgb(c, quality, NULL, src[0][o + j], src[1][o + j], src[2][o + j], &d[ro], &d[go], &d[bo]);
      if (ao >= 0) d[ao] = in_alpha ? src[3][o + j] : 255;
    }
  }
}

/* convert_combineplanes_frame :7593-7640.  X: on padded planes the reference's alpha pointer is never advanced by the row padding
 * (:7625-7637); restated with the stride on all four planes. */
void pe_or_combine_planes(const uint8_t *const src[4], int irow, int width, int height, uint8_t *dest, int orow, int in_alpha,
                          int out_alpha) {
  const int ops = out_alpha ? 4 : 3;
  for (int k = 0; k < height; k++) {
    uint8_t *d = dest + (long)orow * k;
    const long o = (long)irow * k;
    for (int x = 0; x < width; x++, d += ops) {
      d[0] = src[0][o + x]; d[1] = src[1][o + x]; d[2] = src[2][o + x];
      if (out_alpha) d[3] = in_alpha ? src[3][o + x] : 255;
    }
  }
}

/* convert_splitplanes_frame :9198-9252.  X: with a destination alpha plane the reference advances that plane by
 * (stride - width * ipsize) per row (:9233-9236) and writes before its buffer; with a source alpha but no destination alpha the
 * 4th source byte is never skipped (:9238-9246).  Restated as the evident intent (equal to the reference for 3 -> 3 planes). */
void pe_or_split_planes(const uint8_t *src, int irow, int width, int height, uint8_t *const dest[4], const int orows[4],
                        int src_alpha, int dest_alpha) {
  const int ips = src_alpha ? 4 : 3;
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    for (int j = 0; j < width; j++, s += ips) {
      dest[0][(long)orows[0] * i + j] = s[0];
      dest[1][(long)orows[1] * i + j] = s[1];
      dest[2][(long)orows[2] * i + j] = s[2];
      if (dest_alpha) dest[3][(long)orows[3] * i + j] = src_alpha ? s[3] : 255;
    }
  }
}

static const uint8_t *or_avg(int clamping) { /* cavg = cavgc / cavgu, set_conversion_arrays :220-300 */
  static uint8_t avg[2][65536];
  static int ok = 0;
  if (!ok) { pe_or_avg_table(0, avg[0]); pe_or_avg_table(1, avg[1]); ok = 1; }
  return avg[clamping == OR_CLAMPED ? 0 : 1];
}

/* convert_halve_chroma :10578-10609: source chroma rows (2k, 2k+1) -> row k = avg_chroma(row 2k, row 2k+1) (the copy of the
 * even row is the first operand, i.e. the table ROW); a trailing unpaired row is copied */
void pe_or_halve_chroma(const uint8_t *const src[3], const int istrides[3], int cwidth, int cheight, uint8_t *const dest[3],
                        const int ostrides[3], int clamping) {
  const uint8_t *avg = or_avg(clamping);
  for (int p = 1; p <= 2; p++)
    for (int i = 0; i < cheight; i++) {
      const uint8_t *s = src[p] + (long)istrides[p] * i;
      uint8_t *d = dest[p] + (long)ostrides[p] * (i >> 1);
      for (int j = 0; j < cwidth; j++) d[j] = (i & 1) ? avg[(d[j] << 8) + s[j]] : s[j];
    }
}

/* convert_double_chroma :10612-10639: output rows 2k and 2k+1 are copies of source row k; then, when the copy of row k (k > 0)
 * lands in output row 2k, output row 2k-1 becomes avg_chroma(row 2k-1, row 2k) = avg(src k-1, src k) */
void pe_or_double_chroma(const uint8_t *const src[3], const int istrides[3], int cwidth, int cheight, uint8_t *const dest[3],
                         const int ostrides[3], int clamping) {
  const uint8_t *avg = or_avg(clamping);
  for (int p = 1; p <= 2; p++)
    for (int i = 0; i < 2 * cheight; i++) {
      const uint8_t *s = src[p] + (long)istrides[p] * (i >> 1);
      uint8_t *d = dest[p] + (long)ostrides[p] * i;
      memcpy(d, s, (size_t)cwidth);
      if (!(i & 1) && i > 0) {
        uint8_t *pr = d - ostrides[p];
        for (int j = 0; j < cwidth; j++) pr[j] = avg[(pr[j] << 8) + d[j]];
      }
    }
}

static inline void or_mpx(int fmt, const uint8_t *m, uint8_t *y0, uint8_t *u, uint8_t *y1, uint8_t *v) {
  if (fmt == 0) { *u = m[0]; *y0 = m[1]; *v = m[2]; *y1 = m[3]; }  /* uyvy_macropixel colourspace.h:222-227 */
  else { *y0 = m[0]; *u = m[1]; *y1 = m[2]; *v = m[3]; }            /* yuyv_macropixel :229-234 */
}

/* convert_{uyvy,yuyv}_to_yuv422_frame :8093-8127.  The reference walks source and planes densely (only right for unpadded rows;
 * strides honoured here, equal there) and NEVER advances its source pointer (:8103,:8121): with quirks every sample of the frame
 * is the first macropixel's; quirks == 0 is the evident intent. */
void pe_or_packed422_to_yuv422p(int fmt, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *const dest[3],
                                const int orows[3], int quirks) {
  for (int k = 0; k < height; k++)
    for (int x = 0; x < width_mpx; x++) {
      const uint8_t *m = quirks ? src : src + (long)irow * k + 4L * x;
      uint8_t y0, u, y1, v;
      or_mpx(fmt, m, &y0, &u, &y1, &v);
      dest[0][(long)orows[0] * k + 2 * x] = y0; dest[0][(long)orows[0] * k + 2 * x + 1] = y1;
      dest[1][(long)orows[1] * k + x] = u; dest[2][(long)orows[2] * k + x] = v;
    }
}

/* convert_{uyvy,yuyv}_to_yuvp_frame :7800-7842: chroma duplicated onto both pixels, no interpolation ("TODO - avg_chroma").
 * The reference indexes the planes with mixed strides (y with orow[1] / orow[0], u with orow[0]); all planes of a 4:4:4 frame
 * share one stride, where this restatement equals it. */
void pe_or_packed422_to_yuv444p(int fmt, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *const dest[4],
                                const int orows[4], int add_alpha) {
  for (int k = 0; k < height; k++)
    for (int x = 0; x < width_mpx; x++) {
      uint8_t y0, u, y1, v;
      or_mpx(fmt, src + (long)irow * k + 4L * x, &y0, &u, &y1, &v);
      dest[0][(long)orows[0] * k + 2 * x] = y0; dest[0][(long)orows[0] * k + 2 * x + 1] = y1;
      dest[1][(long)orows[1] * k + 2 * x] = dest[1][(long)orows[1] * k + 2 * x + 1] = u;
      dest[2][(long)orows[2] * k + 2 * x] = dest[2][(long)orows[2] * k + 2 * x + 1] = v;
    }
  if (add_alpha) memset(dest[3], 255, (size_t)orows[3] * height);
}

/* convert_{uyvy,yuyv}_to_yuv888_frame :7845-7885 */
void pe_or_packed422_to_yuv888(int fmt, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *dest, int orow,
                               int add_alpha) {
  const int ps = add_alpha ? 4 : 3;
  for (int k = 0; k < height; k++) {
    uint8_t *d = dest + (long)orow * k;
    for (int x = 0; x < width_mpx; x++, d += 2 * ps) {
      uint8_t y0, u, y1, v;
      or_mpx(fmt, src + (long)irow * k + 4L * x, &y0, &u, &y1, &v);
      d[0] = y0; d[1] = u; d[2] = v; d[ps] = y1; d[ps + 1] = u; d[ps + 2] = v;
      if (add_alpha) d[3] = d[ps + 3] = 255;
    }
  }
}

/* convert_yuv_planar_to_{uyvy,yuyv}_frame :7500-7590: one macropixel per pixel pair, chroma = avg_chroma(c[2x], c[2x+1]) (first
 * sample = table row).  Only the reference's dense branch (irowstride == width, orowstride == 2 * width) is right: its strided
 * branch runs `width` macropixels per row (:7527, :7577) -- X; restated per pair with the strides. */
void pe_or_yuv444p_to_packed422(int fmt, const uint8_t *const src[3], int irow, int width, int height, uint8_t *dest, int orow,
                                int clamping) {
  const uint8_t *avg = or_avg(clamping);
  for (int k = 0; k < height; k++) {
    const uint8_t *y = src[0] + (long)irow * k, *u = src[1] + (long)irow * k, *v = src[2] + (long)irow * k;
    uint8_t *d = dest + (long)orow * k;
    for (int x = 0; x < width >> 1; x++, d += 4) {
      const uint8_t cu = avg[(u[2 * x] << 8) + u[2 * x + 1]], cv = avg[(v[2 * x] << 8) + v[2 * x + 1]];
      if (fmt == 0) { d[0] = cu; d[1] = y[2 * x]; d[2] = cv; d[3] = y[2 * x + 1]; }
      else { d[0] = y[2 * x]; d[1] = cu; d[2] = y[2 * x + 1]; d[3] = cv; }
    }
  }
}

/* convert_yuvp_to_yuv420_frame :7690-7752: luma copied; chroma row k = avg_chroma(h(2k), h(2k+1)) with h(r)[j] =
 * avg_chroma(c[r][2j], c[r][2j+1]); a trailing unpaired row leaves h(r) */
void pe_or_yuv444p_to_yuv420p(const uint8_t *const src[3], const int irows[3], int width, int height, uint8_t *const dest[3],
                              const int orows[3], int clamping) {
  const uint8_t *avg = or_avg(clamping);
  for (int i = 0; i < height; i++) memcpy(dest[0] + (long)orows[0] * i, src[0] + (long)irows[0] * i, (size_t)width);
  for (int p = 1; p <= 2; p++)
    for (int i = 0; i < height; i++) {
      const uint8_t *s = src[p] + (long)irows[p] * i;
      uint8_t *d = dest[p] + (long)orows[p] * (i >> 1);
      for (int j = 0; j < width >> 1; j++) {
        const uint8_t x = avg[(s[2 * j] << 8) + s[2 * j + 1]];
        d[j] = (i & 1) ? avg[(d[j] << 8) + x] : x;
      }
    }
}

/* convert_yuv420_to_{uyvy,yuyv}_frame :7104-7198: rows 2k and 2k+1 take chroma row k as it is -- the averaging of :7132-7135 is
 * guarded by `i > 0` and i is never incremented, so it never runs (R).  The chroma pointers are rewound by the full rowstride
 * (:7143-7144), which is only right for unpadded chroma planes, and the YUYV variant never skips the luma row padding (:7181-7195)
 * (X: strides honoured here, equal there).
 * convert_yuv422p_to_{uyvy,yuyv}_frame :6442-6494: plain interleave; the reference's row advance is wrong for every argument
 * (`irows[0] -= width` for 2 * width luma bytes, and the dispatcher passes the pixel width as the macropixel count, :13691), so only
 * its first row is defined (X beyond). */
void pe_or_yuv42xp_to_packed422(int fmt, const uint8_t *const src[3], const int irows[3], int width, int height, int is_422,
                                uint8_t *dest, int orow) {
  for (int k = 0; k < height; k++) {
    const int cr = is_422 ? k : k >> 1;
    const uint8_t *y = src[0] + (long)irows[0] * k, *u = src[1] + (long)irows[1] * cr, *v = src[2] + (long)irows[2] * cr;
    uint8_t *d = dest + (long)orow * k;
    for (int x = 0; x < width >> 1; x++, d += 4) {
      if (fmt == 0) { d[0] = u[x]; d[1] = y[2 * x]; d[2] = v[x]; d[3] = y[2 * x + 1]; }
      else { d[0] = y[2 * x]; d[1] = u[x]; d[2] = y[2 * x + 1]; d[3] = v[x]; }
    }
  }
}

/* convert_quad_chroma :10642-10712.  Even destination row 2k, from chroma row k (s): column 0 = s[0]; column 2m (m > 0) =
 * f(s[m-1], s[m]); column 2m+1 = g(s[m], s[m+1]) -- s[cw] is the byte behind the row (padding / next row's first sample; on the
 * last row of an unpadded plane: the edge sample, X) -- with, for JPEG sampling, f = g = avg_chroma; otherwise U: f = avg_3_1
 * (= avg(x, avg(x, y))), g = avg_1_3 (= avg(avg(x, y), y)) and V the other way round (:10669-10685).  Odd row r is written two
 * rows later as avg_chroma(row r+1, row r-1) (:10688-10694: the LOWER row is the table row); for an odd height the last odd row is
 * fixed up after the loop with the operands the other way round (:10707-10708).
 * X: for an EVEN height the last row is never written by the reference (left as allocated) -- defined here as a copy of the row
 * above; the odd-row pass writes one sample past `width` (:10692-10694), harmless on padded planes, not restated; the
 * dispatcher passes add_alpha = TRUE for YUV444P, whose 4th plane pointer is not valid (:13606). */
void pe_or_quad_chroma(const uint8_t *const src[3], const int istrides[3], int width, int height, uint8_t *const dest[4], int ostride,
                       int add_alpha, int sampling_jpeg, int clamping) {
  const uint8_t *avg = or_avg(clamping);
#define OR_AV(x_, y_) avg[((int)(x_) << 8) + (int)(y_)]
  const int w2 = (width >> 1) << 1, cw = w2 >> 1, ch = (height + 1) >> 1;
  for (int p = 1; p <= 2; p++) {
    for (int i = 0; i < height; i += 2) {
      const int k = i >> 1;
      const uint8_t *s = src[p] + (long)istrides[p] * k;
      uint8_t *d = dest[p] + (long)ostride * i;
      for (int m = 0; m < cw; m++) {
        const uint8_t a = s[m], nx = (m + 1 < cw || cw < istrides[p] || k + 1 < ch) ? s[m + 1] : s[m];
        if (m == 0) d[0] = a;
        else {
          const uint8_t pv = s[m - 1];
          d[2 * m] = sampling_jpeg ? OR_AV(pv, a) : (p == 1 ? OR_AV(pv, OR_AV(pv, a)) : OR_AV(OR_AV(pv, a), a));
        }
        d[2 * m + 1] = sampling_jpeg ? OR_AV(a, nx) : (p == 1 ? OR_AV(OR_AV(a, nx), nx) : OR_AV(a, OR_AV(a, nx)));
      }
    }
    for (int r = 1; r < height; r += 2) {
      uint8_t *d = dest[p] + (long)ostride * r;
      const uint8_t *up = d - ostride, *dn = d + ostride;
      if (r + 1 > height - 1) memcpy(d, up, (size_t)w2);                                         /* X */
      else if ((height & 1) && r == height - 2) for (int j = 0; j < w2; j++) d[j] = OR_AV(up[j], dn[j]);  /* post-loop fix-up */
      else for (int j = 0; j < w2; j++) d[j] = OR_AV(dn[j], up[j]);
    }
  }
#undef OR_AV
  if (add_alpha) memset(dest[3], 255, (size_t)ostride * height);
}

/* convert_yuv888_to_{uyvy,yuyv}_frame :8184-8270, convert_yuv888_to_yuv422_frame :8129-8182, convert_yuv888_to_yuv420_frame
 * :8035-8090.  Per pixel pair: both lumas, chroma = avg_chroma(c of the first pixel, c of the second); 4:2:0 row k =
 * avg_chroma(pair averages of row 2k, pair averages of row 2k+1) (a trailing unpaired row leaves its pair averages).
 * X: the strided branches of the packed targets advance the macropixel pointer by a BYTE count (:8215), the strided 4:2:2 branch
 * subtracts a quarter width from the chroma strides (:8163-8164), the 4:2:0 loop walks luma densely and rewinds chroma by the full
 * rowstride (:8046,:8084) -- all equal to this restatement on unpadded buffers, where they are compared. */
void pe_or_yuv888_subsample(int mode, const uint8_t *src, int irow, int width, int height, int src_alpha, uint8_t *const dest[3],
                            const int orows[3], int clamping) {
  const uint8_t *avg = or_avg(clamping);
  const int ips = src_alpha ? 4 : 3, hw = width >> 1;
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    for (int j = 0; j < hw; j++, s += 2 * ips) {
      const uint8_t y0 = s[0], y1 = s[ips], cu = avg[(s[1] << 8) + s[1 + ips]], cv = avg[(s[2] << 8) + s[2 + ips]];
      if (mode <= 1) {
        uint8_t *d = dest[0] + (long)orows[0] * i + 4L * j;
        if (mode == 0) { d[0] = cu; d[1] = y0; d[2] = cv; d[3] = y1; }
        else { d[0] = y0; d[1] = cu; d[2] = y1; d[3] = cv; }
        continue;
      }
      dest[0][(long)orows[0] * i + 2 * j] = y0; dest[0][(long)orows[0] * i + 2 * j + 1] = y1;
      if (mode == 2) { dest[1][(long)orows[1] * i + j] = cu; dest[2][(long)orows[2] * i + j] = cv; }
      else {
        uint8_t *du = dest[1] + (long)orows[1] * (i >> 1) + j, *dv = dest[2] + (long)orows[2] * (i >> 1) + j;
        *du = (i & 1) ? avg[(*du << 8) + cu] : cu;
        *dv = (i & 1) ? avg[(*dv << 8) + cv] : cv;
      }
    }
  }
}

/* convert_{uyvy,yuyv}_to_yuv420_frame :7887-7970: even source rows write their chroma, odd rows fold theirs in with
 * avg_chroma(stored, new) (the EVEN row is the table row); a trailing unpaired row leaves its own chroma.  The reference walks
 * source and planes densely (X: strides honoured here, equal on unpadded buffers). */
void pe_or_packed422_to_yuv420p(int fmt, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *const dest[3],
                                const int orows[3], int clamping) {
  const uint8_t *avg = or_avg(clamping);
  for (int k = 0; k < height; k++)
    for (int x = 0; x < width_mpx; x++) {
      uint8_t y0, u, y1, v;
      or_mpx(fmt, src + (long)irow * k + 4L * x, &y0, &u, &y1, &v);
      dest[0][(long)orows[0] * k + 2 * x] = y0; dest[0][(long)orows[0] * k + 2 * x + 1] = y1;
      uint8_t *du = dest[1] + (long)orows[1] * (k >> 1) + x, *dv = dest[2] + (long)orows[2] * (k >> 1) + x;
      *du = (k & 1) ? avg[(*du << 8) + u] : u;
      *dv = (k & 1) ? avg[(*dv << 8) + v] : v;
    }
}

/* convert_quad_chroma_packed :10715-10808 (4:2:0) and convert_double_chroma_packed :10811-10873 (4:2:2) -> packed Y U V (A = 255).
 * 4:2:2: row i from chroma row i; column 0 = s[0], every other column c = f(s[(c-1)/2 .. ]) -- precisely: column 2m (m > 0) =
 * f(s[m-1], s[m]) and column 2m+1 = f(s[m], s[m+1]) with the SAME f on both (JPEG: avg_chroma; else U: avg_3_1, V: avg_1_3;
 * :10843-10862), s[cw] being the byte behind the row.
 * 4:2:0: even rows exactly as convert_quad_chroma (f on even columns, g on odd ones), odd row r = avg_chroma(row r+1, row r-1)
 * written two rows later (:10771-10780).
 * X (4:2:2): with add_alpha the alpha byte of the second pixel of every pair is never written (:10848-10863) -- 255 here.
 * X (4:2:0): with add_alpha the odd rows get no alpha byte (:10768-10784) -- 255 here;
 *   the last row of an even-height frame never gets its chroma, and for an odd height the post-loop fix-up addresses the
 * wrong rows (`jj = j - ostride`, :10801-10805: it reads one row past the frame and overwrites the last even row, the odd row
 * above stays unwritten) -- defined here as in convert_quad_chroma: last row of an even height = the row above, last odd row of an
 * odd height = avg_chroma(row above, row below). */
void pe_or_chroma_upsample_packed(int is_420, const uint8_t *const src[3], const int istrides[3], int width, int height, uint8_t *dest,
                                  int orow, int add_alpha, int sampling_jpeg, int clamping) {
  const uint8_t *avg = or_avg(clamping);
#define OR_AV(x_, y_) avg[((int)(x_) << 8) + (int)(y_)]
  const int ps = add_alpha ? 4 : 3, w2 = (width >> 1) << 1, cw = w2 >> 1, ch = is_420 ? (height + 1) >> 1 : height;
  for (int i = 0; i < height; i++) {
    uint8_t *d = dest + (long)orow * i;
    const uint8_t *y = src[0] + (long)istrides[0] * i;
    for (int j = 0; j < w2; j++) { d[j * ps] = y[j]; if (add_alpha) d[j * ps + 3] = 255; }
    if (is_420 && (i & 1)) continue;
    const int k = is_420 ? i >> 1 : i;
    for (int p = 1; p <= 2; p++) {
      const uint8_t *s = src[p] + (long)istrides[p] * k;
      for (int m = 0; m < cw; m++) {
        const uint8_t a = s[m], nx = (m + 1 < cw || cw < istrides[p] || k + 1 < ch) ? s[m + 1] : s[m];
        uint8_t e, o2;
        if (m == 0) e = a;
        else {
          const uint8_t pv = s[m - 1];
          e = sampling_jpeg ? OR_AV(pv, a) : (p == 1 ? OR_AV(pv, OR_AV(pv, a)) : OR_AV(OR_AV(pv, a), a));
        }
        if (is_420) o2 = sampling_jpeg ? OR_AV(a, nx) : (p == 1 ? OR_AV(OR_AV(a, nx), nx) : OR_AV(a, OR_AV(a, nx)));
        else o2 = sampling_jpeg ? OR_AV(a, nx) : (p == 1 ? OR_AV(a, OR_AV(a, nx)) : OR_AV(OR_AV(a, nx), nx));
        d[(2 * m) * ps + p] = e;
        d[(2 * m + 1) * ps + p] = o2;
      }
    }
  }
  if (is_420)
    for (int r = 1; r < height; r += 2) {
      uint8_t *d = dest + (long)orow * r;
      const uint8_t *up = d - orow, *dn = d + orow;
      for (int j = 0; j < w2; j++)
        for (int p = 1; p <= 2; p++) {
          const int q = j * ps + p;
          if (r + 1 > height - 1) d[q] = up[q];
          else if ((height & 1) && r == height - 2) d[q] = OR_AV(up[q], dn[q]);
          else d[q] = OR_AV(dn[q], up[q]);
        }
    }
#undef OR_AV
}

/* convert_swab_frame :10517-10566: swab() of width * 4 bytes per row */
void pe_or_swab(uint8_t *pixels, int irow, int width_mpx, int height) {
  for (int k = 0; k < height; k++) {
    uint8_t *r = pixels + (long)irow * k;
    for (int x = 0; x < width_mpx * 2; x++) { const uint8_t t = r[2 * x]; r[2 * x] = r[2 * x + 1]; r[2 * x + 1] = t; }
  }
}

/* init_YUV_to_YUV_tables :1108-1138 (YUV_CLAMP_MIN 16, Y_CLAMP_MAX 235, UV_CLAMP_MAX 240, colourspace.h:16-18; note `<=` in
 * the first Y loop and `<` in the first UV loop) */
void pe_or_yy_table(int which, uint8_t out[256]) {
  int i;
  switch (which) {
  case 0:
    for (i = 0; i <= 16; i++) out[i] = 0;
    for (; i < 235; i++) out[i] = (uint8_t)or_myround((i - 16.) * 255. / (235. - 16.));
    for (; i < 256; i++) out[i] = 255;
    break;
  case 1:
    for (i = 0; i < 16; i++) out[i] = 0;
    for (; i < 240; i++) out[i] = (uint8_t)or_myround((i - 16.) * 255. / (240. - 16.));
    for (; i < 256; i++) out[i] = 255;
    break;
  case 2:
    for (i = 0; i < 256; i++) out[i] = (uint8_t)or_myround((i / 255.) * (235. - 16.) + 16.);
    break;
  default:
    for (i = 0; i < 256; i++) out[i] = (uint8_t)or_myround((i / 255.) * (240. - 16.) + 16.);
    break;
  }
}

/* switch_yuv_clamping_and_subspace :10929-11100: every byte of the plane through Y_to_Y / U_to_U (= V_to_V), walked densely
 * from the plane start for height * rowstride bytes -- row padding included */
void pe_or_switch_clamping_plane(uint8_t *plane, long nbytes, int kind, int to_unclamped) {
  uint8_t ty[256], tc[256];
  pe_or_yy_table(to_unclamped ? 0 : 2, ty);
  pe_or_yy_table(to_unclamped ? 1 : 3, tc);
  for (long i = 0; i < nbytes; i++) {
    int luma;
    switch (kind) {
    case 0: luma = 1; break;
    case 1: luma = 0; break;
    case 2: luma = (i % 3) == 0; break;
    case 3: if ((i & 3) == 3) continue; luma = (i & 3) == 0; break;
    case 4: luma = (i & 1) == 1; break;
    default: luma = (i & 1) == 0; break;
    }
    plane[i] = luma ? ty[plane[i]] : tc[plane[i]];
  }
}

/* slide_over.c:94,109,124,135: the dividing line, `(float)dim * (1. - transval / 255.)` resp. `(float)dim * (transval / 255.)`,
 * truncated -- AS THE REFERENCE BUILDS IT: the weed plugins are compiled with -ffast-math (lives-plugins/weed-plugins/Makefile.am:49),
 * under which gcc turns the division into a multiplication by the rounded reciprocal of 255 and regroups the second form as
 * (dim * (1 / 255.)) * transval (read off the compiled plugin; pinned against it for all 256 values in
 * tests/test_oracle_vs_reference.py).  E.g. a 37 pixel wide frame keeps one column of the old clip at transval 255. */
int pe_or_slide_over_bound(int direction, int transval, int width, int height) {
  const double r255 = 1. / 255.;
  const double dim = (double)(float)(direction <= 2 ? width : height);
  switch (direction) {
  case 1: case 3: return (int)((1. - (double)transval * r255) * dim);
  case 2: case 4: return (int)((dim * r255) * (double)transval);
  }
  return 0;
}

/* slide_over.c sover_process :55-145, restated per destination byte.  Directions 1 and 3 show in1 before the line and in2
 * behind it, 2 and 4 the other way round; a clip that "moves" is read shifted so that its far edge sits on the line
 * (:95-97, :110-112, :126-127, :137-138), otherwise it is read in place */
void pe_or_slide_over(int direction, int transval, int mvlower, int mvupper, const uint8_t *src1, int irow1, const uint8_t *src2,
                      int irow2, uint8_t *dest, int orow, int width, int height, int psize) {
  const int bound = pe_or_slide_over_bound(direction, transval, width, height);
  const int along_y = direction >= 3, swapped = direction == 2 || direction == 4;
  const uint8_t *first = swapped ? src2 : src1, *second = swapped ? src1 : src2;
  const int rs_f = swapped ? irow2 : irow1, rs_s = swapped ? irow1 : irow2;
  const int mv_f = swapped ? mvlower : mvupper, mv_s = swapped ? mvupper : mvlower;
  const int row_bytes = width * psize;
  if (direction < 1 || direction > 4) return;
  for (int j = 0; j < height; j++)
    for (int x = 0; x < row_bytes; x++) {
      int sj = j, sx = x;
      if (along_y ? j < bound : x < bound * psize) {
        if (mv_f) { if (along_y) sj = j + (height - bound); else sx = x + (width - bound) * psize; }
        dest[(long)orow * j + x] = first[(long)rs_f * sj + sx];
      } else {
        if (mv_s) { if (along_y) sj = j - bound; else sx = x - bound * psize; }
        dest[(long)orow * j + x] = second[(long)rs_s * sj + sx];
      }
    }
}

/* ==== per-frame diagnostics ===================================================================================== */

/* is_all_black_ish src/colourspace.c:2554-2594.  The non-exact branch is a chain of BITWISE ands (not logical ones): with
 * hi(x) = x & 0xE0 it reads  hi(a) & hi(c) & (hi(b) | (b has bits 4 and 3 ? 0x20 : 0)), so "not black" also needs the three
 * bytes to share a set bit among bits 5 .. 7 -- replicated as written. */
int pe_or_is_all_black_ish(int width, int height, int rowstride, int has_alpha, const uint8_t *pixels, int exact) {
  const int psize = has_alpha ? 4 : 3;
  for (int y = 0; y < height; y++) {
    const uint8_t *q = pixels + (long)rowstride * y;
    for (int x = 0; x < width; x++, q += psize) {
      const unsigned a = q[0], b = q[1], c = q[2];
      if (exact) {
        if (a | b | c) return 0;
      } else {
        const unsigned na = (a & 0x1F) ^ a, nc = (c & 0x1F) ^ c, nb = (b & 0x1F) ^ b;
        const unsigned t1 = ((b << 1) & 0x1F) ^ (b << 1), t2 = (((b & 0x0F) << 2) & 0x1F) ^ ((b & 0x0F) << 2);
        if (na & nc & (nb | (t1 & t2))) return 0;
      }
    }
  }
  return 1;
}

/* The reference's MD5 variant (src/maths.c:473-546, macros src/maths.h:42-56).  Padding, length and rounds 2 - 4 are RFC 1321's;
 * round 1 is its BX macro as written: the state words are stepped in the order A B C D with rotations 7, 22, 17, 12, and the
 * fourth step takes the second step's CONSTANT where the function's third operand would be the state word C. */
static uint32_t or_rotl(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
static uint32_t or_ff(uint32_t b, uint32_t c, uint32_t d) { return d ^ (b & (c ^ d)); }

static void or_md5_block(const uint8_t *blk, uint32_t st[4]) {
  static const uint32_t T[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
    0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
    0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
    0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
    0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
    0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
  static const int K[48] = {1, 6, 11, 0, 5, 10, 15, 4, 9, 14, 3, 8, 13, 2, 7, 12, 5, 8, 11, 14, 1, 4, 7, 10, 13, 0, 3, 6, 9, 12, 15, 2,
                            0, 7, 14, 5, 12, 3, 10, 1, 8, 15, 6, 13, 4, 11, 2, 9};
  static const int S[3][4] = {{5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
  uint32_t X[16], v[4] = {st[0], st[1], st[2], st[3]};
  for (int i = 0; i < 16; i++) X[i] = (uint32_t)blk[4 * i] | ((uint32_t)blk[4 * i + 1] << 8) | ((uint32_t)blk[4 * i + 2] << 16) | ((uint32_t)blk[4 * i + 3] << 24);
  for (int q = 0; q < 4; q++) { /* BX */
    static const int rot[4] = {7, 22, 17, 12};
    for (int j = 0; j < 4; j++) {
      /* step j updates word j (A, B, C, D in turn) from the next three words in A B C D order ... */
      uint32_t h = v[(j + 1) & 3], i2 = v[(j + 2) & 3], jj = v[(j + 3) & 3];
      if (j == 3) jj = T[4 * q + 1]; /* ... except that the last step's third operand is the second step's constant */
      v[j] += or_ff(h, i2, jj) + X[4 * q + j] + T[4 * q + j];
      v[j] = or_rotl(v[j], rot[j]) + h;
    }
  }
  for (int r = 0; r < 3; r++)
    for (int i = 0; i < 16; i++) {
      /* RFC 1321 order: A D C B, each from (next, next + 1, next + 2) in A B C D order */
      const int g = (4 - (i & 3)) & 3; /* 0, 3, 2, 1 */
      const uint32_t b = v[(g + 1) & 3], c = v[(g + 2) & 3], d = v[(g + 3) & 3];
      const uint32_t f = r == 0 ? or_ff(d, b, c) : r == 1 ? (b ^ c ^ d) : (c ^ (b | ~d));
      v[g] += f + X[K[16 * r + i]] + T[16 + 16 * r + i];
      v[g] = or_rotl(v[g], S[r][i & 3]) + b;
    }
  for (int i = 0; i < 4; i++) st[i] += v[i];
}

uint64_t pe_or_minimd5(const uint8_t *data, size_t n) {
  uint32_t st[4] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476};
  uint8_t tail[128];
  size_t full = n / 64, rem = n - full * 64, tl;
  for (size_t i = 0; i < full; i++) or_md5_block(data + 64 * i, st);
  memset(tail, 0, sizeof(tail));
  memcpy(tail, data + 64 * full, rem);
  tail[rem] = 0x80;
  tl = rem >= 56 ? 128 : 64;
  {
    const uint64_t bits = (uint64_t)n << 3;
    for (int i = 0; i < 8; i++) tail[tl - 8 + i] = (uint8_t)(bits >> (8 * i));
  }
  or_md5_block(tail, st);
  if (tl == 128) or_md5_block(tail + 64, st);
  return ((uint64_t)st[0] | ((uint64_t)st[1] << 32)) ^ ((uint64_t)st[2] | ((uint64_t)st[3] << 32));
}

uint64_t pe_or_row_hashes(const uint8_t *pixels, int nbytes, int height, int rowstride, uint64_t *out) {
  uint64_t parity = 0;
  for (int y = 0; y < height; y++) {
    out[y] = pe_or_minimd5(pixels + (long)rowstride * y, (size_t)nbytes);
    parity ^= out[y];
  }
  return parity;
}

/* ---- SURVEY 8f rank 3, second batch: softlight.c, layout_blends.c ("triple split"), multi_transitions.c ------------------------ */

/* softlight.c sqrti :33-47: digit-by-digit integer square root = floor(sqrt(n)) */
static uint32_t or_sqrti(uint32_t n) {
  uint32_t root = 0, rem = n, place = 0x40000000u;
  while (place > rem) place >>= 2;
  while (place) {
    const uint32_t t = root + place;
    if (rem >= t) { rem -= t; root += place << 1; }
    root >>= 1;
    place >>= 2;
  }
  return root;
}

/* softlight.c softlight_process :62-162 on the luma plane (the chroma planes are copied, :150-154): rows 0 and height - 1 and columns
 * 0 and width - 1 are copied; elsewhere an edge magnitude (two 3 x 3 differences exactly as written at :114-118 -- the last term of
 * row0 subtracts the lower-LEFT sample from the lower-right one, the last term of row1 ADDS the two lower corners) is scaled, clamped
 * to the luma range and mixed 64 : 192 with the source sample. */
void pe_or_softlight(const uint8_t *src, int irow, uint8_t *dst, int orow, int width, int height, int clamped) {
  const int ymin = clamped ? 16 : 0, ymax = clamped ? 235 : 255, scale = 384, mix = 192;
  memcpy(dst, src, (size_t)width);
  for (int y = 1; y < height - 1; y++) {
    const uint8_t *s = src + (long)irow * y;
    uint8_t *d = dst + (long)orow * y;
    d[0] = s[0];
    for (int x = 1; x < width - 1; x++) {
      const uint8_t *p = s + x;
      const int row0 = (p[irow - 1] - p[-irow - 1]) + ((p[irow] - p[-irow]) << 1) + (p[irow + 1] - p[irow - 1]);
      const int row1 = (p[-irow + 1] - p[-irow - 1]) + ((p[1] - p[-1]) << 1) + (p[irow + 1] + p[irow - 1]);
      int sum = (int)(((3 * or_sqrti((uint32_t)(row0 * row0 + row1 * row1)) / 2) * scale) >> 8);
      sum = sum < ymin ? ymin : sum > ymax ? ymax : sum;
      sum = ((256 - mix) * sum + mix * p[0]) >> 8;
      d[x] = (uint8_t)(sum < ymin ? ymin : sum > ymax ? ymax : sum);
    }
    if (width > 1) d[width - 1] = s[width - 1];
  }
  if (height > 1) memcpy(dst + (long)orow * (height - 1), src + (long)irow * (height - 1), (size_t)width);
}

/* layout_blends.c common_process :19-119 ("triple split", RGB24 / BGR24): per pixel one of src2 (outside the band), src1 (inside)
 * or the border colour, decided by the double comparisons of :92-99 on the byte offset j and the row.  colclass[x] / rowclass[y]:
 * bit 0 = "outside" test, bit 1 = "inside" test of that column / row, exactly as the reference evaluates them (the same function
 * serves the product's host code: the comparisons are per column and per row, never per pixel).  bordercol is in RGB order; a
 * BGR24 frame swaps it (:66-70). */
void pe_or_triple_split_classes(int width, int height, double xstart, int sym, double xend, int vert, double bw, uint8_t *colclass,
                                uint8_t *rowclass) {
  const int wb = width * 3;
  int tbs = height, tbe = height, bbs = height, bbe = height; /* rows; "end" = row `height` */
  if (sym) { xstart /= 2.; xend = 1. - xstart; }
  if (xstart > xend) { const double t = xend; xend = xstart; xstart = t; }
  if (vert) {
    tbs = (int)(height * (xstart - bw) + .5); tbe = (int)(height * (xstart + bw) + .5);
    bbs = (int)(height * (xend - bw) + .5); bbe = (int)(height * (xend + bw) + .5);
    xstart = xend = -bw;
  }
  for (int x = 0; x < width; x++) {
    const int j = 3 * x;
    const int out = (j < wb * (xstart - bw) || j >= wb * (xend + bw)) ? 1 : 0;
    const int in = (j > wb * (xstart + bw) && j < wb * (xend - bw)) ? 2 : 0;
    colclass[x] = (uint8_t)(out | in);
  }
  for (int y = 0; y < height; y++) rowclass[y] = (uint8_t)(((y <= tbs || y >= bbe) ? 1 : 0) | ((y > tbe && y < bbs) ? 2 : 0));
}

void pe_or_triple_split(const uint8_t *src1, int irow1, const uint8_t *src2, int irow2, uint8_t *dst, int orow, int width, int height,
                        int bgr, double xstart, int sym, double xend, int vert, double bw, const int bordercol[3]) {
  uint8_t *cc = (uint8_t *)malloc((size_t)width), *rc = (uint8_t *)malloc((size_t)height);
  const int c0 = bgr ? bordercol[2] : bordercol[0], c1 = bordercol[1], c2 = bgr ? bordercol[0] : bordercol[2];
  pe_or_triple_split_classes(width, height, xstart, sym, xend, vert, bw, cc, rc);
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      uint8_t *d = dst + (long)orow * y + 3 * x;
      if ((cc[x] & 1) && (rc[y] & 1)) memcpy(d, src2 + (long)irow2 * y + 3 * x, 3);
      else if ((cc[x] & 2) || (rc[y] & 2)) { if (d != src1 + (long)irow1 * y + 3 * x) memcpy(d, src1 + (long)irow1 * y + 3 * x, 3); }
      else { d[0] = (uint8_t)c0; d[1] = (uint8_t)c1; d[2] = (uint8_t)c2; }
    }
  free(cc); free(rc);
}

/* multi_transitions.c dissolve_init :42-70: mask[i] = (float)fastrand_dbl_re(1., ...) -- xorshift64 (13, 7, 17) of the host's random
 * seed (libweed/weed-plugin-utils.c:666,686-704), `val / divd / divd * range` with divd = (double)0xFFFFFFFF, which the plugins'
 * -ffast-math build (Makefile.am:49) folds into ONE multiplication by 1 / divd^2 (the constant 0x3BF0000000200000, read off the
 * compiled plugin) */
void pe_or_dissolve_mask(int64_t seed, long n, float *mask) {
  uint64_t x = (uint64_t)seed;
  union { uint64_t u; double d; } k;
  k.u = 0x3BF0000000200000ull;
  for (long i = 0; i < n; i++) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    mask[i] = (float)((double)x * k.d);
  }
}

/* multi_transitions.c common_process :85-225, types 0 "iris rectangle", 1 "iris circle", 2 "4 way split", 3 "dissolve" (type 4, "rand
 * replace", is a whole-frame copy of src1 or src2 decided by the plugin's own random stream: host logic, no pixel arithmetic).
 * Float expressions in the form the plugins' -ffast-math build evaluates them (divisions by loop invariants become multiplications by
 * a reciprocal computed once; read off the compiled plugin, checked against it in tests/test_oracle_vs_reference.py):
 *   0: xx = (int)((double)((float)(int)hwidth_bytes * bfneg) + .5), yy likewise with hheight; src2 inside [xx, wb - xx) x [yy, h - yy)
 *   1: t = (yyf * yyf + xxf * xxf) * (1.f / maxradsq), yyf = (float)(j - ihwidth) * (1.f / (float)psize); src1 when sqrt((double)t) > bf
 *   2: src2 when |i - hheight| * (1.f / hheight) < bf or |j - hwidth| * (1.f / hwidth) < bf or bf == 1; else src1 displaced by
 *      (+-yy bytes, +-xx rows), xx = (int)((double)(hheight * bf) + .5), yy = (int)((double)((bf * (1.f / psize)) * hwidth) + .5) * psize
 *   3: src2 where mask[pixel] < bf
 * dst may alias src1 for types 0, 1, 3 (pixels that keep src1 are not written). */
void pe_or_multi_transition(int type, const uint8_t *src1, int irow1, const uint8_t *src2, int irow2, uint8_t *dst, int orow, int width,
                            int height, int psize, double bfd, const float *mask) {
  const float bf = (float)bfd, bfneg = 1.f - bf;
  const float hheight = (float)height * 0.5f;
  const int wb = width * psize, ihwidth = wb >> 1, ihheight = height >> 1;
  const float hwidth_px = (float)width * 0.5f, hwidth = (float)wb * 0.5f;
  const float maxradsq = hheight * hheight + hwidth_px * hwidth_px;
  const float inv_maxradsq = 1.f / maxradsq, inv_psize = 1.f / (float)psize, inv_hh = 1.f / hheight, inv_hw = 1.f / hwidth;
  int xx = 0, yy = 0;
  if (type == 0) {
    xx = (int)((double)((float)(int)hwidth * bfneg) + .5);
    yy = (int)((double)((float)(int)hheight * bfneg) + .5);
  } else if (type == 2) {
    xx = (int)((double)(hheight * bf) + .5);
    yy = (int)((double)((bf * (psize == 3 ? 0.333333343267440796f : psize == 4 ? 0.25f : inv_psize)) * hwidth) + .5) * psize;
  }
  for (int i = 0; i < height; i++)
    for (int j = 0; j < wb; j += psize) {
      const uint8_t *s1 = src1 + (long)irow1 * i + j, *s2 = src2 + (long)irow2 * i + j;
      uint8_t *d = dst + (long)orow * i + j;
      const uint8_t *from = s1;
      if (type == 0) {
        if (!(j < xx || j >= wb - xx || i < yy || i >= height - yy)) from = s2;
      } else if (type == 1) {
        const float xxf = (float)(i - ihheight), yyf = (float)(j - ihwidth) * inv_psize;
        const float t = (yyf * yyf + xxf * xxf) * inv_maxradsq;
        if (!(sqrt((double)t) > (double)bf)) from = s2;
      } else if (type == 2) {
        if (fabsf((float)i - hheight) * inv_hh < bf || fabsf((float)j - hwidth) * inv_hw < bf || bf == 1.f) from = s2;
        else from = s1 + (j > ihwidth ? -yy : yy) + (long)(i > ihheight ? -xx : xx) * irow1;
      } else {
        if (mask[(long)i * width + j / psize] < bf) from = s2;
      }
      if (from != d) memmove(d, from, (size_t)psize);
    }
}

/* ---- the reference's float ("experimental") YUV -> RGB path: tables colourspace.c:1040-1104 (BT.709 only), clamp0255f :592,
 *      yuv2rgb_float :2367 ------------------------------------------------------------------------------------------------------- */
/* which: 0 RGBf_Y 1 Rf_Cr 2 Gf_Cb 3 Gf_Cr 4 Bf_Cb; clamping 0 clamped, 1 unclamped.  Evaluated in double and stored as float32, as
 * the reference's assignments do.  Quirks replicated: the clamped RGBf_Y keeps 0 from 235 up (the loop at :1051 starts where the
 * integer loop of :1050 ended, so it never runs); the clamped chroma tables saturate at 254 - 128 above 240 (:1083-1086) where the
 * integer ones use 255 - 128. */
void pe_or_float_table(int clamping, int which, float out[256]) {
  const double kr = 0.2126, kb = 0.0722;
  const double c[5] = {1., 2. * (1. - kr), -.5 / (1. + kb + kb), -.5 / (1. - kr), 2. * (1. - kb)};
  for (int i = 0; i < 256; i++) {
    double v;
    if (clamping == 1) v = which == 0 ? (double)i : c[which] * ((double)i - 128.);
    else if (which == 0) v = i <= 16 ? 0. : i < 235 ? ((double)i - 16.) / (235. - 16.) * 255. : 0.;
    else v = i <= 16 ? 0. : i < 240 ? c[which] * ((((double)i - 16.) / (240. - 16.) * 255.) - 128.) : c[which] * (254. - 128.);
    out[i] = (float)v;
  }
}

static uint8_t or_clampf_lc(float f) { /* clamp0255f, the lower-case inline function of colourspace.c:592-596 */
  if (f > 255.) f = 255.;
  if (f < 0.) f = 0;
  return (uint8_t)f;
}

/* mode 0: yuv2rgb_float as written (:2367-2372: `int yy = RGB_Y[y]`, the 16.16 INTEGER table, added to the float chroma tables);
 * mode 1: the form of the commented-out variant at :2398-2400 (RGBf_Y[y] + ...).  rgb_y_int: the BT.709 RGB_Y table of `clamping`
 * (pe_or_conv_table which 9); sums (optional): the float sums before the clamp */
void pe_or_yuv2rgb_float(int mode, int clamping, const int32_t *rgb_y_int, const uint8_t *yuv, uint8_t *rgb, float *sums, long n) {
  float ty[256], rcr[256], gcb[256], gcr[256], bcb[256];
  pe_or_float_table(clamping, 0, ty); pe_or_float_table(clamping, 1, rcr); pe_or_float_table(clamping, 2, gcb);
  pe_or_float_table(clamping, 3, gcr); pe_or_float_table(clamping, 4, bcb);
  for (long i = 0; i < n; i++) {
    const uint8_t y = yuv[i * 3], u = yuv[i * 3 + 1], v = yuv[i * 3 + 2];
    float r, g, b;
    if (mode == 0) {
      const int yy = rgb_y_int[y];
      r = yy + rcr[v]; g = yy + gcb[u] + gcr[v]; b = yy + bcb[u];
    } else {
      r = ty[y] + rcr[v]; g = ty[y] + gcb[u] + gcr[v]; b = ty[y] + bcb[u];
    }
    rgb[i * 3] = or_clampf_lc(r); rgb[i * 3 + 1] = or_clampf_lc(g); rgb[i * 3 + 2] = or_clampf_lc(b);
    if (sums) { sums[i * 3] = r; sums[i * 3 + 1] = g; sums[i * 3 + 2] = b; }
  }
}

/* ---- YUV411 (IYU1) as a conversion source: convert_yuv411_to_{rgb,bgr,argb,yuv888,yuvp,uyvy,yuyv}_frame  colourspace.c:8305-8910 ----
 * macropixel {u2, y0, y1, v2, y2, y3} (colourspace.h:203-210) = 4 pixels.  Per row of w macropixels the reference writes: 2 pixels
 * with macropixel 0's chroma; for j = 1 .. w-1 the 4 pixels (y2, y3 of j-1; y0, y1 of j) with the chroma ladder
 *   h = avg(p, c); qp = avg(h, p); qc = avg(h, c);  u = avg(qp, p), avg(qp, c), avg(qc, p), avg(qc, c)      (p = chroma j-1, c = chroma j)
 * (UYVY / YUYV: one step, avg(h, p) / avg(h, c)); 2 pixels with macropixel w-1's chroma.  avg = avg_chromaf = the cavg table of the
 * frame's clamping (:2100-2104); RGB through yuv2rgb with the YCbCr tables of that clamping.  Replicated: the planar 4:4:4 and packed
 * 4:2:2 variants write the first luma of a pair twice (:8745-8822, :8867-8893; planar also in the first pair :8726-8734).
 * Replicated under `quirks`: the BGR / BGRA variant writes the first pixel of a row and its last two in R, G, B order (:8445, :8514).
 * X (defined here): the RGBA / BGRA variants never write the alpha bytes of the first pixel pair of a loop iteration (:8338-8370):
 * 255; source and (except RGB) destination rows are walked densely: strides honoured.
 * target: 0 RGB (order, add_alpha), 1 packed 4:4:4, 2 planar 4:4:4 (dest[3] = alpha plane when add_alpha), 3 UYVY, 4 YUYV,
 * 5 planar 4:2:2, 6 planar 4:2:0 (dest[1] = Cb, dest[2] = Cr) */
void pe_or_yuv411_to(int target, const uint8_t *src, int irow, int width_mpx, int height, uint8_t *const dest[4], const int orow[4],
                     int order, int add_alpha, int clamping, int quality, int quirks) {
  const or_conv_t *c = or_conv(clamping, OR_SUBSPACE_YCBCR);
  const uint8_t *avg = or_avg(clamping);
  int ro = 0, go = 1, bo = 2, ao = -1, ps = 3;
  if (target == 0) or_order_offsets(order, add_alpha, &ro, &go, &bo, &ao, &ps);
  else if (target == 1) ps = add_alpha ? 4 : 3;
#define OR_AVG(x, y) (avg[((int)(x) << 8) + (int)(y)])
  for (int i = 0; i < height; i++) {
    const uint8_t *m = src + (long)irow * i;
    for (int j = 0; j <= width_mpx; j++) {
      uint8_t ys[4], us[4], vs[4];
      int npx, px0;
      if (j == 0 || j == width_mpx) {
        const uint8_t *q = m + 6L * (j == 0 ? 0 : width_mpx - 1);
        npx = 2; px0 = j == 0 ? 0 : 4 * width_mpx - 2;
        ys[0] = j == 0 ? q[1] : q[4]; ys[1] = j == 0 ? q[2] : q[5];
        if (target == 2 && j == 0) ys[1] = ys[0];
        us[0] = us[1] = q[0]; vs[0] = vs[1] = q[3];
      } else {
        const uint8_t *p = m + 6L * (j - 1), *q = p + 6;
        const uint8_t pu = p[0], pv = p[3], cu = q[0], cv = q[3];
        const uint8_t hu = OR_AVG(pu, cu), hv = OR_AVG(pv, cv);
        npx = 4; px0 = 4 * j - 2;
        ys[0] = p[4]; ys[1] = p[5]; ys[2] = q[1]; ys[3] = q[2];
        if (target >= 3) {
          us[0] = us[1] = OR_AVG(hu, pu); vs[0] = vs[1] = OR_AVG(hv, pv);
          us[2] = us[3] = OR_AVG(hu, cu); vs[2] = vs[3] = OR_AVG(hv, cv);
          if (target < 5) { ys[1] = ys[0]; ys[3] = ys[2]; }
        } else {
          const uint8_t qpu = OR_AVG(hu, pu), qpv = OR_AVG(hv, pv), qcu = OR_AVG(hu, cu), qcv = OR_AVG(hv, cv);
          us[0] = OR_AVG(qpu, pu); vs[0] = OR_AVG(qpv, pv); us[1] = OR_AVG(qpu, cu); vs[1] = OR_AVG(qpv, cv);
          us[2] = OR_AVG(qcu, pu); vs[2] = OR_AVG(qcv, pv); us[3] = OR_AVG(qcu, cu); vs[3] = OR_AVG(qcv, cv);
          if (target == 2) { ys[1] = ys[0]; ys[3] = ys[2]; }
        }
      }
      for (int k = 0; k < npx; k++) {
        const long x = px0 + k;
        if (target == 0) {
          uint8_t *d = dest[0] + (long)orow[0] * i + x * ps;
          /* convert_yuv411_to_bgr_frame hands the row's first pixel and its last two to uyvy2rgb in R, G, B order (:8445, :8514) */
          const int swap = quirks && order == OR_ORDER_BGR && (x == 0 || x >= 4L * width_mpx - 2);
          or_px_yuv2rgb(c, quality, NULL, ys[k], us[k], vs[k], &d[swap ? bo : ro], &d[go], &d[swap ? ro : bo]);
          if (ao >= 0) d[ao] = 255;
        } else if (target == 1) {
          uint8_t *d = dest[0] + (long)orow[0] * i + x * ps;
          d[0] = ys[k]; d[1] = us[k]; d[2] = vs[k];
          if (add_alpha) d[3] = 255;
        } else if (target == 2) {
          dest[0][(long)orow[0] * i + x] = ys[k]; dest[1][(long)orow[1] * i + x] = us[k]; dest[2][(long)orow[2] * i + x] = vs[k];
          if (add_alpha) dest[3][(long)orow[3] * i + x] = 255;
        } else if (target >= 5) {
          /* planar 4:2:2 (:8976-9032, every luma kept) / 4:2:0: chroma row k = avg(4:2:2 row 2k, 4:2:2 row 2k+1), the even row as the
           * table row; a trailing unpaired row keeps its own (the reference's 4:2:0 loop never advances its chroma pointers on odd
           * rows and rewinds them on even ones, :9060-9140: only chroma row 0 is ever written -- X, defined by its evident intent) */
          dest[0][(long)orow[0] * i + x] = ys[k];
          if (!(k & 1)) {
            const long cr = target == 5 ? i : i >> 1;
            uint8_t *du = dest[1] + (long)orow[1] * cr + (x >> 1), *dv = dest[2] + (long)orow[2] * cr + (x >> 1);
            if (target == 5 || !(i & 1)) { *du = us[k]; *dv = vs[k]; }
            else { *du = OR_AVG(*du, us[k]); *dv = OR_AVG(*dv, vs[k]); }
          }
        } else if (!(k & 1)) {
          uint8_t *d = dest[0] + (long)orow[0] * i + (x >> 1) * 4;
          if (target == 3) { d[0] = us[k]; d[1] = ys[k]; d[2] = vs[k]; d[3] = ys[k + 1]; }
          else { d[0] = ys[k]; d[1] = us[k]; d[2] = ys[k + 1]; d[3] = vs[k]; }
        }
      }
    }
  }
#undef OR_AVG
}

/* ---- RGB(A) / BGR(A) / ARGB -> YUV411: convert_{rgb,bgr,argb}_to_yuv411_frame :6499-6614, rgb2_411 :2323-2343.  Whole macropixels
 * only (the rightmost width % 4 pixels are cut); always the YCbCr tables; luma per pixel `(Y_R + Y_G + Y_B) >> 16` clamped, chroma = the
 * sum of the four pixels' `>> 16` terms, `>> 2`, clamped (plain shifts: rgb2_411 does not go through spc_rnd).  The reference writes
 * the macropixels densely (`u++`); the output rowstride is honoured here (X). */
void pe_or_rgb_to_yuv411(const uint8_t *src, int irow, int width, int height, uint8_t *dest, int orow, int order, int in_alpha,
                         int clamping) {
  const or_conv_t *c = or_conv(clamping, OR_SUBSPACE_YCBCR);
  int ro, go, bo, ao, ips;
  or_order_offsets(order, in_alpha, &ro, &go, &bo, &ao, &ips);
  for (int i = 0; i < height; i++) {
    const uint8_t *s = src + (long)irow * i;
    uint8_t *d = dest + (long)orow * i;
    for (int j = 0; j < width >> 2; j++, s += 4 * ips, d += 6) {
      int su = 0, sv = 0;
      uint8_t yy[4];
      for (int k = 0; k < 4; k++) {
        const uint8_t r = s[k * ips + ro], g = s[k * ips + go], b = s[k * ips + bo];
        int a = (c->t[0][r] + c->t[1][g] + c->t[2][b]) >> 16;
        yy[k] = a > c->max_y ? c->max_y : a < c->min_y ? c->min_y : a;
        su += (c->t[3][r] + c->t[4][g] + c->t[5][b]) >> 16;
        sv += (c->t[6][r] + c->t[7][g] + c->t[8][b]) >> 16;
      }
      su >>= 2; sv >>= 2;
      d[0] = su > c->max_uv ? c->max_uv : su < c->min_uv ? c->min_uv : su;
      d[1] = yy[0]; d[2] = yy[1];
      d[3] = sv > c->max_uv ? c->max_uv : sv < c->min_uv ? c->min_uv : sv;
      d[4] = yy[2]; d[5] = yy[3];
    }
  }
}

/* ---- YUV -> YUV411.  mode 0 UYVY / 1 YUYV (convert_{uyvy,yuyv}_to_yuv411_frame :7973-8032): one macropixel from two, chroma =
 * avg_chroma(first, second).  mode 2 YUV420P / 3 YUV422P (convert_yuv420_to_yuv411_frame :9148-9195): chroma = avg_chroma of the two
 * samples under the four pixels; 4:2:0 reads chroma row r >> 1 for luma row r and then folds every EVEN row r >= 2 into the macropixels
 * of row r - 1: u2(r-1) = avg_chroma(u2(r-1), u2(r)) (:9176-9179) -- replicated.  mode 4 YUV888 / 5 YUVA8888
 * (convert_yuv888_to_yuv411_frame :8272-8302): chroma = (sum of the four samples) >> 2, no table; the reference stops after
 * width * height BYTES, a third / a quarter of the frame (:8278) -- X, every row here, pinned on the rows it reaches.  mode 6 planar
 * 4:4:4 (convert_yuvp_to_yuv411_frame :7755-7797): chroma = avg_chroma(avg_chroma(c0, c1), avg_chroma(c2, c3)); the reference never
 * advances its output pointer (every macropixel lands on the first one) nor its luma rows by the padding -- X, pinned macropixel by
 * macropixel.  All walk densely; strides honoured (X).  width in pixels; whole macropixels only. */
void pe_or_to_yuv411(int mode, const uint8_t *const src[3], const int irow[3], int width, int height, uint8_t *dest, int orow,
                     int clamping) {
  const uint8_t *avg = or_avg(clamping);
  const int wm = width >> 2;
#define OR_AVG(x, y) (avg[((int)(x) << 8) + (int)(y)])
  for (int i = 0; i < height; i++) {
    uint8_t *d = dest + (long)orow * i;
    for (int j = 0; j < wm; j++, d += 6) {
      if (mode <= 1) {
        uint8_t y0, u0, y1, v0, y2, u1, y3, v1;
        or_mpx(mode, src[0] + (long)irow[0] * i + 8L * j, &y0, &u0, &y1, &v0);
        or_mpx(mode, src[0] + (long)irow[0] * i + 8L * j + 4, &y2, &u1, &y3, &v1);
        d[0] = OR_AVG(u0, u1); d[1] = y0; d[2] = y1; d[3] = OR_AVG(v0, v1); d[4] = y2; d[5] = y3;
      } else if (mode <= 3) {
        const long cr = mode == 2 ? i >> 1 : i;
        const uint8_t *y = src[0] + (long)irow[0] * i + 4L * j, *u = src[1] + (long)irow[1] * cr + 2L * j, *v = src[2] + (long)irow[2] * cr + 2L * j;
        d[0] = OR_AVG(u[0], u[1]); d[1] = y[0]; d[2] = y[1]; d[3] = OR_AVG(v[0], v[1]); d[4] = y[2]; d[5] = y[3];
        if (mode == 2 && i >= 2 && !(i & 1)) {
          uint8_t *p = d - orow;
          p[0] = OR_AVG(p[0], d[0]); p[3] = OR_AVG(p[3], d[3]);
        }
      } else if (mode <= 5) {
        const int ps = mode == 4 ? 3 : 4;
        const uint8_t *q = src[0] + (long)irow[0] * i + 4L * ps * j;
        d[0] = (uint8_t)((q[1] + q[ps + 1] + q[2 * ps + 1] + q[3 * ps + 1]) >> 2);
        d[1] = q[0]; d[2] = q[ps];
        d[3] = (uint8_t)((q[2] + q[ps + 2] + q[2 * ps + 2] + q[3 * ps + 2]) >> 2);
        d[4] = q[2 * ps]; d[5] = q[3 * ps];
      } else {
        const uint8_t *y = src[0] + (long)irow[0] * i + 4L * j, *u = src[1] + (long)irow[1] * i + 4L * j, *v = src[2] + (long)irow[2] * i + 4L * j;
        d[0] = OR_AVG(OR_AVG(u[0], u[1]), OR_AVG(u[2], u[3])); d[1] = y[0]; d[2] = y[1];
        d[3] = OR_AVG(OR_AVG(v[0], v[1]), OR_AVG(v[2], v[3])); d[4] = y[2]; d[5] = y[3];
      }
    }
  }
#undef OR_AVG
}
