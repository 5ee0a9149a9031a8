// pe_tables.cpp -- host-side table construction (see pe_tables.h).
//
// Bit-parity notes
//  * colour matrices are evaluated in double with the reference's operand order
//    (a * i * clamp_factor * SCALE_FACTOR, left to right) and rounded half away from zero
//    (myround, src/maths.h:118); SCALE_FACTOR is 65793 = 0xFFFFFF / 0xFF (colourspace.h:60).
//  * the BT.709 green/Cb coefficient uses 1 + Kb + Kb and the YCbCr one 1 + Kb + Kr exactly as the
//    reference spells them (colourspace.c:1005,1062) -- they are table constants, not "fixed" here.
//  * gamma LUTs go through float32 / powf and reproduce the reference's loop-carried overwrite of
//    gamma_from (colourspace.c:701): after the first entry the source is treated as linear.
#include "pe_tables.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/pixel_engine.h"

namespace pe {

namespace {

inline int round_half_away(double n) { return n >= 0. ? (int)(n + 0.5) : (int)(n - 0.5); }

constexpr double kScale = 65793.;  // SCALE_FACTOR with USE_EXTEND (colourspace.h:13,60)

struct Primaries { double kr, kb; };
inline Primaries primaries_for(int subspace) {
  // colourspace.h:84-91; WEED_YUV_SUBSPACE_YUV is treated as YCbCr (colourspace.c:226)
  if (subspace == PE_YUV_SUBSPACE_BT709) return {0.2126, 0.0722};
  return {0.299, 0.114};
}

}  // namespace

void build_conv_tables(int clamping, int subspace, ConvTables *out) {
  const Primaries p = primaries_for(subspace);
  const bool clamped = (clamping == PE_YUV_CLAMPING_CLAMPED);
  const bool hd = (subspace == PE_YUV_SUBSPACE_BT709);
  const double kg = 1. - p.kr - p.kb;          // written (1. - KR - KB) in the luma rows
  const double kg_c = 1. - p.kb - p.kr;        // written (1. - KB - KR) in the chroma rows
  const double fy = (235. - 16.) / 255.;       // CLAMP_FACTOR_Y  colourspace.h:120
  const double fuv = (240. - 16.) / 255.;      // CLAMP_FACTOR_UV colourspace.h:121
  const double fac_b = .5 / (1. - p.kb), fac_r = .5 / (1. - p.kr);

  // ---- RGB -> YUV (init_RGB_to_YUV_tables, colourspace.c:851-981)
  for (int i = 0; i < 256; i++) {
    const double d = (double)i;
    if (clamped) {
      out->t[Y_R][i] = round_half_away(p.kr * d * fy * kScale);
      out->t[Y_G][i] = round_half_away(kg * d * fy * kScale);
      out->t[Y_B][i] = round_half_away((p.kb * d * fy + 16.) * kScale);
      out->t[CB_R][i] = round_half_away(-fac_b * p.kr * d * fuv * kScale);
      out->t[CB_G][i] = round_half_away(-fac_b * kg_c * d * fuv * kScale);
      out->t[CB_B][i] = round_half_away((0.5 * d * fuv + 128.) * kScale);
      out->t[CR_R][i] = round_half_away((0.5 * d * fuv + 128.) * kScale);
      out->t[CR_G][i] = round_half_away(-fac_r * kg_c * d * fuv * kScale);
      out->t[CR_B][i] = round_half_away(-fac_r * p.kb * d * fuv * kScale);
    } else {
      out->t[Y_R][i] = round_half_away(p.kr * d * kScale);
      out->t[Y_G][i] = round_half_away(kg * d * kScale);
      out->t[Y_B][i] = round_half_away(p.kb * d * kScale);
      out->t[CB_R][i] = round_half_away(-fac_b * p.kr * d * kScale);
      out->t[CB_G][i] = round_half_away(-fac_b * kg_c * d * kScale);
      out->t[CB_B][i] = round_half_away((0.5 * d + 128.) * kScale);
      out->t[CR_R][i] = round_half_away((0.5 * d + 128.) * kScale);
      out->t[CR_G][i] = round_half_away(-fac_r * kg_c * d * kScale);
      out->t[CR_B][i] = round_half_away(-fac_r * p.kb * d * kScale);
    }
  }
  if (clamped) { out->min_y = out->min_uv = 16; out->max_y = 235; out->max_uv = 240; }
  else { out->min_y = out->min_uv = 0; out->max_y = out->max_uv = 255; }

  // ---- YUV -> RGB (init_YUV_to_RGB_tables, colourspace.c:984-1105)
  const double c_rcr = 2. * (1. - p.kr);
  const double c_gcb = hd ? -.5 / (1. + p.kb + p.kb) : -.5 / (1. + p.kb + p.kr);
  const double c_gcr = -.5 / (1. - p.kr);
  const double c_bcb = 2. * (1. - p.kb);
  auto chroma_row = [&](int i, double centred) {
    out->t[R_CR][i] = round_half_away(c_rcr * centred * kScale);
    out->t[G_CB][i] = round_half_away(c_gcb * centred * kScale);
    out->t[G_CR][i] = round_half_away(c_gcr * centred * kScale);
    out->t[B_CB][i] = round_half_away(c_bcb * centred * kScale);
  };
  if (clamped) {
    for (int i = 0; i < 256; i++) {
      if (i <= 16) out->t[RGB_Y][i] = 0;
      else if (i < 235) out->t[RGB_Y][i] = round_half_away(((double)i - 16.) / (235. - 16.) * 255. * kScale);
      else out->t[RGB_Y][i] = (int)(255 * kScale);
      if (i <= 16) {
        out->t[R_CR][i] = out->t[G_CB][i] = out->t[G_CR][i] = out->t[B_CB][i] = 0;
      } else if (i < 240) {
        chroma_row(i, (((double)i - 16.) / (240. - 16.) * 255.) - 128.);
      } else {
        // above 240: YCbCr saturates at the value for 240 (:1015-1024), BT.709 at 255 - 128 (:1078-1081)
        chroma_row(i, hd ? (255. - 128.) : (((240. - 16.) / (240. - 16.) * 255.) - 128.));
      }
    }
  } else {
    for (int i = 0; i < 256; i++) {
      out->t[RGB_Y][i] = (int)(i * kScale);
      chroma_row(i, (double)i - 128.);
    }
  }
}

// ---- float YUV -> RGB tables (colourspace.c:1040-1104, BT.709 only) ------------------------------------------------------------
// Kept as the reference leaves them: the clamped RGBf_Y holds 0 from 235 up (the loop at :1051 starts where the integer loop of :1050
// ended and never runs -- the statics stay zero); the clamped chroma tables saturate at 254 - 128 above 240 (:1083-1086) where the
// integer tables use 255 - 128.
void build_float_yuv_tables(int clamping, float out[5][256]) {
  const double kr = 0.2126, kb = 0.0722;
  const double c[5] = {1., 2. * (1. - kr), -.5 / (1. + kb + kb), -.5 / (1. - kr), 2. * (1. - kb)};
  const bool clamped = clamping == PE_YUV_CLAMPING_CLAMPED;
  for (int w = 0; w < 5; w++)
    for (int i = 0; i < 256; i++) {
      double v;
      if (!clamped) v = w == 0 ? (double)i : c[w] * ((double)i - 128.);
      else if (w == 0) v = i <= 16 ? 0. : i < 235 ? ((double)i - 16.) / (235. - 16.) * 255. : 0.;
      else v = i <= 16 ? 0. : i < 240 ? c[w] * ((((double)i - 16.) / (240. - 16.) * 255.) - 128.) : c[w] * (254. - 128.);
      out[w][i] = (float)v;
    }
}

// ---- gamma ------------------------------------------------------------------------------------

namespace {

struct GammaTx { float offs, lin, thresh, pf; };

// INIT_GAMMA (colourspace.h:157-161) for sRGB and BT.709 (colourspace.h:168-169)
void gamma_tx_table(GammaTx tx[2]) {
  tx[0] = {0.f, 12.92f, 0.04045f, 2.4f};
  tx[1] = {0.f, 4.5f, 0.018f, (float)(1. / .45)};
  for (int k = 0; k < 2; k++) {
    const float knee = powf((tx[k].thresh / tx[k].lin), (1. / tx[k].pf));
    tx[k].offs = (knee - tx[k].thresh) / (1. - knee);
  }
}

inline int tx_index(int gamma_type) { return gamma_type == PE_GAMMA_BT709 ? 1 : 0; }  // get_gamma_idx :625

struct GammaWalk {
  // state carried from one LUT entry to the next, as the reference's parameter is (colourspace.c:697-713)
  int from;
  const int to;
  const double fileg, screen_gamma;
  float inv_gamma;
  GammaTx tx[2];

  GammaWalk(double fg, int f, int t, double sg) : from(f), to(t), fileg(fg), screen_gamma(sg), inv_gamma(0.f) {
    gamma_tx_table(tx);
    if (to == PE_GAMMA_MONITOR) inv_gamma = 1. / (float)screen_gamma;
  }

  float step(float a) {
    float x = a;
    if (fileg != 1.0) x = powf(a, fileg);
    if (from == PE_GAMMA_MONITOR) {
      x = powf(a, screen_gamma);
      from = PE_GAMMA_SRGB;
    }
    if (from != PE_GAMMA_LINEAR && !(from == PE_GAMMA_SRGB && to == PE_GAMMA_MONITOR)) {
      const GammaTx &g = tx[tx_index(from)];
      a = (a < g.thresh) ? a / g.lin : powf((a + g.offs) / (1. + g.offs), g.pf);
      from = PE_GAMMA_LINEAR;
    }
    if (to != PE_GAMMA_LINEAR) {
      const GammaTx &g = tx[to == PE_GAMMA_MONITOR ? 0 : tx_index(to)];
      x = (a < (g.thresh) / g.lin) ? a * g.lin : powf((1. + g.offs) * a, 1. / g.pf) - g.offs;
    }
    if (to == PE_GAMMA_MONITOR) x = powf(a, inv_gamma);
    return x;
  }
};

inline bool gamma_is_noop(double fileg, int from, int to) {
  return fileg == 1.0 && (to == from || to == PE_GAMMA_UNKNOWN || from == PE_GAMMA_UNKNOWN);  // :662-663
}

}  // namespace

bool build_gamma_lut8(double fileg, int gamma_from, int gamma_to, double screen_gamma, uint8_t out[256]) {
  if (gamma_is_noop(fileg, gamma_from, gamma_to)) return false;
  GammaWalk w(fileg, gamma_from, gamma_to, screen_gamma);
  out[0] = 0;
  for (int i = 1; i < 256; ++i) {
    const float x = w.step((float)i / 255.);
    const int v = (int)(x * 255.);  // CLAMP0_255i, colourspace.h:22-23
    out[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
  }
  return true;
}

bool build_gamma_lut16(double fileg, int gamma_from, int gamma_to, double screen_gamma, uint16_t *out) {
  if (gamma_is_noop(fileg, gamma_from, gamma_to)) return false;
  GammaWalk w(fileg, gamma_from, gamma_to, screen_gamma);
  out[0] = 0;
  for (int i = 1; i < 65536; ++i) {
    const float x = w.step((float)i / 65536.);
    out[i] = x >= 0.99999 ? 65535 : x < 0.00001 ? 0 : (uint16_t)(x * 65535.9999);  // CLAMP16bit colourspace.h:16
  }
  return true;
}

// ---- premultiply (init_unal, colourspace.c:1141-1160) ------------------------------------------

namespace {
// CLAMP0255f (src/maths.h:88) evaluated in the type of the expression handed to it
template <typename T>
inline uint8_t clamp_round_u8(T a) { return a >= 254.5 ? (uint8_t)255 : a < -0.5 ? (uint8_t)0 : (uint8_t)(a + .5); }
}  // namespace

void build_premult_table(int which, uint8_t *out) {
  for (int a = 0; a < 256; a++) {
    const float alpha = (float)255. / (float)a;
    for (int v = 0; v < 256; v++) {
      int r;
      switch (which) {
      case 0: r = clamp_round_u8((float)v / alpha); break;                               // unal
      case 1: r = clamp_round_u8((float)v * alpha); break;                               // al
      case 2: r = (int)((float)v / alpha + .5) > (235. - 16.) ? 235                      // unalcy
                  : (int)((float)(v - 16.) / alpha + 16. + .5); break;
      case 3: r = (int)((float)v / alpha + .5) > (240. - 16.) ? 240                      // alcy
                  : (int)((float)(v - 16.) / alpha + 16. + .5); break;
      case 4: r = clamp_round_u8((float)(v - 16.) * alpha + 16.); break;                 // unalcuv
      default: r = clamp_round_u8((float)(v - 128.) * alpha + 128.); break;              // alcuv
      }
      out[a * 256 + v] = (uint8_t)r;  // the int tables are only ever stored into bytes (:12061,12071)
    }
  }
}

// init_average (colourspace.c:190-217, !MULT_AVG): the chroma averaging tables behind avg_chroma().  clamped (cavgc): float
// maths as the reference spells it, result clamped to 16..240; unclamped (cavgu): short arithmetic
void build_avg_table(bool clamped, uint8_t *out) {
  for (int x = 0; x < 256; x++) {
    const float fa = (float)(x - 128.) * 255. / 244.;
    const short sa = (short)(x - 128);
    for (int y = 0; y < 256; y++) {
      const float fb = (float)(y - 128.) * 255. / 244.;
      const short sb = (short)(y - 128);
      const float fc = (fa + fb) * 224. / 512. + 128.;
      const short c = ((sa + sb) >> 1) + 128;
      out[x * 256 + y] = clamped ? (uint8_t)(fc > 240. ? 240 : fc < 16. ? 16 : fc) : (uint8_t)(c > 255 ? 255 : c < 0 ? 0 : c);
    }
  }
}

bool avg_form_matches(bool clamped, const uint8_t *table) {
  const AvgForm F = avg_form(clamped);
  for (int x = 0; x < 256; x++)
    for (int y = 0; y < 256; y++) {
      const uint32_t n = (uint32_t)(x + y) * F.A + F.B;
      int f = (int)(((unsigned long long)n * F.M) >> 32);
      f = f < F.lo ? F.lo : f > F.hi ? F.hi : f;
      if (table[x * 256 + y] != f) return false;
    }
  return true;
}

// init_YUV_to_YUV_tables (colourspace.c:1108-1138): clamped <-> unclamped, same subspace.  Limits 16 / 235 / 240
// (colourspace.h:96-109); the first luma loop runs to i <= 16, the first chroma loop to i < 16; rounding is myround
// (maths.h:118: half away from zero, in double)
void build_yy_table(int which, uint8_t out[256]) {
  const double lo = 16., ymax = 235., cmax = 240.;
  int i = 0;
  switch (which) {
  case 0:
    for (; i <= 16; i++) out[i] = 0;
    for (; i < 235; i++) out[i] = (uint8_t)round_half_away((i - lo) * 255. / (ymax - lo));
    for (; i < 256; i++) out[i] = 255;
    break;
  case 1:
    for (; i < 16; i++) out[i] = 0;
    for (; i < 240; i++) out[i] = (uint8_t)round_half_away((i - lo) * 255. / (cmax - lo));
    for (; i < 256; i++) out[i] = 255;
    break;
  case 2:
    for (; i < 256; i++) out[i] = (uint8_t)round_half_away((i / 255.) * (ymax - lo) + lo);
    break;
  default:
    for (; i < 256; i++) out[i] = (uint8_t)round_half_away((i / 255.) * (cmax - lo) + lo);
    break;
  }
}

void build_plugin_luma_tables(int32_t yr[256], int32_t yg[256], int32_t yb[256]) {
  for (int i = 0; i < 256; i++) {
    yr[i] = round_half_away(0.299 * (double)i * 65536.);
    yg[i] = round_half_away((1. - 0.299 - 0.114) * (double)i * 65536.);
    yb[i] = round_half_away(0.114 * (double)i * 65536.);
  }
}

// ---- resize filter bank (our contract; the reference delegates to libswscale) --------------------

bool build_resize_filter(int src_n, int dst_n, int shift_bits, ResizeFilter *out) {
  if (src_n <= 0 || dst_n <= 0) return false;
  const double ratio = (double)src_n / (double)dst_n;
  const bool down = ratio > 1.;
  const double support = down ? ratio : 1.;
  const int taps = down ? (int)std::ceil(2. * ratio) + 1 : 2;
  if (taps > 64) return false;
  const int one = 1 << shift_bits;
  out->taps = taps;
  out->first.assign(dst_n, 0);
  out->coef.assign((size_t)dst_n * taps, 0);
  std::vector<double> w(taps);
  for (int i = 0; i < dst_n; i++) {
    const double centre = ((double)i + 0.5) * ratio - 0.5;
    const int left = down ? (int)std::floor(centre - support) + 1 : (int)std::floor(centre);
    double sum = 0.;
    for (int k = 0; k < taps; k++) {
      const double d = std::fabs((double)(left + k) - centre) / support;
      w[k] = d < 1. ? 1. - d : 0.;
      sum += w[k];
    }
    int acc = 0, big = 0;
    for (int k = 0; k < taps; k++) {
      const int q = (int)std::floor(w[k] / sum * one + 0.5);
      out->coef[(size_t)i * taps + k] = (int16_t)q;
      acc += q;
      if (w[k] > w[big]) big = k;
    }
    out->coef[(size_t)i * taps + big] += (int16_t)(one - acc);
    out->first[i] = left;
  }
  return true;
}

// libswscale's coefficient recipes (third-party library the reference calls, src/colourspace.c:15059; not in the reference tree,
// version not pinned by it; flags per LiVESInterpType at :14991-14997).  Restated from the published algorithm (libswscale/utils.c
// initFilter) and checked against libswscale 9.1.100 (tests/test_resize_vs_swscale.py): tap weights at 2^-30 distance precision --
// triangle (SWS_BILINEAR), the (B, C) = (0, 0.6) cubic in 24-bit fixed point (SWS_BICUBIC), 3-lobe Lanczos in double (SWS_LANCZOS),
// or the two-tap bank SWS_FAST_BILINEAR uses vertically --, near-zero taps (cumulated weight < 0.002) dropped from either end, taps
// outside the frame folded onto the edge sample, normalised to 1 << shift_bits with the rounding error carried from tap to tap.
namespace {

struct SwsTaps {              // raw 64-bit weights before shrinking
  int fs = 0;                 // taps per output sample
  int64_t fone = 0;
  std::vector<int64_t> w;     // [dst_n * fs]
  std::vector<int32_t> pos;   // [dst_n]
};

int64_t sws_weight(SwsKind kind, int64_t d, int64_t fone) {
  switch (kind) {
  case SWS_KIND_BICUBIC: {
    const int64_t B = 0, Cq = (int64_t)(0.6 * (1 << 24)), unit = (int64_t)1 << 30;
    int64_t c = 0;
    if (d < 2 * unit) {
      const int64_t dd = (d * d) >> 30, ddd = (dd * d) >> 30;
      c = d < unit ? (12 * (1 << 24) - 9 * B - 6 * Cq) * ddd + (-18 * (1 << 24) + 12 * B + 6 * Cq) * dd + (6 * (1 << 24) - 2 * B) * unit
                   : (-B - 6 * Cq) * ddd + (6 * B + 30 * Cq) * dd + (-12 * B - 48 * Cq) * d + (8 * B + 24 * Cq) * unit;
    }
    return c / (((int64_t)1 << 54) / fone);
  }
  case SWS_KIND_LANCZOS: {
    const double fd = (double)d * (1.0 / (1 << 30)), lobes = 3.0;
    double v = fd == 0.0 ? 1.0 : std::sin(fd * M_PI) * std::sin(fd * M_PI / lobes) / (fd * fd * M_PI * M_PI / lobes);
    if (fd > lobes) v = 0;
    return (int64_t)(v * (double)fone);
  }
  default: {
    const int64_t c = ((int64_t)1 << 30) - d;
    return c < 0 ? 0 : c * (fone >> 30);
  }
  }
}

}  // namespace

bool build_resize_filter_sws(int src_n, int dst_n, int shift_bits, ResizeFilter *out, SwsKind kind) {
  if (src_n <= 0 || dst_n <= 0) return false;
  const int64_t xinc = (((int64_t)src_n << 16) + (dst_n >> 1)) / dst_n, one = (int64_t)1 << shift_bits;
  if (kind == SWS_KIND_FAST_H) {
    // SWS_FAST_BILINEAR's horizontal pass is a 16.16 position walk from the left edge (hyscale_fast): (s[xx] << 7) + (s[xx + 1] -
    // s[xx]) * xalpha with a 7-bit xalpha -- as 14-bit taps {(128 - xalpha) << 7, xalpha << 7}; the tail holds the last sample
    out->taps = 2;
    out->first.assign(dst_n, 0);
    out->coef.assign((size_t)dst_n * 2, 0);
    uint32_t xpos = 0;
    for (int i = 0; i < dst_n; i++, xpos += (uint32_t)xinc) {
      int xx = (int)(xpos >> 16), xalpha = (int)((xpos & 0xFFFF) >> 9);
      if (xx >= src_n - 1) { xx = src_n - 2; xalpha = 128; }
      if (xx < 0) { xx = 0; xalpha = 0; }
      out->first[i] = xx;
      out->coef[(size_t)i * 2] = (int16_t)((128 - xalpha) << 7);
      out->coef[(size_t)i * 2 + 1] = (int16_t)(xalpha << 7);
    }
    return true;
  }
  SwsTaps T;
  int lg = 0;
  for (int r = src_n / dst_n; r > 1; r >>= 1) lg++;
  T.fone = (int64_t)1 << (54 - std::min(lg, 8));
  T.pos.assign(dst_n, 0);
  if (kind == SWS_KIND_FAST_V) {  // two taps around a top-left aligned position, whatever the scale factor
    T.fs = 2;
    T.w.assign((size_t)dst_n * 2, 0);
    int64_t at = (xinc >> 1) - 0x8000;
    for (int i = 0; i < dst_n; i++, at += xinc) {
      const int xx = (int)(at >> 16);
      T.pos[i] = xx;
      for (int j = 0; j < 2; j++) {
        const int64_t c = T.fone - std::llabs(((int64_t)(xx + j) << 16) - at) * (T.fone >> 16);
        T.w[(size_t)i * 2 + j] = c < 0 ? 0 : c;
      }
    }
  } else {
    const int size_factor = kind == SWS_KIND_BICUBIC ? 4 : kind == SWS_KIND_LANCZOS ? 6 : 2;
    T.fs = xinc <= (1 << 16) ? 1 + size_factor : 1 + (size_factor * src_n + dst_n - 1) / dst_n;
    T.fs = std::max(std::min(T.fs, src_n - 2), 1);
    if (T.fs > 64) return false;
    T.w.assign((size_t)dst_n * T.fs, 0);
    int64_t at = xinc - 65536;  // both grids sampled at pixel centres
    for (int i = 0; i < dst_n; i++, at += 2 * xinc) {
      int xx = (int)((at - (int64_t)(T.fs - 2) * 65536) / (1 << 17));  // towards zero
      T.pos[i] = xx;
      for (int j = 0; j < T.fs; j++, xx++) {
        int64_t d = std::llabs((int64_t)xx * (1 << 17) - at) << 13;
        if (xinc > (1 << 16)) d = d * dst_n / src_n;
        T.w[(size_t)i * T.fs + j] = sws_weight(kind, d, T.fone);
      }
    }
  }
  const int fs = T.fs;
  std::vector<int64_t> &f = T.w;
  std::vector<int32_t> &pos = T.pos;
  const double cutoff = 0.002 * (double)T.fone;
  int min_fs = 0;
  for (int i = dst_n - 1; i >= 0; i--) {
    int64_t *r = &f[(size_t)i * fs], cut = 0;
    int mn = fs;
    for (int j = 0; j < fs; j++) {
      cut += std::llabs(r[0]);
      if ((double)cut > cutoff) break;
      if (i < dst_n - 1 && pos[i] >= pos[i + 1]) break;  // positions stay monotonic
      for (int k = 1; k < fs; k++) r[k - 1] = r[k];
      r[fs - 1] = 0;
      pos[i]++;
    }
    cut = 0;
    for (int j = fs - 1; j > 0; j--) {
      cut += std::llabs(r[j]);
      if ((double)cut > cutoff) break;
      mn--;
    }
    min_fs = std::max(min_fs, mn);
  }
  out->taps = min_fs;
  out->first.assign(dst_n, 0);
  out->coef.assign((size_t)dst_n * min_fs, 0);
  std::vector<int64_t> t(min_fs);
  for (int i = 0; i < dst_n; i++) {
    for (int j = 0; j < min_fs; j++) t[j] = f[(size_t)i * fs + j];
    if (pos[i] < 0) {
      for (int j = 1; j < min_fs; j++) {
        const int left = std::max(j + pos[i], 0);
        t[left] += t[j];
        t[j] = 0;
      }
      pos[i] = 0;
    }
    if (pos[i] + min_fs > src_n) {
      const int shift = pos[i] + std::min(min_fs - src_n, 0);
      int64_t acc = 0;
      for (int j = min_fs - 1; j >= 0; j--)
        if (pos[i] + j >= src_n) { acc += t[j]; t[j] = 0; }
      for (int j = min_fs - 1; j >= 0; j--) t[j] = j < shift ? 0 : t[j - shift];
      pos[i] -= shift;
      t[src_n - 1 - pos[i]] += acc;
    }
    int64_t sum = 0, err = 0;
    for (int j = 0; j < min_fs; j++) sum += t[j];
    sum = (sum + one / 2) / one;
    if (!sum) sum = 1;
    for (int j = 0; j < min_fs; j++) {
      const int64_t v = t[j] + err;
      const int64_t q = v >= 0 ? (v + (sum >> 1)) / sum : (v - (sum >> 1)) / sum;
      out->coef[(size_t)i * min_fs + j] = (int16_t)q;
      err = v - q * sum;
    }
    out->first[i] = pos[i];
  }
  return true;
}

}  // namespace pe
