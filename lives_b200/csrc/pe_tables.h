// pe_tables.h -- host-side construction of every lookup table the kernels use.
// Product code (no oracle involvement): the formulas follow the reference's table builders
// (file:line cited per function in pe_tables.cpp) and are evaluated in the same C types
// (double for the colour matrices, float32 + powf for gamma) so that the tables are bit-identical.
#pragma once
#include <cstdint>
#include <vector>

namespace pe {

// index of the 14 conversion tables, the order struct _conv_array lists them (colourspace.h:65-82)
enum ConvTab { Y_R = 0, Y_G, Y_B, CB_R, CB_G, CB_B, CR_R, CR_G, CR_B, RGB_Y, R_CR, G_CB, G_CR, B_CB, N_CONVTAB };

struct ConvTables {
  int32_t t[N_CONVTAB][256];
  int min_y, max_y, min_uv, max_uv;
};

void build_conv_tables(int clamping, int subspace, ConvTables *out);

// the "float - experimental" BT.709 tables of init_YUV_to_RGB_tables (colourspace.c:1040-1104): out[0..4] = RGBf_Y, Rf_Cr, Gf_Cb, Gf_Cr,
// Bf_Cb, evaluated in double and stored as float32 like the reference's assignments (quirks kept: see pe_tables.cpp)
void build_float_yuv_tables(int clamping, float out[5][256]);

// create_gamma_lut8 / create_gamma_lut (colourspace.c:655 / :738); return false when the reference returns NULL
bool build_gamma_lut8(double fileg, int gamma_from, int gamma_to, double screen_gamma, uint8_t out[256]);
bool build_gamma_lut16(double fileg, int gamma_from, int gamma_to, double screen_gamma, uint16_t *out /*65536*/);

// init_unal (colourspace.c:1141): which = 0 unal 1 al 2 unalcy 3 alcy 4 unalcuv 5 alcuv, as uint8
void build_premult_table(int which, uint8_t *out /*65536*/);

// init_average (colourspace.c:190): cavgc (clamped) / cavgu, [x][y] as uint8
void build_avg_table(bool clamped, uint8_t *out /*65536*/);
// avg_chroma(x, y) in closed form: both tables of init_average depend on s = x + y only,
//   table[x][y] = clamp(((s * A + B) * M) >> 32, lo, hi)
// clamped: floor((1785 (s - 256) + 128 * 3904) / 3904) (the float expression of colourspace.c:208 is never closer than 1 / 3904 to an
// integer except at s = 256, where it is exact), limits 16 / 240; unclamped: s >> 1.  avg_form_matches() compares the form with the
// table entry by entry (the engine refuses to use it otherwise).
struct AvgForm {
  uint32_t A, B, M;
  int lo, hi;
};
inline AvgForm avg_form(bool clamped) {
  return clamped ? AvgForm{1785u, 42752u, 1100154u, 16, 240} : AvgForm{1u, 0u, 0x80000000u, 0, 255};
}
bool avg_form_matches(bool clamped, const uint8_t *table);

// init_YUV_to_YUV_tables (colourspace.c:1108): which = 0 Yclamped_to_Yunclamped 1 UVclamped_to_UVunclamped 2 Yunclamped_to_Yclamped
// 3 UVunclamped_to_UVclamped
void build_yy_table(int which, uint8_t out[256]);

// calc_luma tables of libweed/weed-plugin-utils.c:881-886 (16.16, SCALE_FACTOR 65536)
void build_plugin_luma_tables(int32_t yr[256], int32_t yg[256], int32_t yb[256]);

// Resize filter bank (our published contract, DESIGN.md "resize"): taps per output sample, first source
// index and fixed-point coefficients summing to 1 << shift_bits.
struct ResizeFilter {
  int taps = 0;
  std::vector<int32_t> first;  // [dst_n]
  std::vector<int16_t> coef;   // [dst_n * taps]
  bool nonneg() const { for (int16_t c : coef) if (c < 0) return false; return true; }
  int fast_taps() const { return nonneg() ? taps : 1 << 20; }  // what the <= 4-tap kernels may be asked about
};
bool build_resize_filter(int src_n, int dst_n, int shift_bits, ResizeFilter *out);
// libswscale's coefficient recipes for the same two passes (the default, pe_engine_set_resize_recipe): one per flag
// resize_layer_full passes (src/colourspace.c:14991-14997)
enum SwsKind { SWS_KIND_BILINEAR = 1, SWS_KIND_BICUBIC = 2, SWS_KIND_LANCZOS = 3, SWS_KIND_FAST_V = 4, SWS_KIND_FAST_H = 5 };
bool build_resize_filter_sws(int src_n, int dst_n, int shift_bits, ResizeFilter *out, SwsKind kind = SWS_KIND_BILINEAR);

}  // namespace pe
