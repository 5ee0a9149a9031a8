// pe_kernels_fx2.cu -- SURVEY 8f rank 3, second batch of effect plugins:
//   k_softlight   lives-plugins/weed-plugins/softlight.c softlight_process :62 (luma plane of a planar YUV frame)
//   k_select      layout_blends.c common_process :19 ("triple split") and multi_transitions.c common_process :85
//                 ("iris rectangle", "iris circle", "4 way split", "dissolve"): every destination pixel is a copy of a pixel of
//                 in1 or in2 (or a constant colour); what differs is the per-pixel predicate.
// All byte work bounded by HBM: 1 read + 1 write per luma sample (softlight), 2 reads + 1 write per pixel (the selectors read both
// clips with full-width vector loads: a predicated load would save traffic only where whole 32-byte sectors fall on one side).
#include "pe_device.cuh"
#include "pe_kernels.h"

namespace pe {

namespace {

constexpr int kBlock = 256;

inline int grid_for(const Launch &L, long long work_items, int per_sm = 8) {
  long long blocks = (work_items + kBlock - 1) / kBlock;
  long long cap = (long long)L.sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

// ---- softlight -------------------------------------------------------------------------------------------------------------------
// floor(sqrt(n)) for n < 2^24 (softlight.c sqrti :33): n is exact in float, the rounded root is off by at most one
__device__ __forceinline__ uint32_t isqrt24(uint32_t n) {
  uint32_t r = (uint32_t)__fsqrt_rn((float)n);
  if (r * r > n) r--;
  else if ((r + 1) * (r + 1) <= n) r++;
  return r;
}

// samples x0 - 1 .. x0 + 4 of one row (x0 a multiple of 4); columns outside [0, width) are never used by the caller
template <bool WORDS>
__device__ __forceinline__ void fetch6(const uint8_t *__restrict__ row, int x0, int width, int (&s)[6]) {
  if (WORDS) {
    const uint32_t wc = __ldg(reinterpret_cast<const uint32_t *>(row + x0));
    const uint32_t wl = x0 > 0 ? __ldg(reinterpret_cast<const uint32_t *>(row + x0 - 4)) : 0u;
    const uint32_t wr = x0 + 4 < width ? __ldg(reinterpret_cast<const uint32_t *>(row + x0 + 4)) : 0u;
    s[0] = (int)(wl >> 24);
    s[1] = (int)byte_of(wc, 0); s[2] = (int)byte_of(wc, 1); s[3] = (int)byte_of(wc, 2); s[4] = (int)byte_of(wc, 3);
    s[5] = (int)(wr & 0xFFu);
  } else {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int x = x0 - 1 + k;
      s[k] = (x >= 0 && x < width) ? (int)__ldg(row + x) : 0;
    }
  }
}

struct SoftlightParams {
  const uint8_t *src;
  uint8_t *dst;
  int irow, orow, width, height, ymin, ymax;
};

// one thread = 4 adjacent samples of one row; WORDS: source rows are 4-byte aligned and a whole word may be read at the row's end
template <bool WORDS>
__global__ void __launch_bounds__(kBlock) k_softlight(const SoftlightParams P) {
  const int groups = (P.width + 3) >> 2;
  const long long total = (long long)groups * P.height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int y = (int)(it / groups), x0 = (int)(it - (long long)y * groups) << 2;
    const uint8_t *rm = P.src + (long long)P.irow * y;
    uint8_t *d = P.dst + (long long)P.orow * y + x0;
    const int n = min(4, P.width - x0);
    int up[6], mid[6], dn[6];
    fetch6<WORDS>(rm, x0, P.width, mid);
    uint32_t o[4];
    if (y == 0 || y == P.height - 1) {
#pragma unroll
      for (int i = 0; i < 4; i++) o[i] = (uint32_t)mid[i + 1];
    } else {
      fetch6<WORDS>(rm - P.irow, x0, P.width, up);
      fetch6<WORDS>(rm + P.irow, x0, P.width, dn);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int k = i + 1, x = x0 + i;
        // softlight.c:114-118, term for term (the last term of row0 is lower-right minus lower-LEFT, the last of row1 a SUM)
        const int row0 = (dn[k - 1] - up[k - 1]) + ((dn[k] - up[k]) * 2) + (dn[k + 1] - dn[k - 1]);
        const int row1 = (up[k + 1] - up[k - 1]) + ((mid[k + 1] - mid[k - 1]) * 2) + (dn[k + 1] + dn[k - 1]);
        int sum = (int)(((3u * isqrt24((uint32_t)(row0 * row0 + row1 * row1)) / 2u) * 384u) >> 8);
        sum = min(max(sum, P.ymin), P.ymax);
        sum = (64 * sum + 192 * mid[k]) >> 8;
        sum = min(max(sum, P.ymin), P.ymax);
        o[i] = (x == 0 || x >= P.width - 1) ? (uint32_t)mid[k] : (uint32_t)sum;
      }
    }
    if (n == 4 && !(reinterpret_cast<uintptr_t>(d) & 3)) *reinterpret_cast<uint32_t *>(d) = pack4(o[0], o[1], o[2], o[3]);
    else
      for (int i = 0; i < n; i++) d[i] = (uint8_t)o[i];
  }
}

// ---- selectors -------------------------------------------------------------------------------------------------------------------
enum { SEL_TSPLIT = 0, SEL_IRIS_RECT = 1, SEL_IRIS_CIRC = 2, SEL_FOURWAY = 3, SEL_DISSOLVE = 4 };

// which source a pixel takes: 0 = in1, 1 = in2, 2 = the constant colour
__device__ __forceinline__ int select_of(const SelectArgs &P, int mode, int x, int y) {
  const int j = x * P.psize;  // the reference's byte offset inside the row
  switch (mode) {
    case SEL_TSPLIT: {  // layout_blends.c:92-99
      const int cc = __ldg(P.colclass + x), rc = __ldg(P.rowclass + y);
      if ((cc & 1) && (rc & 1)) return 1;
      if ((cc & 2) || (rc & 2)) return 0;
      return 2;
    }
    case SEL_IRIS_RECT:  // multi_transitions.c:152-168
      return (j < P.xx || j >= P.row_bytes - P.xx || y < P.yy || y >= P.height - P.yy) ? 0 : 1;
    case SEL_IRIS_CIRC: {  // :169-182, in the -ffast-math form of the compiled plugin (reciprocals computed once, sqrt in double)
      const float xxf = (float)(y - P.ihheight);
      const float yyf = __fmul_rn((float)(j - P.ihwidth), P.inv_psize);
      const float t = __fmul_rn(__fadd_rn(__fmul_rn(yyf, yyf), __fmul_rn(xxf, xxf)), P.inv_maxradsq);
      return __dsqrt_rn((double)t) > (double)P.bf ? 0 : 1;
    }
    case SEL_FOURWAY:  // :183-191
      return (__fmul_rn(fabsf(__fsub_rn((float)y, P.hheight)), P.inv_hh) < P.bf || __fmul_rn(fabsf(__fsub_rn((float)j, P.hwidth)), P.inv_hw) < P.bf ||
              P.bf == 1.f) ? 1 : 0;
    default:  // SEL_DISSOLVE :192-196
      return __ldg(P.mask + (long long)y * P.width + x) < P.bf ? 1 : 0;
  }
}

// one thread = one pixel: ragged widths, unaligned frames, and the displaced src1 reads of the "4 way split"
template <int PS>
__global__ void __launch_bounds__(kBlock) k_select_px(const SelectArgs P, int mode) {
  const long long total = (long long)P.width * P.height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int y = (int)(it / P.width), x = (int)(it - (long long)y * P.width);
    const int c = select_of(P, mode, x, y);
    uint8_t *d = P.d + (long long)P.rsd * y + x * PS;
    if (c == 2) {
      d[0] = (uint8_t)P.colour[0]; d[1] = (uint8_t)P.colour[1]; d[2] = (uint8_t)P.colour[2];
      continue;
    }
    const uint8_t *s;
    if (c == 1) s = P.s2 + (long long)P.rs2 * y + x * PS;
    else {
      s = P.s1 + (long long)P.rs1 * y + x * PS;
      if (mode == SEL_FOURWAY) s += (x * PS > P.ihwidth ? -P.yy : P.yy) + (long long)(y > P.ihheight ? -P.xx : P.xx) * P.rs1;
    }
    if (s == d) continue;  // in place: the pixel keeps in1
#pragma unroll
    for (int k = 0; k < PS; k++) d[k] = __ldg(s + k);
  }
}

// one thread = 4 pixels of aligned frames: 3 words (PS 3) / one 128-bit vector (PS 4) from each clip, per-byte masks pick the source
template <int PS>
__global__ void __launch_bounds__(kBlock) k_select_vec(const SelectArgs P, int mode) {
  const int groups = P.width >> 2;
  const long long total = (long long)groups * P.height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int y = (int)(it / groups), x0 = (int)(it - (long long)y * groups) << 2;
    uint32_t m1[4], m2[4];  // per pixel: all ones when the pixel takes in1 / in2
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const int c = select_of(P, mode, x0 + p, y);
      m1[p] = c == 0 ? 0xFFFFFFFFu : 0u;
      m2[p] = c == 1 ? 0xFFFFFFFFu : 0u;
    }
    const uint8_t *a = P.s1 + (long long)P.rs1 * y + x0 * PS, *b = P.s2 + (long long)P.rs2 * y + x0 * PS;
    uint8_t *d = P.d + (long long)P.rsd * y + x0 * PS;
    if (PS == 4) {
      const uint4 va = ld_u4(a), vb = ld_stream_u4(b);
      uint4 o;
      o.x = (va.x & m1[0]) | (vb.x & m2[0]); o.y = (va.y & m1[1]) | (vb.y & m2[1]);
      o.z = (va.z & m1[2]) | (vb.z & m2[2]); o.w = (va.w & m1[3]) | (vb.w & m2[3]);
      *reinterpret_cast<uint4 *>(d) = o;
    } else {
      const uint32_t *wa = reinterpret_cast<const uint32_t *>(a), *wb = reinterpret_cast<const uint32_t *>(b);
      const uint32_t a0 = wa[0], a1 = wa[1], a2 = wa[2], b0 = ld_stream_u32(wb), b1 = ld_stream_u32(wb + 1), b2 = ld_stream_u32(wb + 2);
      // bytes 0..2 pixel 0, 3..5 pixel 1, 6..8 pixel 2, 9..11 pixel 3
      const uint32_t k1_0 = (m1[0] & 0x00FFFFFFu) | (m1[1] & 0xFF000000u), k1_1 = (m1[1] & 0x0000FFFFu) | (m1[2] & 0xFFFF0000u),
                     k1_2 = (m1[2] & 0x000000FFu) | (m1[3] & 0xFFFFFF00u);
      const uint32_t k2_0 = (m2[0] & 0x00FFFFFFu) | (m2[1] & 0xFF000000u), k2_1 = (m2[1] & 0x0000FFFFu) | (m2[2] & 0xFFFF0000u),
                     k2_2 = (m2[2] & 0x000000FFu) | (m2[3] & 0xFFFFFF00u);
      uint32_t *wd = reinterpret_cast<uint32_t *>(d);
      wd[0] = (a0 & k1_0) | (b0 & k2_0) | (P.colour_words[0] & ~(k1_0 | k2_0));
      wd[1] = (a1 & k1_1) | (b1 & k2_1) | (P.colour_words[1] & ~(k1_1 | k2_1));
      wd[2] = (a2 & k1_2) | (b2 & k2_2) | (P.colour_words[2] & ~(k1_2 | k2_2));
    }
  }
}

}  // namespace

cudaError_t launch_softlight(const Launch &L, CImg src, Img dst, int width, int height, int ymin, int ymax) {
  if (width <= 0 || height <= 0) return cudaSuccess;
  SoftlightParams P{src.p, dst.p, src.rs, dst.rs, width, height, ymin, ymax};
  // word loads: aligned rows, and the word at the end of a row may reach past `width` only inside the rowstride
  const bool words = !(reinterpret_cast<uintptr_t>(src.p) & 3) && !(src.rs & 3) && ((width + 3) & ~3) <= src.rs;
  const int grid = grid_for(L, (long long)((width + 3) >> 2) * height);
  if (words) k_softlight<true><<<grid, kBlock, 0, L.stream>>>(P);
  else k_softlight<false><<<grid, kBlock, 0, L.stream>>>(P);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_select(const Launch &L, int mode, const SelectArgs &a) {
  if (a.width <= 0 || a.height <= 0) return cudaSuccess;
  if (a.psize != 3 && a.psize != 4) return cudaErrorInvalidValue;
  SelectArgs P = a;
  P.row_bytes = a.width * a.psize;
  for (int w = 0; w < 3; w++) {  // the 12-byte period of the constant colour, as words
    uint32_t v = 0;
    for (int b = 0; b < 4; b++) v |= (uint32_t)(a.colour[(4 * w + b) % 3] & 0xFF) << (8 * b);
    P.colour_words[w] = v;
  }
  const uintptr_t bases = reinterpret_cast<uintptr_t>(a.s1) | reinterpret_cast<uintptr_t>(a.s2) | reinterpret_cast<uintptr_t>(a.d);
  const int strides = a.rs1 | a.rs2 | a.rsd;
  const int al = a.psize == 4 ? 15 : 3;
  const bool vec = mode != SEL_FOURWAY && !(a.width & 3) && !(bases & al) && !(strides & al) && (a.psize == 3 || mode != SEL_TSPLIT);
  if (vec) {
    const int grid = grid_for(L, (long long)(a.width >> 2) * a.height);
    if (a.psize == 4) k_select_vec<4><<<grid, kBlock, 0, L.stream>>>(P, mode);
    else k_select_vec<3><<<grid, kBlock, 0, L.stream>>>(P, mode);
  } else {
    const int grid = grid_for(L, (long long)a.width * a.height);
    if (a.psize == 4) k_select_px<4><<<grid, kBlock, 0, L.stream>>>(P, mode);
    else k_select_px<3><<<grid, kBlock, 0, L.stream>>>(P, mode);
  }
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

}  // namespace pe
