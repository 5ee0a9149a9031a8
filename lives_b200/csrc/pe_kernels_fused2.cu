// pe_kernels_fused2.cu -- the fast fused chain for the case the headline workload is in:
//     planar 4:2:x fg -> RGBA  |  letterbox with NO horizontal scaling (inner_w == fg width), vertical filter of <= 4 taps
//     |  scalar alpha-over a RGBA32 bg  |  optional 8-bit gamma LUT          (everything else: k_fused, pe_kernels_fused.cu)
//
// Same arithmetic, bit for bit, as k_fused and as the unfused ops.  What differs is the data path:
//   * one CTA (128 threads, 4-5 resident per SM) owns an output tile of 128 x tile_h pixels and stages the raw Y / U / V bytes
//     the tile needs in shared memory with 32- / 16-bit loads (the chroma halo columns carry the reference's edge semantics,
//     so the arithmetic below never looks at a frame edge);
//   * conversion works on 4 x 4 pixel units (two of the reference's row pairs): the chroma sums of colourspace.c:3440-3549
//     are shared inside the unit, and the converted bytes are written PLANAR and COLUMN-MAJOR -- one 32-bit word = one channel
//     of 4 vertically adjacent pixels -- so that
//   * the vertical filter is two DP2A instructions per channel and pixel (16-bit coefficient pairs x 4 byte taps), fed by a
//     funnel shift over two such words; (sum c12 * pix + 2^11) >> 12 equals the two-pass contract of k_resize_h / k_resize_v
//     exactly when the horizontal pass is the identity (pix * 16384 >> 7 = pix * 128);
//   * alpha = k / 256 is blended in integers ((bg * (256 - k) + fg * k) >> 8 is exactly trunc(bg * (1 - a) + fg * a) in double
//     for such alpha: every product and the sum are exact); any other alpha goes through the 64 KB [bg][fg] table in shared
//     memory, as in k_fused.
// Bank-conflict notes: converted tile column stride is 17 words (odd), stage-1 lanes are laid out 8 column-quads x 4 row-quads.
#include "pe_device.cuh"
#include "pe_kernels.h"

namespace pe {

namespace {

#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

constexpr int F2_NT = 128;
constexpr int F2_TW = 128;                 // output tile width
constexpr int F2_MAXVR = 60;               // max virtual source rows (first .. first + 3 of the last row) per tile
constexpr int F2_CW = 17;                  // words per column of the converted tile (68 rows)
constexpr int F2_CCOLS = F2_TW + 4;        // columns of the converted tile (alignment slack of one quad)
constexpr int F2_YROWS = 72;               // raw luma rows
constexpr int F2_YRS = 136;                // raw luma row stride in bytes (34 words, == 2 mod 8)
constexpr int F2_CROWS = 72;               // raw chroma rows (4:2:2: one per luma row)
constexpr int F2_CRS = 72;                 // raw chroma row stride in bytes: [1] left halo, [2 ..] interior, then right halo
constexpr int F2_MAXTH = 48;               // max output rows per tile

constexpr int OFF_TAB = 0;                                   // int32 [5][256]
constexpr int OFF_LUT = OFF_TAB + 5 * 1024;                  // u8 [256]
constexpr int OFF_ROW = OFF_LUT + 256;                       // int32 [F2_MAXTH][4]: pos, a0, a1, -
constexpr int OFF_VF = OFF_ROW + F2_MAXTH * 16;              // u8 [F2_CROWS] true column 0 of V per chroma row (+ pad)
constexpr int OFF_Y = OFF_VF + 80;
constexpr int OFF_U = OFF_Y + F2_YROWS * F2_YRS;
constexpr int OFF_V = OFF_U + F2_CROWS * F2_CRS;
constexpr int OFF_C = OFF_V + F2_CROWS * F2_CRS;             // u32 [3][F2_CCOLS][F2_CW]
constexpr int OFF_OVER = OFF_C + 3 * F2_CCOLS * F2_CW * 4;   // u8 [65536] (table blend only)
constexpr int F2_SMEM_ARITH = OFF_OVER;
constexpr int F2_SMEM_TABLE = OFF_OVER + 65536;

struct Fused2Params {
  const FusedArgs *frames;
  int nframes, tiles_x, tiles_y, tile_h;
  int blend_a, blend_ia;     // arithmetic blend: weights of fg / bg, sum 256
  const uint8_t *lut8;       // optional gamma LUT applied after the blend (nullptr: none)
};

__device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }

// chroma sample with the reference's one-past-row read (see k_yuv_planar_to_rgb, pe_kernels_yuv.cu)
__device__ __forceinline__ uint32_t chroma_edge(const uint8_t *__restrict__ p, int stride, int r, int c, int cw, int ch) {
  if (c >= cw) c = (cw < stride || r + 1 < ch) ? cw : cw - 1;
  return p[(long long)stride * r + c];
}

__device__ __forceinline__ uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// quad index of a luma row: 4:2:0 groups the reference's row pairs (1,2)(3,4) | (5,6)(7,8) ..., row 0 sits alone in quad 0
__device__ __forceinline__ int quad_of(int row, int is422) { return is422 ? (row >> 2) : ((row + 3) >> 2); }
__device__ __forceinline__ int quad_first_row(int g, int is422) { return is422 ? 4 * g : 4 * g - 3; }

// yuv2rgb_int / xyuv2rgb (colourspace.c:2345-2356) through the shared-memory tables; byte results
__device__ __forceinline__ void px_rgb(const int32_t *__restrict__ t, int y, int u, int v, uint32_t &r, uint32_t &g, uint32_t &b) {
  const int yy = t[y];
  r = (uint32_t)sat8((yy + t[256 + v]) >> 16);
  g = (uint32_t)sat8((yy + t[512 + u] + t[768 + v]) >> 16);
  b = (uint32_t)sat8((yy + t[1024 + u]) >> 16);
}

template <int MODE>  // 0: arithmetic blend (alpha = k / 256), 1: [bg][fg] table blend
__global__ void __launch_bounds__(F2_NT, MODE == 0 ? 4 : 2) k_fused2(const Fused2Params P) {
  extern __shared__ __align__(16) uint8_t smem[];
  int32_t *s_tab = reinterpret_cast<int32_t *>(smem + OFF_TAB);
  uint8_t *s_lut = smem + OFF_LUT;
  int32_t *s_row = reinterpret_cast<int32_t *>(smem + OFF_ROW);
  uint8_t *s_vf = smem + OFF_VF;
  uint8_t *s_y = smem + OFF_Y;
  uint8_t *s_u = smem + OFF_U;
  uint8_t *s_v = smem + OFF_V;
  uint32_t *s_c = reinterpret_cast<uint32_t *>(smem + OFF_C);
  uint8_t *s_over = smem + OFF_OVER;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t *cur_conv = nullptr;
  const uint8_t *cur_over = nullptr;
  const bool has_lut = P.lut8 != nullptr;
  if (has_lut) for (int i = tid; i < 256; i += F2_NT) s_lut[i] = P.lut8[i];

  const long long tiles_per_frame = (long long)P.tiles_x * P.tiles_y, total_tiles = tiles_per_frame * P.nframes;
  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int f = (int)(tile / tiles_per_frame);
    const int t = (int)(tile - (long long)f * tiles_per_frame);
    const FusedArgs &A = P.frames[f];
    const int is422 = A.is_422, fw = A.fw, fh = A.fh;
    const int tx = t % P.tiles_x, ty = t / P.tiles_x;
    const int x0 = tx * F2_TW, y0 = ty * P.tile_h;
    const int x1 = min(x0 + F2_TW, A.ow), y1 = min(y0 + P.tile_h, A.oh);
    // intersection with the inner rectangle, in inner (= source column) coordinates
    const int ix0 = max(x0 - A.ox, 0), ix1 = min(x1 - A.ox, A.iw);
    const int iy0 = max(y0 - A.oy, 0), iy1 = min(y1 - A.oy, A.ih);
    const bool has_inner = ix0 < ix1 && iy0 < iy1;

    __syncthreads();  // the previous tile is done with shared memory
    if (A.conv.t != cur_conv) {
      for (int i = tid; i < 5 * 256; i += F2_NT) s_tab[i] = A.conv.t[9 * 256 + i];
      cur_conv = A.conv.t;
    }
    if (MODE == 1 && A.over_table != cur_over) {
      for (int i = tid; i < 4096; i += F2_NT) reinterpret_cast<uint4 *>(s_over)[i] = reinterpret_cast<const uint4 *>(A.over_table)[i];
      cur_over = A.over_table;
    }

    int base_g = 0, cq0 = 0;
    if (has_inner) {
      const int vr0 = A.fy.first[iy0], vr1 = A.fy.first[iy1 - 1] + 3;     // virtual source rows of the tile
      const int ar0 = min(max(vr0, 0), fh - 1), ar1 = min(max(vr1, 0), fh - 1);
      base_g = quad_of(vr0, is422);                                       // (vr0 >= -3 always: first >= -support)
      const int ga = quad_of(ar0, is422), gb = quad_of(ar1, is422);       // quads that hold real rows
      cq0 = ix0 >> 2;
      const int cq1 = (ix1 - 1) >> 2, ncq = cq1 - cq0 + 1;
      const int yrow0 = max(quad_first_row(ga, is422), 0), yrow1 = min(quad_first_row(gb, is422) + 3, fh - 1);
      const int nyr = yrow1 - yrow0 + 1;
      const int cw = A.fg.cw, ch = A.fg.ch;
      const int crow0 = is422 ? yrow0 : max(2 * ga - 2, 0), crow1 = is422 ? yrow1 : min(2 * gb, ch - 1);
      const int ncr = crow1 - crow0 + 1;

      // ---- per-row filter data: window position inside the converted tile, coefficient pairs
      for (int i = tid; i < iy1 - iy0; i += F2_NT) {
        const int iy = iy0 + i;
        const int first = A.fy.first[iy];
        const int16_t *c = A.fy.coef + (long long)iy * A.fy.taps;
        uint32_t cc[4] = {0, 0, 0, 0};
        for (int k = 0; k < A.fy.taps; k++) cc[k] = (uint16_t)c[k];
        s_row[4 * i + 0] = first + (is422 ? 0 : 3) - 4 * base_g;
        s_row[4 * i + 1] = (int)(cc[0] | (cc[1] << 16));
        s_row[4 * i + 2] = (int)(cc[2] | (cc[3] << 16));
      }
      // ---- stage raw luma: rows yrow0..yrow1, 32-bit words cq0..cq1
      {
        const int nw = ncq;
        for (int i = tid; i < nyr * nw; i += F2_NT) {
          const int r = i / nw, w = i - r * nw;
          const uint32_t v = ld_stream_u32(A.fg.y + (long long)A.fg.rs_y * (yrow0 + r) + 4 * (cq0 + w));
          *reinterpret_cast<uint32_t *>(s_y + r * F2_YRS + 4 * w) = v;
        }
      }
      // ---- stage raw chroma: interior columns 2*cq0 .. 2*cq1+1 as 16-bit pairs, the two halo columns with the reference's
      //      edge rules: column -1 replicates column 0 (last = this at the start of a row); column >= cw reads the byte at
      //      plane[stride * r + cw] -- padding or the first sample of the next row -- except on the last chroma row of a
      //      plane without padding, where it is the replicated edge sample (colourspace.c:3508-3512, DESIGN.md "edge read");
      //      4:2:2 with ref_quirks: columns <= 0 take column 0 of chroma row (r >> 1) (the seed slip, :3600)
      {
        const int npair = ncq;  // 16-bit pairs per row
        const int per_row = npair + 2;
        for (int i = tid; i < ncr * per_row * 2; i += F2_NT) {
          const int plane = i / (ncr * per_row);
          const int j = i - plane * (ncr * per_row);
          const int r = j / per_row, e = j - r * per_row;
          const int cr = crow0 + r;
          const uint8_t *src = plane ? A.fg.v : A.fg.u;
          const int rs = plane ? A.fg.rs_v : A.fg.rs_u;
          uint8_t *dst = (plane ? s_v : s_u) + r * F2_CRS;
          const bool seed = is422 && A.quirks;
          if (e < npair) {
            const int c = 2 * (cq0 + e);  // even column; c + 1 <= cw may be the one-past column
            uint32_t b0, b1;
            if (c + 1 < cw) {
              const uint32_t v = *reinterpret_cast<const uint16_t *>(src + (long long)rs * cr + c);
              b0 = v & 0xFFu; b1 = v >> 8;
            } else {
              b0 = chroma_edge(src, rs, cr, c, cw, ch);
              b1 = chroma_edge(src, rs, cr, c + 1, cw, ch);
            }
            if (seed && c == 0) b0 = src[(long long)rs * (cr >> 1)];
            *reinterpret_cast<uint16_t *>(dst + 2 + 2 * e) = (uint16_t)(b0 | (b1 << 8));
          } else if (e == npair) {  // left halo
            const int c = 2 * cq0 - 1;
            uint32_t b;
            if (c < 0) b = seed ? src[(long long)rs * (cr >> 1)] : src[(long long)rs * cr];
            else b = src[(long long)rs * cr + c];
            dst[1] = (uint8_t)b;
          } else {                  // right halo
            const int c = 2 * cq1 + 2;
            dst[2 + 2 * npair] = (uint8_t)chroma_edge(src, rs, cr, c, cw, ch);
          }
        }
        for (int r = tid; r < ncr; r += F2_NT) s_vf[r] = A.fg.v[(long long)A.fg.rs_v * (crow0 + r)];
      }
      __syncthreads();

      // ---- stage 1: convert 4 x 4 units.  Warp lanes: 8 column quads x 4 row quads.
      {
        const int ngr = gb - ga + 1;
        const int cq_groups = (ncq + 7) >> 3, rq_groups = (ngr + 3) >> 2;
        const int lo = A.clamped ? 16 : 0, hi = A.clamped ? 240 : 255;
        for (int wt = warp; wt < cq_groups * rq_groups; wt += F2_NT / 32) {
          const int cg = wt % cq_groups, rg = wt / cq_groups;
          const int q = cg * 8 + (lane & 7), gl = rg * 4 + (lane >> 3);
          if (q >= ncq || gl >= ngr) continue;
          const int g = ga + gl;
          const int jc0 = 2 * (cq0 + q);  // absolute chroma column of the unit's first pair
          uint32_t acc[4][3];
#pragma unroll
          for (int k = 0; k < 4; k++) acc[k][0] = acc[k][1] = acc[k][2] = 0;
          // chroma words: columns jc0-1 .. jc0+2 of one staged chroma row
          auto cword = [&](const uint8_t *pl, int cr) -> uint32_t {
            const uint8_t *rowp = pl + (cr - crow0) * F2_CRS;
            const int idx = 2 * q + 1;  // byte index of column jc0 - 1
            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(rowp + (idx & ~3));
            const uint32_t w1 = *reinterpret_cast<const uint32_t *>(rowp + (idx & ~3) + 4);
            return __funnelshift_r(w0, w1, 8 * (idx & 3));
          };
          auto yword = [&](int row) -> uint32_t { return *reinterpret_cast<const uint32_t *>(s_y + (row - yrow0) * F2_YRS + 4 * q); };
          // a single row: horizontal average only (row 0, an even frame's last row, every 4:2:2 row)
          auto do_single = [&](int row, int cr, int bytepos) {
            const uint32_t yw = yword(row), uw = cword(s_u, cr), vw = cword(s_v, cr);
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int p = k >> 1;
              const int ua = byte_of(uw, p + 1), va = byte_of(vw, p + 1);
              const int ub = (k & 1) ? byte_of(uw, p + 2) : byte_of(uw, p), vb = (k & 1) ? byte_of(vw, p + 2) : byte_of(vw, p);
              const int u = clamp_i((ua + ub) >> 1, lo, hi), v = clamp_i((va + vb) >> 1, lo, hi);
              uint32_t r, gg, b;
              px_rgb(s_tab, byte_of(yw, k), u, v, r, gg, b);
              acc[k][0] |= r << (8 * bytepos); acc[k][1] |= gg << (8 * bytepos); acc[k][2] |= b << (8 * bytepos);
            }
          };
          // an interior row pair (colourspace.c:3440-3549)
          auto do_pair = [&](int row_a, int cr_a, int bytepos) {
            const int cr_b = cr_a + 1;
            const uint32_t ya = yword(row_a), yb = yword(row_a + 1);
            const uint32_t u1w = cword(s_u, cr_a), u2w = cword(s_u, cr_b), v1w = cword(s_v, cr_a), v2w = cword(s_v, cr_b);
            const int v2_first = s_vf[cr_b - crow0];
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int p = k >> 1;
              int u1, u2, v1, v2;
              if (k & 1) {
                u1 = byte_of(u1w, p + 1) + byte_of(u1w, p + 2); u2 = byte_of(u2w, p + 1) + byte_of(u2w, p + 2);
                v1 = byte_of(v1w, p + 1) + byte_of(v1w, p + 2); v2 = byte_of(v2w, p + 1) + byte_of(v2w, p + 2);
              } else {
                u1 = byte_of(u1w, p + 1) + byte_of(u1w, p); u2 = byte_of(u2w, p + 1) + byte_of(u2w, p);
                v1 = byte_of(v1w, p + 1) + byte_of(v1w, p); v2 = byte_of(v2w, p + 1) + byte_of(v2w, p);
                if (A.quirks) {
                  u2 = u1;                                                          // colourspace.c:3461
                  if (jc0 + p > 0) v1 = byte_of(v1w, p + 1) + byte_of(v2w, p);      // :3544
                  v2 = byte_of(v2w, p + 1) + v2_first;                              // last_v2 never advanced
                }
              }
              int u3, u4, v3, v4;
              if (!A.low_quality) {
                u3 = third_round(u1 + (u2 >> 1)); u4 = third_round((u1 >> 1) + u2);
                v3 = third_round(v1 + (v2 >> 1)); v4 = third_round((v1 >> 1) + v2);
              } else {
                u3 = u1 >> 1; u4 = u2 >> 1; v3 = v1 >> 1; v4 = v2 >> 1;
              }
              u3 = clamp_i(u3, lo, hi); u4 = clamp_i(u4, lo, hi); v3 = clamp_i(v3, lo, hi); v4 = clamp_i(v4, lo, hi);
              uint32_t r, gg, b;
              px_rgb(s_tab, byte_of(ya, k), u3, v3, r, gg, b);
              acc[k][0] |= r << (8 * bytepos); acc[k][1] |= gg << (8 * bytepos); acc[k][2] |= b << (8 * bytepos);
              px_rgb(s_tab, byte_of(yb, k), u4, v4, r, gg, b);
              acc[k][0] |= r << (8 * bytepos + 8); acc[k][1] |= gg << (8 * bytepos + 8); acc[k][2] |= b << (8 * bytepos + 8);
            }
          };
          if (is422) {
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
              const int row = 4 * g + rr;
              if (row < fh) do_single(row, row, rr);
            }
          } else {
            // quad g holds rows 4g-3 .. 4g: pairs (4g-3, 4g-2) and (4g-1, 4g); chroma rows (2g-2, 2g-1) and (2g-1, 2g)
#pragma unroll
            for (int pp = 0; pp < 2; pp++) {
              const int row_a = 4 * g - 3 + 2 * pp, cr_a = 2 * g - 2 + pp;
              if (row_a + 1 == 0) do_single(0, 0, 2 * pp + 1);                       // row 0 (lower half of the "pair" -1, 0)
              else if (row_a >= 0 && row_a + 1 < fh) do_pair(row_a, cr_a, 2 * pp);
              else if (row_a == fh - 1 && row_a >= 0) do_single(row_a, ch - 1, 2 * pp);  // even height: last row alone
            }
          }
          uint32_t *dst = s_c + (4 * q) * F2_CW + (g - base_g);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            dst[k * F2_CW] = acc[k][0];
            dst[F2_CCOLS * F2_CW + k * F2_CW] = acc[k][1];
            dst[2 * F2_CCOLS * F2_CW + k * F2_CW] = acc[k][2];
          }
        }
      }
      __syncthreads();
      // ---- edge fix-up: virtual rows outside the frame replicate row 0 / row fh-1 (source indices are clamped in the contract)
      if (vr0 < 0 || vr1 > fh - 1) {
        const int shift = is422 ? 0 : 3;
        const int ncols = 4 * ncq;
        uint8_t *cb = reinterpret_cast<uint8_t *>(s_c);
        for (int i = tid; i < 3 * ncols; i += F2_NT) {
          const int chn = i / ncols, col = i - chn * ncols;
          uint8_t *colp = cb + ((chn * F2_CCOLS + col) * F2_CW) * 4;
          if (vr0 < 0) {
            const uint8_t v = colp[0 + shift - 4 * base_g];
            for (int r = vr0; r < 0; r++) colp[r + shift - 4 * base_g] = v;
          }
          if (vr1 > fh - 1) {
            const uint8_t v = colp[fh - 1 + shift - 4 * base_g];
            for (int r = fh; r <= vr1; r++) colp[r + shift - 4 * base_g] = v;
          }
        }
        __syncthreads();
      }
    }

    // ---- stage 3: vertical filter + letterbox + alpha-over (+ gamma), one thread = one column x 4 consecutive rows
    {
      const int th = y1 - y0;
      const int nquads = (th + 3) >> 2;
      for (int task = warp; task < nquads * (F2_TW / 32); task += F2_NT / 32) {
        const int qd = task / (F2_TW / 32), cwp = task - qd * (F2_TW / 32);
        const int x = x0 + cwp * 32 + lane;
        if (x >= A.ow) continue;
        const int ix = x - A.ox;
        const bool col_in = has_inner && ix >= ix0 && ix < ix1;
        const int lcol = ix - 4 * cq0;
        uint32_t bgw[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int oy = y0 + qd * 4 + r;
          bgw[r] = oy < y1 ? ld_stream_u32(A.bg.p + (long long)A.bg.rs * oy + 4ll * x) : 0u;
        }
        // cached words of the converted column (per channel): positions wi, wi + 1
        uint32_t wl[3] = {0, 0, 0}, wh[3] = {0, 0, 0};
        int cur = -100;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int oy = y0 + qd * 4 + r;
          if (oy >= y1) break;
          const int iy = oy - A.oy;
          uint32_t fr = 0, fg_ = 0, fb = 0;  // letterbox border: black (blank_pixel, colourspace.c:11169)
          if (col_in && iy >= iy0 && iy < iy1) {
            const int pos = s_row[4 * (iy - iy0)];
            const uint32_t a0 = (uint32_t)s_row[4 * (iy - iy0) + 1], a1 = (uint32_t)s_row[4 * (iy - iy0) + 2];
            const int wi = pos >> 2, sh = 8 * (pos & 3);
            if (wi != cur) {  // warp-uniform
#pragma unroll
              for (int c = 0; c < 3; c++) {
                const uint32_t *colp = s_c + (c * F2_CCOLS + lcol) * F2_CW;
                wl[c] = (wi == cur + 1) ? wh[c] : colp[wi];
                wh[c] = colp[wi + 1];
              }
              cur = wi;
            }
            const uint32_t b0 = __funnelshift_r(wl[0], wh[0], sh), b1 = __funnelshift_r(wl[1], wh[1], sh),
                           b2 = __funnelshift_r(wl[2], wh[2], sh);
            fr = dp2a_hi(a1, b0, dp2a_lo(a0, b0, 2048u)) >> 12;
            fg_ = dp2a_hi(a1, b1, dp2a_lo(a0, b1, 2048u)) >> 12;
            fb = dp2a_hi(a1, b2, dp2a_lo(a0, b2, 2048u)) >> 12;
          }
          const uint32_t b = bgw[r];
          uint32_t o0, o1, o2;
          if (MODE == 0) {
            const uint32_t ka = (uint32_t)P.blend_a, kia = (uint32_t)P.blend_ia;
            o0 = ((b & 0xFFu) * kia + fr * ka) >> 8;
            o1 = (((b >> 8) & 0xFFu) * kia + fg_ * ka) >> 8;
            o2 = (((b >> 16) & 0xFFu) * kia + fb * ka) >> 8;
            if (has_lut) { o0 = s_lut[o0]; o1 = s_lut[o1]; o2 = s_lut[o2]; }
          } else {  // the table already contains the gamma LUT (launch_over_table)
            o0 = s_over[((b & 0xFFu) << 8) | fr];
            o1 = s_over[(((b >> 8) & 0xFFu) << 8) | fg_];
            o2 = s_over[(((b >> 16) & 0xFFu) << 8) | fb];
          }
          st_stream_u32(A.out.p + (long long)A.out.rs * oy + 4ll * x, o0 | (o1 << 8) | (o2 << 16) | 0xFF000000u);
        }
      }
    }
  }
}

}  // namespace

// Can the fast kernel take this job?  (checked per launch by the engine; everything else goes to k_fused)
bool fused2_supported(const FusedArgs &a, int fy_taps, int max_virtual_rows_per_tile_h_probe) {
  (void)max_virtual_rows_per_tile_h_probe;
  if (a.iw != a.fw) return false;                        // horizontal pass must be the identity
  if (fy_taps > 4) return false;
  if (a.fw & 3) return false;                            // whole 32-bit luma words
  if (((uintptr_t)a.fg.y & 3) || (a.fg.rs_y & 3)) return false;
  if (((uintptr_t)a.fg.u & 1) || ((uintptr_t)a.fg.v & 1) || (a.fg.rs_u & 1) || (a.fg.rs_v & 1)) return false;
  if (((uintptr_t)a.bg.p & 3) || (a.bg.rs & 3) || ((uintptr_t)a.out.p & 3) || (a.out.rs & 3)) return false;
  if (a.fh < 2) return false;
  return true;
}

int fused2_max_virtual_rows() { return F2_MAXVR; }
int fused2_max_tile_h() { return F2_MAXTH; }

// blend_a < 0: table blend (FusedArgs::over_table, gamma folded in); else arithmetic blend with weights blend_a / 256 - blend_a
cudaError_t launch_fused2_dev(const Launch &L, const FusedArgs *frames_dev, int nframes, int ow, int oh, int tile_h, int blend_a,
                              const uint8_t *lut8_dev) {
  Fused2Params P;
  P.frames = frames_dev; P.nframes = nframes;
  P.tiles_x = (ow + F2_TW - 1) / F2_TW; P.tiles_y = (oh + tile_h - 1) / tile_h; P.tile_h = tile_h;
  P.blend_a = blend_a; P.blend_ia = 256 - blend_a; P.lut8 = lut8_dev;
  const long long total = (long long)P.tiles_x * P.tiles_y * nframes;
  static bool attr0 = false, attr1 = false;
  if (blend_a >= 0) {
    if (!attr0) {
      cudaError_t e = cudaFuncSetAttribute(k_fused2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_ARITH);
      if (e != cudaSuccess) return e;
      attr0 = true;
    }
    const int grid = (int)(total < (long long)L.sm_count * 4 ? total : (long long)L.sm_count * 4);
    k_fused2<0><<<grid, F2_NT, F2_SMEM_ARITH, L.stream>>>(P);
  } else {
    if (!attr1) {
      cudaError_t e = cudaFuncSetAttribute(k_fused2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_TABLE);
      if (e != cudaSuccess) return e;
      attr1 = true;
    }
    P.lut8 = nullptr;
    const int grid = (int)(total < (long long)L.sm_count * 2 ? total : (long long)L.sm_count * 2);
    k_fused2<1><<<grid, F2_NT, F2_SMEM_TABLE, L.stream>>>(P);
  }
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

}  // namespace pe
