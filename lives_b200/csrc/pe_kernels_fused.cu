// pe_kernels_fused.cu -- resize kernels and the fused convert -> letterbox/resize -> alpha-over -> gamma kernel.
//
// Resize arithmetic is OUR contract (the reference hands resizing to libswscale, which is not in its tree):
// separable, swscale-shaped data path
//     horizontal: tmp = min((sum_k coef14[k] * pix[first + k]) >> 7, 32767)            (15-bit intermediate)
//     vertical  : out = clip_u8((sum_k coef12[k] * tmp[first + k] + (1 << 18)) >> 19)
// with source indices clamped to the frame; see pe_tables.cpp build_resize_filter and DESIGN.md.
//
// k_fused computes, per 64 x 16 output tile held by one CTA:
//   1. the source pixels the tile needs, converted YUV4:2:x planar -> RGBA once each into shared memory
//      (same arithmetic as k_yuv_planar_to_rgb),
//   2. the horizontally scaled 15-bit rows in shared memory,
//   3. vertical scale, letterbox placement (black opaque border), alpha-over against the background through
//      the 64 KB [bg][fg] table (which already contains the gamma LUT) and a 128-bit store.
// HBM traffic is therefore the algorithmic minimum: fg planes + bg read once, out written once.
#include <cstdlib>

#include "pe_device.cuh"
#include "pe_kernels.h"

namespace pe {

namespace {

constexpr int kBlock = 256;
#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

inline int grid_for(const Launch &L, long long work_items, int per_sm = 8) {
  long long blocks = (work_items + kBlock - 1) / kBlock;
  long long cap = (long long)L.sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__global__ void __launch_bounds__(kBlock) k_resize_h(const uint8_t *__restrict__ src, int srs, int sw, int sh, int16_t *__restrict__ tmp,
                                                     int dw, int psize, DevFilter fx) {
  const long long total = (long long)dw * sh;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int y = (int)(it / dw), x = (int)(it - (long long)y * dw);
    const uint8_t *row = src + (long long)srs * y;
    const int first = fx.first[x];
    int acc[4] = {0, 0, 0, 0};
    for (int k = 0; k < fx.taps; k++) {
      const int sx = min(max(first + k, 0), sw - 1);
      const int c = fx.coef[(long long)x * fx.taps + k];
      for (int ch = 0; ch < psize; ch++) acc[ch] += c * row[sx * psize + ch];
    }
    for (int ch = 0; ch < psize; ch++) tmp[((long long)y * dw + x) * psize + ch] = (int16_t)min(acc[ch] >> 7, 32767);
  }
}

__global__ void __launch_bounds__(kBlock) k_resize_v(const int16_t *__restrict__ tmp, int sh, uint8_t *__restrict__ dst, int drs, int dw,
                                                     int dh, int psize, DevFilter fy) {
  const int rowlen = dw * psize;
  const long long total = (long long)rowlen * dh;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int y = (int)(it / rowlen), x = (int)(it - (long long)y * rowlen);
    const int first = fy.first[y];
    int acc = 1 << 18;
    for (int k = 0; k < fy.taps; k++) {
      const int sy = min(max(first + k, 0), sh - 1);
      acc += fy.coef[(long long)y * fy.taps + k] * tmp[(long long)sy * rowlen + x];
    }
    dst[(long long)drs * y + x] = (uint8_t)min(max(acc >> 19, 0), 255);
  }
}

// ------------------------------------------------------------------------------------------------------
// k_resize_tile: both passes of the contract in one kernel.  One CTA = one output tile (tw x th):
//   1. the source rectangle the tile needs is staged in shared memory (source indices clamped to the frame),
//   2. horizontal pass into a 15-bit shared-memory intermediate (4 x int16 per pixel),
//   3. vertical pass, 8-bit store.
// Same arithmetic as k_resize_h + k_resize_v (which remain the fallback for scale factors whose source rectangle
// does not fit in shared memory); the intermediate never touches HBM.
// ------------------------------------------------------------------------------------------------------

struct ResizeTileParams {
  const uint8_t *src;
  uint8_t *dst;
  int srs, drs, sw, sh, dw, dh;
  int tw, th;                 // output tile
  int max_rows, max_cols;     // largest source extent of a tile
  int vec_src;                // k_resize_tile4: source rows are 16-byte aligned (128-bit staging loads)
  DevFilter fx, fy;
};

// Shared memory: [raw source rectangle, edge-replicated][tmp: 4 x int16 per (row, out column)][x taps][y taps]
//   * the staged rectangle covers the UNCLAMPED tap range of the tile (virtual rows / columns outside the frame hold the
//     replicated edge pixel), so the two passes below index it without clamping;
//   * the coefficients of the tile's columns / rows are staged once per tile (int16, taps padded to even).
template <int PS>
__global__ void __launch_bounds__(kBlock) k_resize_tile(const ResizeTileParams P) {
  extern __shared__ __align__(16) uint8_t rsm[];
  const int raw_stride = (P.max_cols * PS + 3) & ~3;
  const int tx_taps = P.fx.taps, ty_taps = P.fy.taps;
  uint8_t *s_raw = rsm;                                                    // [max_rows][raw_stride]
  size_t off = ((size_t)P.max_rows * raw_stride + 15) & ~(size_t)15;
  uint2 *s_tmp = reinterpret_cast<uint2 *>(rsm + off);                     // [max_rows][tw]
  off += (size_t)P.max_rows * P.tw * 8;
  int16_t *s_cx = reinterpret_cast<int16_t *>(rsm + off);                  // [tw][tx_taps]
  off += ((size_t)P.tw * tx_taps * 2 + 15) & ~(size_t)15;
  int16_t *s_cy = reinterpret_cast<int16_t *>(rsm + off);                  // [th][ty_taps]
  off += ((size_t)P.th * ty_taps * 2 + 15) & ~(size_t)15;
  int *s_fx = reinterpret_cast<int *>(rsm + off);                          // [tw] first - vc0
  int *s_fy = s_fx + P.tw;                                                 // [th] first - vr0

  const int tiles_x = (P.dw + P.tw - 1) / P.tw;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int x0 = tx * P.tw, y0 = ty * P.th;
  const int x1 = min(x0 + P.tw, P.dw), y1 = min(y0 + P.th, P.dh);
  const int ncol = x1 - x0, nrow = y1 - y0;
  // virtual (unclamped) source range of the tile
  const int vr0 = P.fy.first[y0], vr1 = P.fy.first[y1 - 1] + ty_taps - 1;
  const int vc0 = P.fx.first[x0], vc1 = P.fx.first[x1 - 1] + tx_taps - 1;
  const int nvr = vr1 - vr0 + 1, nvc = vc1 - vc0 + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- 0. filter data of the tile
  for (int i = threadIdx.x; i < ncol * tx_taps; i += kBlock) s_cx[i] = P.fx.coef[(size_t)x0 * tx_taps + i];
  for (int i = threadIdx.x; i < nrow * ty_taps; i += kBlock) s_cy[i] = P.fy.coef[(size_t)y0 * ty_taps + i];
  for (int i = threadIdx.x; i < ncol; i += kBlock) s_fx[i] = P.fx.first[x0 + i] - vc0;
  for (int i = threadIdx.x; i < nrow; i += kBlock) s_fy[i] = P.fy.first[y0 + i] - vr0;
  // ---- 1. stage the source rectangle, replicating the frame edges
  {
    const int cin0 = max(vc0, 0), cin1 = min(vc1, P.sw - 1);       // columns that exist
    const int lead = cin0 - vc0;                                   // replicated columns on the left
    for (int r = warp; r < nvr; r += kBlock / 32) {
      const int sy = min(max(vr0 + r, 0), P.sh - 1);
      const uint8_t *rp = P.src + (size_t)P.srs * sy;
      uint8_t *dp = s_raw + r * raw_stride;
      if (PS == 4 && (((uintptr_t)rp) & 3) == 0) {
        for (int c = lane; c < nvc; c += 32) {
          const int sx = min(max(vc0 + c, 0), P.sw - 1);
          reinterpret_cast<uint32_t *>(dp)[c] = ld_stream_u32(rp + 4 * sx);
        }
      } else {
        for (int b = lane; b < nvc * PS; b += 32) {
          const int c = b / PS, k = b - c * PS;
          const int sx = min(max(vc0 + c, 0), P.sw - 1);
          dp[b] = rp[sx * PS + k];
        }
      }
    }
    (void)cin1; (void)lead;
  }
  __syncthreads();
  // ---- 2. horizontal pass: tmp = min((sum c14 * pix) >> 7, 32767); lanes walk the output columns of one source row
  for (int r = warp; r < nvr; r += kBlock / 32) {
    const uint8_t *row = s_raw + r * raw_stride;
    for (int xo = lane; xo < ncol; xo += 32) {
      const int16_t *cf = s_cx + xo * tx_taps;
      const int f0 = s_fx[xo];
      int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (int k = 0; k < tx_taps; k++) {
        const int c = cf[k];
        if (PS == 4) {
          const uint32_t p = reinterpret_cast<const uint32_t *>(row)[f0 + k];
          a0 += c * (int)(p & 0xFF); a1 += c * (int)((p >> 8) & 0xFF); a2 += c * (int)((p >> 16) & 0xFF); a3 += c * (int)(p >> 24);
        } else if (PS == 3) {
          const uint8_t *q = row + 3 * (f0 + k);
          a0 += c * q[0]; a1 += c * q[1]; a2 += c * q[2];
        } else {
          a0 += c * row[f0 + k];
        }
      }
      a0 = min(a0 >> 7, 32767); a1 = min(a1 >> 7, 32767); a2 = min(a2 >> 7, 32767); a3 = min(a3 >> 7, 32767);
      s_tmp[r * P.tw + xo] = make_uint2(((uint32_t)a0 & 0xFFFFu) | ((uint32_t)a1 << 16), ((uint32_t)a2 & 0xFFFFu) | ((uint32_t)a3 << 16))  /* signed 16-bit halves: bicubic / Lanczos lobes go negative */;
    }
  }
  __syncthreads();
  // ---- 3. vertical pass: out = clip_u8((sum c12 * tmp + 2^18) >> 19); lanes walk the columns of one output row
  for (int yo = warp; yo < nrow; yo += kBlock / 32) {
    const int16_t *cf = s_cy + yo * ty_taps;
    const int f0 = s_fy[yo];
    uint8_t *drow = P.dst + (size_t)P.drs * (y0 + yo) + (size_t)x0 * PS;
    for (int xo = lane; xo < ncol; xo += 32) {
      int a0 = 1 << 18, a1 = 1 << 18, a2 = 1 << 18, a3 = 1 << 18;
      for (int k = 0; k < ty_taps; k++) {
        const int c = cf[k];
        const uint2 hv = s_tmp[(f0 + k) * P.tw + xo];
        a0 += c * (int)(short)hv.x;
        if (PS > 1) { a1 += c * ((int)hv.x >> 16); a2 += c * (int)(short)hv.y; }
        if (PS == 4) a3 += c * ((int)hv.y >> 16);
      }
      const uint32_t o0 = (uint32_t)min(max(a0 >> 19, 0), 255), o1 = (uint32_t)min(max(a1 >> 19, 0), 255),
                     o2 = (uint32_t)min(max(a2 >> 19, 0), 255), o3 = (uint32_t)min(max(a3 >> 19, 0), 255);
      uint8_t *d = drow + xo * PS;
      if (PS == 4) {
        if (((uintptr_t)d & 3) == 0) *reinterpret_cast<uint32_t *>(d) = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
        else { d[0] = (uint8_t)o0; d[1] = (uint8_t)o1; d[2] = (uint8_t)o2; d[3] = (uint8_t)o3; }
      } else if (PS == 3) {
        d[0] = (uint8_t)o0; d[1] = (uint8_t)o1; d[2] = (uint8_t)o2;
      } else {
        d[0] = (uint8_t)o0;
      }
    }
  }
}

// k_resize_tile4: the same tile kernel specialised for 4-byte pixels and <= 4 taps per axis (bilinear up to 1.5x down), where
// the generic loops above spend ~350 instructions per output pixel.  Horizontal pass: the four tap pixels are byte-transposed
// with 8 PRMT into one word per channel and filtered with two DP2A per channel (16-bit coefficient pairs x 4 bytes); vertical
// pass: 4 x 64-bit tmp loads, unrolled multiply-adds, one saturating pack.  Coefficients are staged per tile as 16-bit pairs
// padded to 4 taps.  Same arithmetic, bit for bit (tests/test_gpu_parity.py::test_resize_*).
__device__ __forceinline__ uint32_t rz_dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t rz_dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// frames of one batched launch: same geometry and strides, blockIdx.y selects the frame
struct ResizeFrameList {
  const uint8_t *src[32];
  uint8_t *dst[32];
};

__global__ void __launch_bounds__(kBlock) k_resize_tile4(const ResizeTileParams P, const __grid_constant__ ResizeFrameList FL) {
  const uint8_t *const f_src = FL.src[blockIdx.y];
  uint8_t *const f_dst = FL.dst[blockIdx.y];
  extern __shared__ __align__(16) uint8_t rsm[];
  const int raw_stride = ((P.max_cols + 10) * 4 + 15) & ~15;                 // window start aligned down to 4 columns, + 3 words of slack for the 4-word tap window, rounded to whole 16-byte groups
  const int tx_taps = P.fx.taps, ty_taps = P.fy.taps;
  uint8_t *s_raw = rsm;                                                    // [max_rows][raw_stride]
  size_t off = ((size_t)P.max_rows * raw_stride + 15) & ~(size_t)15;
  uint2 *s_tmp = reinterpret_cast<uint2 *>(rsm + off);                     // [max_rows + 3][tw]
  off += (size_t)(P.max_rows + 3) * P.tw * 8;
  uint2 *s_cx = reinterpret_cast<uint2 *>(rsm + off);                      // [tw]: c0 | c1 << 16, c2 | c3 << 16
  off += (size_t)P.tw * 8;
  int4 *s_cy = reinterpret_cast<int4 *>(rsm + off);                        // [th]: c0..c3
  off += (size_t)P.th * 16;
  int *s_fx = reinterpret_cast<int *>(rsm + off);                          // [tw] first - vc0
  int *s_fy = s_fx + P.tw;                                                 // [th] first - vr0

  const int tiles_x = (P.dw + P.tw - 1) / P.tw;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int x0 = tx * P.tw, y0 = ty * P.th;
  const int x1 = min(x0 + P.tw, P.dw), y1 = min(y0 + P.th, P.dh);
  const int ncol = x1 - x0, nrow = y1 - y0;
  const int vr0 = P.fy.first[y0], vr1 = P.fy.first[y1 - 1] + ty_taps - 1;
  const int vc0 = P.fx.first[x0], vc1 = P.fx.first[x1 - 1] + tx_taps - 1;
  const int nvr = vr1 - vr0 + 1, nvc = vc1 - vc0 + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- 0. filter data of the tile, padded to 4 taps
  for (int i = threadIdx.x; i < ncol; i += kBlock) {
    uint32_t c[4] = {0, 0, 0, 0};
    for (int k = 0; k < tx_taps; k++) c[k] = (uint16_t)P.fx.coef[(size_t)(x0 + i) * tx_taps + k];
    s_cx[i] = make_uint2(c[0] | (c[1] << 16), c[2] | (c[3] << 16));
    s_fx[i] = P.fx.first[x0 + i] - (vc0 & ~3);
  }
  for (int i = threadIdx.x; i < nrow; i += kBlock) {
    int c[4] = {0, 0, 0, 0};
    for (int k = 0; k < ty_taps; k++) c[k] = P.fy.coef[(size_t)(y0 + i) * ty_taps + k];
    s_cy[i] = make_int4(c[0], c[1], c[2], c[3]);
    s_fy[i] = P.fy.first[y0 + i] - vr0;
  }
  // ---- 1. stage the source rectangle, replicating the frame edges: whole 16-byte groups of 4 pixels, starting at a column that
  //         is a multiple of 4 (so that global and shared addresses are 16-byte aligned together) and reaching 3 pixels past
  //         the last tap (the 4-word window of the last column); groups that touch a frame edge go pixel by pixel
  {
    const int vca = vc0 & ~3;
    const int ngroups = (vc1 + 4 - vca + 3) >> 2;
    const bool vec = P.vec_src != 0;
    for (int r = warp; r < nvr; r += kBlock / 32) {
      const int sy = min(max(vr0 + r, 0), P.sh - 1);
      const uint8_t *rp = f_src + (size_t)P.srs * sy;
      uint4 *dp = reinterpret_cast<uint4 *>(s_raw + r * raw_stride);
      for (int g = lane; g < ngroups; g += 32) {
        const int col = vca + 4 * g;
        if (vec && col >= 0 && col + 3 <= P.sw - 1) {
          dp[g] = ld_stream_u4(rp + 4 * col);
        } else {
          uint4 v;
          v.x = ld_stream_u32(rp + 4 * min(max(col, 0), P.sw - 1));
          v.y = ld_stream_u32(rp + 4 * min(max(col + 1, 0), P.sw - 1));
          v.z = ld_stream_u32(rp + 4 * min(max(col + 2, 0), P.sw - 1));
          v.w = ld_stream_u32(rp + 4 * min(max(col + 3, 0), P.sw - 1));
          dp[g] = v;
        }
      }
    }
  }
  __syncthreads();
  // ---- 2. horizontal pass: tmp = min((sum c14 * pix) >> 7, 32767)
  for (int r = warp; r < nvr; r += kBlock / 32) {
    const uint32_t *row = reinterpret_cast<const uint32_t *>(s_raw + r * raw_stride);
    for (int xo = lane; xo < ncol; xo += 32) {
      const uint2 cf = s_cx[xo];
      const uint32_t *q = row + s_fx[xo];
      const uint32_t p0 = q[0], p1 = q[1], p2 = q[2], p3 = q[3];
      // 4 x 4 byte transpose: one word per channel holding the four taps
      const uint32_t a01 = __byte_perm(p0, p1, 0x5140), b01 = __byte_perm(p0, p1, 0x7362);
      const uint32_t a23 = __byte_perm(p2, p3, 0x5140), b23 = __byte_perm(p2, p3, 0x7362);
      const uint32_t c0 = __byte_perm(a01, a23, 0x5410), c1 = __byte_perm(a01, a23, 0x7632);
      const uint32_t c2 = __byte_perm(b01, b23, 0x5410), c3 = __byte_perm(b01, b23, 0x7632);
      const uint32_t t0 = min(rz_dp2a_hi(cf.y, c0, rz_dp2a_lo(cf.x, c0, 0u)) >> 7, 32767u);
      const uint32_t t1 = min(rz_dp2a_hi(cf.y, c1, rz_dp2a_lo(cf.x, c1, 0u)) >> 7, 32767u);
      const uint32_t t2 = min(rz_dp2a_hi(cf.y, c2, rz_dp2a_lo(cf.x, c2, 0u)) >> 7, 32767u);
      const uint32_t t3 = min(rz_dp2a_hi(cf.y, c3, rz_dp2a_lo(cf.x, c3, 0u)) >> 7, 32767u);
      s_tmp[r * P.tw + xo] = make_uint2(t0 | (t1 << 16), t2 | (t3 << 16));
    }
  }
  __syncthreads();
  // ---- 3. vertical pass: out = clip_u8((sum c12 * tmp + 2^18) >> 19)   (rows past nvr - 1 are only ever multiplied by the zero padding taps)
  for (int yo = warp; yo < nrow; yo += kBlock / 32) {
    const int4 cf = s_cy[yo];
    const uint2 *t = s_tmp + s_fy[yo] * P.tw;
    uint8_t *drow = f_dst + (size_t)P.drs * (y0 + yo) + (size_t)x0 * 4;
    for (int xo = lane; xo < ncol; xo += 32) {
      const uint2 h0 = t[xo], h1 = t[P.tw + xo], h2 = t[2 * P.tw + xo], h3 = t[3 * P.tw + xo];
      int a0 = 1 << 18, a1 = 1 << 18, a2 = 1 << 18, a3 = 1 << 18;
      a0 += cf.x * (int)(h0.x & 0xFFFF) + cf.y * (int)(h1.x & 0xFFFF) + cf.z * (int)(h2.x & 0xFFFF) + cf.w * (int)(h3.x & 0xFFFF);
      a1 += cf.x * (int)(h0.x >> 16) + cf.y * (int)(h1.x >> 16) + cf.z * (int)(h2.x >> 16) + cf.w * (int)(h3.x >> 16);
      a2 += cf.x * (int)(h0.y & 0xFFFF) + cf.y * (int)(h1.y & 0xFFFF) + cf.z * (int)(h2.y & 0xFFFF) + cf.w * (int)(h3.y & 0xFFFF);
      a3 += cf.x * (int)(h0.y >> 16) + cf.y * (int)(h1.y >> 16) + cf.z * (int)(h2.y >> 16) + cf.w * (int)(h3.y >> 16);
      uint32_t lo, px;
      asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(a1 >> 19), "r"(a0 >> 19), "r"(0u));
      asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(px) : "r"(a3 >> 19), "r"(a2 >> 19), "r"(0u));
      *reinterpret_cast<uint32_t *>(drow + 4 * xo) = lo | (px << 16);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// fused kernel
// ------------------------------------------------------------------------------------------------------

constexpr int kTileW = 64, kTileH = 16;

struct FusedSmem {
  // carved from dynamic shared memory: [over table 64 KB][yuv tables 5 KB][rgba tile][hscaled tile]
  uint8_t *over;
  int32_t *tabs;   // [5][256]
  uint32_t *rgba;  // [src_rows][src_cols]
  uint2 *hs;       // [src_rows][kTileW]  4 x int16 per pixel
};

__device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }

__device__ __forceinline__ int chroma_at(const uint8_t *__restrict__ p, int stride, int r, int c, int cw, int ch) {
  if (c >= cw) c = (cw < stride || r + 1 < ch) ? cw : cw - 1;
  return __ldg(p + (long long)stride * r + c);
}

// (u, v) of source pixel (sx, sy): the per-pixel form of the row / row-pair loops of colourspace.c:3391-3642,
// identical to k_yuv_planar_to_rgb
__device__ __forceinline__ void chroma_for_pixel(const FusedArgs &A, int sx, int sy, int &u, int &v) {
  const Planes &S = A.fg;
  const int cw = S.cw, ch = S.ch, h = A.fh;
  const int lo = A.clamped ? 16 : 0, hi = A.clamped ? 240 : 255;
  const int jc = sx >> 1, right = sx & 1;
  bool pair = false;
  int cr_a, cr_b = 0, upper = 0;
  if (A.is_422) cr_a = sy;
  else if (sy == 0) cr_a = 0;
  else if (!(h & 1) && sy == h - 1) cr_a = ch - 1;
  else { pair = true; const int k = (sy + 1) >> 1; cr_a = k - 1; cr_b = k; upper = sy & 1; }
  if (!pair) {
    const int seed_row = (A.is_422 && A.quirks) ? (sy >> 1) : cr_a;
    // column <= 0 is the seed sample
    const int ca = jc, cb = right ? jc + 1 : jc - 1;
    const int ua = ca <= 0 ? __ldg(S.u + (long long)S.rs_u * seed_row) : chroma_at(S.u, S.rs_u, cr_a, ca, cw, ch);
    const int va = ca <= 0 ? __ldg(S.v + (long long)S.rs_v * seed_row) : chroma_at(S.v, S.rs_v, cr_a, ca, cw, ch);
    const int ub = cb <= 0 ? __ldg(S.u + (long long)S.rs_u * seed_row) : chroma_at(S.u, S.rs_u, cr_a, cb, cw, ch);
    const int vb = cb <= 0 ? __ldg(S.v + (long long)S.rs_v * seed_row) : chroma_at(S.v, S.rs_v, cr_a, cb, cw, ch);
    u = clamp_i((ua + ub) >> 1, lo, hi);
    v = clamp_i((va + vb) >> 1, lo, hi);
    return;
  }
  const int cn = right ? jc + 1 : max(jc - 1, 0);  // neighbour column
  const int u1t = chroma_at(S.u, S.rs_u, cr_a, jc, cw, ch), u1n = chroma_at(S.u, S.rs_u, cr_a, cn, cw, ch);
  const int u2t = chroma_at(S.u, S.rs_u, cr_b, jc, cw, ch), u2n = chroma_at(S.u, S.rs_u, cr_b, cn, cw, ch);
  const int v1t = chroma_at(S.v, S.rs_v, cr_a, jc, cw, ch), v1n = chroma_at(S.v, S.rs_v, cr_a, cn, cw, ch);
  const int v2t = chroma_at(S.v, S.rs_v, cr_b, jc, cw, ch), v2n = chroma_at(S.v, S.rs_v, cr_b, cn, cw, ch);
  int u1 = u1t + u1n, u2 = u2t + u2n, v1 = v1t + v1n, v2 = v2t + v2n;
  if (!right && A.quirks) {
    u2 = u1;                                                           // colourspace.c:3461
    if (jc > 0) v1 = v1t + v2n;                                        // :3544
    v2 = v2t + __ldg(S.v + (long long)S.rs_v * cr_b);                  // last_v2 never advanced
  }
  if (!A.low_quality) {
    u = upper ? third_round(u1 + (u2 >> 1)) : third_round((u1 >> 1) + u2);
    v = upper ? third_round(v1 + (v2 >> 1)) : third_round((v1 >> 1) + v2);
  } else {
    u = upper ? (u1 >> 1) : (u2 >> 1);
    v = upper ? (v1 >> 1) : (v2 >> 1);
  }
  u = clamp_i(u, lo, hi);
  v = clamp_i(v, lo, hi);
}

__global__ void __launch_bounds__(kBlock) k_fused(const FusedArgs *__restrict__ frames, int nframes, int tiles_x, int tiles_y,
                                                  int max_src_rows, int max_src_cols) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FusedSmem sm;
  sm.over = smem_raw;
  sm.tabs = (int32_t *)(smem_raw + 65536);
  sm.rgba = (uint32_t *)(smem_raw + 65536 + 5 * 1024);
  sm.hs = (uint2 *)(smem_raw + 65536 + 5 * 1024 + (((size_t)max_src_rows * max_src_cols * 4 + 15) & ~(size_t)15));

  const uint8_t *cur_over = nullptr;
  const int32_t *cur_conv = nullptr;
  const long long tiles_per_frame = (long long)tiles_x * tiles_y, total_tiles = tiles_per_frame * nframes;

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int f = (int)(tile / tiles_per_frame);
    const int t = (int)(tile - (long long)f * tiles_per_frame);
    const FusedArgs A = frames[f];
    __syncthreads();  // previous tile done with shared memory
    if (A.over_table != cur_over) {
      for (int i = threadIdx.x; i < 4096; i += blockDim.x) ((uint4 *)sm.over)[i] = ((const uint4 *)A.over_table)[i];
      cur_over = A.over_table;
    }
    if (A.conv.t != cur_conv) {
      for (int i = threadIdx.x; i < 5 * 256; i += blockDim.x) sm.tabs[i] = A.conv.t[9 * 256 + i];
      cur_conv = A.conv.t;
    }
    const int tx = t % tiles_x, ty = t / tiles_x;
    const int x0 = tx * kTileW, y0 = ty * kTileH;
    const int x1 = min(x0 + kTileW, A.ow), y1 = min(y0 + kTileH, A.oh);
    // intersection with the inner rectangle, in inner coordinates
    const int ix0 = max(x0 - A.ox, 0), ix1 = min(x1 - A.ox, A.iw);
    const int iy0 = max(y0 - A.oy, 0), iy1 = min(y1 - A.oy, A.ih);
    const bool has_inner = ix0 < ix1 && iy0 < iy1;
    int sr0 = 0, sr1 = -1, sc0 = 0, sc1 = -1;
    if (has_inner) {
      sr0 = min(max(A.fy.first[iy0], 0), A.fh - 1);
      sr1 = min(max(A.fy.first[iy1 - 1] + A.fy.taps - 1, 0), A.fh - 1);
      sc0 = min(max(A.fx.first[ix0], 0), A.fw - 1);
      sc1 = min(max(A.fx.first[ix1 - 1] + A.fx.taps - 1, 0), A.fw - 1);
    }
    const int nsr = sr1 - sr0 + 1, nsc = sc1 - sc0 + 1;
    __syncthreads();
    // ---- stage 1: convert the needed source pixels once
    for (int i = threadIdx.x; i < nsr * nsc; i += blockDim.x) {
      const int r = i / nsc, c = i - r * nsc;
      const int sy = sr0 + r, sx = sc0 + c;
      int u, v;
      chroma_for_pixel(A, sx, sy, u, v);
      const int y = __ldg(A.fg.y + (long long)A.fg.rs_y * sy + sx);
      const int yy = sm.tabs[y];
      const int rr = sat8((yy + sm.tabs[256 + v]) >> 16);
      const int gg = sat8((yy + sm.tabs[512 + u] + sm.tabs[768 + v]) >> 16);
      const int bb = sat8((yy + sm.tabs[1024 + u]) >> 16);
      sm.rgba[r * max_src_cols + c] = (uint32_t)rr | ((uint32_t)gg << 8) | ((uint32_t)bb << 16) | 0xFF000000u;
    }
    __syncthreads();
    // ---- stage 2: horizontal scale of every needed source row, for the tile's inner columns
    const int niw = ix1 - ix0;
    for (int i = threadIdx.x; i < nsr * niw; i += blockDim.x) {
      const int r = i / niw, xi = ix0 + (i - r * niw);
      const int first = A.fx.first[xi];
      int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (int k = 0; k < A.fx.taps; k++) {
        const int sx = min(max(first + k, 0), A.fw - 1) - sc0;
        const int c = A.fx.coef[(long long)xi * A.fx.taps + k];
        const uint32_t p = sm.rgba[r * max_src_cols + sx];
        a0 += c * (int)(p & 0xFF); a1 += c * (int)((p >> 8) & 0xFF); a2 += c * (int)((p >> 16) & 0xFF); a3 += c * (int)(p >> 24);
      }
      a0 = min(a0 >> 7, 32767); a1 = min(a1 >> 7, 32767); a2 = min(a2 >> 7, 32767); a3 = min(a3 >> 7, 32767);
      sm.hs[r * kTileW + (xi - ix0)] = make_uint2(((uint32_t)a0 & 0xFFFFu) | ((uint32_t)a1 << 16), ((uint32_t)a2 & 0xFFFFu) | ((uint32_t)a3 << 16))  /* signed 16-bit halves: bicubic / Lanczos lobes go negative */;
    }
    __syncthreads();
    // ---- stage 3: vertical scale, letterbox, alpha-over (+ gamma) and store; one thread = 4 output pixels
    const int gw = kTileW / 4;
    for (int i = threadIdx.x; i < gw * (y1 - y0); i += blockDim.x) {
      const int ry = i / gw, gx = i - ry * gw;
      const int oy_ = y0 + ry, ox_ = x0 + gx * 4;
      if (ox_ >= A.ow) continue;
      const int npx = min(4, A.ow - ox_);
      const uint8_t *bgp = A.bg.p + (long long)A.bg.rs * oy_ + (long long)ox_ * 4;
      uint8_t *dp = A.out.p + (long long)A.out.rs * oy_ + (long long)ox_ * 4;
      const bool vec = npx == 4 && ((((uintptr_t)bgp | (uintptr_t)dp) & 15) == 0);
      uint32_t bgw[4], outw[4];
      if (vec) { const uint4 b = ld_stream_u4(bgp); bgw[0] = b.x; bgw[1] = b.y; bgw[2] = b.z; bgw[3] = b.w; }
      else for (int k = 0; k < npx; k++) bgw[k] = *(const uint32_t *)(bgp + 4 * k);
      const int iy = oy_ - A.oy;
      const bool row_in = iy >= iy0 && iy < iy1;
      int vfirst = 0;
      if (row_in) vfirst = A.fy.first[iy];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (k >= npx) break;
        const int ix = ox_ + k - A.ox;
        uint32_t fgp = 0xFF000000u;  // letterbox border: black, opaque (blank_pixel, colourspace.c:11169)
        if (row_in && ix >= ix0 && ix < ix1) {
          int a0 = 1 << 18, a1 = 1 << 18, a2 = 1 << 18, a3 = 1 << 18;
          for (int t2 = 0; t2 < A.fy.taps; t2++) {
            const int sy = min(max(vfirst + t2, 0), A.fh - 1) - sr0;
            const int c = A.fy.coef[(long long)iy * A.fy.taps + t2];
            const uint2 hv = sm.hs[sy * kTileW + (ix - ix0)];
            a0 += c * (int)(short)hv.x; a1 += c * ((int)hv.x >> 16);
            a2 += c * (int)(short)hv.y; a3 += c * ((int)hv.y >> 16);
          }
          fgp = (uint32_t)sat8(a0 >> 19) | ((uint32_t)sat8(a1 >> 19) << 8) | ((uint32_t)sat8(a2 >> 19) << 16) |
                ((uint32_t)sat8(a3 >> 19) << 24);
        }
        const uint32_t b = bgw[k];
        outw[k] = (uint32_t)sm.over[((b & 0xFF) << 8) | (fgp & 0xFF)] |
                  ((uint32_t)sm.over[(((b >> 8) & 0xFF) << 8) | ((fgp >> 8) & 0xFF)] << 8) |
                  ((uint32_t)sm.over[(((b >> 16) & 0xFF) << 8) | ((fgp >> 16) & 0xFF)] << 16) | 0xFF000000u;
      }
      if (vec) st_stream_u4(dp, make_uint4(outw[0], outw[1], outw[2], outw[3]));
      else for (int k = 0; k < npx; k++) *(uint32_t *)(dp + 4 * k) = outw[k];
    }
  }
}

}  // namespace

cudaError_t launch_resize_h(const Launch &L, CImg src, int sw, int sh, int16_t *tmp, int dw, int psize, DevFilter fx) {
  k_resize_h<<<grid_for(L, (long long)dw * sh), kBlock, 0, L.stream>>>(src.p, src.rs, sw, sh, tmp, dw, psize, fx);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_resize_v(const Launch &L, const int16_t *tmp, int sh, Img dst, int dw, int dh, int psize, DevFilter fy) {
  k_resize_v<<<grid_for(L, (long long)dw * psize * dh), kBlock, 0, L.stream>>>(tmp, sh, dst.p, dst.rs, dw, dh, psize, fy);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// hx / hy: host copies of the filter banks (to size the tile).  Returns cudaErrorInvalidConfiguration when no tile fits in
// shared memory (the caller then runs the two-kernel path).
// nbatch > 0: srcs / dsts hold nbatch frames of the same geometry and strides (as src / dst describe frame 0): one launch per 32
// frames of k_resize_tile4; cudaErrorInvalidConfiguration when the specialised kernel cannot take the job (the caller then
// issues them one by one)
static cudaError_t launch_resize_tile_impl(const Launch &L, CImg src, int sw, int sh, Img dst, int dw, int dh, int psize, DevFilter fx,
                                           DevFilter fy, const int32_t *hx_first, const int32_t *hy_first, const uint8_t *const *srcs,
                                           uint8_t *const *dsts, int nbatch);
cudaError_t launch_resize_tile(const Launch &L, CImg src, int sw, int sh, Img dst, int dw, int dh, int psize, DevFilter fx,
                               DevFilter fy, const int32_t *hx_first, const int32_t *hy_first) {
  return launch_resize_tile_impl(L, src, sw, sh, dst, dw, dh, psize, fx, fy, hx_first, hy_first, nullptr, nullptr, 0);
}
cudaError_t launch_resize_tile_batch(const Launch &L, const uint8_t *const *srcs, int srs, int sw, int sh, uint8_t *const *dsts, int drs,
                                     int dw, int dh, int psize, DevFilter fx, DevFilter fy, const int32_t *hx_first,
                                     const int32_t *hy_first, int n) {
  return launch_resize_tile_impl(L, CImg{srcs[0], srs}, sw, sh, Img{dsts[0], drs}, dw, dh, psize, fx, fy, hx_first, hy_first, srcs, dsts, n);
}
static cudaError_t launch_resize_tile_impl(const Launch &L, CImg src, int sw, int sh, Img dst, int dw, int dh, int psize, DevFilter fx,
                                           DevFilter fy, const int32_t *hx_first, const int32_t *hy_first, const uint8_t *const *srcs,
                                           uint8_t *const *dsts, int nbatch) {
  auto span = [](const int32_t *first, int taps, int dst_n, int tile) {  // unclamped tap range of the widest tile
    int worst = 1;
    for (int i0 = 0; i0 < dst_n; i0 += tile) {
      const int i1 = i0 + tile - 1 < dst_n - 1 ? i0 + tile - 1 : dst_n - 1;
      const int n = first[i1] + taps - 1 - first[i0] + 1;
      if (n > worst) worst = n;
    }
    return worst;
  };
  ResizeTileParams P;
  P.src = src.p; P.dst = dst.p; P.srs = src.rs; P.drs = dst.rs; P.sw = sw; P.sh = sh; P.dw = dw; P.dh = dh; P.fx = fx; P.fy = fy;
  P.vec_src = 0;
  size_t smem = 0;
  bool ok = false;
  static int tw0 = 0, th0 = 0;
  if (!tw0) {
    tw0 = getenv("PE_RESIZE_TW") ? atoi(getenv("PE_RESIZE_TW")) : 128;  // measured on cfg2: 128 x 16 45.1k fps, 64 x 32 42.6k, 32 x 16 39.2k
    th0 = getenv("PE_RESIZE_TH") ? atoi(getenv("PE_RESIZE_TH")) : 16;
    if (tw0 < 8 || tw0 > 128) tw0 = 128;
    if (th0 < 8 || th0 > 64) th0 = 16;
  }
  for (int tw = tw0, th = th0; tw >= 8 && !ok; tw >>= 1, th = th > 8 ? th >> 1 : th) {
    P.tw = tw; P.th = th;
    P.max_cols = span(hx_first, fx.taps, dw, tw);
    P.max_rows = span(hy_first, fy.taps, dh, th);
    const size_t raw = (((size_t)P.max_rows * ((P.max_cols * psize + 3) & ~3)) + 15) & ~(size_t)15;
    smem = raw + (size_t)P.max_rows * tw * 8 + (((size_t)tw * fx.taps * 2 + 15) & ~(size_t)15) +
           (((size_t)th * fy.taps * 2 + 15) & ~(size_t)15) + (size_t)(tw + th) * 4;
    ok = smem <= 96 * 1024;
  }
  if (!ok) return cudaErrorInvalidConfiguration;
  const int tiles = ((dw + P.tw - 1) / P.tw) * ((dh + P.th - 1) / P.th);
  static PerDevice attr[5];
  cudaError_t e = cudaSuccess;
  if (smem > attr[psize].cur()) {
    if (psize == 4) e = cudaFuncSetAttribute(k_resize_tile<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    else if (psize == 3) e = cudaFuncSetAttribute(k_resize_tile<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    else e = cudaFuncSetAttribute(k_resize_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    attr[psize].cur() = 96 * 1024;
  }
  if (psize == 4 && fx.taps <= 4 && fy.taps <= 4 && fx.nonneg && fy.nonneg && ((((uintptr_t)dst.p | (uintptr_t)src.p) | (uint32_t)dst.rs | (uint32_t)src.rs) & 3) == 0 &&
      getenv("PE_RESIZE_GENERIC") == nullptr) {
    // the specialised kernel has its own (slightly larger) shared-memory layout
    const size_t raw4 = (size_t)P.max_rows * ((((size_t)P.max_cols + 10) * 4 + 15) & ~(size_t)15);
    P.vec_src = ((uint32_t)src.rs & 15) == 0 && getenv("PE_RESIZE_NOVEC") == nullptr;
    for (int i = 0; i < (nbatch > 0 ? nbatch : 1) && P.vec_src; i++) P.vec_src = (((uintptr_t)(nbatch > 0 ? srcs[i] : src.p)) & 15) == 0;
    const size_t smem4 = raw4 + (size_t)(P.max_rows + 3) * P.tw * 8 + (size_t)P.tw * 8 + (size_t)P.th * 16 + (size_t)(P.tw + P.th) * 4;
    static PerDevice attr4;
    if (smem4 <= 96 * 1024) {
      if (!attr4.cur()) {
        if ((e = cudaFuncSetAttribute(k_resize_tile4, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)) != cudaSuccess) return e;
        attr4.cur() = 1;
      }
      const int nf = nbatch > 0 ? nbatch : 1;
      for (int base = 0; base < nf; base += 32) {
        ResizeFrameList fl;
        const int cnt = nf - base < 32 ? nf - base : 32;
        for (int i = 0; i < 32; i++) {
          const int k = base + (i < cnt ? i : 0);
          fl.src[i] = nbatch > 0 ? srcs[k] : src.p;
          fl.dst[i] = nbatch > 0 ? dsts[k] : dst.p;
        }
        k_resize_tile4<<<dim3(tiles, cnt), kBlock, smem4, L.stream>>>(P, fl);
        PE_COUNT_LAUNCH(L);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
      }
      return cudaSuccess;
    }
  }
  if (nbatch > 0) return cudaErrorInvalidConfiguration;  // only the specialised kernel takes frame lists
  if (psize == 4) k_resize_tile<4><<<tiles, kBlock, smem, L.stream>>>(P);
  else if (psize == 3) k_resize_tile<3><<<tiles, kBlock, smem, L.stream>>>(P);
  else if (psize == 1) k_resize_tile<1><<<tiles, kBlock, smem, L.stream>>>(P);
  else return cudaErrorInvalidValue;
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// frames_dev: FusedArgs array in DEVICE memory; src extents computed by the caller (engine) from the filter banks
cudaError_t launch_fused_dev(const Launch &L, const FusedArgs *frames_dev, int nframes, int ow, int oh, int max_src_rows,
                             int max_src_cols) {
  const int tiles_x = (ow + kTileW - 1) / kTileW, tiles_y = (oh + kTileH - 1) / kTileH;
  const size_t smem = 65536 + 5 * 1024 + (((size_t)max_src_rows * max_src_cols * 4 + 15) & ~(size_t)15) + (size_t)max_src_rows * kTileW * 8;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  static PerDevice attr_set;
  if (smem > attr_set.cur()) {
    cudaError_t e = cudaFuncSetAttribute(k_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set.cur() = smem;
  }
  const int per_sm = smem <= 110 * 1024 ? 2 : 1;
  long long total = (long long)tiles_x * tiles_y * nframes;
  int grid = (int)(total < (long long)L.sm_count * per_sm ? total : (long long)L.sm_count * per_sm);
  k_fused<<<grid, kBlock, smem, L.stream>>>(frames_dev, nframes, tiles_x, tiles_y, max_src_rows, max_src_cols);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

int fused_tile_w() { return kTileW; }
int fused_tile_h() { return kTileH; }

}  // namespace pe
