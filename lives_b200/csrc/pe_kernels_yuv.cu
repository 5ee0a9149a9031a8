// pe_kernels_yuv.cu -- YUV <-> RGB conversion kernels (sm_100a).
//
//   k_yuv_planar_to_rgb   convert_yuv420p_to_{rgb,bgr,argb}_frame  colourspace.c:3260 / 3927 / 4527
//                         (4:2:0 and 4:2:2 planar, chroma super-sampling, optional inline 16-bit gamma LUT)
//   k_packed422_to_rgb    convert_{uyvy,yuyv}_to_*_frame            colourspace.c:6616-7103
//   k_yuv888_to_rgb       convert_yuv888 / yuva8888_to_*_frame      colourspace.c:2750-3258
//   k_rgb_to_yuv888       convert_{rgb,bgr,argb}_to_yuv_frame       colourspace.c:5700-6239
//
// Arithmetic contract (bit exact with the reference):
//   R = clamp((RGB_Y[y] + R_Cr[v]) >> 16, 0, 255) etc. -- the reference's float32 spc_rnd(HIGH) and
//   CLAMP0255f are byte-identical to this integer form for all 2^24 inputs (tests/test_oracle_vs_reference.py),
//   chroma weights (int)(n / 3. + .5) == (2n + 3) / 6.
// The five 256-entry int tables live in shared memory.
#include <algorithm>

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

// shared-memory copy of the YUV->RGB tables: [0] RGB_Y [1] R_Cr [2] G_Cb [3] G_Cr [4] B_Cb
struct SmemYuvTabs {
  int32_t t[5][256];
};

__device__ __forceinline__ void load_yuv_tabs(SmemYuvTabs &s, const int32_t *conv_t) {
  for (int i = threadIdx.x; i < 5 * 256; i += blockDim.x) s.t[i >> 8][i & 255] = conv_t[(9 + (i >> 8)) * 256 + (i & 255)];
}

__device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }

// yuv2rgb_int colourspace.c:2345 / xyuv2rgb :2351 (+ xyuv2rgb_with_gamma :2386 when lut16 != nullptr)
__device__ __forceinline__ void yuv_px(const SmemYuvTabs &s, const uint16_t *lut16, int y, int u, int v, int &r, int &g,
                                       int &b) {
  const int yy = s.t[0][y];
  const int rr = yy + s.t[1][v], gg = yy + s.t[2][u] + s.t[3][v], bb = yy + s.t[4][u];
  if (!lut16) {
    r = sat8(rr >> 16); g = sat8(gg >> 16); b = sat8(bb >> 16);
  } else {
    r = __ldg(lut16 + min(max(rr >> 8, 0), 65535)) >> 8;
    g = __ldg(lut16 + min(max(gg >> 8, 0), 65535)) >> 8;
    b = __ldg(lut16 + min(max(bb >> 8, 0), 65535)) >> 8;
  }
}

__device__ __forceinline__ uint32_t pack_px(const RgbLayout &o, int r, int g, int b) {
  uint32_t w = ((uint32_t)r << (8 * o.r)) | ((uint32_t)g << (8 * o.g)) | ((uint32_t)b << (8 * o.b));
  if (o.a >= 0) w |= 0xFFu << (8 * o.a);
  return w;
}

// store 4 pixels (already packed, one per word) as 16 or 12 bytes
__device__ __forceinline__ void store_px4(uint8_t *d, int psize, const uint32_t px[4], bool vec) {
  if (psize == 4) {
    if (vec) st_stream_u4(d, make_uint4(px[0], px[1], px[2], px[3]));
    else for (int k = 0; k < 4; k++) *(uint32_t *)(d + 4 * k) = px[k];
  } else {
    const uint32_t w0 = __byte_perm(px[0], px[1], 0x4210), w1 = __byte_perm(px[1], px[2], 0x5421),
                   w2 = __byte_perm(px[2], px[3], 0x6542);
    if (vec) { st_stream_u32(d, w0); st_stream_u32(d + 4, w1); st_stream_u32(d + 8, w2); }
    else for (int k = 0; k < 12; k++) d[k] = (uint8_t)((k < 4 ? w0 : k < 8 ? w1 : w2) >> (8 * (k & 3)));
  }
}

__device__ __forceinline__ void store_px_n(uint8_t *d, int psize, const uint32_t *px, int n) {
  for (int k = 0; k < n; k++)
    for (int b = 0; b < psize; b++) d[k * psize + b] = (uint8_t)(px[k] >> (8 * b));
}

// chroma sample with the reference's one-past-row read (colourspace.c:3508-3512, :3613): column == cw reads the
// byte at plane[stride * r + cw] -- padding or the first sample of row r+1 -- except on the last chroma row of a
// plane whose stride equals its width, where it is defined as the replicated edge sample (DESIGN.md "edge read").
__device__ __forceinline__ int chroma_at(const uint8_t *__restrict__ p, int stride, int r, int c, int cw, int ch) {
  if (c >= cw) c = (cw < stride || r + 1 < ch) ? cw : cw - 1;
  return __ldg(p + (long long)stride * r + c);
}

// d = (c[15:0] << 16) | (sat_u8(a) << 8) | sat_u8(b)
__device__ __forceinline__ uint32_t pack_sat(int a, int b, uint32_t c) {
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// One thread = 4 luma columns (two chroma columns) of one "job":
//   4:2:0  job 0           : luma row 0      with chroma row 0              (single-row average)
//          job k, 1..ch-1  : luma rows 2k-1, 2k with chroma rows k-1, k     (2/3 - 1/3 vertical weights)
//          job ch (h even) : luma row h-1    with chroma row ch-1           (single-row average)
//   4:2:2  job i           : luma row i      with chroma row i
// The chroma tables in shared memory are the extended ones (DevConv::ext): indexed by the un-divided chroma sum, they
// already contain the (int)(n / 3. + .5) rounding and the CLAMP16_240 / CLAMP0_255 of colourspace.c:3465-3469.
// Interior threads fetch the 4 chroma samples they need (columns jc0-1 .. jc0+2) with two aligned 32-bit loads and a funnel
// shift; threads at the left / right frame edge take the byte path with the reference's edge rules (chroma_at).
__global__ void __launch_bounds__(kBlock) k_yuv_planar_to_rgb(const YuvToRgbArgs A) {
  __shared__ int32_t s_t[256 + 4 * kExtN];
  for (int i = threadIdx.x; i < 256 + 4 * kExtN; i += kBlock) s_t[i] = A.conv.ext[i];
  __syncthreads();
  const Planes &S = A.src;
  const int w = A.width, h = A.height, cw = S.cw, ch = S.ch;
  const int groups = (w + 3) >> 2;
  const int njobs = A.is_422 ? h : (h >= 2 && !(h & 1) ? ch + 1 : ch);
  const bool vec = ((uintptr_t)A.dst.p % 16 == 0) && (A.dst.rs % 16 == 0);
  const bool yvec = ((uintptr_t)S.y % 4 == 0) && (S.rs_y % 4 == 0);
  const bool cvec = ((((uintptr_t)S.u | (uintptr_t)S.v) & 3) == 0) && (((S.rs_u | S.rs_v) & 3) == 0);
  // canonical pixel word = [r, g, b, 255]; sel moves its bytes to the palette's order
  uint32_t sel = 0;
  {
    uint32_t nib[4] = {3, 3, 3, 3};
    nib[A.out.r] = 0; nib[A.out.g] = 1; nib[A.out.b] = 2;
    if (A.out.a >= 0) nib[A.out.a] = 3;
    sel = nib[0] | (nib[1] << 4) | (nib[2] << 8) | (nib[3] << 12);
  }
  const uint16_t *lut16 = A.lut16;
  auto emit = [&](int y, int nu, int nv) -> uint32_t {
    const int yy = s_t[y];
    const int rr = yy + s_t[256 + nv], gg = yy + s_t[256 + kExtN + nu] + s_t[256 + 2 * kExtN + nv], bb = yy + s_t[256 + 3 * kExtN + nu];
    uint32_t px;
    if (!lut16) {
      px = pack_sat(gg >> 16, rr >> 16, pack_sat(255, bb >> 16, 0u));
    } else {  // xyuv2rgb_with_gamma colourspace.c:2386
      const uint32_t r = __ldg(lut16 + min(max(rr >> 8, 0), 65535)) >> 8, g = __ldg(lut16 + min(max(gg >> 8, 0), 65535)) >> 8,
                     b = __ldg(lut16 + min(max(bb >> 8, 0), 65535)) >> 8;
      px = r | (g << 8) | (b << 16) | 0xFF000000u;
    }
    return __byte_perm(px, 0u, sel);
  };
  // chroma samples of columns jc0-1 .. jc0+2 of row r as one word
  auto cword = [&](const uint8_t *__restrict__ p, int stride, int r, int jc0, bool interior) -> uint32_t {
    if (interior) {
      const uint8_t *q = p + (size_t)stride * r + ((jc0 - 1) & ~3);
      return __funnelshift_r(__ldg(reinterpret_cast<const uint32_t *>(q)), __ldg(reinterpret_cast<const uint32_t *>(q + 4)),
                             8 * ((jc0 - 1) & 3));
    }
    uint32_t wv = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) wv |= (uint32_t)chroma_at(p, stride, r, max(jc0 - 1 + k, 0), cw, ch) << (8 * k);
    return wv;
  };
  const long long total = (long long)groups * njobs;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int job = (int)(it / groups), g = (int)(it - (long long)job * groups);
    const int x0 = g * 4, npx = min(4, w - x0), jc0 = g * 2;
    int row_a, row_b = -1, cr_a, cr_b = -1;  // luma rows / chroma rows
    bool pair = false;
    if (A.is_422) { row_a = job; cr_a = job; }
    else if (job == 0) { row_a = 0; cr_a = 0; }
    else if (job < ch) { pair = true; row_a = 2 * job - 1; row_b = 2 * job; cr_a = job - 1; cr_b = job; }
    else { row_a = h - 1; cr_a = ch - 1; }
    // columns jc0-1 .. jc0+2 all inside the plane, and the two aligned words inside the row
    const bool interior = cvec && jc0 >= 1 && jc0 + 2 < cw && (((jc0 - 1) & ~3) + 8 <= min(S.rs_u, S.rs_v));

    // luma
    uint32_t ya, yb = 0;
    {
      const uint8_t *py = S.y + (size_t)S.rs_y * row_a + x0;
      if (yvec && npx == 4) ya = ld_stream_u32(py);
      else { ya = 0; for (int k = 0; k < npx; k++) ya |= (uint32_t)py[k] << (8 * k); }
      if (pair) {
        py = S.y + (size_t)S.rs_y * row_b + x0;
        if (yvec && npx == 4) yb = ld_stream_u32(py);
        else for (int k = 0; k < npx; k++) yb |= (uint32_t)py[k] << (8 * k);
      }
    }

    uint32_t out_a[4], out_b[4];
    if (!pair) {
      // horizontal average only (:3394-3438 row 0, :3551-3596 last row -- X rows: intended arithmetic -- and the
      // whole 4:2:2 branch :3598-3642).  4:2:2 quirk: the running pair is seeded from chroma row (i >> 1) (:3600)
      const int seed_row = (A.is_422 && A.quirks) ? (row_a >> 1) : cr_a;
      uint32_t uw = cword(S.u, S.rs_u, cr_a, jc0, interior), vw = cword(S.v, S.rs_v, cr_a, jc0, interior);
      if (jc0 == 0) {  // columns <= 0 are the seed sample (last = this = seed at the start of a row)
        const uint32_t su = __ldg(S.u + (size_t)S.rs_u * seed_row), sv = __ldg(S.v + (size_t)S.rs_v * seed_row);
        uw = (uw & 0xFFFF0000u) | su | (su << 8);
        vw = (vw & 0xFFFF0000u) | sv | (sv << 8);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        // pixel x0+k: pair index p = k>>1 (chroma column jc0+p = byte p+1); left pixel averages with the previous
        // column, right pixel with the next one
        const int p = k >> 1;
        const int ua = byte_of(uw, p + 1), va = byte_of(vw, p + 1);
        const int ub = (k & 1) ? byte_of(uw, p + 2) : byte_of(uw, p), vb = (k & 1) ? byte_of(vw, p + 2) : byte_of(vw, p);
        out_a[k] = emit(byte_of(ya, k), 3 * ((ua + ub) >> 1), 3 * ((va + vb) >> 1));
      }
    } else {
      // interior row pair (:3440-3549)
      const uint32_t u1w = cword(S.u, S.rs_u, cr_a, jc0, interior), u2w = cword(S.u, S.rs_u, cr_b, jc0, interior);
      const uint32_t v1w = cword(S.v, S.rs_v, cr_a, jc0, interior), v2w = cword(S.v, S.rs_v, cr_b, jc0, interior);
      const int v2_first = A.quirks ? __ldg(S.v + (size_t)S.rs_v * cr_b) : 0;
#pragma unroll
      for (int p = 0; p < 2; p++) {
        const int U1[3] = {(int)byte_of(u1w, p), (int)byte_of(u1w, p + 1), (int)byte_of(u1w, p + 2)};
        const int U2[3] = {(int)byte_of(u2w, p), (int)byte_of(u2w, p + 1), (int)byte_of(u2w, p + 2)};
        const int V1[3] = {(int)byte_of(v1w, p), (int)byte_of(v1w, p + 1), (int)byte_of(v1w, p + 2)};
        const int V2[3] = {(int)byte_of(v2w, p), (int)byte_of(v2w, p + 1), (int)byte_of(v2w, p + 2)};
#pragma unroll
        for (int lr = 0; lr < 2; lr++) {
          const int k = 2 * p + lr;
          int u1, u2, v1, v2;
          if (lr) {      // right pixel: this + next
            u1 = U1[1] + U1[2]; u2 = U2[1] + U2[2]; v1 = V1[1] + V1[2]; v2 = V2[1] + V2[2];
          } else {       // left pixel: this + last
            u1 = U1[1] + U1[0]; u2 = U2[1] + U2[0]; v1 = V1[1] + V1[0]; v2 = V2[1] + V2[0];
            if (A.quirks) {
              u2 = u1;                                        // `u2 = this_u1 + last_u1`          (:3461)
              if (jc0 + p > 0) v1 = V1[1] + V2[0];            // `last_v1 = this_v2`               (:3544)
              v2 = V2[1] + v2_first;                          // last_v2 is never advanced         (:3543-3546)
            }
          }
          int n3u, n4u, n3v, n4v;
          if (!A.low_quality) {
            n3u = u1 + (u2 >> 1); n4u = (u1 >> 1) + u2; n3v = v1 + (v2 >> 1); n4v = (v1 >> 1) + v2;
          } else {       // PB_QUALITY_LOW: u3 = u1 >> 1, u4 = u2 >> 1 (:3470-3474)
            n3u = 3 * (u1 >> 1); n4u = 3 * (u2 >> 1); n3v = 3 * (v1 >> 1); n4v = 3 * (v2 >> 1);
          }
          out_a[k] = emit(byte_of(ya, k), n3u, n3v);
          out_b[k] = emit(byte_of(yb, k), n4u, n4v);
        }
      }
    }
    if (A.blend2) {
      // fused crossfade: dst = (bf * in2 + (255 - bf) * converted) >> 8 per byte (make_blend_table, simple_blend.c:31-35)
      const uint32_t bf = (uint32_t)A.blend_bf & 0xFFu, nb = 255u - bf;
      auto xfade = [&](uint32_t *px, int row) {
        const uint8_t *q = A.blend2 + (size_t)A.blend2_rs * row + (size_t)x0 * 3;
        uint32_t o[4];
        if (npx == 4 && (((uintptr_t)A.blend2 | (uint32_t)A.blend2_rs) & 3) == 0) {
          const uint32_t w0 = ld_stream_u32(q), w1 = ld_stream_u32(q + 4), w2 = ld_stream_u32(q + 8);
          o[0] = w0; o[1] = __byte_perm(w0, w1, 0x0543); o[2] = __byte_perm(w1, w2, 0x0432); o[3] = w2 >> 8;
        } else {
          for (int k = 0; k < 4; k++) o[k] = k < npx ? (uint32_t)q[3 * k] | ((uint32_t)q[3 * k + 1] << 8) | ((uint32_t)q[3 * k + 2] << 16) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t even = (((px[k] & 0x00FF00FFu) * nb + (o[k] & 0x00FF00FFu) * bf) >> 8) & 0x00FF00FFu;
          const uint32_t mid = (((px[k] >> 8) & 0xFFu) * nb + ((o[k] >> 8) & 0xFFu) * bf) & 0xFF00u;
          px[k] = even | mid;
        }
      };
      xfade(out_a, row_a);
      if (pair) xfade(out_b, row_b);
    }
    uint8_t *da = A.dst.p + (size_t)A.dst.rs * row_a + (size_t)x0 * A.out.psize;
    if (npx == 4) store_px4(da, A.out.psize, out_a, vec); else store_px_n(da, A.out.psize, out_a, npx);
    if (pair) {
      uint8_t *db = A.dst.p + (size_t)A.dst.rs * row_b + (size_t)x0 * A.out.psize;
      if (npx == 4) store_px4(db, A.out.psize, out_b, vec); else store_px_n(db, A.out.psize, out_b, npx);
    }
  }
}

// packed 4:2:2: both pixels of a macropixel share u0, v0 (uyvy2rgb colourspace.c:2410-2415). One thread = 2
// macropixels = 4 output pixels.
__global__ void __launch_bounds__(kBlock) k_packed422_to_rgb(int fmt, const uint8_t *__restrict__ src, int irow, uint8_t *dst,
                                                             int orow, int width_mpx, int height, RgbLayout out, DevConv conv) {
  __shared__ SmemYuvTabs s;
  load_yuv_tabs(s, conv.t);
  __syncthreads();
  const int groups = (width_mpx + 1) >> 1;
  const bool vec = ((uintptr_t)dst % 16 == 0) && (orow % 16 == 0);
  const long long total = (long long)groups * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int nm = min(2, width_mpx - g * 2);
    uint32_t px[4];
    for (int m = 0; m < nm; m++) {
      const uint32_t mp = *(const uint32_t *)(src + (long long)irow * row + (long long)(g * 2 + m) * 4);
      int y0, y1, u, v, r, gg, b;
      if (fmt == 0) { u = byte_of(mp, 0); y0 = byte_of(mp, 1); v = byte_of(mp, 2); y1 = byte_of(mp, 3); }
      else { y0 = byte_of(mp, 0); u = byte_of(mp, 1); y1 = byte_of(mp, 2); v = byte_of(mp, 3); }
      yuv_px(s, nullptr, y0, u, v, r, gg, b); px[2 * m] = pack_px(out, r, gg, b);
      yuv_px(s, nullptr, y1, u, v, r, gg, b); px[2 * m + 1] = pack_px(out, r, gg, b);
    }
    uint8_t *d = dst + (long long)orow * row + (long long)g * 4 * out.psize;
    if (nm == 2) store_px4(d, out.psize, px, vec); else store_px_n(d, out.psize, px, 2 * nm);
  }
}

__global__ void __launch_bounds__(kBlock) k_yuv888_to_rgb(const uint8_t *__restrict__ src, int irow, uint8_t *dst, int orow,
                                                          int width, int height, int in_alpha, RgbLayout out, DevConv conv) {
  __shared__ SmemYuvTabs s;
  load_yuv_tabs(s, conv.t);
  __syncthreads();
  const int ips = in_alpha ? 4 : 3;
  const long long total = (long long)width * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / width), x = (int)(it - (long long)row * width);
    const uint8_t *q = src + (long long)irow * row + (long long)x * ips;
    int r, g, b;
    yuv_px(s, nullptr, q[0], q[1], q[2], r, g, b);
    uint32_t w = ((uint32_t)r << (8 * out.r)) | ((uint32_t)g << (8 * out.g)) | ((uint32_t)b << (8 * out.b));
    if (out.a >= 0) w |= (in_alpha ? (uint32_t)q[3] : 255u) << (8 * out.a);
    uint8_t *d = dst + (long long)orow * row + (long long)x * out.psize;
    for (int k = 0; k < out.psize; k++) d[k] = (uint8_t)(w >> (8 * k));
  }
}

// rgb2yuv colourspace.c:2119-2127: always the YCbCr tables (:5710); the reference rounds the width down to even (:5750)
__global__ void __launch_bounds__(kBlock) k_rgb_to_yuv888(const uint8_t *__restrict__ src, int irow, uint8_t *dst, int orow,
                                                          int width, int height, RgbLayout in, int out_alpha, DevConv conv) {
  __shared__ int32_t t[9][256];
  for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) t[i >> 8][i & 255] = conv.t[i];
  __syncthreads();
  const int ops = out_alpha ? 4 : 3;
  const long long total = (long long)width * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / width), x = (int)(it - (long long)row * width);
    const uint8_t *q = src + (long long)irow * row + (long long)x * in.psize;
    const int r = q[in.r], g = q[in.g], b = q[in.b];
    // the reference stores the rounded sum in a short before comparing (:2120); sums are < 2^15 so no wrap
    const int y = clamp_i((t[0][r] + t[1][g] + t[2][b]) >> 16, conv.min_y, conv.max_y);
    const int u = clamp_i((t[3][r] + t[4][g] + t[5][b]) >> 16, conv.min_uv, conv.max_uv);
    const int v = clamp_i((t[6][r] + t[7][g] + t[8][b]) >> 16, conv.min_uv, conv.max_uv);
    uint8_t *d = dst + (long long)orow * row + (long long)x * ops;
    d[0] = (uint8_t)y; d[1] = (uint8_t)u; d[2] = (uint8_t)v;
    if (out_alpha) d[3] = in.a >= 0 ? q[in.a] : 255;
  }
}

// RGB(A) -> UYVY / YUYV (convert_{rgb,bgr,argb}_to_{uyvy,yuyv}_frame colourspace.c:5129-5700): one thread = one macropixel.
// Cb from the first pixel, Cr from the second, YCbCr tables; YUYV keeps the reference's missing `else` (no max clamp on U, V,
// :2183-2190); optional 16-bit gamma LUT on the 16.8 sums (rgb2uyvy_with_gamma :2146).
__global__ void __launch_bounds__(kBlock) k_rgb_to_packed422(int fmt, const uint8_t *__restrict__ src, int irow, uint8_t *dst, int orow,
                                                             int width_mpx, int height, RgbLayout in, DevConv conv,
                                                             const uint16_t *__restrict__ lut16) {
  __shared__ int32_t t[9][256];
  for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) t[i >> 8][i & 255] = conv.t[i];
  __syncthreads();
  const long long total = (long long)width_mpx * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / width_mpx), m = (int)(it - (long long)row * width_mpx);
    const uint8_t *q = src + (size_t)irow * row + (size_t)m * 2 * in.psize;
    const int r0 = q[in.r], g0 = q[in.g], b0 = q[in.b], r1 = q[in.psize + in.r], g1 = q[in.psize + in.g], b1 = q[in.psize + in.b];
    int au = t[3][r0] + t[4][g0] + t[5][b0], ay0 = t[0][r0] + t[1][g0] + t[2][b0];
    int av = t[6][r1] + t[7][g1] + t[8][b1], ay1 = t[0][r1] + t[1][g1] + t[2][b1];
    if (lut16) {
      au = __ldg(lut16 + (au >> 8)) >> 8; ay0 = __ldg(lut16 + (ay0 >> 8)) >> 8;
      av = __ldg(lut16 + (av >> 8)) >> 8; ay1 = __ldg(lut16 + (ay1 >> 8)) >> 8;
    } else {
      au >>= 16; ay0 >>= 16; av >>= 16; ay1 >>= 16;
    }
    const uint32_t y0 = (uint32_t)clamp_i(ay0, conv.min_y, conv.max_y), y1 = (uint32_t)clamp_i(ay1, conv.min_y, conv.max_y);
    uint32_t w;
    if (fmt == 0) {
      const uint32_t u = (uint32_t)clamp_i(au, conv.min_uv, conv.max_uv), v = (uint32_t)clamp_i(av, conv.min_uv, conv.max_uv);
      w = u | (y0 << 8) | (v << 16) | (y1 << 24);
    } else {
      const uint32_t u = au < conv.min_uv ? (uint32_t)conv.min_uv : ((uint32_t)au & 0xFFu);
      const uint32_t v = av < conv.min_uv ? (uint32_t)conv.min_uv : ((uint32_t)av & 0xFFu);
      w = y0 | (u << 8) | (y1 << 16) | (v << 24);
    }
    *reinterpret_cast<uint32_t *>(dst + (size_t)orow * row + 4 * (size_t)m) = w;
  }
}

// RGB(A) -> planar 4:4:4 (+ alpha plane) (convert_{rgb,bgr,argb}_to_yuvp_frame colourspace.c:5786-6240): one thread = 4 pixels,
// one 32-bit store per plane.
__global__ void __launch_bounds__(kBlock) k_rgb_to_yuv444p(const uint8_t *__restrict__ src, int irow, uint8_t *py, uint8_t *pu, uint8_t *pv,
                                                           uint8_t *pa, int orow, int width, int height, RgbLayout in, DevConv conv) {
  __shared__ int32_t t[9][256];
  for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) t[i >> 8][i & 255] = conv.t[i];
  __syncthreads();
  const int groups = (width + 3) >> 2;
  const bool vec = (((uintptr_t)py | (uintptr_t)pu | (uintptr_t)pv | (uintptr_t)pa) & 3) == 0 && (orow & 3) == 0;
  const long long total = (long long)groups * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int npx = min(4, width - 4 * g);
    const uint8_t *q = src + (size_t)irow * row + (size_t)g * 4 * in.psize;
    uint32_t wy = 0, wu = 0, wv = 0, wa = 0;
    for (int k = 0; k < npx; k++, q += in.psize) {
      const int r = q[in.r], gg = q[in.g], b = q[in.b];
      wy |= (uint32_t)clamp_i((t[0][r] + t[1][gg] + t[2][b]) >> 16, conv.min_y, conv.max_y) << (8 * k);
      wu |= (uint32_t)clamp_i((t[3][r] + t[4][gg] + t[5][b]) >> 16, conv.min_uv, conv.max_uv) << (8 * k);
      wv |= (uint32_t)clamp_i((t[6][r] + t[7][gg] + t[8][b]) >> 16, conv.min_uv, conv.max_uv) << (8 * k);
      wa |= (in.a >= 0 ? (uint32_t)q[in.a] : 255u) << (8 * k);
    }
    const size_t o = (size_t)orow * row + 4 * (size_t)g;
    if (vec && npx == 4) {
      *reinterpret_cast<uint32_t *>(py + o) = wy; *reinterpret_cast<uint32_t *>(pu + o) = wu; *reinterpret_cast<uint32_t *>(pv + o) = wv;
      if (pa) *reinterpret_cast<uint32_t *>(pa + o) = wa;
    } else {
      for (int k = 0; k < npx; k++) {
        py[o + k] = (uint8_t)(wy >> (8 * k)); pu[o + k] = (uint8_t)(wu >> (8 * k)); pv[o + k] = (uint8_t)(wv >> (8 * k));
        if (pa) pa[o + k] = (uint8_t)(wa >> (8 * k));
      }
    }
  }
}

// YUV411 (IYU1, weed-palettes.h:96: macropixel {u2, y0, y1, v2, y2, y3} = 4 pixels sharing one chroma sample) -> RGB(A) / packed
// 4:4:4 / planar 4:4:4 / UYVY / YUYV: convert_yuv411_to_{rgb,bgr,argb,yuv888,yuvp,uyvy,yuyv}_frame colourspace.c:8305-8910.
// A row of w macropixels is w + 1 UNITS: unit 0 = the first two pixels with macropixel 0's chroma, unit w = the last two with
// macropixel w-1's, unit j in between = four pixels (y2, y3 of macropixel j-1, then y0, y1 of macropixel j) whose chroma climbs a
// ladder of averages between p = chroma(j-1) and c = chroma(j), each step one lookup in the reference's 64 KB averaging table
// (avg_chroma(x, y) = cavg[x][y], the table of the frame's clamping):  h = avg(p, c);  qp = avg(h, p);  qc = avg(h, c);
//   pixel 0: avg(qp, p)   pixel 1: avg(qp, c)   pixel 2: avg(qc, p)   pixel 3: avg(qc, c)          (RGB, packed / planar 4:4:4)
//   first macropixel: avg(h, p)   second: avg(h, c)                                                  (UYVY / YUYV)
// Replicated: the planar 4:4:4 and the packed 4:2:2 converters write the FIRST luma of each pair twice (`y0` for both samples,
// :8745-8822, :8867-8893); the planar one also in unit 0.  One thread = one unit.
struct Yuv411Args {
  const uint8_t *src;
  int irow, wmp, height;
  uint8_t *dst[4];
  int orow[4];
  int target;   // 0 RGB (layout `out`), 1 packed 4:4:4 (alpha: 4 bytes), 2 planar 4:4:4 (alpha: plane 3), 3 UYVY, 4 YUYV, 5 planar 4:2:2,
                // 6 planar 4:2:0 (convert_yuv411_to_yuv422_frame :8976 / _to_yuv420_frame :9037, the latter by its intent)
  int alpha;
  int bgr_quirk;   // BGR / BGRA: the row's first pixel and its last two in R, G, B order (convert_yuv411_to_bgr_frame :8445, :8514)
  RgbLayout out;
};

__global__ void __launch_bounds__(kBlock) k_yuv411_to(Yuv411Args A, DevConv conv, const uint8_t *__restrict__ cavg) {
  __shared__ SmemYuvTabs s;
  if (A.target == 0) {
    load_yuv_tabs(s, conv.t);
    __syncthreads();
  }
  auto avg = [&](uint32_t x, uint32_t y) -> uint32_t { return __ldg(cavg + ((x << 8) | y)); };
  // the pixels of unit j of a row: lumas, chromas, how many, the first pixel's index
  auto unit = [&](int row, int j, uint32_t (&ys)[4], uint32_t (&us)[4], uint32_t (&vs)[4], int &npx, int &px0) {
    const uint8_t *r0 = A.src + (long long)A.irow * row;
    if (j == 0 || j == A.wmp) {
      const uint8_t *m = r0 + 6LL * (j == 0 ? 0 : A.wmp - 1);
      npx = 2; px0 = j == 0 ? 0 : 4 * A.wmp - 2;
      ys[0] = j == 0 ? m[1] : m[4];
      ys[1] = j == 0 ? m[2] : m[5];
      if (A.target == 2 && j == 0) ys[1] = ys[0];   // convert_yuv411_to_yuvp_frame writes y0 twice at the row start (:8726-8734)
      us[0] = us[1] = m[0]; vs[0] = vs[1] = m[3];
    } else {
      const uint8_t *mp = r0 + 6LL * (j - 1), *mc = mp + 6;
      npx = 4; px0 = 4 * j - 2;
      const uint32_t pu = mp[0], pv = mp[3], cu = mc[0], cv = mc[3];
      ys[0] = mp[4]; ys[1] = mp[5]; ys[2] = mc[1]; ys[3] = mc[2];
      const uint32_t hu = avg(pu, cu), hv = avg(pv, cv);
      if (A.target >= 3) {       // 4:2:2 chroma: one ladder step per pixel pair; the packed variants write the first luma twice
        us[0] = us[1] = avg(hu, pu); vs[0] = vs[1] = avg(hv, pv);
        us[2] = us[3] = avg(hu, cu); vs[2] = vs[3] = avg(hv, cv);
        if (A.target < 5) { ys[1] = ys[0]; ys[3] = ys[2]; }
      } else {
        const uint32_t qpu = avg(hu, pu), qpv = avg(hv, pv), qcu = avg(hu, cu), qcv = avg(hv, cv);
        us[0] = avg(qpu, pu); vs[0] = avg(qpv, pv);
        us[1] = avg(qpu, cu); vs[1] = avg(qpv, cv);
        us[2] = avg(qcu, pu); vs[2] = avg(qcv, pv);
        us[3] = avg(qcu, cu); vs[3] = avg(qcv, cv);
        if (A.target == 2) { ys[1] = ys[0]; ys[3] = ys[2]; }
      }
    }
  };
  const int units = A.wmp + 1;
  // planar 4:2:0: one thread = one unit of a ROW PAIR (chroma row k = avg_chroma(4:2:2 row 2k, 4:2:2 row 2k + 1), the even row as the table row)
  const int vrows = A.target == 6 ? (A.height + 1) >> 1 : A.height;
  const long long total = (long long)units * vrows;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int vrow = (int)(it / units), j = (int)(it - (long long)vrow * units);
    const int row = A.target == 6 ? 2 * vrow : vrow;
    uint32_t ys[4], us[4], vs[4];
    int npx, px0;
    unit(row, j, ys, us, vs, npx, px0);
    if (A.target == 0) {
      uint8_t *d = A.dst[0] + (long long)A.orow[0] * row + (long long)px0 * A.out.psize;
      for (int k = 0; k < npx; k++, d += A.out.psize) {
        int r, g, b;
        yuv_px(s, nullptr, (int)ys[k], (int)us[k], (int)vs[k], r, g, b);
        const bool swap = A.bgr_quirk && (px0 + k == 0 || px0 + k >= 4 * A.wmp - 2);
        d[swap ? A.out.b : A.out.r] = (uint8_t)r; d[A.out.g] = (uint8_t)g; d[swap ? A.out.r : A.out.b] = (uint8_t)b;
        if (A.out.a >= 0) d[A.out.a] = 255;
      }
    } else if (A.target == 1) {
      const int ps = A.alpha ? 4 : 3;
      uint8_t *d = A.dst[0] + (long long)A.orow[0] * row + (long long)px0 * ps;
      for (int k = 0; k < npx; k++, d += ps) {
        d[0] = (uint8_t)ys[k]; d[1] = (uint8_t)us[k]; d[2] = (uint8_t)vs[k];
        if (A.alpha) d[3] = 255;
      }
    } else if (A.target == 2) {
      for (int k = 0; k < npx; k++) {
        A.dst[0][(long long)A.orow[0] * row + px0 + k] = (uint8_t)ys[k];
        A.dst[1][(long long)A.orow[1] * row + px0 + k] = (uint8_t)us[k];
        A.dst[2][(long long)A.orow[2] * row + px0 + k] = (uint8_t)vs[k];
        if (A.alpha) A.dst[3][(long long)A.orow[3] * row + px0 + k] = 255;
      }
    } else if (A.target >= 5) {
      for (int k = 0; k < npx; k++) A.dst[0][(long long)A.orow[0] * row + px0 + k] = (uint8_t)ys[k];
      uint32_t cu[2] = {us[0], us[2]}, cv[2] = {vs[0], vs[2]};
      if (A.target == 6 && row + 1 < A.height) {
        uint32_t yb[4], ub[4], vb[4];
        int nb, pb;
        unit(row + 1, j, yb, ub, vb, nb, pb);
        for (int k = 0; k < nb; k++) A.dst[0][(long long)A.orow[0] * (row + 1) + pb + k] = (uint8_t)yb[k];
        cu[0] = avg(cu[0], ub[0]); cv[0] = avg(cv[0], vb[0]);
        if (npx == 4) { cu[1] = avg(cu[1], ub[2]); cv[1] = avg(cv[1], vb[2]); }
      }
      for (int k = 0; k < npx >> 1; k++) {
        A.dst[1][(long long)A.orow[1] * vrow + (px0 >> 1) + k] = (uint8_t)cu[k];
        A.dst[2][(long long)A.orow[2] * vrow + (px0 >> 1) + k] = (uint8_t)cv[k];
      }
    } else {
      uint8_t *d = A.dst[0] + (long long)A.orow[0] * row + (long long)(px0 >> 1) * 4;
      for (int k = 0; k < npx; k += 2, d += 4) {
        if (A.target == 3) { d[0] = (uint8_t)us[k]; d[1] = (uint8_t)ys[k]; d[2] = (uint8_t)vs[k]; d[3] = (uint8_t)ys[k + 1]; }
        else { d[0] = (uint8_t)ys[k]; d[1] = (uint8_t)us[k]; d[2] = (uint8_t)ys[k + 1]; d[3] = (uint8_t)vs[k]; }
      }
    }
  }
}

// YUV -> YUV411, one thread = one output macropixel.  mode 0 UYVY / 1 YUYV (convert_{uyvy,yuyv}_to_yuv411_frame :7973-8032), 2 YUV420P /
// 3 YUV422P (convert_yuv420_to_yuv411_frame :9148: 4:2:0 folds every even row r >= 2 into the chroma of row r - 1), 4 YUV888 / 5 YUVA8888
// (convert_yuv888_to_yuv411_frame :8272: plain (sum of four) >> 2), 6 planar 4:4:4 (convert_yuvp_to_yuv411_frame :7755)
struct ToYuv411Args {
  const uint8_t *src[3];
  int irow[3];
  int wmp, height, mode;
  uint8_t *dst;
  int orow;
};

__global__ void __launch_bounds__(kBlock) k_to_yuv411(ToYuv411Args A, const uint8_t *__restrict__ cavg) {
  auto avg = [&](uint32_t x, uint32_t y) -> uint32_t { return __ldg(cavg + ((x << 8) | y)); };
  auto chroma_42x = [&](int row, int j, uint32_t &u, uint32_t &v) {   // modes 2 / 3: the macropixel's own chroma
    const long long cr = A.mode == 2 ? row >> 1 : row;
    const uint8_t *pu = A.src[1] + A.irow[1] * cr + 2LL * j, *pv = A.src[2] + A.irow[2] * cr + 2LL * j;
    u = avg(pu[0], pu[1]); v = avg(pv[0], pv[1]);
  };
  const long long total = (long long)A.wmp * A.height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / A.wmp), j = (int)(it - (long long)row * A.wmp);
    uint32_t u, v, y[4];
    if (A.mode <= 1) {
      const uint8_t *m = A.src[0] + (long long)A.irow[0] * row + 8LL * j;
      const int yo = A.mode == 0 ? 1 : 0, uo = A.mode == 0 ? 0 : 1;
      y[0] = m[yo]; y[1] = m[yo + 2]; y[2] = m[4 + yo]; y[3] = m[6 + yo];
      u = avg(m[uo], m[4 + uo]); v = avg(m[uo + 2], m[6 + uo]);
    } else if (A.mode <= 3) {
      const uint8_t *py = A.src[0] + (long long)A.irow[0] * row + 4LL * j;
      y[0] = py[0]; y[1] = py[1]; y[2] = py[2]; y[3] = py[3];
      chroma_42x(row, j, u, v);
      if (A.mode == 2 && (row & 1) && row + 1 < A.height) {   // the fold of row + 1 (even, >= 2) into this row (:9176-9179)
        uint32_t u2, v2;
        chroma_42x(row + 1, j, u2, v2);
        u = avg(u, u2); v = avg(v, v2);
      }
    } else if (A.mode <= 5) {
      const int ps = A.mode == 4 ? 3 : 4;
      const uint8_t *q = A.src[0] + (long long)A.irow[0] * row + 4LL * ps * j;
      y[0] = q[0]; y[1] = q[ps]; y[2] = q[2 * ps]; y[3] = q[3 * ps];
      u = ((uint32_t)q[1] + q[ps + 1] + q[2 * ps + 1] + q[3 * ps + 1]) >> 2;
      v = ((uint32_t)q[2] + q[ps + 2] + q[2 * ps + 2] + q[3 * ps + 2]) >> 2;
    } else {
      const long long o = 4LL * j;
      const uint8_t *py = A.src[0] + (long long)A.irow[0] * row + o, *pu = A.src[1] + (long long)A.irow[1] * row + o,
                    *pv = A.src[2] + (long long)A.irow[2] * row + o;
      y[0] = py[0]; y[1] = py[1]; y[2] = py[2]; y[3] = py[3];
      u = avg(avg(pu[0], pu[1]), avg(pu[2], pu[3])); v = avg(avg(pv[0], pv[1]), avg(pv[2], pv[3]));
    }
    uint8_t *d = A.dst + (long long)A.orow * row + 6LL * j;
    d[0] = (uint8_t)u; d[1] = (uint8_t)y[0]; d[2] = (uint8_t)y[1]; d[3] = (uint8_t)v; d[4] = (uint8_t)y[2]; d[5] = (uint8_t)y[3];
  }
}

}  // namespace

cudaError_t launch_yuv_planar_to_rgb(const Launch &L, const YuvToRgbArgs &a) {
  const int groups = (a.width + 3) >> 2;
  const int njobs = a.is_422 ? a.height : a.src.ch + 1;
  k_yuv_planar_to_rgb<<<grid_for(L, (long long)groups * njobs), kBlock, 0, L.stream>>>(a);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_packed422_to_rgb(const Launch &L, int fmt, CImg src, Img dst, int width_mpx, int height,
                                    RgbLayout out, DevConv conv) {
  k_packed422_to_rgb<<<grid_for(L, (long long)((width_mpx + 1) >> 1) * height), kBlock, 0, L.stream>>>(
      fmt, src.p, src.rs, dst.p, dst.rs, width_mpx, height, out, conv);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_yuv888_to_rgb(const Launch &L, CImg src, Img dst, int width, int height, int in_alpha,
                                 RgbLayout out, DevConv conv) {
  k_yuv888_to_rgb<<<grid_for(L, (long long)width * height), kBlock, 0, L.stream>>>(src.p, src.rs, dst.p, dst.rs, width, height,
                                                                                 in_alpha, out, conv);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_rgb_to_yuv888(const Launch &L, CImg src, Img dst, int width, int height, RgbLayout in,
                                 int out_alpha, DevConv conv) {
  width = (width >> 1) << 1;
  k_rgb_to_yuv888<<<grid_for(L, (long long)width * height), kBlock, 0, L.stream>>>(src.p, src.rs, dst.p, dst.rs, width, height, in,
                                                                                 out_alpha, conv);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// RGB(A) -> planar 4:2:0 / 4:2:2 (convert_{rgb,bgr}_to_yuv420_frame colourspace.c:6250 / :6385).  One rgb2uyvy macropixel
// (:2162) per pixel pair: Y of both pixels, Cb of the first, Cr of the second.  4:2:0 chroma as the reference's pointer dance
// leaves it (:6291-6306):  C[c] = cavg[C(2c+2)][C(2c+1)],  last row C(h-1).
// One thread = two macropixels (4 pixels) of the luma rows (2g-1, 2g) of row group g = 0 .. h/2 (4:2:0) or of one row (4:2:2):
// one 32-bit luma store per row, one 16-bit store per chroma plane.
__global__ void __launch_bounds__(kBlock) k_rgb_to_yuv420p(const uint8_t *__restrict__ src, int irow, uint8_t *py, uint8_t *pu, uint8_t *pv,
                                                           int rs_y, int rs_u, int rs_v, int width, int height, RgbLayout in, int is_422,
                                                           DevConv conv, const uint8_t *__restrict__ cavg) {
  __shared__ int32_t t[9][256];
  for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) t[i >> 8][i & 255] = conv.t[i];
  __syncthreads();
  const int quads = (width + 3) >> 2;                       // width is even: the last quad may hold one macropixel
  const int groups = is_422 ? height : (height >> 1) + 1;
  const long long total = (long long)quads * groups;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int g = (int)(it / quads), q = (int)(it - (long long)g * quads);
    const int x = 4 * q, nm = min(2, (width - x) >> 1);     // macropixels of this thread
    const int ra = is_422 ? g : 2 * g - 1, rb = is_422 ? -1 : 2 * g;
    uint32_t cu[2][2], cv[2][2];                            // [row a / b][macropixel]
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const int row = rr ? rb : ra;
      if (row < 0 || row >= height) continue;
      const uint8_t *p = src + (size_t)irow * row + (size_t)x * in.psize;
      uint32_t yw = 0;
#pragma unroll
      for (int m = 0; m < 2; m++) {
        if (m < nm) {
          const uint8_t *p0 = p + 2 * m * in.psize, *p1 = p0 + in.psize;
          const int r0 = p0[in.r], g0 = p0[in.g], b0 = p0[in.b], r1 = p1[in.r], g1 = p1[in.g], b1 = p1[in.b];
          const int y0 = clamp_i((t[0][r0] + t[1][g0] + t[2][b0]) >> 16, conv.min_y, conv.max_y);
          const int y1 = clamp_i((t[0][r1] + t[1][g1] + t[2][b1]) >> 16, conv.min_y, conv.max_y);
          cu[rr][m] = (uint32_t)clamp_i((t[3][r0] + t[4][g0] + t[5][b0]) >> 16, conv.min_uv, conv.max_uv);
          cv[rr][m] = (uint32_t)clamp_i((t[6][r1] + t[7][g1] + t[8][b1]) >> 16, conv.min_uv, conv.max_uv);
          yw |= ((uint32_t)y0 | ((uint32_t)y1 << 8)) << (16 * m);
        }
      }
      uint8_t *yo = py + (size_t)rs_y * row + x;
      if (nm == 2) *reinterpret_cast<uint32_t *>(yo) = yw;
      else *reinterpret_cast<uint16_t *>(yo) = (uint16_t)yw;
    }
    // chroma row of this group
    int crow;
    bool have_a, have_b;
    if (is_422) { crow = g; have_a = true; have_b = false; }
    else { crow = g - 1; have_a = ra >= 0; have_b = rb < height; }
    if (crow < 0 || !have_a) continue;
    uint32_t uo = 0, vo = 0;
#pragma unroll
    for (int m = 0; m < 2; m++) {
      if (m < nm) {
        uint32_t u = cu[0][m], v = cv[0][m];
        if (have_b) {  // avg_chroma(new = row 2c+2, old = row 2c+1)
          u = __ldg(cavg + ((cu[1][m] << 8) | u));
          v = __ldg(cavg + ((cv[1][m] << 8) | v));
        }
        uo |= u << (8 * m); vo |= v << (8 * m);
      }
    }
    uint8_t *uop = pu + (size_t)rs_u * crow + (x >> 1), *vop = pv + (size_t)rs_v * crow + (x >> 1);
    if (nm == 2) { *reinterpret_cast<uint16_t *>(uop) = (uint16_t)uo; *reinterpret_cast<uint16_t *>(vop) = (uint16_t)vo; }
    else { *uop = (uint8_t)uo; *vop = (uint8_t)vo; }
  }
}

}  // namespace pe

namespace pe {
cudaError_t launch_rgb_to_packed422(const Launch &L, int fmt, CImg src, Img dst, int width_px, int height, RgbLayout in, DevConv conv,
                                    const uint16_t *lut16_dev) {
  const int mpx = width_px >> 1;
  if (mpx <= 0 || height <= 0) return cudaSuccess;
  k_rgb_to_packed422<<<(int)std::min<long long>(((long long)mpx * height + 255) / 256, (long long)L.sm_count * 8), 256, 0, L.stream>>>(
      fmt, src.p, src.rs, dst.p, dst.rs, mpx, height, in, conv, lut16_dev);
  if (L.launch_counter) ++*L.launch_counter;
  return cudaGetLastError();
}

cudaError_t launch_rgb_to_yuv444p(const Launch &L, CImg src, uint8_t *const planes[4], int orow, int width, int height, RgbLayout in,
                                  DevConv conv) {
  width = (width >> 1) << 1;  // the reference drops an odd last column (colourspace.c:5866)
  if (width <= 0 || height <= 0) return cudaSuccess;
  const long long work = (long long)((width + 3) >> 2) * height;
  k_rgb_to_yuv444p<<<(int)std::min<long long>((work + 255) / 256, (long long)L.sm_count * 8), 256, 0, L.stream>>>(
      src.p, src.rs, planes[0], planes[1], planes[2], planes[3], orow, width, height, in, conv);
  if (L.launch_counter) ++*L.launch_counter;
  return cudaGetLastError();
}
cudaError_t launch_rgb_to_yuv420p(const Launch &L, CImg src, uint8_t *const planes[3], const int rowstrides[3], int width, int height,
                                  RgbLayout in, int is_422, DevConv conv, const uint8_t *cavg_dev) {
  width &= ~1; height &= ~1;
  if (width <= 0 || height <= 0) return cudaSuccess;
  const long long work = (long long)((width + 3) >> 2) * (is_422 ? height : (height >> 1) + 1);
  k_rgb_to_yuv420p<<<(int)std::min<long long>((work + 255) / 256, (long long)L.sm_count * 8), 256, 0, L.stream>>>(
      src.p, src.rs, planes[0], planes[1], planes[2], rowstrides[0], rowstrides[1], rowstrides[2], width, height, in, is_422, conv,
      cavg_dev);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

}  // namespace pe


namespace pe {
namespace {
// RGB(A) / BGR(A) / ARGB -> YUV411: convert_{rgb,bgr,argb}_to_yuv411_frame colourspace.c:6499-6614, rgb2_411 :2323.  One thread = one
// macropixel = 4 pixels: luma per pixel, chroma = (sum of the four pixels' >> 16 terms) >> 2, clamped.
__global__ void __launch_bounds__(kBlock) k_rgb_to_yuv411(const uint8_t *__restrict__ src, int irow, uint8_t *dst, int orow, int wmp, int height,
                                                          RgbLayout in, DevConv conv) {
  __shared__ int32_t t[9][256];
  for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) t[i >> 8][i & 255] = conv.t[i];
  __syncthreads();
  const long long total = (long long)wmp * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / wmp), j = (int)(it - (long long)row * wmp);
    const uint8_t *q = src + (long long)irow * row + (long long)j * 4 * in.psize;
    int su = 0, sv = 0, yy[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int r = q[k * in.psize + in.r], g = q[k * in.psize + in.g], b = q[k * in.psize + in.b];
      yy[k] = min(max((t[0][r] + t[1][g] + t[2][b]) >> 16, conv.min_y), conv.max_y);
      su += (t[3][r] + t[4][g] + t[5][b]) >> 16;
      sv += (t[6][r] + t[7][g] + t[8][b]) >> 16;
    }
    uint8_t *d = dst + (long long)orow * row + 6LL * j;
    d[0] = (uint8_t)min(max(su >> 2, conv.min_uv), conv.max_uv);
    d[1] = (uint8_t)yy[0]; d[2] = (uint8_t)yy[1];
    d[3] = (uint8_t)min(max(sv >> 2, conv.min_uv), conv.max_uv);
    d[4] = (uint8_t)yy[2]; d[5] = (uint8_t)yy[3];
  }
}
}  // namespace

cudaError_t launch_rgb_to_yuv411(const Launch &L, CImg src, Img dst, int width_mpx, int height, RgbLayout in, DevConv conv) {
  if (width_mpx <= 0 || height <= 0) return cudaSuccess;
  k_rgb_to_yuv411<<<grid_for(L, (long long)width_mpx * height), kBlock, 0, L.stream>>>(src.p, src.rs, dst.p, dst.rs, width_mpx, height, in, conv);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_to_yuv411(const Launch &L, int mode, const uint8_t *const src[3], const int irow[3], int width_mpx, int height, Img dst,
                             const uint8_t *cavg_dev) {
  if (width_mpx <= 0 || height <= 0) return cudaSuccess;
  ToYuv411Args A;
  for (int i = 0; i < 3; i++) { A.src[i] = src[i]; A.irow[i] = irow[i]; }
  A.wmp = width_mpx; A.height = height; A.mode = mode; A.dst = dst.p; A.orow = dst.rs;
  k_to_yuv411<<<grid_for(L, (long long)width_mpx * height), kBlock, 0, L.stream>>>(A, cavg_dev);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_yuv411_to(const Launch &L, CImg src, int width_mpx, int height, uint8_t *const dst[4], const int orow[4], int target,
                             int alpha, RgbLayout out, int bgr_quirk, DevConv conv, const uint8_t *cavg_dev) {
  if (width_mpx <= 0 || height <= 0) return cudaSuccess;
  Yuv411Args A;
  A.bgr_quirk = bgr_quirk;
  A.src = src.p; A.irow = src.rs; A.wmp = width_mpx; A.height = height;
  for (int i = 0; i < 4; i++) { A.dst[i] = dst[i]; A.orow[i] = orow[i]; }
  A.target = target; A.alpha = alpha; A.out = out;
  k_yuv411_to<<<grid_for(L, (long long)(width_mpx + 1) * (target == 6 ? (height + 1) / 2 : height)), kBlock, 0, L.stream>>>(A, conv, cavg_dev);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}
}  // namespace pe
