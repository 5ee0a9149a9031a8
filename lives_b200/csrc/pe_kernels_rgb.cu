// pe_kernels_rgb.cu -- packed-RGB kernels: palette permutation, gamma LUT, premultiply, letterbox,
// effect blends, alpha-over, frame statistics.  sm_100a.
//
// All of these are HBM-bound byte streams (<= ~20 integer ops per byte): the design rules are
// 128-bit coalesced accesses, L1-bypassing streaming loads/stores, grids sized as a multiple of the
// SM count with grid-stride loops, and look-up tables staged in shared memory.
#include <cstdlib>

#include "pe_device.cuh"
#include "pe_kernels.h"

namespace pe {

namespace {

constexpr int kBlock = 256;

inline int grid_for(const Launch &L, long long work_items, int per_sm = 8) {
  long long blocks = (work_items + kBlock - 1) / kBlock;
  long long cap = (long long)L.sm_count * per_sm;  // a whole number of waves of resident CTAs
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

// =====================================================================================================
// RGB <-> RGB.  One thread = 4 pixels.  3-byte pixels travel as 3 x u32, 4-byte pixels as one uint4.
// Replaces convert_swap3 / swap4 / addpost / addpre / delpost / delpre / swap3* / swapprepost
// (colourspace.c:9259-10515) as dispatched by convert_layer_palette_full (:12370-12556).
// =====================================================================================================

struct PermuteParams {
  const uint8_t *src;
  uint8_t *dst;
  int irow, orow, width, height;
  int in_r, in_g, in_b, in_a;      // byte offsets inside an input pixel (in_a = -1: none)
  int out_r, out_g, out_b, out_a;  // byte offsets inside an output pixel
  uint32_t sel;                    // PRMT selector building an output pixel from (in_pixel, 0xFFFFFFFF)
  const uint8_t *lut;              // device LUT or nullptr
  int vec_ok;                      // rows are 4-byte (3 bpp) / 16-byte (4 bpp) aligned on both sides
  int vec16_ok;                    // 3 bpp -> 3 bpp: rows 16-byte aligned on both sides, width a multiple of 16
};

template <bool HAS_LUT>
__device__ __forceinline__ uint32_t permute_pixel(uint32_t pix, const PermuteParams &P, const uint8_t *s_lut) {
  if (!HAS_LUT) return __byte_perm(pix, 0xFFFFFFFFu, P.sel);
  uint32_t r = s_lut[byte_of(pix, P.in_r)], g = s_lut[byte_of(pix, P.in_g)], b = s_lut[byte_of(pix, P.in_b)];
  uint32_t o = (r << (8 * P.out_r)) | (g << (8 * P.out_g)) | (b << (8 * P.out_b));
  if (P.out_a >= 0) o |= (P.in_a >= 0 ? byte_of(pix, P.in_a) : 255u) << (8 * P.out_a);
  return o;
}

// frames of one batched launch (same geometry, strides and palettes): blockIdx.y selects the frame
constexpr int kRgbBatch = 256;  // frames per launch: 4 KB of kernel parameters (CUDA 12.1+ takes up to 32 764 bytes on sm_70+)
struct RgbFrameList {
  const uint8_t *src[kRgbBatch];
  uint8_t *dst[kRgbBatch];
};

template <int IPS, int OPS, bool HAS_LUT>
__global__ void __launch_bounds__(kBlock) k_rgb_to_rgb(const PermuteParams P, const __grid_constant__ RgbFrameList FL) {
  __shared__ uint8_t s_lut[256];
  const uint8_t *const f_src = FL.src[blockIdx.y];
  uint8_t *const f_dst = FL.dst[blockIdx.y];
  if (HAS_LUT) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = P.lut[i];
    __syncthreads();
  }
  if (IPS == 3 && OPS == 3 && P.vec16_ok) {
    // 3-byte pixels with 16-byte aligned rows of whole 16-pixel groups: a thread moves 48 bytes as 3 x 128 bits (a warp's 32-bit
    // accesses at a 12-byte stride cost three times the LSU wavefronts of the same bytes moved as 128-bit vectors)
    const int groups16 = P.width >> 4;
    const long long total16 = (long long)groups16 * P.height;
    // (32-bit index arithmetic: a 64-bit division per 48 bytes was a third of this loop's instructions; frames are far below 2^31 groups)
    const uint32_t tot = (uint32_t)total16, ug = (uint32_t)groups16, T = (uint32_t)gridDim.x * blockDim.x;
    for (uint32_t it = (uint32_t)blockIdx.x * blockDim.x + threadIdx.x; it < tot; it += T) {
      const uint32_t row = it / ug, g = it - row * ug;
      const uint8_t *s = f_src + (long long)row * P.irow + (long long)g * 48;
      uint8_t *d = f_dst + (long long)row * P.orow + (long long)g * 48;
      const uint4 a = ld_u4(s), b = ld_u4(s + 16), c = ld_u4(s + 32);  // may be in place: plain loads
      uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
        const uint32_t p0 = permute_pixel<HAS_LUT>(w0, P, s_lut), p1 = permute_pixel<HAS_LUT>(__byte_perm(w0, w1, 0x0543), P, s_lut);
        const uint32_t p2 = permute_pixel<HAS_LUT>(__byte_perm(w1, w2, 0x0432), P, s_lut), p3 = permute_pixel<HAS_LUT>(w2 >> 8, P, s_lut);
        w[3 * q] = __byte_perm(p0, p1, 0x4210); w[3 * q + 1] = __byte_perm(p1, p2, 0x5421); w[3 * q + 2] = __byte_perm(p2, p3, 0x6542);
      }
      *(uint4 *)d = make_uint4(w[0], w[1], w[2], w[3]);
      *(uint4 *)(d + 16) = make_uint4(w[4], w[5], w[6], w[7]);
      *(uint4 *)(d + 32) = make_uint4(w[8], w[9], w[10], w[11]);
    }
    return;
  }
  const int groups = (P.width + 3) >> 2;
  const long long total = (long long)groups * P.height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const uint8_t *s = f_src + (long long)row * P.irow + (long long)g * 4 * IPS;
    uint8_t *d = f_dst + (long long)row * P.orow + (long long)g * 4 * OPS;
    const int npx = min(4, P.width - g * 4);
    uint32_t pix[4];
    if (P.vec_ok && npx == 4) {
      if (IPS == 4) {
        const uint4 v = ld_u4(s);  // may be in place: plain load
        pix[0] = v.x; pix[1] = v.y; pix[2] = v.z; pix[3] = v.w;
      } else {
        const uint32_t w0 = *(const uint32_t *)(s), w1 = *(const uint32_t *)(s + 4), w2 = *(const uint32_t *)(s + 8);
        pix[0] = w0;
        pix[1] = __byte_perm(w0, w1, 0x0543);
        pix[2] = __byte_perm(w1, w2, 0x0432);
        pix[3] = w2 >> 8;
      }
#pragma unroll
      for (int k = 0; k < 4; k++) pix[k] = permute_pixel<HAS_LUT>(pix[k], P, s_lut);
      if (OPS == 4) {
        *(uint4 *)d = make_uint4(pix[0], pix[1], pix[2], pix[3]);
      } else {
        *(uint32_t *)(d) = __byte_perm(pix[0], pix[1], 0x4210);
        *(uint32_t *)(d + 4) = __byte_perm(pix[1], pix[2], 0x5421);
        *(uint32_t *)(d + 8) = __byte_perm(pix[2], pix[3], 0x6542);
      }
    } else {
      // ragged tail / unaligned frame: byte path
      for (int k = 0; k < npx; k++) {
        uint32_t p = 0;
        for (int b = 0; b < IPS; b++) p |= (uint32_t)s[k * IPS + b] << (8 * b);
        p = permute_pixel<HAS_LUT>(p, P, s_lut);
        for (int b = 0; b < OPS; b++) d[k * OPS + b] = (uint8_t)(p >> (8 * b));
      }
    }
  }
}

}  // namespace

// n frames of the same geometry, strides and palettes in one launch per 128 (srcs[i] may equal dsts[i]: in place)
cudaError_t launch_rgb_to_rgb_batch(const Launch &L, const uint8_t *const *srcs, int irow, uint8_t *const *dsts, int orow, int n, int width,
                                    int height, RgbLayout in, RgbLayout out, const uint8_t *lut8_dev) {
  PermuteParams P;
  P.src = srcs[0]; P.dst = dsts[0]; P.irow = irow; P.orow = orow; P.width = width; P.height = height;
  P.in_r = in.r; P.in_g = in.g; P.in_b = in.b; P.in_a = in.a;
  P.out_r = out.r; P.out_g = out.g; P.out_b = out.b; P.out_a = out.a;
  P.lut = lut8_dev;
  // selector nibble per output byte: index of the source byte, or 4 (= a byte of 0xFFFFFFFF) for opaque alpha
  uint32_t nib[4] = {0, 0, 0, 0};
  nib[out.r] = in.r; nib[out.g] = in.g; nib[out.b] = in.b;
  if (out.a >= 0) nib[out.a] = in.a >= 0 ? in.a : 4;
  P.sel = nib[0] | (nib[1] << 4) | (nib[2] << 8) | (nib[3] << 12);
  const int ia = in.psize == 4 ? 16 : 4, oa = out.psize == 4 ? 16 : 4;
  bool vec = (irow % ia == 0) && (orow % oa == 0);
  for (int i = 0; i < n && vec; i++) vec = ((uintptr_t)srcs[i] % ia == 0) && ((uintptr_t)dsts[i] % oa == 0);
  P.vec_ok = vec;
  bool vec16 = in.psize == 3 && out.psize == 3 && !(width & 15) && !(irow & 15) && !(orow & 15) && getenv("PE_RGB_NO_VEC16") == nullptr &&
               (long long)(width >> 4) * height < (1ll << 31);   // (the kernel's 32-bit group index)
  for (int i = 0; i < n && vec16; i++) vec16 = ((uintptr_t)srcs[i] % 16 == 0) && ((uintptr_t)dsts[i] % 16 == 0);
  P.vec16_ok = vec16;
  const long long work = vec16 ? (long long)(width >> 4) * height : (long long)((width + 3) >> 2) * height;
  for (int base = 0; base < n; base += kRgbBatch) {
    RgbFrameList fl;
    const int cnt = n - base < kRgbBatch ? n - base : kRgbBatch;
    for (int i = 0; i < kRgbBatch; i++) { fl.src[i] = srcs[base + (i < cnt ? i : 0)]; fl.dst[i] = dsts[base + (i < cnt ? i : 0)]; }
    // the frames of a launch share the grid: about 8 CTAs per SM over all of them
    int gx = grid_for(L, work);
    const int cap = (L.sm_count * 8 + cnt - 1) / cnt;
    if (gx > cap) gx = cap < 1 ? 1 : cap;
    const dim3 grid(gx, cnt);
#define PE_PERM(I, O) \
  do { if (lut8_dev) k_rgb_to_rgb<I, O, true><<<grid, kBlock, 0, L.stream>>>(P, fl); \
       else k_rgb_to_rgb<I, O, false><<<grid, kBlock, 0, L.stream>>>(P, fl); } while (0)
    if (in.psize == 3 && out.psize == 3) PE_PERM(3, 3);
    else if (in.psize == 3 && out.psize == 4) PE_PERM(3, 4);
    else if (in.psize == 4 && out.psize == 3) PE_PERM(4, 3);
    else PE_PERM(4, 4);
#undef PE_PERM
    PE_COUNT_LAUNCH(L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_rgb_to_rgb(const Launch &L, CImg src, Img dst, int width, int height, RgbLayout in, RgbLayout out,
                              const uint8_t *lut8_dev) {
  const uint8_t *s1[1] = {src.p};
  uint8_t *d1[1] = {dst.p};
  return launch_rgb_to_rgb_batch(L, s1, src.rs, d1, dst.rs, 1, width, height, in, out, lut8_dev);
}

// =====================================================================================================
// 8-bit gamma LUT over a rectangle, in place (gamma_convert_layer_thread colourspace.c:14034-14062):
// the first min(3, psize) bytes of every pixel, starting one byte later for ARGB.
// =====================================================================================================

namespace {

struct LutRectParams {
  uint8_t *p;
  int rs, psize, x, y, width, height;
  uint32_t keep_mask;  // 4-byte palettes: byte lanes that are NOT transformed (the alpha byte)
  const uint8_t *lut;
};

__device__ __forceinline__ uint32_t lut_word(uint32_t w, const uint8_t *s_lut, uint32_t keep_mask) {
  uint32_t o = pack4(s_lut[byte_of(w, 0)], s_lut[byte_of(w, 1)], s_lut[byte_of(w, 2)], s_lut[byte_of(w, 3)]);
  return (o & ~keep_mask) | (w & keep_mask);
}

__global__ void __launch_bounds__(kBlock) k_lut8_rect(const LutRectParams P) {
  __shared__ uint8_t s_lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = P.lut[i];
  __syncthreads();
  // byte span of the rectangle inside a row
  const int b0 = P.x * P.psize, b1 = (P.x + P.width) * P.psize;
  // 16-byte chunks; for 3-byte pixels every byte is colour so chunks need no pixel alignment, for 4-byte pixels
  // a 16-byte aligned chunk always holds 4 whole pixels (rows are 16-byte aligned when vec is used)
  const bool vec = ((uintptr_t)P.p % 16 == 0) && (P.rs % 16 == 0);
  const int c0 = vec ? (b0 & ~15) : b0, nchunks = vec ? ((b1 - c0 + 15) >> 4) : 0;
  if (vec) {
    const long long total = (long long)nchunks * P.height;
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / nchunks), c = (int)(it - (long long)row * nchunks);
      uint8_t *q = P.p + (long long)(P.y + row) * P.rs + c0 + c * 16;
      const int lo = c0 + c * 16;
      if (lo >= b0 && lo + 16 <= b1) {
        uint4 v = *(uint4 *)q;
        v.x = lut_word(v.x, s_lut, P.keep_mask); v.y = lut_word(v.y, s_lut, P.keep_mask);
        v.z = lut_word(v.z, s_lut, P.keep_mask); v.w = lut_word(v.w, s_lut, P.keep_mask);
        *(uint4 *)q = v;
      } else {
        for (int b = max(lo, b0); b < min(lo + 16, b1); b++) {
          if (P.psize == 4 && ((P.keep_mask >> (8 * (b & 3))) & 0xFF)) continue;
          q[b - lo] = s_lut[q[b - lo]];
        }
      }
    }
  } else {
    const long long total = (long long)(b1 - b0) * P.height;
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / (b1 - b0)), b = b0 + (int)(it - (long long)row * (b1 - b0));
      if (P.psize == 4 && ((P.keep_mask >> (8 * (b & 3))) & 0xFF)) continue;
      uint8_t *q = P.p + (long long)(P.y + row) * P.rs + b;
      *q = s_lut[*q];
    }
  }
}

}  // namespace

cudaError_t launch_lut8_rect(const Launch &L, Img img, RgbLayout lay, int x, int y, int width, int height,
                             const uint8_t *lut8_dev) {
  LutRectParams P;
  P.p = img.p; P.rs = img.rs; P.psize = lay.psize; P.x = x; P.y = y; P.width = width; P.height = height;
  P.keep_mask = lay.a >= 0 ? (0xFFu << (8 * lay.a)) : 0u;
  P.lut = lut8_dev;
  const long long work = ((long long)width * lay.psize + 15) / 16 * height;
  k_lut8_rect<<<grid_for(L, work), kBlock, 0, L.stream>>>(P);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// =====================================================================================================
// alpha premultiply / un-premultiply, in place (alpha_premult colourspace.c:11968-12106).
// tab0/1/2: the 64 KB [alpha][value] table for colour byte 0/1/2 of the pixel (RGB: all the same table;
// clamped YUVA8888: Y table + UV table twice).  yuva_fwd_quirk: index the U and V tables with the pixel's
// Y byte, as colourspace.c:12093-12094 does.
// =====================================================================================================

namespace {

__global__ void __launch_bounds__(kBlock) k_premult(uint8_t *p, int rs, int width, int height, int coffs, int aoffs,
                                                    const uint8_t *__restrict__ t0, const uint8_t *__restrict__ t1,
                                                    const uint8_t *__restrict__ t2, int quirk) {
  const long long total = (long long)width * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / width), x = (int)(it - (long long)row * width);
    uint32_t *q = (uint32_t *)(p + (long long)row * rs) + x;
    const uint32_t w = *q;
    const uint32_t a = byte_of(w, aoffs);
    const uint32_t c0 = byte_of(w, coffs), c1 = byte_of(w, coffs + 1), c2 = byte_of(w, coffs + 2);
    const uint32_t n0 = __ldg(t0 + a * 256 + c0);
    // quirk: U and V are looked up with the Y byte AFTER it was rewritten (colourspace.c:12092-12094)
    const uint32_t n1 = __ldg(t1 + a * 256 + (quirk ? n0 : c1));
    const uint32_t n2 = __ldg(t2 + a * 256 + (quirk ? n0 : c2));
    *q = (a << (8 * aoffs)) | (n0 << (8 * coffs)) | (n1 << (8 * (coffs + 1))) | (n2 << (8 * (coffs + 2)));
  }
}

}  // namespace

namespace {
// YUVA4444P (colourspace.c:12001-12049): 4 samples of each plane per thread, every plane through its own table with the sample's alpha
__global__ void __launch_bounds__(kBlock) k_premult_planar(uint8_t *py, uint8_t *pu, uint8_t *pv, const uint8_t *__restrict__ pa, int rs_y,
                                                           int rs_u, int rs_v, int rs_a, int width, int height,
                                                           const uint8_t *__restrict__ ty, const uint8_t *__restrict__ tc) {
  const int groups = (width + 3) >> 2;
  const long long total = (long long)groups * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), x0 = 4 * (int)(it - (long long)row * groups);
    const int n = min(4, width - x0);
    for (int k = 0; k < n; k++) {
      const uint32_t a = pa[(long long)rs_a * row + x0 + k] * 256u;
      uint8_t *y = py + (long long)rs_y * row + x0 + k, *u = pu + (long long)rs_u * row + x0 + k, *v = pv + (long long)rs_v * row + x0 + k;
      *y = __ldg(ty + a + *y); *u = __ldg(tc + a + *u); *v = __ldg(tc + a + *v);
    }
  }
}
}  // namespace

cudaError_t launch_premult_planar(const Launch &L, uint8_t *const planes[4], const int rowstrides[4], int width, int height,
                                  const uint8_t *tab_y, const uint8_t *tab_c) {
  if (width <= 0 || height <= 0) return cudaSuccess;
  k_premult_planar<<<grid_for(L, (long long)((width + 3) >> 2) * height), kBlock, 0, L.stream>>>(
      planes[0], planes[1], planes[2], planes[3], rowstrides[0], rowstrides[1], rowstrides[2], rowstrides[3], width, height, tab_y, tab_c);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_premult(const Launch &L, Img img, int width, int height, int coffs, int ncol, int aoffs,
                           const uint8_t *tab0, const uint8_t *tab1, const uint8_t *tab2, int yuva_fwd_quirk) {
  (void)ncol;
  k_premult<<<grid_for(L, (long long)width * height), kBlock, 0, L.stream>>>(img.p, img.rs, width, height, coffs, aoffs,
                                                                           tab0, tab1, tab2, yuva_fwd_quirk);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// =====================================================================================================
// fill / 2-D copy / letterbox (blank_frame colourspace.c:11213, letterbox_layer :15343-15567)
// =====================================================================================================

namespace {

__global__ void __launch_bounds__(kBlock) k_fill(uint8_t *p, int rs, int width, int height, int psize, uint32_t pixel) {
  const long long total = (long long)width * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / width), x = (int)(it - (long long)row * width);
    uint8_t *q = p + (long long)row * rs + (long long)x * psize;
    if (psize == 4) *(uint32_t *)q = pixel;
    else { q[0] = (uint8_t)pixel; q[1] = (uint8_t)(pixel >> 8); q[2] = (uint8_t)(pixel >> 16); }
  }
}

__global__ void __launch_bounds__(kBlock) k_copy2d(const uint8_t *__restrict__ src, int srs, uint8_t *__restrict__ dst,
                                                   int drs, int row_bytes, int rows, int fill, uint8_t fv) {
  const bool vec = ((uintptr_t)dst % 16 == 0) && (drs % 16 == 0) && (fill || (((uintptr_t)src % 16 == 0) && (srs % 16 == 0)));
  if (vec) {
    const int chunks = (row_bytes + 15) >> 4;
    const long long total = (long long)chunks * rows;
    const uint32_t f4 = fv * 0x01010101u;
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / chunks), c = (int)(it - (long long)row * chunks);
      uint8_t *d = dst + (long long)row * drs + c * 16;
      if (c * 16 + 16 <= row_bytes) {
        st_stream_u4(d, fill ? make_uint4(f4, f4, f4, f4) : ld_stream_u4(src + (long long)row * srs + c * 16));
      } else {
        for (int b = c * 16; b < row_bytes; b++) dst[(long long)row * drs + b] = fill ? fv : src[(long long)row * srs + b];
      }
    }
  } else {
    const long long total = (long long)row_bytes * rows;
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / row_bytes), b = (int)(it - (long long)row * row_bytes);
      dst[(long long)row * drs + b] = fill ? fv : src[(long long)row * srs + b];
    }
  }
}

// every outer pixel is written exactly once: inner pixel or black border
__global__ void __launch_bounds__(kBlock) k_letterbox(const uint8_t *__restrict__ inner, int irs, int iw, int ih,
                                                      uint8_t *__restrict__ outer, int ors, int ow, int oh, int psize,
                                                      int ox, int oy, uint32_t black) {
  const long long total = (long long)ow * oh;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / ow), x = (int)(it - (long long)row * ow);
    const int ix = x - ox, iy = row - oy;
    const bool in = ix >= 0 && ix < iw && iy >= 0 && iy < ih;
    uint8_t *d = outer + (long long)row * ors + (long long)x * psize;
    if (psize == 4) {
      *(uint32_t *)d = in ? *(const uint32_t *)(inner + (long long)iy * irs + (long long)ix * 4) : black;
    } else {
      const uint8_t *s = inner + (long long)iy * irs + (long long)ix * 3;
      d[0] = in ? s[0] : (uint8_t)black; d[1] = in ? s[1] : (uint8_t)(black >> 8); d[2] = in ? s[2] : (uint8_t)(black >> 16);
    }
  }
}

}  // namespace

cudaError_t launch_fill(const Launch &L, Img dst, int width, int height, int psize, uint32_t pixel) {
  k_fill<<<grid_for(L, (long long)width * height), kBlock, 0, L.stream>>>(dst.p, dst.rs, width, height, psize, pixel);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_copy2d(const Launch &L, const uint8_t *src, int srs, uint8_t *dst, int drs, int row_bytes, int rows,
                          int fill, uint8_t fill_value) {
  if (rows <= 0 || row_bytes <= 0) return cudaSuccess;
  k_copy2d<<<grid_for(L, (long long)((row_bytes + 15) >> 4) * rows), kBlock, 0, L.stream>>>(src, srs, dst, drs, row_bytes,
                                                                                          rows, fill, fill_value);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_letterbox(const Launch &L, CImg inner, int iw, int ih, Img outer, int ow, int oh, int psize,
                             uint32_t black_pixel) {
  const int ox = (ow - iw + 1) >> 1, oy = (oh - ih + 1) >> 1;  // colourspace.c:15522-15523
  k_letterbox<<<grid_for(L, (long long)ow * oh), kBlock, 0, L.stream>>>(inner.p, inner.rs, iw, ih, outer.p, outer.rs, ow, oh,
                                                                       psize, ox, oy, black_pixel);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// =====================================================================================================
// simple_blend.c (chroma blend + luma overlays).  A batch of frames shares one launch: blockIdx.y = frame.
// =====================================================================================================

namespace {

// (bf * a + bfn * b) >> 8 on the four bytes of a word, two 16-bit lanes at a time.
// bf + bfn == 255, so every lane stays below 65536 and nothing carries between lanes.
__device__ __forceinline__ uint32_t blend4(uint32_t a, uint32_t b, uint32_t bf, uint32_t bfn) {
  const uint32_t lo = (a & 0x00FF00FFu) * bf + (b & 0x00FF00FFu) * bfn;
  const uint32_t hi = ((a >> 8) & 0x00FF00FFu) * bf + ((b >> 8) & 0x00FF00FFu) * bfn;
  return ((lo >> 8) & 0x00FF00FFu) | (hi & 0xFF00FF00u);
}

struct BlendParams {
  const BlendFrame *frames;
  int width, height, psize, bf, type;
  int r_off, g_off, b_off, a_off;  // a_off: 3 (RGBA/BGRA), 0 (ARGB), -1
  const int32_t *luma;             // [3][256] plugin-side 16.16 luma tables
};

// chroma blend on 3-byte pixels: a pure byte stream (simple_blend.c:117-125)
__global__ void __launch_bounds__(kBlock) k_chroma_blend3(const BlendParams P) {
  const BlendFrame F = P.frames[blockIdx.y];
  const uint32_t bf = (uint8_t)P.bf, bfn = 0xFFu - bf;
  const int row_bytes = P.width * 3;
  const bool vec = (((uintptr_t)F.s1 | (uintptr_t)F.s2 | (uintptr_t)F.d) % 16 == 0) && ((F.rs1 | F.rs2 | F.rsd) % 16 == 0);
  if (vec) {
    const int chunks = (row_bytes + 15) >> 4;
    const long long total = (long long)chunks * P.height;
    if (total < (1ll << 31) && F.d != F.s2 && (long long)P.height * max(max(F.rs1, F.rs2), F.rsd) < (1ll << 32)) {
      // the streaming form: 32-bit index arithmetic (the 64-bit division per 16-byte chunk was a third of the kernel's instructions),
      // four independent chunks in flight per thread, L1-bypassing loads and stores.  Every byte is read before the same thread
      // writes it, so the in-place call (d == s1, effects-weed.c:2304-2314) keeps plain loads for s1 and is otherwise identical.
      const bool inplace = F.d == F.s1;
      const uint32_t T = (uint32_t)gridDim.x * blockDim.x, tot = (uint32_t)total, uch = (uint32_t)chunks;
      for (uint32_t it0 = (uint32_t)blockIdx.x * blockDim.x + threadIdx.x; it0 < tot; it0 += 4u * T) {
        uint4 a[4], b[4];
        uint32_t od[4];
        bool full[4], ok[4];
        uint32_t o1s[4], o2s[4], rem[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t it = it0 + (uint32_t)j * T;
          ok[j] = it < tot;
          const uint32_t row = ok[j] ? it / uch : 0u, c = ok[j] ? it - row * uch : 0u;
          o1s[j] = row * (uint32_t)F.rs1 + c * 16u; o2s[j] = row * (uint32_t)F.rs2 + c * 16u; od[j] = row * (uint32_t)F.rsd + c * 16u;
          rem[j] = (uint32_t)row_bytes - c * 16u;
          full[j] = ok[j] && rem[j] >= 16u;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (full[j]) {
            a[j] = ld_stream_u4(F.s2 + o2s[j]);
            b[j] = inplace ? ld_u4(F.s1 + o1s[j]) : ld_stream_u4(F.s1 + o1s[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (full[j]) {
            uint4 r;
            r.x = blend4(a[j].x, b[j].x, bf, bfn); r.y = blend4(a[j].y, b[j].y, bf, bfn);
            r.z = blend4(a[j].z, b[j].z, bf, bfn); r.w = blend4(a[j].w, b[j].w, bf, bfn);
            st_stream_u4(F.d + od[j], r);
          } else if (ok[j]) {
            for (uint32_t k = 0; k < rem[j]; k++)
              F.d[od[j] + k] = (uint8_t)((bf * F.s2[o2s[j] + k] + bfn * F.s1[o1s[j] + k]) >> 8);
          }
        }
      }
      return;
    }
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / chunks), c = (int)(it - (long long)row * chunks);
      const long long o1 = (long long)row * F.rs1 + c * 16, o2 = (long long)row * F.rs2 + c * 16,
                      od = (long long)row * F.rsd + c * 16;
      if (c * 16 + 16 <= row_bytes) {
        const uint4 a = ld_u4(F.s2 + o2), b = ld_u4(F.s1 + o1);
        uint4 r;
        r.x = blend4(a.x, b.x, bf, bfn); r.y = blend4(a.y, b.y, bf, bfn);
        r.z = blend4(a.z, b.z, bf, bfn); r.w = blend4(a.w, b.w, bf, bfn);
        *(uint4 *)(F.d + od) = r;
      } else {
        for (int k = 0; k < row_bytes - c * 16; k++)
          F.d[od + k] = (uint8_t)((bf * F.s2[o2 + k] + bfn * F.s1[o1 + k]) >> 8);
      }
    }
  } else {
    const long long total = (long long)row_bytes * P.height;
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / row_bytes), k = (int)(it - (long long)row * row_bytes);
      F.d[(long long)row * F.rsd + k] =
          (uint8_t)((bf * F.s2[(long long)row * F.rs2 + k] + bfn * F.s1[(long long)row * F.rs1 + k]) >> 8);
    }
  }
}

// chroma blend on 4-byte pixels (simple_blend.c:127-148): opaque src2 pixels blend bytewise, others scale both
// operands by alpha in float32 first; the alpha byte of dst is never written.  For ARGB the plugin starts at
// byte 1 (start = 1, :80) so the "alpha" it tests is byte 0 of the NEXT pixel: replicated via a_next.
__global__ void __launch_bounds__(kBlock) k_chroma_blend4(const BlendParams P) {
  const BlendFrame F = P.frames[blockIdx.y];
  const uint32_t bf = (uint8_t)P.bf, bfn = 0xFFu - bf;
  const bool argb = (P.a_off == 0);
  const long long total = (long long)P.width * P.height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / P.width), x = (int)(it - (long long)row * P.width);
    const long long o1 = (long long)row * F.rs1 + x * 4, o2 = (long long)row * F.rs2 + x * 4, od = (long long)row * F.rsd + x * 4;
    const uint32_t p1 = *(const uint32_t *)(F.s1 + o1), p2 = *(const uint32_t *)(F.s2 + o2);
    uint32_t a2;
    if (!argb) a2 = p2 >> 24;
    else a2 = (o2 + 4 < F.s2_bytes) ? F.s2[o2 + 4] : 255u;
    const uint32_t cmask = argb ? 0xFFFFFF00u : 0x00FFFFFFu;
    uint32_t res;
    if (a2 == 255u) {
      res = blend4(p2, p1, bf, bfn);
    } else {
      const float alpha = (float)((double)(float)a2 / 255.), inv_alpha = (float)(1. - (double)alpha);
      uint32_t q2 = 0, q1 = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        q2 |= ((uint32_t)(uint8_t)__fmul_rn((float)byte_of(p2, k), alpha)) << (8 * k);
        q1 |= ((uint32_t)(uint8_t)__fmul_rn((float)byte_of(p1, k), inv_alpha)) << (8 * k);
      }
      res = blend4(q2, q1, bf, bfn);
    }
    // bytes outside the colour mask keep whatever dst holds (in place: src1's alpha)
    const uint32_t keep = (F.d == F.s1) ? p1 : *(const uint32_t *)(F.d + od);
    *(uint32_t *)(F.d + od) = (res & cmask) | (keep & ~cmask);
  }
}

__device__ __forceinline__ uint32_t plugin_luma(const int32_t *luma, uint32_t r, uint32_t g, uint32_t b) {
  return (uint32_t)((luma[r] + luma[256 + g] + luma[512 + b]) >> 16) & 0xFFu;  // calc_luma returns uint8_t
}

// luma overlay / underlay / negative overlay (simple_blend.c:153-197): whole-pixel select.
// ARGB: the plugin walks the row from byte 1 (start = 1, :80) and hands calc_luma that pointer, so the luma it
// compares is lumaR[G] + lumaG[B] + lumaB[alpha byte of the NEXT pixel] -- replicated; past the end of the buffer
// that byte is taken as 255 (s2_bytes; both inputs share src2's geometry).
__global__ void __launch_bounds__(kBlock) k_luma_select(const BlendParams P) {
  __shared__ int32_t s_luma[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_luma[i] = P.luma[i];
  __syncthreads();
  const BlendFrame F = P.frames[blockIdx.y];
  const uint32_t bf = (uint8_t)P.bf, bfn = 0xFFu - bf;
  const bool argb = (P.a_off == 0);
  const int start = argb ? 1 : 0;
  const long long total = (long long)P.width * P.height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / P.width), x = (int)(it - (long long)row * P.width);
    const long long o1 = (long long)row * F.rs1 + x * P.psize + start, o2 = (long long)row * F.rs2 + x * P.psize + start;
    const uint8_t *s1 = F.s1 + o1, *s2 = F.s2 + o2;
    uint8_t *d = F.d + (long long)row * F.rsd + x * P.psize + start;
    const uint8_t *sl = (P.type == 2) ? s2 : s1;  // the input whose luma is tested
    uint32_t l;
    if (!argb) {
      l = plugin_luma(s_luma, sl[P.r_off], sl[P.g_off], sl[P.b_off]);
    } else {
      const long long o3 = ((P.type == 2) ? o2 : o1) + 3;
      const uint32_t nxt = o3 < F.s2_bytes ? sl[3] : 255u;
      l = plugin_luma(s_luma, sl[1], sl[2], nxt);
    }
    const bool take2 = (P.type == 1) ? (l < bf) : (l > bfn);
    const uint8_t *s = take2 ? s2 : s1;
    if (take2 || F.d != F.s1) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }
  }
}

}  // namespace

cudaError_t launch_simple_blend(const Launch &L, int type, const BlendFrame *frames_dev, int nframes, int width,
                                int height, RgbLayout lay, int bf, const int32_t *luma_tabs_dev) {
  BlendParams P;
  P.frames = frames_dev; P.width = width; P.height = height; P.psize = lay.psize; P.bf = bf; P.type = type;
  P.r_off = lay.r; P.g_off = lay.g; P.b_off = lay.b; P.a_off = lay.a; P.luma = luma_tabs_dev;
  const long long per_frame = type == 0 && lay.psize == 3 ? (long long)((width * 3 + 15) >> 4) * height
                                                          : (long long)width * height;
  // split the SMs between the frames of the batch
  int gx = grid_for(L, per_frame, 8);
  if (nframes > 1) gx = max(1, min(gx, (L.sm_count * 8 + nframes - 1) / nframes));
  dim3 grid(gx, nframes);
  if (type == 0 && lay.psize == 3) k_chroma_blend3<<<grid, kBlock, 0, L.stream>>>(P);
  else if (type == 0) k_chroma_blend4<<<grid, kBlock, 0, L.stream>>>(P);
  else k_luma_select<<<grid, kBlock, 0, L.stream>>>(P);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// =====================================================================================================
// multi_blends.c (multiply, screen, darken, lighten, overlay, dodge, burn), RGB24 / BGR24
// =====================================================================================================

namespace {

__global__ void __launch_bounds__(kBlock) k_multi_blend(int type, BlendFrame F, int width, int height, int bgr, int bf,
                                                        const int32_t *__restrict__ luma) {
  __shared__ int32_t s_luma[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_luma[i] = luma[i];
  __syncthreads();
  const uint8_t blend_factor = (uint8_t)bf;
  // unsigned char arithmetic of multi_blends.c:56-60 (wraps modulo 256)
  const uint32_t blend1 = (uint8_t)(blend_factor * 2), blendneg1 = (uint8_t)(255 - blend_factor * 2);
  const uint32_t blend2 = (uint8_t)((255 - blend_factor) * 2), blendneg2 = (uint8_t)((blend_factor - 128) * 2);
  const long long total = (long long)width * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / width), x = (int)(it - (long long)row * width);
    const uint8_t *s1 = F.s1 + (long long)row * F.rs1 + x * 3, *s2 = F.s2 + (long long)row * F.rs2 + x * 3;
    uint8_t *d = F.d + (long long)row * F.rsd + x * 3;
    int a[3] = {s1[0], s1[1], s1[2]}, b[3] = {s2[0], s2[1], s2[2]}, px[3];
    const int r1 = bgr ? a[2] : a[0], bl1 = bgr ? a[0] : a[2], r2 = bgr ? b[2] : b[0], bl2 = bgr ? b[0] : b[2];
    bool mpy = false, scr = false;
    switch (type) {
    case 0: mpy = true; break;
    case 1: scr = true; break;
    case 2: case 3: {
      const uint32_t l1 = plugin_luma(s_luma, r1, a[1], bl1), l2 = plugin_luma(s_luma, r2, b[1], bl2);
      const bool first = type == 2 ? (l1 <= l2) : (l1 >= l2);
#pragma unroll
      for (int k = 0; k < 3; k++) px[k] = first ? a[k] : b[k];
      break;
    }
    case 4: if (plugin_luma(s_luma, r1, a[1], bl1) < 128) mpy = true; else scr = true; break;
    case 5:
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (b[k] == 255) px[k] = 255;
        else { const int v = (a[k] << 8) / (255 - b[k]); px[k] = v > 255 ? 255 : (v & 0xFF); }
      }
      break;
    default:
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (b[k] == 0) px[k] = 0;
        else { const int v = 255 - (255 - (a[k] << 8)) / b[k]; px[k] = v < 0 ? 0 : (v & 0xFF); }
      }
      break;
    }
    if (mpy) {
#pragma unroll
      for (int k = 0; k < 3; k++) px[k] = (b[k] * a[k]) >> 8;
    }
    if (scr) {
#pragma unroll
      for (int k = 0; k < 3; k++) px[k] = (255 - (((255 - b[k]) * (255 - a[k])) >> 8)) & 0xFF;
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
      d[k] = blend_factor < 128 ? (uint8_t)((blend1 * px[k] + blendneg1 * a[k]) >> 8)
                                : (uint8_t)((blend2 * px[k] + blendneg2 * b[k]) >> 8);
  }
}

}  // namespace

cudaError_t launch_multi_blend(const Launch &L, int type, BlendFrame f, int width, int height, int bgr, int bf,
                               const int32_t *luma_tabs_dev) {
  k_multi_blend<<<grid_for(L, (long long)width * height), kBlock, 0, L.stream>>>(type, f, width, height, bgr, bf, luma_tabs_dev);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// =====================================================================================================
// slide over (slide_over.c sover_process :55-145): every destination byte is a copy of one byte of in1 or in2 -- the side of the
// dividing line picks the clip, a "moving" clip is read with a constant offset.  Pure copy, 1 load + 1 store per byte: one thread
// moves 16 bytes, as one 128-bit access when both ends of the chunk lie on the same side and source and destination are aligned.
// =====================================================================================================
namespace {

__global__ void __launch_bounds__(kBlock) k_slide_over(const SlideArgs P) {
  const int chunks = (P.row_bytes + 15) >> 4;
  const long long total = (long long)chunks * P.height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int j = (int)(it / chunks), x0 = (int)(it - (long long)j * chunks) << 4;
    const int n = min(16, P.row_bytes - x0);
    uint8_t *d = P.d + (long long)P.rsd * j + x0;
    auto is_first = [&](int x) -> bool { return P.along_y ? j < P.bound : x < P.bound; };
    auto src_of = [&](int x, bool f) -> const uint8_t * {
      return f ? P.first + (long long)P.rs_first * j + P.off_first + x : P.second + (long long)P.rs_second * j + P.off_second + x;
    };
    const bool f0 = is_first(x0), f1 = is_first(x0 + n - 1);
    const uint8_t *s = src_of(x0, f0);
    if (n == 16 && f0 == f1 && !(reinterpret_cast<uintptr_t>(d) & 15)) {
      const uintptr_t sa = reinterpret_cast<uintptr_t>(s);
      if (!(sa & 15)) {
        *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(s);
        continue;
      }
      if (P.word_safe) {
        // a moving clip is read at an offset of (width - bound) * psize bytes: any alignment.  Aligned 32-bit loads of the 4 or 5
        // words that hold the 16 bytes, re-aligned with funnel shifts, one 128-bit store (the 5th word stays inside the row
        // stride: word_safe = every source stride is a multiple of 4)
        const uint32_t *w = reinterpret_cast<const uint32_t *>(sa & ~uintptr_t(3));
        const uint32_t sh = (uint32_t)(sa & 3) * 8u;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        uint4 v;
        if (!sh) v = make_uint4(w0, w1, w2, w3);
        else {
          const uint32_t w4 = w[4];
          v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
        *reinterpret_cast<uint4 *>(d) = v;
        continue;
      }
    }
    for (int k = 0; k < n; k++) d[k] = *src_of(x0 + k, is_first(x0 + k));
  }
}

}  // namespace

cudaError_t launch_slide_over(const Launch &L, const SlideArgs &a) {
  if (a.row_bytes <= 0 || a.height <= 0) return cudaSuccess;
  SlideArgs b = a;
  b.word_safe = !((a.rs_first | a.rs_second) & 3) && !((reinterpret_cast<uintptr_t>(a.first) | reinterpret_cast<uintptr_t>(a.second)) & 3);
  k_slide_over<<<grid_for(L, (long long)((a.row_bytes + 15) >> 4) * a.height), kBlock, 0, L.stream>>>(b);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// =====================================================================================================
// alpha-over: dst = (u8)(bg * (1 - alpha) + fg * alpha) in double, truncating store
// (gdk/compositor.c paint_pixel :120-125), optionally followed by the 8-bit gamma LUT in the same pass.
//
// The scalar alpha makes the result a pure function of the (bg, fg) byte pair.  k_over_table evaluates the
// reference's double expression (no FMA contraction: __dmul_rn / __dadd_rn) once for all 65536 pairs --
// composed with the gamma LUT when there is one -- into a 64 KB table in HBM (cached by the engine per
// (alpha, LUT)); k_alpha_over stages that table in shared memory and the per-pixel work becomes one gather
// per colour byte, with 128-bit loads / stores of 4 pixels per thread.
// =====================================================================================================

namespace {

__global__ void __launch_bounds__(kBlock) k_over_table(double alpha, const uint8_t *__restrict__ lut, uint8_t *__restrict__ tab) {
  const double invalpha = 1. - alpha;
  for (int i = (int)global_tid(); i < 65536; i += (int)global_threads()) {
    const double v = __dadd_rn(__dmul_rn((double)(i >> 8), invalpha), __dmul_rn((double)(i & 255), alpha));
    uint8_t r = (uint8_t)(int)v;  // C conversion double -> unsigned char: truncation
    if (lut) r = lut[r];
    tab[i] = r;  // [bg][fg]
  }
}

struct OverParams {
  const uint8_t *bg, *fg;
  uint8_t *dst;
  int rs_bg, rs_fg, rs_d, width, height, psize;
  const uint8_t *tab;
  int force_opaque;  // compositor fills alpha with 0xFF (compositor.c:184) and never paints it
  int a_off;
};

__global__ void __launch_bounds__(1024, 1) k_alpha_over(const OverParams P) {
  extern __shared__ __align__(16) uint8_t s_tab[];  // [bg][fg]
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) ((uint4 *)s_tab)[i] = ((const uint4 *)P.tab)[i];
  __syncthreads();
  if (P.psize == 4) {
    const uint32_t amask = 0xFFu << (8 * P.a_off);
    const bool vec = (((uintptr_t)P.bg | (uintptr_t)P.fg | (uintptr_t)P.dst) % 16 == 0) && ((P.rs_bg | P.rs_fg | P.rs_d) % 16 == 0);
    const int groups = (P.width + 3) >> 2;
    const long long total = (long long)groups * P.height;
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
      const int npx = min(4, P.width - g * 4);
      const uint8_t *b = P.bg + (long long)row * P.rs_bg + g * 16, *f = P.fg + (long long)row * P.rs_fg + g * 16;
      uint8_t *d = P.dst + (long long)row * P.rs_d + g * 16;
      uint32_t bw[4], fw[4], ow[4];
      if (vec && npx == 4) {
        const uint4 vb = ld_u4(b), vf = ld_u4(f);
        bw[0] = vb.x; bw[1] = vb.y; bw[2] = vb.z; bw[3] = vb.w;
        fw[0] = vf.x; fw[1] = vf.y; fw[2] = vf.z; fw[3] = vf.w;
      } else {
        for (int k = 0; k < npx; k++) { bw[k] = *(const uint32_t *)(b + 4 * k); fw[k] = *(const uint32_t *)(f + 4 * k); }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (k < npx) {
          uint32_t o = 0;
#pragma unroll
          for (int c = 0; c < 4; c++) o |= (uint32_t)s_tab[(byte_of(bw[k], c) << 8) | byte_of(fw[k], c)] << (8 * c);
          ow[k] = (o & ~amask) | (P.force_opaque ? amask : (bw[k] & amask));
        }
      }
      if (vec && npx == 4) *(uint4 *)d = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      else for (int k = 0; k < npx; k++) *(uint32_t *)(d + 4 * k) = ow[k];
    }
  } else {
    const int row_bytes = P.width * 3;
    const long long total = (long long)row_bytes * P.height;
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / row_bytes), k = (int)(it - (long long)row * row_bytes);
      P.dst[(long long)row * P.rs_d + k] = s_tab[((uint32_t)P.bg[(long long)row * P.rs_bg + k] << 8) | P.fg[(long long)row * P.rs_fg + k]];
    }
  }
}

}  // namespace

namespace {

// alpha = k / 256: (bg * (256 - k) + fg * k) >> 8 is exactly trunc(bg * (1 - a) + fg * a) in double (all terms exact), so the
// table is not needed: two channels per multiply, 4 pixels (128 bits) per thread, optional 8-bit gamma LUT afterwards.
struct OverArithParams {
  const uint8_t *bg, *fg;
  uint8_t *dst;
  int rs_bg, rs_fg, rs_d, width, height, psize;
  uint32_t ka, kia;
  const uint8_t *lut;  // nullptr: none
  int force_opaque;
};

__device__ __forceinline__ uint32_t over_px4(uint32_t b, uint32_t f, uint32_t ka, uint32_t kia, const uint8_t *s_lut, bool has_lut,
                                             bool force_opaque) {
  const uint32_t rb = (b & 0x00FF00FFu) * kia + (f & 0x00FF00FFu) * ka;          // R | B, 16 bits each, no carry: kia + ka = 256
  const uint32_t g = ((b >> 8) & 0xFFu) * kia + ((f >> 8) & 0xFFu) * ka;
  uint32_t o0 = (rb >> 8) & 0xFFu, o1 = g >> 8, o2 = rb >> 24;
  if (has_lut) { o0 = s_lut[o0]; o1 = s_lut[o1]; o2 = s_lut[o2]; }
  return o0 | (o1 << 8) | (o2 << 16) | (force_opaque ? 0xFF000000u : (b & 0xFF000000u));
}

// frames of one batched launch: same geometry, strides, alpha and LUT; blockIdx.y selects the frame
struct OverFrameList {
  const uint8_t *bg[32], *fg[32];
  uint8_t *dst[32];
};

__global__ void __launch_bounds__(kBlock) k_alpha_over_arith(const OverArithParams P0, const __grid_constant__ OverFrameList FL) {
  OverArithParams P = P0;
  P.bg = FL.bg[blockIdx.y]; P.fg = FL.fg[blockIdx.y]; P.dst = FL.dst[blockIdx.y];
  __shared__ uint8_t s_lut[256];
  const bool has_lut = P.lut != nullptr;
  if (has_lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = P.lut[i];
    __syncthreads();
  }
  if (P.psize == 4) {
    const bool vec = (((uintptr_t)P.bg | (uintptr_t)P.fg | (uintptr_t)P.dst) % 16 == 0) && ((P.rs_bg | P.rs_fg | P.rs_d) % 16 == 0);
    const int groups = (P.width + 3) >> 2;
    const long long total = (long long)groups * P.height;
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
      const int npx = min(4, P.width - g * 4);
      const uint8_t *b = P.bg + (size_t)row * P.rs_bg + g * 16, *f = P.fg + (size_t)row * P.rs_fg + g * 16;
      uint8_t *d = P.dst + (size_t)row * P.rs_d + g * 16;
      if (vec && npx == 4) {
        const uint4 vb = ld_u4(b), vf = ld_u4(f);  // dst may alias bg: plain loads
        uint4 o;
        o.x = over_px4(vb.x, vf.x, P.ka, P.kia, s_lut, has_lut, P.force_opaque);
        o.y = over_px4(vb.y, vf.y, P.ka, P.kia, s_lut, has_lut, P.force_opaque);
        o.z = over_px4(vb.z, vf.z, P.ka, P.kia, s_lut, has_lut, P.force_opaque);
        o.w = over_px4(vb.w, vf.w, P.ka, P.kia, s_lut, has_lut, P.force_opaque);
        *reinterpret_cast<uint4 *>(d) = o;
      } else {
        for (int k = 0; k < npx; k++)
          *reinterpret_cast<uint32_t *>(d + 4 * k) = over_px4(*reinterpret_cast<const uint32_t *>(b + 4 * k),
                                                               *reinterpret_cast<const uint32_t *>(f + 4 * k), P.ka, P.kia, s_lut,
                                                               has_lut, P.force_opaque);
      }
    }
  } else {
    const int row_bytes = P.width * 3;
    const long long total = (long long)row_bytes * P.height;
    for (long long it = global_tid(); it < total; it += global_threads()) {
      const int row = (int)(it / row_bytes), k = (int)(it - (long long)row * row_bytes);
      uint32_t o = ((uint32_t)P.bg[(size_t)row * P.rs_bg + k] * P.kia + (uint32_t)P.fg[(size_t)row * P.rs_fg + k] * P.ka) >> 8;
      if (has_lut) o = s_lut[o];
      P.dst[(size_t)row * P.rs_d + k] = (uint8_t)o;
    }
  }
}

}  // namespace

cudaError_t launch_alpha_over_arith(const Launch &L, CImg bg, CImg fg, Img dst, int width, int height, int psize, int k256,
                                    const uint8_t *lut8_dev, int force_opaque) {
  OverArithParams P;
  P.bg = bg.p; P.fg = fg.p; P.dst = dst.p; P.rs_bg = bg.rs; P.rs_fg = fg.rs; P.rs_d = dst.rs;
  P.width = width; P.height = height; P.psize = psize; P.ka = (uint32_t)k256; P.kia = 256u - (uint32_t)k256;
  P.lut = lut8_dev; P.force_opaque = force_opaque;
  const long long work = psize == 4 ? (long long)((width + 3) >> 2) * height : (long long)width * 3 * height;
  OverFrameList fl;
  for (int i = 0; i < 32; i++) { fl.bg[i] = bg.p; fl.fg[i] = fg.p; fl.dst[i] = dst.p; }
  k_alpha_over_arith<<<grid_for(L, work, 8), kBlock, 0, L.stream>>>(P, fl);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// n frames that share geometry, strides, alpha and LUT: one launch per 32 (blockIdx.y = frame)
cudaError_t launch_alpha_over_arith_batch(const Launch &L, const uint8_t *const *bgs, const uint8_t *const *fgs, uint8_t *const *dsts, int n,
                                          int rs_bg, int rs_fg, int rs_d, int width, int height, int psize, int k256,
                                          const uint8_t *lut8_dev, int force_opaque) {
  OverArithParams P;
  P.bg = nullptr; P.fg = nullptr; P.dst = nullptr; P.rs_bg = rs_bg; P.rs_fg = rs_fg; P.rs_d = rs_d;
  P.width = width; P.height = height; P.psize = psize; P.ka = (uint32_t)k256; P.kia = 256u - (uint32_t)k256;
  P.lut = lut8_dev; P.force_opaque = force_opaque;
  const long long work = psize == 4 ? (long long)((width + 3) >> 2) * height : (long long)width * 3 * height;
  for (int base = 0; base < n; base += 32) {
    const int m = n - base < 32 ? n - base : 32;
    OverFrameList fl;
    for (int i = 0; i < 32; i++) { const int j = base + (i < m ? i : 0); fl.bg[i] = bgs[j]; fl.fg[i] = fgs[j]; fl.dst[i] = dsts[j]; }
    // the frames share the grid: keep ~8 CTAs per SM in total, at least one column of CTAs per frame
    int gx = grid_for(L, work, 8) / m;
    if (gx < 1) gx = 1;
    k_alpha_over_arith<<<dim3(gx, m), kBlock, 0, L.stream>>>(P, fl);
    PE_COUNT_LAUNCH(L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_over_table(const Launch &L, double alpha, const uint8_t *lut8_dev, uint8_t *table_dev) {
  k_over_table<<<64, kBlock, 0, L.stream>>>(alpha, lut8_dev, table_dev);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_alpha_over(const Launch &L, CImg bg, CImg fg, Img dst, int width, int height, int psize,
                              const uint8_t *over_table_dev, int force_opaque) {
  OverParams P;
  P.bg = bg.p; P.fg = fg.p; P.dst = dst.p; P.rs_bg = bg.rs; P.rs_fg = fg.rs; P.rs_d = dst.rs;
  P.width = width; P.height = height; P.psize = psize; P.tab = over_table_dev;
  P.force_opaque = force_opaque; P.a_off = 3;
  static PerDevice attr_set;
  if (!attr_set.cur()) {
    cudaError_t e = cudaFuncSetAttribute(k_alpha_over, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    if (e != cudaSuccess) return e;
    attr_set.cur() = 1;
  }
  k_alpha_over<<<L.sm_count, 1024, 65536, L.stream>>>(P);  // persistent: one CTA per SM (64 KB table each)
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

// =====================================================================================================
// per-frame diagnostics: min / max per byte position, histogram of the colour bytes, byte sum, black test.
// Warp-shuffle reductions, one atomic per warp (per CTA for the histogram).
// =====================================================================================================

namespace {

// k_stats: min / max / sum per byte position, histogram of the colour bytes, and both branches of is_all_black_ish
// (colourspace.c:2554-2594) in one pass.  A CTA stages a chunk of a row in shared memory with 128-bit streaming loads (every byte
// crosses HBM once, as whole sectors); a thread then owns whole pixels of the chunk, so its channel indices are static; the histogram
// goes to a per-WARP private copy in shared memory (no contention between warps), everything else is reduced with warp shuffles and
// leaves the CTA as one atomic per quantity.
constexpr int ST_CHUNK3 = 4080, ST_CHUNK4 = 4096;   // bytes of a row per step: whole pixels and whole 16-byte vectors

// the reference's "ish" test on the first three bytes a, b, c of a pixel, bit for bit (:2583-2587): nonzero = not black-ish
__device__ __forceinline__ unsigned int ish_expr(unsigned int a, unsigned int b, unsigned int c) {
  const unsigned int na = (a & 0x1Fu) ^ a, nc = (c & 0x1Fu) ^ c, nb = (b & 0x1Fu) ^ b;
  const unsigned int t1 = ((b << 1) & 0x1Fu) ^ (b << 1), t2 = (((b & 0x0Fu) << 2) & 0x1Fu) ^ ((b & 0x0Fu) << 2);
  return na & nc & (nb | (t1 & t2));
}

template <int PS>
__global__ void __launch_bounds__(kBlock) k_stats(const uint8_t *__restrict__ p, int rs, int width, int height, int a_off, DevStats *out) {
  constexpr int CHUNK = PS == 3 ? ST_CHUNK3 : ST_CHUNK4;
  __shared__ __align__(16) uint8_t s_buf[ST_CHUNK4];
  __shared__ unsigned int s_hist[kBlock / 32][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kBlock / 32) * 256; i += kBlock) (&s_hist[0][0])[i] = 0u;
  unsigned int mn[4] = {255, 255, 255, 255}, mx[4] = {0, 0, 0, 0}, notblack = 0, notish = 0;
  unsigned long long sum = 0;
  const int row_bytes = width * PS;
  const int chunks_per_row = (row_bytes + CHUNK - 1) / CHUNK;
  const long long units = (long long)chunks_per_row * height;
  unsigned int *hist = s_hist[warp];
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const int row = (int)(u / chunks_per_row), c0 = (int)(u - (long long)row * chunks_per_row) * CHUNK;
    const int nb = min(CHUNK, row_bytes - c0);
    const uint8_t *src = p + (long long)rs * row + c0;
    __syncthreads();  // the previous chunk has been consumed
    if ((((uintptr_t)src) & 15) == 0) {
      for (int i = threadIdx.x; i * 16 < nb; i += kBlock) {
        if (i * 16 + 16 <= nb) reinterpret_cast<uint4 *>(s_buf)[i] = ld_stream_u4(src + 16 * i);
        else for (int k = i * 16; k < nb; k++) s_buf[k] = src[k];
      }
    } else {
      for (int k = threadIdx.x; k < nb; k += kBlock) s_buf[k] = src[k];
    }
    __syncthreads();
    if (PS == 1) {  // planar: one byte per sample, 4 at a time
      for (int i = threadIdx.x * 4; i < nb; i += kBlock * 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (i + k < nb) {
            const unsigned int v = s_buf[i + k];
            mn[0] = min(mn[0], v); mx[0] = max(mx[0], v); sum += v;
            atomicAdd(&hist[v], 1u);
            notblack |= v;
          }
        }
      }
    } else {
      const int npx = nb / PS;
      for (int i = threadIdx.x; i < npx; i += kBlock) {
        unsigned int v[4];
        if (PS == 4) {
          const unsigned int w = reinterpret_cast<const unsigned int *>(s_buf)[i];
          v[0] = w & 255u; v[1] = (w >> 8) & 255u; v[2] = (w >> 16) & 255u; v[3] = w >> 24;
        } else {
          v[0] = s_buf[3 * i]; v[1] = s_buf[3 * i + 1]; v[2] = s_buf[3 * i + 2]; v[3] = 0;
        }
#pragma unroll
        for (int k = 0; k < PS; k++) {
          mn[k] = min(mn[k], v[k]); mx[k] = max(mx[k], v[k]); sum += v[k];
          if (k != a_off) atomicAdd(&hist[v[k]], 1u);
        }
        // is_all_black_ish looks at bytes 0, 1, 2 of the pixel whatever the palette (:2575-2577; its caller passes offs = 0)
        notblack |= v[0] | v[1] | v[2];
        notish |= ish_expr(v[0], v[1], v[2]);
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      mn[k] = min(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], off));
      mx[k] = max(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
    notblack |= __shfl_xor_sync(0xffffffffu, notblack, off);
    notish |= __shfl_xor_sync(0xffffffffu, notish, off);
  }
  if (lane == 0) {
    for (int k = 0; k < 4; k++) { atomicMin(&out->minv[k], mn[k]); atomicMax(&out->maxv[k], mx[k]); }
    atomicAdd(&out->sum, sum);
    if (notblack) atomicOr(&out->not_black, 1u);
    if (notish) atomicOr(&out->not_black_ish, 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += kBlock) {
    unsigned int t = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; w++) t += s_hist[w][i];
    if (t) atomicAdd(&out->hist[i], t);
  }
}

// k_row_hash: the row hashes of hash_cmp_layer (colourspace.c:16044-16075): minimd5 (src/maths.c:575) of the first `nbytes` bytes of
// every row.  minimd5 is the reference's OWN MD5 variant (src/maths.h:42-56): rounds 2 - 4 are RFC 1321's, round 1 is not -- its BX
// macro runs the four state words in the order A B C D with the rotations 7 22 17 12, and its fourth step mixes the step-2 CONSTANT
// where RFC 1321 has the state word C (a macro parameter named X shadows nothing but reads as the constant).  Restated as written.
// One warp per row: the lanes load the row 128 bytes at a time (coalesced), the 16 words of each 64-byte block are broadcast with
// shuffles and every lane runs the same 64 steps (no divergence); the tail block(s) with the padding and the bit count are built in
// registers.
__device__ __forceinline__ uint32_t md5_rotl(uint32_t x, int s) { return __funnelshift_l(x, x, s); }
__device__ __forceinline__ uint32_t md5_f(uint32_t b, uint32_t c, uint32_t d) { return d ^ (b & (c ^ d)); }

__device__ __forceinline__ void lives_md5_block(const uint32_t (&X)[16], uint32_t &A, uint32_t &B, uint32_t &C, uint32_t &D) {
  const uint32_t T1[16] = {0xd76aa478u, 0xe8c7b756u, 0x242070dbu, 0xc1bdceeeu, 0xf57c0fafu, 0x4787c62au, 0xa8304613u, 0xfd469501u,
                           0x698098d8u, 0x8b44f7afu, 0xffff5bb1u, 0x895cd7beu, 0x6b901122u, 0xfd987193u, 0xa679438eu, 0x49b40821u};
  const uint32_t a0 = A, b0 = B, c0 = C, d0 = D;
#pragma unroll
  for (int q = 0; q < 4; q++) {  // BX(W, X, Y, Z), src/maths.h:48
    A += md5_f(B, C, D) + X[4 * q] + T1[4 * q]; A = md5_rotl(A, 7) + B;
    B += md5_f(C, D, A) + X[4 * q + 1] + T1[4 * q + 1]; B = md5_rotl(B, 22) + C;
    C += md5_f(D, A, B) + X[4 * q + 2] + T1[4 * q + 2]; C = md5_rotl(C, 17) + D;
    D += md5_f(A, B, T1[4 * q + 1]) + X[4 * q + 3] + T1[4 * q + 3]; D = md5_rotl(D, 12) + A;
  }
#define PE_MD5_STEP(fn, a, b, c, d, k, s, t) a += fn(b, c, d) + X[k] + t; a = md5_rotl(a, s) + b;
#define PE_MD5_G(b, c, d) md5_f(d, b, c)
#define PE_MD5_H(b, c, d) ((b) ^ (c) ^ (d))
#define PE_MD5_I(b, c, d) ((c) ^ ((b) | ~(d)))
  PE_MD5_STEP(PE_MD5_G, A, B, C, D, 1, 5, 0xf61e2562u) PE_MD5_STEP(PE_MD5_G, D, A, B, C, 6, 9, 0xc040b340u)
  PE_MD5_STEP(PE_MD5_G, C, D, A, B, 11, 14, 0x265e5a51u) PE_MD5_STEP(PE_MD5_G, B, C, D, A, 0, 20, 0xe9b6c7aau)
  PE_MD5_STEP(PE_MD5_G, A, B, C, D, 5, 5, 0xd62f105du) PE_MD5_STEP(PE_MD5_G, D, A, B, C, 10, 9, 0x02441453u)
  PE_MD5_STEP(PE_MD5_G, C, D, A, B, 15, 14, 0xd8a1e681u) PE_MD5_STEP(PE_MD5_G, B, C, D, A, 4, 20, 0xe7d3fbc8u)
  PE_MD5_STEP(PE_MD5_G, A, B, C, D, 9, 5, 0x21e1cde6u) PE_MD5_STEP(PE_MD5_G, D, A, B, C, 14, 9, 0xc33707d6u)
  PE_MD5_STEP(PE_MD5_G, C, D, A, B, 3, 14, 0xf4d50d87u) PE_MD5_STEP(PE_MD5_G, B, C, D, A, 8, 20, 0x455a14edu)
  PE_MD5_STEP(PE_MD5_G, A, B, C, D, 13, 5, 0xa9e3e905u) PE_MD5_STEP(PE_MD5_G, D, A, B, C, 2, 9, 0xfcefa3f8u)
  PE_MD5_STEP(PE_MD5_G, C, D, A, B, 7, 14, 0x676f02d9u) PE_MD5_STEP(PE_MD5_G, B, C, D, A, 12, 20, 0x8d2a4c8au)
  PE_MD5_STEP(PE_MD5_H, A, B, C, D, 5, 4, 0xfffa3942u) PE_MD5_STEP(PE_MD5_H, D, A, B, C, 8, 11, 0x8771f681u)
  PE_MD5_STEP(PE_MD5_H, C, D, A, B, 11, 16, 0x6d9d6122u) PE_MD5_STEP(PE_MD5_H, B, C, D, A, 14, 23, 0xfde5380cu)
  PE_MD5_STEP(PE_MD5_H, A, B, C, D, 1, 4, 0xa4beea44u) PE_MD5_STEP(PE_MD5_H, D, A, B, C, 4, 11, 0x4bdecfa9u)
  PE_MD5_STEP(PE_MD5_H, C, D, A, B, 7, 16, 0xf6bb4b60u) PE_MD5_STEP(PE_MD5_H, B, C, D, A, 10, 23, 0xbebfbc70u)
  PE_MD5_STEP(PE_MD5_H, A, B, C, D, 13, 4, 0x289b7ec6u) PE_MD5_STEP(PE_MD5_H, D, A, B, C, 0, 11, 0xeaa127fau)
  PE_MD5_STEP(PE_MD5_H, C, D, A, B, 3, 16, 0xd4ef3085u) PE_MD5_STEP(PE_MD5_H, B, C, D, A, 6, 23, 0x04881d05u)
  PE_MD5_STEP(PE_MD5_H, A, B, C, D, 9, 4, 0xd9d4d039u) PE_MD5_STEP(PE_MD5_H, D, A, B, C, 12, 11, 0xe6db99e5u)
  PE_MD5_STEP(PE_MD5_H, C, D, A, B, 15, 16, 0x1fa27cf8u) PE_MD5_STEP(PE_MD5_H, B, C, D, A, 2, 23, 0xc4ac5665u)
  PE_MD5_STEP(PE_MD5_I, A, B, C, D, 0, 6, 0xf4292244u) PE_MD5_STEP(PE_MD5_I, D, A, B, C, 7, 10, 0x432aff97u)
  PE_MD5_STEP(PE_MD5_I, C, D, A, B, 14, 15, 0xab9423a7u) PE_MD5_STEP(PE_MD5_I, B, C, D, A, 5, 21, 0xfc93a039u)
  PE_MD5_STEP(PE_MD5_I, A, B, C, D, 12, 6, 0x655b59c3u) PE_MD5_STEP(PE_MD5_I, D, A, B, C, 3, 10, 0x8f0ccc92u)
  PE_MD5_STEP(PE_MD5_I, C, D, A, B, 10, 15, 0xffeff47du) PE_MD5_STEP(PE_MD5_I, B, C, D, A, 1, 21, 0x85845dd1u)
  PE_MD5_STEP(PE_MD5_I, A, B, C, D, 8, 6, 0x6fa87e4fu) PE_MD5_STEP(PE_MD5_I, D, A, B, C, 15, 10, 0xfe2ce6e0u)
  PE_MD5_STEP(PE_MD5_I, C, D, A, B, 6, 15, 0xa3014314u) PE_MD5_STEP(PE_MD5_I, B, C, D, A, 13, 21, 0x4e0811a1u)
  PE_MD5_STEP(PE_MD5_I, A, B, C, D, 4, 6, 0xf7537e82u) PE_MD5_STEP(PE_MD5_I, D, A, B, C, 11, 10, 0xbd3af235u)
  PE_MD5_STEP(PE_MD5_I, C, D, A, B, 2, 15, 0x2ad7d2bbu) PE_MD5_STEP(PE_MD5_I, B, C, D, A, 9, 21, 0xeb86d391u)
#undef PE_MD5_STEP
#undef PE_MD5_G
#undef PE_MD5_H
#undef PE_MD5_I
  A += a0; B += b0; C += c0; D += d0;
}

// word `wi` of the padded message of a row of n bytes: data, then 0x80, zeros, the bit count in the last two words of the last block
__device__ __forceinline__ uint32_t md5_msg_word(const uint8_t *__restrict__ row, int n, int wi, int total_words) {
  const int o = 4 * wi;
  if (o + 4 <= n) {
    if ((((uintptr_t)row) & 3) == 0) return *reinterpret_cast<const uint32_t *>(row + o);
    return (uint32_t)row[o] | ((uint32_t)row[o + 1] << 8) | ((uint32_t)row[o + 2] << 16) | ((uint32_t)row[o + 3] << 24);
  }
  if (wi == total_words - 2) return (uint32_t)n << 3;
  if (wi == total_words - 1) return (uint32_t)n >> 29;
  uint32_t w = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int b = o + k;
    const uint32_t v = b < n ? row[b] : (b == n ? 0x80u : 0u);
    w |= v << (8 * k);
  }
  return w;
}

__global__ void __launch_bounds__(kBlock) k_row_hash(const uint8_t *__restrict__ p, int rs, int nbytes, int height, unsigned long long *out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  if (row >= height) return;
  const uint8_t *r = p + (long long)rs * row;
  const int nblocks = (nbytes + 8) / 64 + 1;       // 0x80 and the 8 length bytes must fit
  const int total_words = nblocks * 16;
  uint32_t A = 0x67452301u, B = 0xefcdab89u, C = 0x98badcfeu, D = 0x10325476u;
  for (int blk = 0; blk < nblocks; blk += 2) {     // the warp fetches two blocks (32 words) at a time
    const int wi = blk * 16 + lane;
    const uint32_t mine = wi < total_words ? md5_msg_word(r, nbytes, wi, total_words) : 0u;
    uint32_t X[16];
#pragma unroll
    for (int k = 0; k < 16; k++) X[k] = __shfl_sync(0xffffffffu, mine, k);
    lives_md5_block(X, A, B, C, D);
    if (blk + 1 < nblocks) {
#pragma unroll
      for (int k = 0; k < 16; k++) X[k] = __shfl_sync(0xffffffffu, mine, 16 + k);
      lives_md5_block(X, A, B, C, D);
    }
  }
  // minimd5: U[0] ^ U[1] of the 16 digest bytes (little endian)
  if (lane == 0) out[row] = ((unsigned long long)A | ((unsigned long long)B << 32)) ^ ((unsigned long long)C | ((unsigned long long)D << 32));
}

}  // namespace

cudaError_t launch_stats(const Launch &L, CImg img, int width, int height, int psize, int a_off, DevStats *out_dev) {
  const int row_bytes = width * psize, chunk = psize == 3 ? ST_CHUNK3 : ST_CHUNK4;
  const long long units = (long long)((row_bytes + chunk - 1) / chunk) * height;
  long long g = (long long)L.sm_count * 6;
  if (g > units) g = units;
  if (g < 1) g = 1;
  if (psize == 4) k_stats<4><<<(int)g, kBlock, 0, L.stream>>>(img.p, img.rs, width, height, a_off, out_dev);
  else if (psize == 3) k_stats<3><<<(int)g, kBlock, 0, L.stream>>>(img.p, img.rs, width, height, a_off, out_dev);
  else if (psize == 1) k_stats<1><<<(int)g, kBlock, 0, L.stream>>>(img.p, img.rs, width, height, a_off, out_dev);
  else return cudaErrorInvalidValue;
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_row_hash(const Launch &L, CImg img, int nbytes, int height, unsigned long long *out_dev) {
  k_row_hash<<<(height + kBlock / 32 - 1) / (kBlock / 32), kBlock, 0, L.stream>>>(img.p, img.rs, nbytes, height, out_dev);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

}  // namespace pe
