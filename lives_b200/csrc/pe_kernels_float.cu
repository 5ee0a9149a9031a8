// pe_kernels_float.cu -- the reference's float ("experimental") YUV -> RGB arithmetic (colourspace.c:101-172 float tables, :592
// clamp0255f, :2367 yuv2rgb_float) on packed YUV888 / YUVA8888 frames.
//   mode 0: yuv2rgb_float exactly as written -- `int yy = RGB_Y[y]` (the 16.16 INTEGER table) added to the float chroma tables;
//   mode 1: the form of the commented-out variant at :2398-2400 -- RGBf_Y[y] + Rf_Cr[v], ...
// Every sum is a left-to-right chain of IEEE single-precision additions (__fadd_rn: no FMA contraction, no reassociation), the clamp
// compares and the conversion truncates as C does: bit-identical to the compiled reference (0 ULP on the float sums), not merely
// within 1 ULP.  HBM-bound byte work (6 - 8 bytes per pixel against 4 additions): CUDA cores; a tensor-core formulation of the 3 x 3
// matrix would have to round the coefficients to tf32 / bf16 and could not reproduce these sums.
#include "pe_device.cuh"
#include "pe_kernels.h"

namespace pe {

namespace {

constexpr int kBlock = 256;
#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

struct FloatTabs {
  float ty[256], rcr[256], gcb[256], gcr[256], bcb[256];
  int32_t yi[256];
};

__device__ __forceinline__ uint32_t clamp0255f_dev(float f) {  // colourspace.c:592-596
  if (f > 255.f) f = 255.f;
  if (f < 0.f) f = 0.f;
  return (uint32_t)f;  // truncation, as the implicit float -> uint8_t conversion
}

template <int MODE>
__device__ __forceinline__ void px_float(const FloatTabs &t, uint32_t y, uint32_t u, uint32_t v, float &r, float &g, float &b) {
  const float yy = MODE == 0 ? (float)t.yi[y] : t.ty[y];  // int + float: the int operand is converted first
  r = __fadd_rn(yy, t.rcr[v]);
  g = __fadd_rn(__fadd_rn(yy, t.gcb[u]), t.gcr[v]);
  b = __fadd_rn(yy, t.bcb[u]);
}

// one thread = one pixel (any alignment / width); sums: optional dense [height][width][3] float output of the unclamped sums
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_yuv888_to_rgb_float(const uint8_t *__restrict__ src, int irow, uint8_t *dst, int orow, int width,
                                                                int height, int in_alpha, RgbLayout out, const float *__restrict__ ftab,
                                                                const int32_t *__restrict__ rgb_y, float *__restrict__ sums) {
  __shared__ FloatTabs t;
  for (int i = threadIdx.x; i < 5 * 256; i += blockDim.x) (&t.ty[0])[i] = ftab[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) t.yi[i] = rgb_y[i];
  __syncthreads();
  const int ips = in_alpha ? 4 : 3;
  const bool vec4 = ips == 4 && out.psize == 4 && !(reinterpret_cast<uintptr_t>(src) & 3) && !(reinterpret_cast<uintptr_t>(dst) & 3) && !((irow | orow) & 3);
  const long long total = (long long)width * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / width), x = (int)(it - (long long)row * width);
    const uint8_t *q = src + (long long)irow * row + (long long)x * ips;
    uint32_t y, u, v, a = 255u;
    if (vec4) {
      const uint32_t w = ld_stream_u32(q);
      y = w & 0xFFu; u = (w >> 8) & 0xFFu; v = (w >> 16) & 0xFFu; a = w >> 24;
    } else {
      y = q[0]; u = q[1]; v = q[2];
      if (in_alpha) a = q[3];
    }
    float r, g, b;
    px_float<MODE>(t, y, u, v, r, g, b);
    if (sums) {
      float *sp = sums + 3 * it;
      sp[0] = r; sp[1] = g; sp[2] = b;
    }
    uint32_t w = (clamp0255f_dev(r) << (8 * out.r)) | (clamp0255f_dev(g) << (8 * out.g)) | (clamp0255f_dev(b) << (8 * out.b));
    if (out.a >= 0) w |= a << (8 * out.a);
    uint8_t *d = dst + (long long)orow * row + (long long)x * out.psize;
    if (vec4) st_stream_u32(d, w);
    else
      for (int k = 0; k < out.psize; k++) d[k] = (uint8_t)(w >> (8 * k));
  }
}

}  // namespace

cudaError_t launch_yuv888_to_rgb_float(const Launch &L, int mode, CImg src, Img dst, int width, int height, int in_alpha, RgbLayout out,
                                       const float *ftab_dev, const int32_t *rgb_y_dev, float *sums_dev) {
  if (width <= 0 || height <= 0) return cudaSuccess;
  long long blocks = ((long long)width * height + kBlock - 1) / kBlock;
  const long long cap = (long long)L.sm_count * 8;
  const int grid = (int)(blocks > cap ? cap : blocks);
  if (mode == 0) k_yuv888_to_rgb_float<0><<<grid, kBlock, 0, L.stream>>>(src.p, src.rs, dst.p, dst.rs, width, height, in_alpha, out, ftab_dev, rgb_y_dev, sums_dev);
  else k_yuv888_to_rgb_float<1><<<grid, kBlock, 0, L.stream>>>(src.p, src.rs, dst.p, dst.rs, width, height, in_alpha, out, ftab_dev, rgb_y_dev, sums_dev);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

}  // namespace pe
