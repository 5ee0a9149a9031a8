// pe_device.cuh -- device-side helpers shared by the kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pe {

// ---- vector loads / stores with streaming hints (data is touched once: keep it out of L1) ----

__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
// (the streaming stores carry no "memory" clobber: they only ever write output pixels that the same kernel never reads,
// and a clobber would force every loop to reload its constants and shared-memory values after each store)
// plain (coherent) variants for in-place kernels, where .nc would be illegal
__device__ __forceinline__ uint4 ld_u4(const void *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void st_stream_u4(void *p, const uint4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ void st_stream_u2(void *p, const uint2 &v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" :: "l"(p), "r"(v.x), "r"(v.y));
}
__device__ __forceinline__ void st_stream_u32(void *p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v));
}

// ---- byte helpers --------------------------------------------------------------------------------

// byte k of a word, zero extended: one PRMT (selector 0x444k: bytes 1..3 come from the zero operand)
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int k) { return __byte_perm(w, 0u, 0x4440u | (uint32_t)k); }
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
  return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}
__device__ __forceinline__ int clamp_i(int v, int lo, int hi) { return min(max(v, lo), hi); }

// (int)(n / 3. + .5) for the chroma weights of colourspace.c:3465; n = u1 + (u2 >> 1) <= 765.
// (2n + 3) / 6 with the division as a multiply-shift (exact for n < 21845, tests/test_host_logic.py)
__device__ __forceinline__ int third_round(int n) { return ((2 * n + 3) * 43691) >> 18; }

// Bank-replicated YUV -> RGB tables in shared memory (k_fused3, k_yuv_planar_to_rgb_fast): RGB_Y as [256][32 lanes] u32,
// {R_Cr, G_Cr} and {G_Cb, B_Cb} as [256][16] pairs.  cv = the 14 x 256 conversion tables in ConvTab order (RGB_Y = 9, R_CR = 10,
// G_CB = 11, G_CR = 12, B_CB = 13).  Every table value is loaded ONCE and stored to all its copies with 128-bit stores
// (a value-per-word loop is a chain of L2 round trips that costs several microseconds per launch).
template <int NT>
__device__ __forceinline__ void fill_replicated_yuv_tables(uint8_t *ty, uint8_t *tv, uint8_t *tu, const int32_t *__restrict__ cv, int tid) {
  // 8 consecutive lanes write the 8 x 16-byte chunks of one 128-byte entry: a warp's 128-bit store covers 512 contiguous bytes (4
  // wavefronts, the minimum; a thread writing its entry's 128 bytes alone is a 32-way bank conflict per store), and EVERY global
  // load is issued before the first store (the stores go through generic pointers: the compiler keeps a later load behind them).
  // Measured on k_fused3's copy of this fill: 7.4 -> 2.5 us per launch (profiles/r02zc_f3_split_variants.log).
  static_assert(256 * 8 % NT == 0, "table fill: whole rounds");
  constexpr int R = 256 * 8 / NT;
  uint32_t y[R], rcr[R], gcb[R], gcr[R], bcb[R];
#pragma unroll
  for (int q = 0; q < R; q++) {
    const int m = (tid + q * NT) >> 3;
    y[q] = (uint32_t)__ldg(cv + 9 * 256 + m); rcr[q] = (uint32_t)__ldg(cv + 10 * 256 + m); gcb[q] = (uint32_t)__ldg(cv + 11 * 256 + m);
    gcr[q] = (uint32_t)__ldg(cv + 12 * 256 + m); bcb[q] = (uint32_t)__ldg(cv + 13 * 256 + m);
  }
#pragma unroll
  for (int q = 0; q < R; q++) {
    const int i = tid + q * NT, m = i >> 3, j = i & 7;
    reinterpret_cast<uint4 *>(ty + 128 * m)[j] = make_uint4(y[q], y[q], y[q], y[q]);
    reinterpret_cast<uint4 *>(tv + 128 * m)[j] = make_uint4(rcr[q], gcr[q], rcr[q], gcr[q]);
    reinterpret_cast<uint4 *>(tu + 128 * m)[j] = make_uint4(gcb[q], bcb[q], gcb[q], bcb[q]);
  }
}

// grid-stride helpers
__device__ __forceinline__ long long global_tid() { return (long long)blockIdx.x * blockDim.x + threadIdx.x; }
__device__ __forceinline__ long long global_threads() { return (long long)gridDim.x * blockDim.x; }

}  // namespace pe
