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
__device__ __forceinline__ void fill_replicated_yuv_tables(uint8_t *ty, uint8_t *tv, uint8_t *tu, const int32_t *__restrict__ cv,
                                                           int tid, int nthreads, int ty_stride = 128) {
  for (int m = tid; m < 256; m += nthreads) {
    const uint32_t y = (uint32_t)cv[9 * 256 + m], rcr = (uint32_t)cv[10 * 256 + m], gcb = (uint32_t)cv[11 * 256 + m],
                   gcr = (uint32_t)cv[12 * 256 + m], bcb = (uint32_t)cv[13 * 256 + m];
    uint4 *py = reinterpret_cast<uint4 *>(ty + ty_stride * m), *pv = reinterpret_cast<uint4 *>(tv + 128 * m),
          *pu = reinterpret_cast<uint4 *>(tu + 128 * m);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      py[j] = make_uint4(y, y, y, y);
      pv[j] = make_uint4(rcr, gcr, rcr, gcr);
      pu[j] = make_uint4(gcb, bcb, gcb, bcb);
    }
  }
}

// grid-stride helpers
__device__ __forceinline__ long long global_tid() { return (long long)blockIdx.x * blockDim.x + threadIdx.x; }
__device__ __forceinline__ long long global_threads() { return (long long)gridDim.x * blockDim.x; }

}  // namespace pe
