// pe_kernels_mc.cu -- the one exchange step of the pixel path (BASELINE config 5): the shared transition operand of a
// multitrack crossfade lives on ONE rank and every rank needs it for every output frame.  k_mc_publish is the owner's side of that
// exchange as a single kernel: it reads the operand frames from the owner's HBM once and writes them through an NVSwitch MULTICAST
// address (multimem.st: the switch replicates the store into the mapped buffer of every GPU of the group), so the owner's NVLink
// egress carries the operand once whatever the number of receivers, nothing is staged, and no receiver runs a copy kernel.
// The kernel is small on purpose (256 threads, <= 32 registers, no shared memory): one CTA per SM fits BESIDE the persistent
// 512-thread conversion / crossfade CTAs (k_yuv_march: 107 registers, 96 KB), so publishing group t + 1 overlaps the owner's own
// crossfade of group t on the same SMs.  The multicast mapping itself (cuMulticast*, symmetric allocation, rendezvous) is plumbing
// done by the host (lives_b200/shard.py through torch's symmetric memory); this file only needs the address.
#include "pe_device.cuh"
#include "pe_kernels.h"

namespace pe {

namespace {

constexpr int MC_NT = 256;
#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

__device__ __forceinline__ void multimem_st_v4(void *mc, const uint4 &v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

// n16 16-byte vectors from src (local HBM) to mc_dst (multicast address); four independent vectors in flight per thread
__global__ void __launch_bounds__(MC_NT) k_mc_publish(const uint4 *__restrict__ src, uint4 *mc_dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * MC_NT;
  size_t i = (size_t)blockIdx.x * MC_NT + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const uint4 a = ld_stream_u4(src + i), b = ld_stream_u4(src + i + stride), c = ld_stream_u4(src + i + 2 * stride),
                d = ld_stream_u4(src + i + 3 * stride);
    multimem_st_v4(mc_dst + i, a);
    multimem_st_v4(mc_dst + i + stride, b);
    multimem_st_v4(mc_dst + i + 2 * stride, c);
    multimem_st_v4(mc_dst + i + 3 * stride, d);
  }
  for (; i < n16; i += stride) multimem_st_v4(mc_dst + i, ld_stream_u4(src + i));
}

}  // namespace

cudaError_t launch_mc_publish(const Launch &L, const void *src, void *mc_dst, size_t bytes, int max_ctas) {
  if (!bytes) return cudaSuccess;
  if ((bytes & 15) || ((uintptr_t)src & 15) || ((uintptr_t)mc_dst & 15)) return cudaErrorMisalignedAddress;
  const size_t n16 = bytes >> 4;
  long long grid = (long long)((n16 + MC_NT - 1) / MC_NT);
  const int cap = max_ctas > 0 ? max_ctas : L.sm_count;
  if (grid > cap) grid = cap;
  k_mc_publish<<<(int)grid, MC_NT, 0, L.stream>>>(reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(mc_dst), n16);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

}  // namespace pe
