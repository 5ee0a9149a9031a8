// pe_kernels_yuv2.cu -- the fast planar 4:2:0 / 4:2:2 -> RGB converter (convert_yuv420p_to_{rgb,bgr,argb}_frame,
// colourspace.c:3260-5127), built like the conversion stage of k_fused3 (pe_kernels_fused3.cu):
//   * RGB_Y, {R_Cr, G_Cr} and {G_Cb, B_Cb} REPLICATED ACROSS BANKS in shared memory (96 KB, one 512-thread CTA per SM): a lane
//     only touches its own bank (pair), so the five lookups of a pixel cost 1 + 2 + 2 wavefronts whatever the pixel values are
//     (the old kernel, k_yuv_planar_to_rgb, paid ~16 on random data);
//   * the chroma sums of colourspace.c:3440-3549 two columns at a time in the 16-bit halves of a register (Q = 2 n + 3), the
//     (int)(n / 3. + .5) rounding as one multiply-high that directly yields the table offset;
//   * one thread = 4 columns x one reference row pair (4:2:0) or one row (4:2:2); luma as one 32-bit word per row, chroma as
//     two words per row and plane, 128-bit (RGBA) or 3 x 32-bit (RGB24) streaming stores;
//   * optional fused crossfade with an operand frame (pe_fx_convert_crossfade, BASELINE config 5).
// Same results, bit for bit, as k_yuv_planar_to_rgb (which keeps PB_QUALITY_LOW, the inline 16-bit gamma LUT, widths that are
// not a multiple of 4 and unaligned planes).  Rows and columns at the frame edges follow the reference's edge rules through
// the scalar slow path below (row 0, the last row of an even frame, the last chroma row of a plane without padding, the
// 4:2:2 seed slip of column 0).
#include <cstdlib>

#include "pe_device.cuh"
#include "pe_kernels.h"
#include "pe_tables.h"

namespace pe {

namespace {

#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

constexpr int Y2_NT = 512;
constexpr int S2_TY = 0;                  // u32 [256][32]
constexpr int S2_TV = 32768;              // uint2 [256][16]: {R_Cr, G_Cr}
constexpr int S2_TU = 65536;              // uint2 [256][16]: {G_Cb, B_Cb}
constexpr int S2_RING = 98304;            // u32 [Y2_DEPTH][Y2_WORDS][Y2_NT]: per-thread cp.async ring of raw input words
constexpr int Y2_DEPTH = 3, Y2_WORDS = 17;  // 2 luma, 8 chroma, first V word, 2 x 3 operand words (crossfade)
constexpr int S2_BYTES = S2_RING + Y2_DEPTH * Y2_WORDS * Y2_NT * 4;

constexpr uint32_t MSK = 0xFFFEFFFEu;     // clears bit 0 of both halves
constexpr uint32_t K3 = 0x00030003u;

// d = (c[15:0] << 16) | (sat_u8(a) << 8) | sat_u8(b)
__device__ __forceinline__ uint32_t pack_sat(int a, int b, uint32_t c) {
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ void cp_async4(uint32_t smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_u32nc(const void *p) {  // chroma words are shared by neighbouring lanes / strips: keep them in L2
  uint32_t r;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_u8nc(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
// 128 * third_round(n) for the n in the high / low half of a packed Q = 2 n + 3 (see pe_kernels_fused3.cu, tests/test_host_logic.py)
__device__ __forceinline__ uint32_t idx_hi(uint32_t q) { return __umulhi(q, 10923u * 128u) & 0x7F80u; }
__device__ __forceinline__ uint32_t idx_lo(uint32_t q) { return __umulhi(q << 16, 10923u * 128u) & 0x7F80u; }

struct RowC {
  uint32_t a, b, c;  // [c(jc0), c(jc0+1)], [c(jc0-1), c(jc0)], [c(jc0+1), c(jc0+2)] as 16-bit halves
};
__device__ __forceinline__ RowC unpack_row(uint32_t w0, uint32_t w1, uint32_t sel) {
  const uint32_t cw4 = __byte_perm(w0, w1, sel);
  RowC r;
  r.a = __byte_perm(cw4, 0u, 0x4241u);
  r.b = __byte_perm(cw4, 0u, 0x4140u);
  r.c = __byte_perm(cw4, 0u, 0x4342u);
  return r;
}

// chroma sample with the reference's edge rules (column -1 replicates column 0; column cw: the byte behind the row)
__device__ __forceinline__ uint32_t chroma_at(const uint8_t *__restrict__ p, int stride, int r, int c, int cw, int ch) {
  if (c < 0) c = 0;
  if (c >= cw) c = (cw < stride || r + 1 < ch) ? cw : cw - 1;
  return p[(size_t)stride * r + c];
}

template <bool QUIRKS>
__global__ void __launch_bounds__(Y2_NT, 1) k_yuv_planar_to_rgb_fast(const __grid_constant__ YuvToRgbArgs A, int k_fast_max) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  fill_replicated_yuv_tables<Y2_NT>(smem + S2_TY, smem + S2_TV, smem + S2_TU, A.conv.t, tid);
  __syncthreads();

  const Planes &S = A.src;
  const int w = A.width, h = A.height, cw = S.cw, ch = S.ch;
  const int is422 = A.is_422;
  const uint32_t lane4 = 4u * (uint32_t)lane, lane8 = 8u * (uint32_t)(lane & 15);
  // canonical pixel word = [r, g, b, 255]; sel moves its bytes to the palette's order
  uint32_t osel;
  {
    uint32_t nib[4] = {3, 3, 3, 3};
    nib[A.out.r] = 0; nib[A.out.g] = 1; nib[A.out.b] = 2;
    if (A.out.a >= 0) nib[A.out.a] = 3;
    osel = nib[0] | (nib[1] << 4) | (nib[2] << 8) | (nib[3] << 12);
  }
  // one pixel: yuv2rgb_int (colourspace.c:2345-2356) through the replicated tables, packed in the palette's byte order
  auto px = [&](uint32_t y, uint32_t ou, uint32_t ov) -> uint32_t {
    const int yy = (int)*reinterpret_cast<const uint32_t *>(smem + S2_TY + (y * 128u + lane4));
    const uint2 tv = *reinterpret_cast<const uint2 *>(smem + S2_TV + (ov | lane8));
    const uint2 tu = *reinterpret_cast<const uint2 *>(smem + S2_TU + (ou | lane8));
    const int r = (yy + (int)tv.x) >> 16, g = (yy + (int)tu.x + (int)tv.y) >> 16, b = (yy + (int)tu.y) >> 16;
    return __byte_perm(pack_sat(g, r, pack_sat(255, b, 0u)), 0u, osel);
  };
  const uint32_t bf = (uint32_t)A.blend_bf & 0xFFu, nb = 255u - bf;
  const bool xf_vec = A.blend2 && ((((uintptr_t)A.blend2) | (uint32_t)A.blend2_rs) & 3) == 0;
  // store 4 pixels of one row (after the optional crossfade with the operand)
  auto store_row = [&](uint32_t *p4, int row, int x0, uint32_t opslot = 0u) {
    if (A.blend2) {  // dst = (bf * in2 + (255 - bf) * converted) >> 8 per byte (make_blend_table, simple_blend.c:31-35)
      const uint8_t *q = A.blend2 + (size_t)A.blend2_rs * row + (size_t)x0 * 3;
      uint32_t o[4];
      if (xf_vec) {
        uint32_t w0, w1, w2;
        if (opslot) { w0 = lds_u32(opslot); w1 = lds_u32(opslot + Y2_NT * 4); w2 = lds_u32(opslot + 2 * Y2_NT * 4); }  // prefetched
        else { w0 = ld_stream_u32(q); w1 = ld_stream_u32(q + 4); w2 = ld_stream_u32(q + 8); }
        o[0] = w0; o[1] = __byte_perm(w0, w1, 0x0543); o[2] = __byte_perm(w1, w2, 0x0432); o[3] = w2 >> 8;
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = (uint32_t)q[3 * k] | ((uint32_t)q[3 * k + 1] << 8) | ((uint32_t)q[3 * k + 2] << 16);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t even = (((p4[k] & 0x00FF00FFu) * nb + (o[k] & 0x00FF00FFu) * bf) >> 8) & 0x00FF00FFu;
        const uint32_t mid = (((p4[k] >> 8) & 0xFFu) * nb + ((o[k] >> 8) & 0xFFu) * bf) & 0xFF00u;
        p4[k] = even | mid;
      }
    }
    uint8_t *d = A.dst.p + (size_t)A.dst.rs * row + (size_t)x0 * A.out.psize;
    if (A.out.psize == 4) {
      st_stream_u4(d, make_uint4(p4[0], p4[1], p4[2], p4[3]));
    } else {
      st_stream_u32(d, __byte_perm(p4[0], p4[1], 0x4210));
      st_stream_u32(d + 4, __byte_perm(p4[1], p4[2], 0x5421));
      st_stream_u32(d + 8, __byte_perm(p4[2], p4[3], 0x6542));
    }
  };

  const int quads = w >> 2;
  const int njobs = is422 ? h : (!(h & 1) ? ch + 1 : ch);
  const long long total = (long long)quads * njobs;
  const uint32_t rs_y = (uint32_t)S.rs_y, rs_u = (uint32_t)S.rs_u, rs_v = (uint32_t)S.rs_v;
  // The raw words of a job (2 luma, 8 chroma, the row's first V word) travel through a per-thread cp.async ring in shared
  // memory, Y2_DEPTH - 1 grid strides ahead of their use: 16 warps per SM do not hide DRAM latency by themselves, and a register
  // pipeline that deep would spill.  Slot layout [depth][word][thread]: conflict-free, and a thread only reads what it copied.
  auto job_class = [&](int job, int x0) -> int {  // 0: 4:2:0 interior pair, 1: 4:2:2 row, 2: slow path
    if (!is422) return (job >= 1 && job <= k_fast_max) ? 0 : 2;
    return (!(QUIRKS && x0 == 0) && !(job == ch - 1 && k_fast_max < ch - 1)) ? 1 : 2;
  };
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem) + S2_RING + 4u * (uint32_t)tid;
  auto slot = [&](int d, int word) -> uint32_t { return ring + (uint32_t)((d * Y2_WORDS + word) * (Y2_NT * 4)); };
  auto issue = [&](int job, int g, int d) {
    const int x0 = 4 * g, o = 2 * g - 1;
    const int cls = job_class(job, x0);
    if (cls != 2) {
      const size_t off0 = x0 == 0 ? 0 : (size_t)(o & ~3);
      if (cls == 0) {
        const uint8_t *yr = S.y + (size_t)rs_y * (uint32_t)(2 * job - 1) + x0;
        cp_async4(slot(d, 0), yr); cp_async4(slot(d, 1), yr + rs_y);
        const uint8_t *up = S.u + (size_t)rs_u * (uint32_t)(job - 1) + off0, *vp = S.v + (size_t)rs_v * (uint32_t)(job - 1) + off0;
        cp_async4(slot(d, 2), up); cp_async4(slot(d, 3), up + 4); cp_async4(slot(d, 4), up + rs_u); cp_async4(slot(d, 5), up + rs_u + 4);
        cp_async4(slot(d, 6), vp); cp_async4(slot(d, 7), vp + 4); cp_async4(slot(d, 8), vp + rs_v); cp_async4(slot(d, 9), vp + rs_v + 4);
        if (QUIRKS) cp_async4(slot(d, 10), S.v + (size_t)rs_v * (uint32_t)job);
      } else {
        cp_async4(slot(d, 0), S.y + (size_t)rs_y * (uint32_t)job + x0);
        const uint8_t *up = S.u + (size_t)rs_u * (uint32_t)job + off0, *vp = S.v + (size_t)rs_v * (uint32_t)job + off0;
        cp_async4(slot(d, 2), up); cp_async4(slot(d, 3), up + 4);
        cp_async4(slot(d, 6), vp); cp_async4(slot(d, 7), vp + 4);
      }
    }
    if (xf_vec && cls != 2) {  // crossfade operand: the 12 bytes under the job's 4 pixels, per row
      const int ra = cls == 0 ? 2 * job - 1 : job;
      const uint8_t *q = A.blend2 + (size_t)A.blend2_rs * ra + (size_t)x0 * 3;
      cp_async4(slot(d, 11), q); cp_async4(slot(d, 12), q + 4); cp_async4(slot(d, 13), q + 8);
      if (cls == 0) {
        q += A.blend2_rs;
        cp_async4(slot(d, 14), q); cp_async4(slot(d, 15), q + 4); cp_async4(slot(d, 16), q + 8);
      }
    }
    cp_async_commit();  // one group per job, empty for the slow class: the wait below counts groups
  };
  struct Raw {
    uint32_t yA, yB, u00, u01, u10, u11, v00, v01, v10, v11, vf;
  };
  // (job, quad) advance by one grid stride without a division per job
  const long long stride = global_threads();
  const int d_job = (int)(stride / quads), d_g = (int)(stride - (long long)d_job * quads);
  auto advance = [&](int &job, int &g) {
    job += d_job; g += d_g;
    if (g >= quads) { g -= quads; job++; }
  };
  long long it = global_tid();
  int job = (int)(it / quads), g = (int)(it - (long long)job * quads);
  int pj = job, pg = g;            // the job the next issue is for
  long long pit = it;
#pragma unroll
  for (int d = 0; d < Y2_DEPTH - 1; d++) {
    if (pit < total) issue(pj, pg, d); else cp_async_commit();
    advance(pj, pg); pit += stride;
  }
  int d = 0;
  for (; it < total; it += stride, d = d + 1 == Y2_DEPTH ? 0 : d + 1) {
    if (pit < total) issue(pj, pg, d == 0 ? Y2_DEPTH - 1 : d - 1); else cp_async_commit();
    advance(pj, pg); pit += stride;
    cp_async_wait<Y2_DEPTH - 1>();  // all but the newest Y2_DEPTH - 1 groups have landed: this job's words are in slot d
    Raw cur;
    cur.yA = lds_u32(slot(d, 0)); cur.yB = lds_u32(slot(d, 1));
    cur.u00 = lds_u32(slot(d, 2)); cur.u01 = lds_u32(slot(d, 3)); cur.u10 = lds_u32(slot(d, 4)); cur.u11 = lds_u32(slot(d, 5));
    cur.v00 = lds_u32(slot(d, 6)); cur.v01 = lds_u32(slot(d, 7)); cur.v10 = lds_u32(slot(d, 8)); cur.v11 = lds_u32(slot(d, 9));
    cur.vf = lds_u32(slot(d, 10)) & 0xFFu;
    const int jobn = job + d_job + (g + d_g >= quads ? 1 : 0), gn = g + d_g >= quads ? g + d_g - quads : g + d_g;
    const int x0 = 4 * g, jc0 = 2 * g;
    const int cls = job_class(job, x0);
    uint32_t pa[4], pb[4];
    if (cls == 0) {
      // ===== 4:2:0 interior row pair (colourspace.c:3440-3549): rows 2k-1, 2k with chroma rows k-1, k
      const int k = job;
      const int o = jc0 - 1;
      const uint32_t sel = x0 == 0 ? 0x2100u : ((o & 3) == 3 ? 0x6543u : 0x4321u);
      const uint32_t yA = cur.yA, yB = cur.yB;
      const RowC U0 = unpack_row(cur.u00, cur.u01, sel), U1 = unpack_row(cur.u10, cur.u11, sel);
      const RowC V0 = unpack_row(cur.v00, cur.v01, sel), V1 = unpack_row(cur.v10, cur.v11, sel);
      // right pixel of both chroma columns: this + next
      const uint32_t RU0 = U0.a + U0.c, RU1 = U1.a + U1.c, RV0 = V0.a + V0.c, RV1 = V1.a + V1.c;
      const uint32_t QUR_up = RU0 * 2u + (RU1 & MSK) + K3, QUR_lo = (RU0 & MSK) + RU1 * 2u + K3;
      const uint32_t QVR_up = RV0 * 2u + (RV1 & MSK) + K3, QVR_lo = (RV0 & MSK) + RV1 * 2u + K3;
      // left pixel: this + last, with the reference's slips under QUIRKS
      uint32_t QUL_up, QUL_lo, QVL_up, QVL_lo;
      const uint32_t LU0 = U0.a + U0.b;
      if (QUIRKS) {
        QUL_up = QUL_lo = LU0 * 2u + (LU0 & MSK) + K3;                              // u2 = this_u1 + last_u1 (:3461)
        const uint32_t bq = x0 == 0 ? __byte_perm(V1.b, V0.a, 0x3254u) : V1.b;      // last_v1 = this_v2 (:3544), except at column 0
        const uint32_t v1 = V0.a + bq;
        const uint32_t v2 = V1.a + cur.vf * 0x10001u;                               // last_v2 never advanced: the row's first V sample
        QVL_up = v1 * 2u + (v2 & MSK) + K3; QVL_lo = (v1 & MSK) + v2 * 2u + K3;
      } else {
        const uint32_t LU1 = U1.a + U1.b, LV0 = V0.a + V0.b, LV1 = V1.a + V1.b;
        QUL_up = LU0 * 2u + (LU1 & MSK) + K3; QUL_lo = (LU0 & MSK) + LU1 * 2u + K3;
        QVL_up = LV0 * 2u + (LV1 & MSK) + K3; QVL_lo = (LV0 & MSK) + LV1 * 2u + K3;
      }
#pragma unroll
      for (int col = 0; col < 4; col++) {
        const bool hi_half = col >> 1, right = col & 1;
        const uint32_t qu_up = right ? QUR_up : QUL_up, qu_lo = right ? QUR_lo : QUL_lo;
        const uint32_t qv_up = right ? QVR_up : QVL_up, qv_lo = right ? QVR_lo : QVL_lo;
        const uint32_t ou_up = hi_half ? idx_hi(qu_up) : idx_lo(qu_up), ov_up = hi_half ? idx_hi(qv_up) : idx_lo(qv_up);
        const uint32_t ov_lo = hi_half ? idx_hi(qv_lo) : idx_lo(qv_lo);
        const uint32_t ou_lo = (QUIRKS && !right) ? ou_up : (hi_half ? idx_hi(qu_lo) : idx_lo(qu_lo));
        pa[col] = px(byte_of(yA, col), ou_up, ov_up);
        pb[col] = px(byte_of(yB, col), ou_lo, ov_lo);
      }
      store_row(pa, 2 * k - 1, x0, slot(d, 11));
      store_row(pb, 2 * k, x0, slot(d, 14));
      job = jobn; g = gn;
      continue;
    }
    if (cls == 1) {
      // ===== 4:2:2 row (:3598-3642): horizontal average only, m = (this + neighbour) >> 1
      const int o = jc0 - 1;
      const uint32_t sel = x0 == 0 ? 0x2100u : ((o & 3) == 3 ? 0x6543u : 0x4321u);
      const uint32_t yA = cur.yA;
      const RowC U = unpack_row(cur.u00, cur.u01, sel), V = unpack_row(cur.v00, cur.v01, sel);
      const uint32_t LU = U.a + U.b, RU = U.a + U.c, LV = V.a + V.b, RV = V.a + V.c;   // sums <= 510 per half
#pragma unroll
      for (int col = 0; col < 4; col++) {
        const bool hi_half = col >> 1, right = col & 1;
        const uint32_t su = right ? RU : LU, sv = right ? RV : LV;
        // 128 * (s >> 1) = (s & ~1) << 6
        const uint32_t ou = hi_half ? ((su >> 10) & 0x7F80u) : ((su << 6) & 0x7F80u);
        const uint32_t ov = hi_half ? ((sv >> 10) & 0x7F80u) : ((sv << 6) & 0x7F80u);
        pa[col] = px(byte_of(yA, col), ou, ov);
      }
      store_row(pa, job, x0, slot(d, 11));
      job = jobn; g = gn;
      continue;
    }
    // ===== slow path: scalar code with the reference's edge rules
    {
      auto single = [&](int row, int cr, int seed_row) {
        const uint32_t yw = *reinterpret_cast<const uint32_t *>(S.y + (size_t)rs_y * row + x0);
#pragma unroll
        for (int col = 0; col < 4; col++) {
          const int jc = jc0 + (col >> 1), jo = (col & 1) ? jc + 1 : jc - 1;
          uint32_t ua = chroma_at(S.u, rs_u, cr, jc, cw, ch), ub = chroma_at(S.u, rs_u, cr, jo, cw, ch);
          uint32_t va = chroma_at(S.v, rs_v, cr, jc, cw, ch), vb = chroma_at(S.v, rs_v, cr, jo, cw, ch);
          if (jc0 == 0 && seed_row != cr) {  // 4:2:2 seed slip (:3600): columns <= 0 are column 0 of chroma row (i >> 1)
            const uint32_t su = S.u[(size_t)rs_u * seed_row], sv = S.v[(size_t)rs_v * seed_row];
            if (jc == 0) { ua = su; va = sv; }
            if (jo <= 0) { ub = su; vb = sv; }
          }
          pa[col] = px(byte_of(yw, col), ((ua + ub) >> 1) * 128u, ((va + vb) >> 1) * 128u);
        }
        store_row(pa, row, x0);
      };
      if (is422) {
        single(job, job, QUIRKS ? (job >> 1) : job);
      } else if (job == 0) {
        single(0, 0, 0);
      } else if (job < ch) {
        const int k = job, ca = k - 1, cb = k;
        const uint32_t ya = *reinterpret_cast<const uint32_t *>(S.y + (size_t)rs_y * (2 * k - 1) + x0);
        const uint32_t yb = *reinterpret_cast<const uint32_t *>(S.y + (size_t)rs_y * (2 * k) + x0);
        const uint32_t vfirst = S.v[(size_t)rs_v * cb];
#pragma unroll
        for (int col = 0; col < 4; col++) {
          const int jc = jc0 + (col >> 1), jo = (col & 1) ? jc + 1 : jc - 1;
          uint32_t u1 = chroma_at(S.u, rs_u, ca, jc, cw, ch) + chroma_at(S.u, rs_u, ca, jo, cw, ch);
          uint32_t u2 = chroma_at(S.u, rs_u, cb, jc, cw, ch) + chroma_at(S.u, rs_u, cb, jo, cw, ch);
          uint32_t v1 = chroma_at(S.v, rs_v, ca, jc, cw, ch) + chroma_at(S.v, rs_v, ca, jo, cw, ch);
          uint32_t v2 = chroma_at(S.v, rs_v, cb, jc, cw, ch) + chroma_at(S.v, rs_v, cb, jo, cw, ch);
          if (QUIRKS && !(col & 1)) {
            u2 = u1;
            if (jc > 0) v1 = chroma_at(S.v, rs_v, ca, jc, cw, ch) + chroma_at(S.v, rs_v, cb, jo, cw, ch);
            v2 = chroma_at(S.v, rs_v, cb, jc, cw, ch) + vfirst;
          }
          const uint32_t mu3 = (uint32_t)third_round((int)(u1 + (u2 >> 1))), mu4 = (uint32_t)third_round((int)((u1 >> 1) + u2));
          const uint32_t mv3 = (uint32_t)third_round((int)(v1 + (v2 >> 1))), mv4 = (uint32_t)third_round((int)((v1 >> 1) + v2));
          pa[col] = px(byte_of(ya, col), mu3 * 128u, mv3 * 128u);
          pb[col] = px(byte_of(yb, col), mu4 * 128u, mv4 * 128u);
        }
        store_row(pa, 2 * k - 1, x0);
        store_row(pb, 2 * k, x0);
      } else {
        single(h - 1, ch - 1, ch - 1);
      }
    }
    job = jobn; g = gn;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------------------------------------
// k_yuv_march: the same conversion, organised like k_fused3 (pe_kernels_fused3.cu) for frames tall enough to march through: a
// warp owns a strip of 128 columns and walks down the rows one reference row pair per step.  What the job kernel above pays per
// 4 x 2 pixels -- (job, quad) bookkeeping, 64-bit addresses of ten cp.async, both chroma rows of the pair unpacked -- is paid
// once per strip here: addresses advance by the row strides, the sums of the previous chroma row are CARRIED from step to step
// (4:2:0), and the raw words of step k + 1 are loaded into registers while step k is computed.
// 4:2:2: a step is two independent rows (horizontal average only); the lane of column 0 takes the scalar path when the seed slip
// (colourspace.c:3600) is on.
// frames of one batched launch (same geometry, strides, palettes and tables: everything else comes from the common YuvToRgbArgs)
struct YuvFrameList {
  const uint8_t *y[32], *u[32], *v[32];
  uint8_t *dst[32];
  const uint8_t *blend2[32];   // crossfade operand of each frame (all null or all set: YuvToRgbArgs::blend2 says which)
  int n;
};

struct MarchPre {
  uint32_t yA, yB, u0, u1, v0, v1, u2, u3, v2, v3, vf, sd;   // 4:2:0: chroma row k in u0 u1 v0 v1; 4:2:2: rows 2k-1 / 2k in (u0 u1 v0 v1) / (u2 u3 v2 v3)
};
struct MarchCarry {
  uint32_t DUr, MUr, DVr, MVr, DUl, MUl, DVl, MVl, QUL, aV;
};

template <bool QUIRKS, bool IS422>
__global__ void __launch_bounds__(Y2_NT, 1) k_yuv_march(const __grid_constant__ YuvToRgbArgs A, const __grid_constant__ YuvFrameList FL, int k_fast_max, int band_h) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  fill_replicated_yuv_tables<Y2_NT>(smem + S2_TY, smem + S2_TV, smem + S2_TU, A.conv.t, tid);
  __syncthreads();

  const Planes &S = A.src;
  const int w = A.width, h = A.height, cw = S.cw, ch = S.ch;
  const uint32_t lane4 = 4u * (uint32_t)lane, lane8 = 8u * (uint32_t)(lane & 15);
  uint32_t osel;
  {
    uint32_t nib[4] = {3, 3, 3, 3};
    nib[A.out.r] = 0; nib[A.out.g] = 1; nib[A.out.b] = 2;
    if (A.out.a >= 0) nib[A.out.a] = 3;
    osel = nib[0] | (nib[1] << 4) | (nib[2] << 8) | (nib[3] << 12);
  }
  auto px = [&](uint32_t y, uint32_t ou, uint32_t ov) -> uint32_t {
    const int yy = (int)*reinterpret_cast<const uint32_t *>(smem + S2_TY + (y * 128u + lane4));
    const uint2 tv = *reinterpret_cast<const uint2 *>(smem + S2_TV + (ov | lane8));
    const uint2 tu = *reinterpret_cast<const uint2 *>(smem + S2_TU + (ou | lane8));
    const int r = (yy + (int)tv.x) >> 16, g = (yy + (int)tu.x + (int)tv.y) >> 16, b = (yy + (int)tu.y) >> 16;
    return __byte_perm(pack_sat(g, r, pack_sat(255, b, 0u)), 0u, osel);
  };
  const uint32_t bf = (uint32_t)A.blend_bf & 0xFFu, nb = 255u - bf;
  const bool xf = A.blend2 != nullptr;
  const uint32_t rs_y = (uint32_t)S.rs_y, rs_u = (uint32_t)S.rs_u, rs_v = (uint32_t)S.rs_v, rs_d = (uint32_t)A.dst.rs,
                 rs_x = (uint32_t)A.blend2_rs;

  // ---- the warp's share of the (band, strip, row) sequence: every row costs the same
  const int nstrips = (w + 127) >> 7;
  const long long per_frame = (long long)nstrips * h;
  const long long total = per_frame * FL.n;   // FL.n frames of the same geometry (a batch is one launch)
  const long long gw = (long long)blockIdx.x * (Y2_NT / 32) + warp, nwarps = (long long)gridDim.x * (Y2_NT / 32);
  long long pos = total * gw / nwarps;
  const long long pos_end = total * (gw + 1) / nwarps;
  while (pos < pos_end) {
    const int fidx = (int)(pos / per_frame);
    const long long fpos = pos - (long long)fidx * per_frame;
    const long long band_units = (long long)band_h * nstrips;
    const int b = (int)(fpos / band_units);
    const int r0 = b * band_h, bh = min(band_h, h - r0);
    const long long rem = fpos - (long long)b * band_units;
    const int s = (int)(rem / bh), within = (int)(rem - (long long)s * bh);
    const long long unit0 = pos - within;
    const int hi = (int)min((long long)bh, pos_end - unit0);
    const int ra = r0 + within, rb = r0 + hi;
    pos = unit0 + bh;
    const int x = 128 * s + 4 * lane;
    if (x >= w || ra >= rb) continue;

    const int jc0 = x >> 1, o = jc0 - 1;
    const size_t off0 = x == 0 ? 0 : (size_t)(o & ~3);
    const uint32_t sel = x == 0 ? 0x2100u : ((o & 3) == 3 ? 0x6543u : 0x4321u);
    const uint32_t selB = x == 0 ? 0x3254u : 0x3210u;
    const uint8_t *const Fy = FL.y[fidx], *const Fu = FL.u[fidx], *const Fv = FL.v[fidx];
    const uint8_t *yp = Fy + x, *up = Fu + off0, *vp = Fv + off0;
    uint8_t *dp = FL.dst[fidx] + (size_t)x * A.out.psize;
    const uint8_t *xp = xf ? FL.blend2[fidx] + (size_t)x * 3 : nullptr;
    const bool seed_lane = IS422 && QUIRKS && x == 0;

    auto store_row = [&](uint32_t *p4, int row, const uint32_t *op) {
      uint8_t *d = dp + (size_t)rs_d * (uint32_t)row;
      if (A.out.psize == 4) {
        st_stream_u4(d, make_uint4(p4[0], p4[1], p4[2], p4[3]));
        return;
      }
      uint32_t w3[3] = {__byte_perm(p4[0], p4[1], 0x4210), __byte_perm(p4[1], p4[2], 0x5421), __byte_perm(p4[2], p4[3], 0x6542)};
      if (xf) {
        // dst = (bf * in2 + (255 - bf) * converted) >> 8 per byte (make_blend_table, simple_blend.c:31-35), on the 12 PACKED bytes:
        // even / odd bytes of a word as 16-bit pairs (one PRMT each), two multiply-adds per pair (<= 255 * 255: no carry between
        // the halves), and ONE PRMT that picks byte 1 of every product.  5 ALU-pipe + 4 FMA-pipe instructions per 4 bytes; the
        // per-pixel mask / shift form this replaces cost 10 + 4 per 3 bytes, and the kernel is bound by the ALU pipe
        // (profiles/r01u_k_yuv_march_cfg5_ncu_full.txt: 72 %)
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const uint32_t E = __byte_perm(w3[k], 0u, 0x4240) * nb + __byte_perm(op[k], 0u, 0x4240) * bf;
          const uint32_t O = __byte_perm(w3[k], 0u, 0x4341) * nb + __byte_perm(op[k], 0u, 0x4341) * bf;
          w3[k] = __byte_perm(E, O, 0x7351);
        }
      }
      st_stream_u32(d, w3[0]);
      st_stream_u32(d + 4, w3[1]);
      st_stream_u32(d + 8, w3[2]);
    };
    auto load_op = [&](int row, uint32_t *op) {   // crossfade operand: the 12 bytes under the lane's 4 pixels
      if (!xf) return;
      const uint8_t *q = xp + (size_t)rs_x * (uint32_t)row;
      op[0] = ld_stream_u32(q); op[1] = ld_stream_u32(q + 4); op[2] = ld_stream_u32(q + 8);
    };
    // scalar single row with the reference's edge rules (row 0, the last row of an even 4:2:0 frame, unsafe rows, the 4:2:2 seed lane)
    auto single = [&](int row, int cr, int seed_row, uint32_t *out4) {
      const uint32_t yw = *reinterpret_cast<const uint32_t *>(Fy + (size_t)rs_y * row + x);
#pragma unroll
      for (int col = 0; col < 4; col++) {
        const int jc = jc0 + (col >> 1), jo = (col & 1) ? jc + 1 : jc - 1;
        uint32_t ua = chroma_at(Fu, rs_u, cr, jc, cw, ch), ub = chroma_at(Fu, rs_u, cr, jo, cw, ch);
        uint32_t va = chroma_at(Fv, rs_v, cr, jc, cw, ch), vb = chroma_at(Fv, rs_v, cr, jo, cw, ch);
        if (jc0 == 0 && seed_row != cr) {  // 4:2:2 seed slip (:3600): columns <= 0 are column 0 of chroma row (i >> 1)
          const uint32_t su = Fu[(size_t)rs_u * seed_row], sv = Fv[(size_t)rs_v * seed_row];
          if (jc == 0) { ua = su; va = sv; }
          if (jo <= 0) { ub = su; vb = sv; }
        }
        out4[col] = px(byte_of(yw, col), ((ua + ub) >> 1) * 128u, ((va + vb) >> 1) * 128u);
      }
    };
    auto emit_single = [&](int row, int cr, int seed_row) {
      uint32_t p4[4], op[3] = {0u, 0u, 0u};
      load_op(row, op);
      single(row, cr, seed_row, p4);
      store_row(p4, row, op);
    };

    auto load_pre = [&](int k, MarchPre &p) {
      const uint8_t *yr = yp + (size_t)rs_y * (uint32_t)(2 * k - 1);
      p.yA = ld_stream_u32(yr); p.yB = ld_stream_u32(yr + rs_y);
      if (IS422) {
        const uint8_t *ur = up + (size_t)rs_u * (uint32_t)(2 * k - 1), *vr = vp + (size_t)rs_v * (uint32_t)(2 * k - 1);
        p.u0 = ld_u32nc(ur); p.u1 = ld_u32nc(ur + 4); p.u2 = ld_u32nc(ur + rs_u); p.u3 = ld_u32nc(ur + rs_u + 4);
        p.v0 = ld_u32nc(vr); p.v1 = ld_u32nc(vr + 4); p.v2 = ld_u32nc(vr + rs_v); p.v3 = ld_u32nc(vr + rs_v + 4);
        if (seed_lane)  // the seed slip (:3600): columns <= 0 of row i are column 0 of chroma row (i >> 1) = k - 1 / k
          p.sd = ld_u8nc(Fu + (size_t)rs_u * (uint32_t)(k - 1)) | (ld_u8nc(Fv + (size_t)rs_v * (uint32_t)(k - 1)) << 8) |
                 (ld_u8nc(Fu + (size_t)rs_u * (uint32_t)k) << 16) | (ld_u8nc(Fv + (size_t)rs_v * (uint32_t)k) << 24);
      } else {
        const uint8_t *ur = up + (size_t)rs_u * (uint32_t)k, *vr = vp + (size_t)rs_v * (uint32_t)k;
        p.u0 = ld_u32nc(ur); p.u1 = ld_u32nc(ur + 4);
        p.v0 = ld_u32nc(vr); p.v1 = ld_u32nc(vr + 4);
        if (QUIRKS) p.vf = ld_u8nc(Fv + (size_t)rs_v * (uint32_t)k);
      }
    };
    auto init_carry = [&](int r, MarchCarry &c) {  // sums of chroma row r as a fast 4:2:0 step leaves them
      const uint8_t *ur = up + (size_t)rs_u * (uint32_t)r, *vr = vp + (size_t)rs_v * (uint32_t)r;
      const RowC U = unpack_row(ld_u32nc(ur), ld_u32nc(ur + 4), sel), V = unpack_row(ld_u32nc(vr), ld_u32nc(vr + 4), sel);
      const uint32_t RU = U.a + U.c, RV = V.a + V.c, LU = U.a + U.b, LV = V.a + V.b;
      c.DUr = RU * 2u; c.MUr = RU & MSK; c.DVr = RV * 2u; c.MVr = RV & MSK;
      c.DUl = LU * 2u; c.MUl = LU & MSK; c.DVl = LV * 2u; c.MVl = LV & MSK;
      c.QUL = LU * 2u + (LU & MSK) + K3;
      c.aV = V.a;
    };

    const int k_first = (ra + 1) >> 1, k_last = rb >> 1;
    MarchPre pre;
    MarchCarry C;
    pre.yA = pre.yB = pre.u0 = pre.u1 = pre.v0 = pre.v1 = pre.u2 = pre.u3 = pre.v2 = pre.v3 = pre.vf = pre.sd = 0u;
    C.DUr = C.MUr = C.DVr = C.MVr = C.DUl = C.MUl = C.DVl = C.MVl = C.QUL = C.aV = 0u;
    bool carry_ok = false;
    auto is_fast = [&](int k) { return k >= 1 && k <= k_fast_max; };
    if (is_fast(k_first)) {
      load_pre(k_first, pre);
      if (!IS422) { init_carry(k_first - 1, C); carry_ok = true; }
    }
    for (int k = k_first; k <= k_last; k++) {
      const int rowA = 2 * k - 1, rowB = 2 * k;
      const bool stA = rowA >= ra && rowA < rb, stB = rowB >= ra && rowB < rb && rowB < h;
      MarchPre nxt = pre;   // (consumed at the end of the step: see k_fused3)
      const bool next_fast = k + 1 <= k_last && is_fast(k + 1);
      if (next_fast) load_pre(k + 1, nxt);
      if (is_fast(k)) {
        uint32_t opA[3] = {0u, 0u, 0u}, opB[3] = {0u, 0u, 0u};
        if (stA) load_op(rowA, opA);
        if (stB) load_op(rowB, opB);
        uint32_t pa[4], pb[4];
        if (IS422) {
          {
            // the seed lane (column 0 under the seed slip) stays on the fast path: sample 0 of its chroma words is replaced by the
            // seed sample, which column -1 then replicates through the lane's PRMT selector like any frame edge
            uint32_t u0 = pre.u0, v0 = pre.v0, u2 = pre.u2, v2 = pre.v2;
            if (seed_lane) {
              u0 = __byte_perm(u0, pre.sd, 0x3214); v0 = __byte_perm(v0, pre.sd, 0x3215);
              u2 = __byte_perm(u2, pre.sd, 0x3216); v2 = __byte_perm(v2, pre.sd, 0x3217);
            }
            const RowC UA = unpack_row(u0, pre.u1, sel), VA = unpack_row(v0, pre.v1, sel);
            const RowC UB = unpack_row(u2, pre.u3, sel), VB = unpack_row(v2, pre.v3, sel);
            const uint32_t LUA = UA.a + UA.b, RUA = UA.a + UA.c, LVA = VA.a + VA.b, RVA = VA.a + VA.c;
            const uint32_t LUB = UB.a + UB.b, RUB = UB.a + UB.c, LVB = VB.a + VB.b, RVB = VB.a + VB.c;
#pragma unroll
            for (int col = 0; col < 4; col++) {
              const bool hi_half = col >> 1, right = col & 1;
              const uint32_t sua = right ? RUA : LUA, sva = right ? RVA : LVA, sub = right ? RUB : LUB, svb = right ? RVB : LVB;
              // 128 * (s >> 1) = (s & ~1) << 6
              pa[col] = px(byte_of(pre.yA, col), hi_half ? ((sua >> 10) & 0x7F80u) : ((sua << 6) & 0x7F80u),
                           hi_half ? ((sva >> 10) & 0x7F80u) : ((sva << 6) & 0x7F80u));
              pb[col] = px(byte_of(pre.yB, col), hi_half ? ((sub >> 10) & 0x7F80u) : ((sub << 6) & 0x7F80u),
                           hi_half ? ((svb >> 10) & 0x7F80u) : ((svb << 6) & 0x7F80u));
            }
          }
        } else {
          const RowC U = unpack_row(pre.u0, pre.u1, sel), V = unpack_row(pre.v0, pre.v1, sel);
          const uint32_t RU = U.a + U.c, RV = V.a + V.c;
          const uint32_t DUn = RU * 2u, MUn = RU & MSK, DVn = RV * 2u, MVn = RV & MSK;
          const uint32_t QUR_up = C.DUr + MUn + K3, QUR_lo = C.MUr + DUn + K3;
          const uint32_t QVR_up = C.DVr + MVn + K3, QVR_lo = C.MVr + DVn + K3;
          C.DUr = DUn; C.MUr = MUn; C.DVr = DVn; C.MVr = MVn;
          uint32_t QUL_up, QUL_lo, QVL_up, QVL_lo;
          const uint32_t LU = U.a + U.b;
          if (QUIRKS) {
            QUL_up = QUL_lo = C.QUL;                                  // u2 = this_u1 + last_u1 (colourspace.c:3461)
            C.QUL = LU * 2u + (LU & MSK) + K3;
            const uint32_t bq = __byte_perm(V.b, C.aV, selB);         // last_v1 = this_v2 (:3544), except at column 0
            const uint32_t v1 = C.aV + bq;
            const uint32_t v2 = V.a + pre.vf * 0x10001u;              // last_v2 is never advanced
            QVL_up = v1 * 2u + (v2 & MSK) + K3; QVL_lo = (v1 & MSK) + v2 * 2u + K3;
            C.aV = V.a;
          } else {
            const uint32_t LV = V.a + V.b;
            const uint32_t DUn_l = LU * 2u, MUn_l = LU & MSK, DVn_l = LV * 2u, MVn_l = LV & MSK;
            QUL_up = C.DUl + MUn_l + K3; QUL_lo = C.MUl + DUn_l + K3;
            QVL_up = C.DVl + MVn_l + K3; QVL_lo = C.MVl + DVn_l + K3;
            C.DUl = DUn_l; C.MUl = MUn_l; C.DVl = DVn_l; C.MVl = MVn_l;
          }
#pragma unroll
          for (int col = 0; col < 4; col++) {
            const bool hi_half = col >> 1, right = col & 1;
            const uint32_t qu_up = right ? QUR_up : QUL_up, qu_lo = right ? QUR_lo : QUL_lo;
            const uint32_t qv_up = right ? QVR_up : QVL_up, qv_lo = right ? QVR_lo : QVL_lo;
            const uint32_t ou_up = hi_half ? idx_hi(qu_up) : idx_lo(qu_up), ov_up = hi_half ? idx_hi(qv_up) : idx_lo(qv_up);
            const uint32_t ov_lo = hi_half ? idx_hi(qv_lo) : idx_lo(qv_lo);
            const uint32_t ou_lo = (QUIRKS && !right) ? ou_up : (hi_half ? idx_hi(qu_lo) : idx_lo(qu_lo));
            pa[col] = px(byte_of(pre.yA, col), ou_up, ov_up);
            pb[col] = px(byte_of(pre.yB, col), ou_lo, ov_lo);
          }
        }
        if (stA) store_row(pa, rowA, opA);
        if (stB) store_row(pb, rowB, opB);
      } else {
        // ---- slow step: frame edges
        if (IS422) {
          if (stA && rowA >= 0) emit_single(rowA, rowA, QUIRKS ? rowA >> 1 : rowA);
          if (stB) emit_single(rowB, rowB, QUIRKS ? rowB >> 1 : rowB);
        } else if (k <= 0) {
          if (stB) emit_single(0, 0, 0);
        } else if (2 * k <= h - 1) {
          // the last chroma row of a plane without padding: scalar pair with the edge rules
          const int ca = k - 1, cb = k;
          const uint32_t ya = *reinterpret_cast<const uint32_t *>(Fy + (size_t)rs_y * rowA + x);
          const uint32_t yb = *reinterpret_cast<const uint32_t *>(Fy + (size_t)rs_y * rowB + x);
          const uint32_t vfirst = Fv[(size_t)rs_v * cb];
          uint32_t pa[4], pb[4], opA[3] = {0u, 0u, 0u}, opB[3] = {0u, 0u, 0u};
          if (stA) load_op(rowA, opA);
          if (stB) load_op(rowB, opB);
#pragma unroll
          for (int col = 0; col < 4; col++) {
            const int jc = jc0 + (col >> 1), jo = (col & 1) ? jc + 1 : jc - 1;
            uint32_t u1 = chroma_at(Fu, rs_u, ca, jc, cw, ch) + chroma_at(Fu, rs_u, ca, jo, cw, ch);
            uint32_t u2 = chroma_at(Fu, rs_u, cb, jc, cw, ch) + chroma_at(Fu, rs_u, cb, jo, cw, ch);
            uint32_t v1 = chroma_at(Fv, rs_v, ca, jc, cw, ch) + chroma_at(Fv, rs_v, ca, jo, cw, ch);
            uint32_t v2 = chroma_at(Fv, rs_v, cb, jc, cw, ch) + chroma_at(Fv, rs_v, cb, jo, cw, ch);
            if (QUIRKS && !(col & 1)) {
              u2 = u1;
              if (jc > 0) v1 = chroma_at(Fv, rs_v, ca, jc, cw, ch) + chroma_at(Fv, rs_v, cb, jo, cw, ch);
              v2 = chroma_at(Fv, rs_v, cb, jc, cw, ch) + vfirst;
            }
            const uint32_t mu3 = (uint32_t)third_round((int)(u1 + (u2 >> 1))), mu4 = (uint32_t)third_round((int)((u1 >> 1) + u2));
            const uint32_t mv3 = (uint32_t)third_round((int)(v1 + (v2 >> 1))), mv4 = (uint32_t)third_round((int)((v1 >> 1) + v2));
            pa[col] = px(byte_of(ya, col), mu3 * 128u, mv3 * 128u);
            pb[col] = px(byte_of(yb, col), mu4 * 128u, mv4 * 128u);
          }
          if (stA) store_row(pa, rowA, opA);
          if (stB) store_row(pb, rowB, opB);
        } else if (rowA == h - 1) {
          if (stA) emit_single(h - 1, ch - 1, ch - 1);
        }
        carry_ok = false;
      }
      if (!IS422 && next_fast && !carry_ok) {
        init_carry(k, C);
        carry_ok = true;
      }
      pre = nxt;
    }
  }
}

}  // namespace

// Can the fast converter take this frame?  (checked by launch_yuv_planar_to_rgb; everything else: k_yuv_planar_to_rgb)
bool yuv_planar_fast_ok(const YuvToRgbArgs &a, const ConvTables *host_tables) {
  if (a.low_quality || a.lut16) return false;
  if ((a.width & 3) || a.width < 4 || a.height < 2) return false;
  if (!a.is_422 && (a.height & 1)) return false;                         // 4:2:0 frames are even (colourspace.c:11603)
  if (a.src.cw != a.width / 2 || a.src.ch != (a.is_422 ? a.height : (a.height + 1) / 2)) return false;
  if (((uintptr_t)a.src.y | (uintptr_t)a.src.u | (uintptr_t)a.src.v) & 3) return false;
  if ((a.src.rs_y | a.src.rs_u | a.src.rs_v) & 3) return false;
  if (a.out.psize == 4 ? ((((uintptr_t)a.dst.p) | (uint32_t)a.dst.rs) & 15) != 0 : ((((uintptr_t)a.dst.p) | (uint32_t)a.dst.rs) & 3) != 0)
    return false;
  if (a.blend2 && a.out.psize != 3) return false;
  return host_tables && fused3_tables_ok(*host_tables);
}

static void launch_march(const Launch &L, const YuvToRgbArgs &a, const YuvFrameList &fl, int kmax, int band_h) {
  const int grid = L.sm_count;
  if (a.quirks) {
    if (a.is_422) k_yuv_march<true, true><<<grid, Y2_NT, S2_RING, L.stream>>>(a, fl, kmax, band_h);
    else k_yuv_march<true, false><<<grid, Y2_NT, S2_RING, L.stream>>>(a, fl, kmax, band_h);
  } else {
    if (a.is_422) k_yuv_march<false, true><<<grid, Y2_NT, S2_RING, L.stream>>>(a, fl, kmax, band_h);
    else k_yuv_march<false, false><<<grid, Y2_NT, S2_RING, L.stream>>>(a, fl, kmax, band_h);
  }
}

static cudaError_t march_attrs() {
  static PerDevice march_attr;
  if (!march_attr.cur()) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_yuv_march<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_RING)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_yuv_march<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_RING)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_yuv_march<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_RING)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_yuv_march<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_RING)) != cudaSuccess) return e;
    march_attr.cur() = 1;
  }
  return cudaSuccess;
}

// A batch of frames that share everything but their pointers (geometry, strides, palettes, tables, flags) in ONE launch of the
// marching kernel: the per-launch costs (table fill, first loads, tail) are paid once, and the warps get shares long enough to
// march.  frames[i] must all pass yuv_planar_fast_ok and compare equal under yuv_planar_same_shape (which admits a crossfade
// operand per frame: pe_fx_convert_crossfade_batch / _batchv).
bool yuv_planar_same_shape(const YuvToRgbArgs &a, const YuvToRgbArgs &b) {
  return a.width == b.width && a.height == b.height && a.is_422 == b.is_422 && a.clamped == b.clamped && a.low_quality == b.low_quality &&
         a.quirks == b.quirks && a.conv.t == b.conv.t && a.lut16 == b.lut16 && a.src.rs_y == b.src.rs_y && a.src.rs_u == b.src.rs_u &&
         a.src.rs_v == b.src.rs_v && a.src.cw == b.src.cw && a.src.ch == b.src.ch && a.dst.rs == b.dst.rs && a.out.r == b.out.r &&
         a.out.g == b.out.g && a.out.b == b.out.b && a.out.a == b.out.a && a.out.psize == b.out.psize &&
         // crossfade operands: all frames of the run have one or none (each its own, or all the same), word aligned, same stride / factor
         (a.blend2 != nullptr) == (b.blend2 != nullptr) && a.blend2_rs == b.blend2_rs && a.blend_bf == b.blend_bf &&
         (!a.blend2 || (((((uintptr_t)a.blend2) | ((uintptr_t)b.blend2)) | (uint32_t)a.blend2_rs) & 3) == 0);
}
cudaError_t launch_yuv_planar_to_rgb_batch(const Launch &L, const YuvToRgbArgs *frames, int n) {
  cudaError_t e = march_attrs();
  if (e != cudaSuccess) return e;
  const YuvToRgbArgs &a = frames[0];
  const bool last_row_unsafe = a.src.rs_u < a.src.cw + 4 || a.src.rs_v < a.src.cw + 4;
  const int kmax = a.is_422 ? (a.height - 1 - (last_row_unsafe ? 1 : 0)) / 2 : a.src.ch - 1 - (last_row_unsafe ? 1 : 0);
  const int nstrips = (a.width + 127) / 128;
  for (int base = 0; base < n; base += 32) {
    YuvFrameList fl;
    fl.n = n - base < 32 ? n - base : 32;
    for (int i = 0; i < 32; i++) {
      const YuvToRgbArgs &f = frames[base + (i < fl.n ? i : 0)];
      fl.y[i] = f.src.y; fl.u[i] = f.src.u; fl.v[i] = f.src.v; fl.dst[i] = f.dst.p; fl.blend2[i] = f.blend2;
    }
    long long rows_per_warp = (long long)nstrips * a.height * fl.n / ((long long)L.sm_count * (Y2_NT / 32));
    int band_h = (int)(rows_per_warp < 8 ? 8 : rows_per_warp);
    if (band_h > a.height) band_h = a.height;
    launch_march(L, a, fl, kmax, band_h);
    PE_COUNT_LAUNCH(L);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_yuv_planar_to_rgb_fast(const Launch &L, const YuvToRgbArgs &a) {
  static PerDevice attr_set;
  if (!attr_set.cur()) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_yuv_planar_to_rgb_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_BYTES)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_yuv_planar_to_rgb_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_BYTES)) != cudaSuccess) return e;
    attr_set.cur() = 1;
  }
  // the fast paths read whole words behind chroma column cw: on the last chroma row that needs 4 bytes of row padding
  const bool last_row_unsafe = a.src.rs_u < a.src.cw + 4 || a.src.rs_v < a.src.cw + 4;
  const int k_fast_max = a.src.ch - 1 - (last_row_unsafe ? 1 : 0);
  // frames tall enough for every warp to march through a few dozen rows of a strip take k_yuv_march; small ones the job kernel.
  // Measured on a single 4K 4:2:2 clip + crossfade (27 rows per warp): march 35.5 us and 15.5 M warp instructions, job kernel
  // 34.6 us and 20.3 M -- at that size both are bounded by the per-launch overheads (table fill, first loads, tail), so the
  // marching kernel only takes over where its lower instruction count can show
  const int nstrips = (a.width + 127) / 128;
  const long long rows_per_warp = (long long)nstrips * a.height / ((long long)L.sm_count * (Y2_NT / 32));
  const bool blend_ok = !a.blend2 || ((((uintptr_t)a.blend2) | (uint32_t)a.blend2_rs) & 3) == 0;
  const char *mr = getenv("PE_YUV_MARCH_MIN_ROWS");  // tests force the marching kernel onto small frames with 0
  if (rows_per_warp >= (mr ? atoi(mr) : 48) && blend_ok && getenv("PE_YUV_NO_MARCH") == nullptr) {
    { cudaError_t e = march_attrs(); if (e != cudaSuccess) return e; }
    // 4:2:2: a fast step reads chroma rows 2k - 1 and 2k; the last one needs 4 bytes of padding behind it
    const int kmax = a.is_422 ? (a.height - 1 - (last_row_unsafe ? 1 : 0)) / 2 : k_fast_max;
    int band_h = (int)rows_per_warp;
    if (band_h < 8) band_h = 8;
    if (band_h > a.height) band_h = a.height;
    YuvFrameList fl;
    fl.n = 1; fl.y[0] = a.src.y; fl.u[0] = a.src.u; fl.v[0] = a.src.v; fl.dst[0] = a.dst.p; fl.blend2[0] = a.blend2;
    for (int i = 1; i < 32; i++) { fl.y[i] = fl.y[0]; fl.u[i] = fl.u[0]; fl.v[i] = fl.v[0]; fl.dst[i] = fl.dst[0]; fl.blend2[i] = fl.blend2[0]; }
    launch_march(L, a, fl, kmax, band_h);
    PE_COUNT_LAUNCH(L);
    return cudaGetLastError();
  }
  const long long work = (long long)(a.width >> 2) * (a.is_422 ? a.height : a.src.ch + 1);
  long long blocks = (work + Y2_NT - 1) / Y2_NT;
  if (blocks > L.sm_count) blocks = L.sm_count;
  if (a.quirks) k_yuv_planar_to_rgb_fast<true><<<(int)blocks, Y2_NT, S2_BYTES, L.stream>>>(a, k_fast_max);
  else k_yuv_planar_to_rgb_fast<false><<<(int)blocks, Y2_NT, S2_BYTES, L.stream>>>(a, k_fast_max);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

}  // namespace pe
