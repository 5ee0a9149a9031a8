// pe_kernels_fused3.cu -- the register-resident fused chain (the kernel bench.py's headline runs):
//     planar 4:2:0 fg -> RGBA  |  letter- / pillarbox without horizontal scaling (inner width == fg width, inner offset a
//     multiple of 4), vertical filter of <= 4 taps
//     |  scalar alpha-over (alpha = k / 256) a RGBA32 bg  |  optional 8-bit gamma LUT
// Same arithmetic, bit for bit, as k_fused2 / k_fused and as the unfused ops (tests/test_gpu_parity.py); everything outside
// this envelope runs k_fused2 (pe_kernels_fused2.cu).  What differs is where the data lives:
//
//   * NO shared-memory image tile and NO block barrier in the main loop.  A warp owns a strip of 128 columns (one lane =
//     4 adjacent columns = one 32-bit luma word = one 16-byte bg / out vector) and marches down the source rows one
//     reference row pair (colourspace.c:3440-3549) per step.  The converted pixels of the last steps live in REGISTERS,
//     packed vertically: W = [row 2k, row 2k-1, row 2k-2, row 2k-3] per column and channel, pushed two rows at a time by
//     the saturating pack (cvt.pack.sat) that the conversion needs anyway.  An output row is emitted as soon as its four
//     source rows are in the window: the <= 4-tap vertical filter is two DP2A per channel on W (or on a byte-permute of
//     W and the previous W), then letterbox / alpha-over / gamma / one 128-bit streaming store.
//   * the raw luma / chroma words of step k+1 and the bg vector of the next output row are loaded into registers while step k
//     is computed (ld.global.nc, L1 no-allocate): each byte of fg and bg crosses the memory system once, as whole sectors.
//   * every table is REPLICATED ACROSS BANKS in shared memory so that a lookup never conflicts, whatever the pixel values:
//       {R_Cr,G_Cr}[256][16]        2 x u32    64-bit loads, lane l reads bank pair l & 15
//       {G_Cb,B_Cb}[256][16]        2 x u32
//       gamma LUT  [256][32 lanes]  u32 = v * 0x010101 | 0xFF000000   } interleaved: one 256-byte entry per value, so that the
//       RGB_Y      [256][32 lanes]  u32        lane l reads bank l     } offset of a lookup is (byte << 8) | 4 * lane
//     (128 KB per SM, one 512-thread CTA per SM).  The chroma tables are indexed by m = third_round(n) (the (int)(n / 3. + .5)
//     of colourspace.c:3465); CLAMP16_240 / CLAMP0_255 are no-ops on these tables (flat outside the range; checked on the
//     host by fused3_tables_ok) and third_round is one multiply-high on the PACKED pair of sums (see idx_hi / idx_lo).
//   * chroma sums are computed two columns at a time in the 16-bit halves of a register (Q = 2 n + 3, <= 1533).
//   * work is split by COST, not by tiles: the (frame, band, strip, row) sequence is cut into one contiguous share per warp
//     (border rows are ~4x cheaper than inner rows), so all warps finish together; consecutive warps get neighbouring strips
//     of the same band.
// Rows that cannot take the fast step (row 0, the last row of an even frame, virtual rows beyond the frame, the last chroma
// row of a plane without padding) go through slow_step(): scalar code with the reference's edge rules, a few steps per strip.
#include <cstdlib>
#include <type_traits>
#include <vector>
#include <algorithm>
#include <cstdio>

#include "pe_device.cuh"
#include "pe_kernels.h"
#include "pe_tables.h"

namespace pe {

namespace {

#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

#ifndef PE_F3_NT
#define PE_F3_NT 512
#endif
constexpr int F3_NT = PE_F3_NT;         // threads per CTA (one CTA per SM).  Measured: 512 (128 regs) 46.6k fps, 640 (96 regs) 42.7k, 768 (80 regs) 36.8k
constexpr int F3_NW = F3_NT / 32;
constexpr int F3_MAXF = 32;               // frames per launch (their pointers travel as kernel parameters)
// Shared-memory layout, as ABSOLUTE addresses of the CTA's shared window (dynamic shared memory starts at 0x400 on sm_100: the
// first KB is the system's; the kernel traps if it does not).  Every table region starts at a multiple of its own size, so a
// lookup address is ONE LOP3 / PRMT -- (index bits) | (region base | the lane's bank offset) -- with no add and no base register:
//   0x08000  uint2 [256][16]  {R_Cr, G_Cr}                 entry stride 128: index bits 7..14
//   0x10000  u32   [256][64]  words 0..31 of an entry = the gamma LUT copies, words 32..63 = the RGB_Y copies
//                             (entry stride 256: the index is byte 1 of the address -- RGB_Y straight from the luma word with one
//                             PRMT, the LUT from the blended 16-bit value with one LOP3)
//   0x20000  uint2 [256][16]  {G_Cb, B_Cb}
//   0x28000  per warp: F3_RING bg rows of 512 bytes, filled by cp.async ahead of the emit (2 KB-aligned per warp: slot | base)
//   then     int4 per inner output row (+ 1): first source row, c3 | c2 << 16, c1 | c0 << 16, 0; then the TMA variant's mbarriers
// (measured with separate, 128-byte-stride RGB_Y / LUT tables: 49.0k vs 49.7k fps)
constexpr uint32_t A_DYN = 0x400u;        // where dynamic shared memory starts
constexpr uint32_t A_TV = 0x8000u;
constexpr uint32_t A_LUT = 0x10000u;
constexpr uint32_t A_TY = 0x10080u;
constexpr uint32_t A_TU = 0x20000u;
constexpr uint32_t A_RING = 0x28000u;
#ifndef PE_F3_RING
#define PE_F3_RING 4
#endif
constexpr int F3_RING = PE_F3_RING;      // power of two
constexpr uint32_t A_ROWS = A_RING + F3_NW * F3_RING * 512;
constexpr int F3_SMEM_MAX = 227 * 1024;   // opt-in limit of dynamic shared memory per CTA on sm_100
constexpr int F3_MAX_IH = ((int)A_DYN + F3_SMEM_MAX - (int)A_ROWS - 8 * F3_NW * F3_RING) / 16 - 1;
__host__ __device__ constexpr int f3_smem_bytes(int ih) { return (int)(A_ROWS - A_DYN) + 16 * (ih + 1) + 8 * F3_NW * F3_RING; }
#ifndef PE_F3_TMA_DEFAULT
#define PE_F3_TMA_DEFAULT 0
#endif

struct Fused3Frame {
  const uint8_t *y, *u, *v, *bg;
  uint8_t *out;
};

struct Fused3Params {
  Fused3Frame fr[F3_MAXF];
  int nframes;
  int fw, fh, cw, ch;                      // fg luma / chroma size
  int rs_y, rs_u, rs_v, rs_bg, rs_out;     // shared by all frames of the launch
  int oh, oy, ih;                          // outer height, first inner row, inner height
  int ow, ox;                              // outer width, first inner column (inner width == fw)
  int nstrips, band_h;
  long long nshares;                       // static shares (<= warps of the grid; the warps beyond it only take part in the dynamic tail)
  int k_fast_max;                          // steps 1 .. k_fast_max take the fast path
  int cost_b, cost_i;                      // cost of a border / an inner output row
  long long frame_cost, total_cost, static_cost, chunk_cost;
  unsigned int *sched;                     // [2] device counters, zero between launches
  uint32_t ka, kia;                        // blend weights of fg / bg, sum 256
  int spread;                              // share numbering, see the kernel
  uint32_t zero;                           // 0 (orders the loads of the marching loop, see step())
  const int4 *rows4;                       // [ih]: first source row, c3 | c2 << 16, c1 | c0 << 16, 0 (coefficients x 16 under C16)
  const int32_t *conv;                     // [14][256] (ConvTab order)
  const uint8_t *lut8;                     // optional
#ifdef PE_F3_TIMELINE
  unsigned long long *tl;                  // [grid][2 + F3_NW] globaltimer stamps: CTA start, tables filled, every warp's end (tools/f3_timeline.py)
#endif
};

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
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
// d = (c[15:0] << 16) | (sat_u8(a) << 8) | sat_u8(b)
__device__ __forceinline__ uint32_t pack_sat(int a, int b, uint32_t c) {
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: the PE_F3_TMA variant of the bg ring
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PE_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PE_MBAR_DONE;\n"
      "bra PE_MBAR_WAIT;\n"
      "PE_MBAR_DONE:\n"
      "}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ int4 lds128_ro(uint32_t a) {   // read-only after the kernel's barrier (filter rows)
  int4 v;
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
// chroma words are read by the lanes of two neighbouring strips (one sector of overlap on each side): plain read-only loads,
// so that L2 keeps the shared sectors until the neighbour has asked for them (the streaming hint marks them evict-first)
__device__ __forceinline__ uint32_t ld_keep_u32(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t *p) {
  uint32_t r;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// address of the lane's RGB_Y copy for luma byte `col` of word w: the byte becomes byte 1 of tyl = A_TY | 4 * lane (one PRMT)
__device__ __forceinline__ uint32_t yaddr(uint32_t w, int col, uint32_t tyl) { return __byte_perm(w, tyl, 0x7604u | ((uint32_t)col << 4)); }

// acc >> 12 of the vertical filter.  PE_F3_SHR12_IMAD: as a multiply-high on the FMA-heavy pipe instead of a shift on the ALU pipe
__device__ __forceinline__ uint32_t shr12(uint32_t x) {
#ifdef PE_F3_SHR12_IMAD
  uint32_t d;
  asm("mul.hi.u32 %0, %1, 1048576;" : "=r"(d) : "r"(x));
  return d;
#else
  return x >> 12;
#endif
}

constexpr uint32_t MSK = 0xFFFEFFFEu;   // clears bit 0 of both halves: 2 * (s >> 1) = s & ~1
constexpr uint32_t K3 = 0x00030003u;    // + 3 in both halves (Q = 2 n + 3)

// m = third_round(n) = (2 n + 3) / 6 for the n in the HIGH / LOW half of a packed Q = 2 n + 3 (each half <= 1533).
// 10923 / 65536 = 1 / 6 + 5e-6: the product is off by < 0.008 (plus < 0.004 from the low half leaking into the high one),
// and (2 n + 3) / 6 is never closer than 1 / 6 to an integer, so the floor is exact.
// The kernel wants the table offset 128 * m: with the multiplier scaled by 128 the product is floor(128 * m_real), whose bits
// 7.. are 128 * floor(m_real) (the bits below are masked off together with the OR of the lane's column offset).
// (tl = region base | 8 * (lane & 15): the OR with the lane's bank pair and the mask are one LOP3)
__device__ __forceinline__ uint32_t idx_hi(uint32_t q, uint32_t tl) { return (__umulhi(q, 10923u * 128u) & 0x7F80u) | tl; }
__device__ __forceinline__ uint32_t idx_lo(uint32_t q, uint32_t tl) { return (__umulhi(q << 16, 10923u * 128u) & 0x7F80u) | tl; }

// the three 'this / last / next' views of one chroma row for the lane's two chroma columns jc0, jc0 + 1, as 16-bit halves
struct RowC {
  uint32_t a, b, c;  // a = [c(jc0), c(jc0+1)], b = [c(jc0-1), c(jc0)], c = [c(jc0+1), c(jc0+2)]
};
__device__ __forceinline__ RowC unpack_cw4(uint32_t cw4) {   // cw4 bytes: columns jc0-1, jc0, jc0+1, jc0+2
  RowC r;
  r.a = __byte_perm(cw4, 0u, 0x4241u);
  r.b = __byte_perm(cw4, 0u, 0x4140u);
  r.c = __byte_perm(cw4, 0u, 0x4342u);
  return r;
}
__device__ __forceinline__ RowC unpack_row(uint32_t w0, uint32_t w1, uint32_t sel) {
  const uint32_t cw4 = __byte_perm(w0, w1, sel);  // bytes: columns jc0-1, jc0, jc0+1, jc0+2
  RowC r;
  r.a = __byte_perm(cw4, 0u, 0x4241u);
  r.b = __byte_perm(cw4, 0u, 0x4140u);
  r.c = __byte_perm(cw4, 0u, 0x4342u);
  return r;
}

// state carried from one row pair to the next: the sums of the previous chroma row
struct Carry {
  uint32_t DUr, MUr, DVr, MVr;   // right pixel (this + next): doubled / bit-0-cleared sums
  uint32_t DUl, MUl, DVl, MVl;   // left pixel (this + last), intended stencil (no quirks)
  uint32_t QUL, aV;              // quirks: Q of the left U (u2 = u1, colourspace.c:3461); 'this' V (:3544)
};

// raw words of one step, loaded one step ahead
struct Pre {
  uint32_t yA, yB, u0, u1, v0, v1, vf;
  uint32_t u2, u3, v2, v3, sd;   // 4:2:2 only: the chroma words of row 2k (u0 .. v1 are row 2k - 1's), the seed samples of the lane of column 0
};

struct Lane {
  // per-lane constants of the current segment
  const uint8_t *yp, *up0, *vp0, *vfp;
  uint32_t sel, selB;
  int x;
};

// chroma sample with the reference's edge rules: column -1 replicates column 0; column cw reads the byte behind the row
// (padding or the first sample of the next row, colourspace.c:3508) except on the last row of a plane without padding
__device__ __forceinline__ uint32_t chroma_at(const uint8_t *__restrict__ p, int stride, int r, int c, int cw, int ch) {
  if (c < 0) c = 0;
  if (c >= cw) c = (cw < stride || r + 1 < ch) ? cw : cw - 1;
  return p[(size_t)stride * r + c];
}

// ---- slow step, out of line (a few steps per strip; keeping it out of the marching loop keeps the loop's code small): rows
//      A = 2k-1, B = 2k of the lane's 4 columns with the reference's edge rules, as (sat(A) << 8) | sat(B) per column-channel.
//      k <= 0: row 0 alone (horizontal average only, colourspace.c:3421-3428); 2k-1 == fh-1: the last row of an even frame alone;
//      otherwise an interior pair whose last chroma row has no padding behind it (one-past-row read, :3508).
struct SlowRows {
  uint32_t ab[12];
};
template <bool QUIRKS>
__device__ __noinline__ SlowRows slow_rows(const uint8_t *py, const uint8_t *pu, const uint8_t *pv, uint32_t rs_y,
                                           uint32_t rs_u, uint32_t rs_v, int x0, int k, int fh, int cw, int ch, int lane) {
  const uint32_t lane4 = 4u * (uint32_t)lane, lane8 = 8u * (uint32_t)(lane & 15);
  int rA[12], rB[12];
  auto rgb = [&](uint32_t y, uint32_t mu, uint32_t mv, int &r, int &g, int &b) {   // y: byte value, mu / mv: table indices
    const int yy = (int)lds32(A_TY + (y * 256u + lane4));
    const uint2 tv = lds64(A_TV + (mv * 128u + lane8));
    const uint2 tu = lds64(A_TU + (mu * 128u + lane8));
    r = (yy + (int)tv.x) >> 16; g = (yy + (int)tu.x + (int)tv.y) >> 16; b = (yy + (int)tu.y) >> 16;
  };
  auto single = [&](int row, int cr) {
    const uint32_t yw = *reinterpret_cast<const uint32_t *>(py + (size_t)rs_y * row + x0);
#pragma unroll
    for (int col = 0; col < 4; col++) {
      const int jc = (x0 >> 1) + (col >> 1), jo = (col & 1) ? jc + 1 : jc - 1;
      const uint32_t mu = (chroma_at(pu, rs_u, cr, jc, cw, ch) + chroma_at(pu, rs_u, cr, jo, cw, ch)) >> 1;
      const uint32_t mv = (chroma_at(pv, rs_v, cr, jc, cw, ch) + chroma_at(pv, rs_v, cr, jo, cw, ch)) >> 1;
      rgb(byte_of(yw, col), mu, mv, rA[3 * col], rA[3 * col + 1], rA[3 * col + 2]);
      rB[3 * col] = rA[3 * col]; rB[3 * col + 1] = rA[3 * col + 1]; rB[3 * col + 2] = rA[3 * col + 2];
    }
  };
  if (k <= 0) {
    single(0, 0);
  } else if (2 * k <= fh - 1) {
    const int ca = k - 1, cbr = k;
    const uint32_t ya = *reinterpret_cast<const uint32_t *>(py + (size_t)rs_y * (2 * k - 1) + x0);
    const uint32_t yb = *reinterpret_cast<const uint32_t *>(py + (size_t)rs_y * (2 * k) + x0);
    const uint32_t vfirst = pv[(size_t)rs_v * cbr];
#pragma unroll
    for (int col = 0; col < 4; col++) {
      const int jc = (x0 >> 1) + (col >> 1), jo = (col & 1) ? jc + 1 : jc - 1;
      uint32_t u1 = chroma_at(pu, rs_u, ca, jc, cw, ch) + chroma_at(pu, rs_u, ca, jo, cw, ch);
      uint32_t u2 = chroma_at(pu, rs_u, cbr, jc, cw, ch) + chroma_at(pu, rs_u, cbr, jo, cw, ch);
      uint32_t v1 = chroma_at(pv, rs_v, ca, jc, cw, ch) + chroma_at(pv, rs_v, ca, jo, cw, ch);
      uint32_t v2 = chroma_at(pv, rs_v, cbr, jc, cw, ch) + chroma_at(pv, rs_v, cbr, jo, cw, ch);
      if (QUIRKS && !(col & 1)) {
        u2 = u1;
        if (jc > 0) v1 = chroma_at(pv, rs_v, ca, jc, cw, ch) + chroma_at(pv, rs_v, cbr, jo, cw, ch);
        v2 = chroma_at(pv, rs_v, cbr, jc, cw, ch) + vfirst;
      }
      const uint32_t mu3 = (uint32_t)third_round((int)(u1 + (u2 >> 1))), mu4 = (uint32_t)third_round((int)((u1 >> 1) + u2));
      const uint32_t mv3 = (uint32_t)third_round((int)(v1 + (v2 >> 1))), mv4 = (uint32_t)third_round((int)((v1 >> 1) + v2));
      rgb(byte_of(ya, col), mu3, mv3, rA[3 * col], rA[3 * col + 1], rA[3 * col + 2]);
      rgb(byte_of(yb, col), mu4, mv4, rB[3 * col], rB[3 * col + 1], rB[3 * col + 2]);
    }
  } else {
    single(fh - 1, ch - 1);
  }
  SlowRows out;
#pragma unroll
  for (int i = 0; i < 12; i++) out.ab[i] = pack_sat(rA[i], rB[i], 0u);
  return out;
}

// the 4:2:2 form (convert_yuv420p_to_rgb_frame with is_422, colourspace.c:3598-3642): every luma row has its own chroma row, the only
// averaging is horizontal.  QUIRKS: the seed slip (:3600) -- columns <= 0 of row i take column 0 of chroma row i >> 1.
template <bool QUIRKS>
__device__ __noinline__ SlowRows slow_rows422(const uint8_t *py, const uint8_t *pu, const uint8_t *pv, uint32_t rs_y, uint32_t rs_u,
                                              uint32_t rs_v, int x0, int k, int fh, int cw, int ch, int lane) {
  const uint32_t lane4 = 4u * (uint32_t)lane, lane8 = 8u * (uint32_t)(lane & 15);
  int rA[12], rB[12];
  auto single = [&](int row, int *out) {
    const int seed_row = QUIRKS ? row >> 1 : row;
    const uint32_t yw = *reinterpret_cast<const uint32_t *>(py + (size_t)rs_y * row + x0);
#pragma unroll
    for (int col = 0; col < 4; col++) {
      const int jc = (x0 >> 1) + (col >> 1), jo = (col & 1) ? jc + 1 : jc - 1;
      uint32_t ua = chroma_at(pu, rs_u, row, jc, cw, ch), ub = chroma_at(pu, rs_u, row, jo, cw, ch);
      uint32_t va = chroma_at(pv, rs_v, row, jc, cw, ch), vb = chroma_at(pv, rs_v, row, jo, cw, ch);
      if (x0 == 0 && seed_row != row) {
        const uint32_t su = pu[(size_t)rs_u * seed_row], sv = pv[(size_t)rs_v * seed_row];
        if (jc == 0) { ua = su; va = sv; }
        if (jo <= 0) { ub = su; vb = sv; }
      }
      const int yy = (int)lds32(A_TY + (byte_of(yw, col) * 256u + lane4));
      const uint2 tv = lds64(A_TV + (((va + vb) >> 1) * 128u + lane8));
      const uint2 tu = lds64(A_TU + (((ua + ub) >> 1) * 128u + lane8));
      out[3 * col] = (yy + (int)tv.x) >> 16; out[3 * col + 1] = (yy + (int)tu.x + (int)tv.y) >> 16; out[3 * col + 2] = (yy + (int)tu.y) >> 16;
    }
  };
  if (k <= 0) {
    single(0, rA);
#pragma unroll
    for (int i = 0; i < 12; i++) rB[i] = rA[i];
  } else if (2 * k <= fh - 1) {
    single(2 * k - 1, rA);
    single(2 * k, rB);
  } else {
    single(fh - 1, rA);
#pragma unroll
    for (int i = 0; i < 12; i++) rB[i] = rA[i];
  }
  SlowRows out;
#pragma unroll
  for (int i = 0; i < 12; i++) out.ab[i] = pack_sat(rA[i], rB[i], 0u);
  return out;
}

// C16: the filter coefficients arrive scaled by 16 (sum 65536; only banks without a 4096 tap): the filtered value is byte 2 of
// the accumulator (byte 3 is zero), so R | B pack with one PRMT and two of the three shifts per pixel disappear
// TMA: the bg ring is filled by ONE elected lane per row with a 512-byte cp.async.bulk (UBLKCP) that completes on the slot's
// mbarrier, instead of 32 per-lane 16-byte cp.async (LDGSTS) + commit / wait groups (full-width strips only: ow % 128 == 0, ox == 0)
// IS422: a planar 4:2:2 fg (a step = two luma rows with their own chroma rows; no carried sums)
template <bool QUIRKS, bool HAS_LUT, bool C16, bool TMA, bool IS422 = false>
__global__ void __launch_bounds__(F3_NT, 1) k_fused3(const __grid_constant__ Fused3Params P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef PE_F3_TIMELINE
  auto stamp = [&](int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.tl[(size_t)blockIdx.x * (2 + F3_NW) + slot] = t;
  };
  if (tid == 0) stamp(0);
#endif
  if ((uint32_t)__cvta_generic_to_shared(smem) != A_DYN) __trap();   // the absolute layout above assumes it
  uint8_t *const sm0 = smem - A_DYN;   // sm0 + absolute address = generic pointer (table fill only)

  // ---- replicated tables
  {
    // 16 consecutive lanes write the 16 x 16-byte chunks of one 256-byte {LUT | RGB_Y} entry, 8 lanes the chunks of a 128-byte
    // chroma entry: a warp's 128-bit store covers 512 contiguous bytes = 4 wavefronts, the minimum.  (A thread per entry writing
    // its 128 bytes alone is a 32-way bank conflict per store: 4.4 us per launch, profiles/r02c_k_fused3_single_frame_ncu.txt.)
    // every global load of the fill is issued before the first shared-memory store: the stores go through a generic pointer, and
    // the compiler keeps a load behind a store it cannot prove disjoint -- the fill was eight dependent L2 round trips, 7.4 us of a
    // single-frame launch (tools/f3_timeline.sh)
    static_assert(256 * 16 % F3_NT == 0 && 256 * 8 % F3_NT == 0, "table fill: whole rounds");
    constexpr int R1 = 256 * 16 / F3_NT, R2 = 256 * 8 / F3_NT;
    uint32_t e1[R1], rcr[R2], gcb[R2], gcr[R2], bcb[R2];
#pragma unroll
    for (int q = 0; q < R1; q++) {
      const int i = tid + q * F3_NT, m = i >> 4, j = i & 15;
      if (j < 8) e1[q] = HAS_LUT ? ((uint32_t)__ldg(P.lut8 + m) * 0x010101u | 0xFF000000u) : 0u;
      else e1[q] = (uint32_t)__ldg(P.conv + 9 * 256 + m);
    }
#pragma unroll
    for (int q = 0; q < R2; q++) {
      const int m = (tid + q * F3_NT) >> 3;
      rcr[q] = (uint32_t)__ldg(P.conv + 10 * 256 + m); gcb[q] = (uint32_t)__ldg(P.conv + 11 * 256 + m);
      gcr[q] = (uint32_t)__ldg(P.conv + 12 * 256 + m); bcb[q] = (uint32_t)__ldg(P.conv + 13 * 256 + m);
    }
#pragma unroll
    for (int q = 0; q < R1; q++) {
      const int i = tid + q * F3_NT, m = i >> 4, j = i & 15;
      reinterpret_cast<uint4 *>(sm0 + A_LUT + 256 * m)[j] = make_uint4(e1[q], e1[q], e1[q], e1[q]);
    }
#pragma unroll
    for (int q = 0; q < R2; q++) {
      const int i = tid + q * F3_NT, m = i >> 3, j = i & 7;
      reinterpret_cast<uint4 *>(sm0 + A_TV + 128 * m)[j] = make_uint4(rcr[q], gcr[q], rcr[q], gcr[q]);
      reinterpret_cast<uint4 *>(sm0 + A_TU + 128 * m)[j] = make_uint4(gcb[q], bcb[q], gcb[q], bcb[q]);
    }
    int4 *sr = reinterpret_cast<int4 *>(sm0 + A_ROWS);
    for (int i = tid; i <= P.ih; i += F3_NT) sr[i] = P.rows4[min(i, P.ih - 1)];   // (one entry past the end: the emit reads one row ahead)
    if (TMA) {  // one mbarrier per (warp, ring slot), behind the filter rows
      if (tid < F3_NW * F3_RING) mbar_init(A_ROWS + 16u * (uint32_t)(P.ih + 1) + 8u * (uint32_t)tid, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  __syncthreads();  // the only barrier of the kernel
#ifdef PE_F3_TIMELINE
  if (tid == 0) stamp(1);
#endif

  Lane L;
  // region base | the lane's bank offset: OR-ed into every lookup address
  const uint32_t tyl = A_TY | (4u * (uint32_t)lane), lutl = A_LUT | (4u * (uint32_t)lane);
  const uint32_t tul = A_TU | (8u * (uint32_t)(lane & 15)), tvl = A_TV | (8u * (uint32_t)(lane & 15));

  const int fw = P.fw, fh = P.fh, cw = P.cw, ch = P.ch;
  const int oy = P.oy, ih = P.ih, oh = P.oh;
  const uint32_t ka = P.ka, kia = P.kia;
  const uint32_t kk = kia | (ka << 16);   // both blend weights as the 16-bit halves of a DP2A operand
  const uint32_t zero = P.zero;

  // yuv2rgb_int (colourspace.c:2345-2356) through the replicated tables; results UNSATURATED (saturated by pack_sat)
  // (ya, ou, ov: complete shared-memory addresses, see yaddr / idx_hi / idx_lo)
  auto rgb = [&](uint32_t ya, uint32_t ou, uint32_t ov, int &r, int &g, int &b) {
    const int yy = (int)lds32(ya);
    const uint2 tv = lds64(ov);
    const uint2 tu = lds64(ou);
    r = (yy + (int)tv.x) >> 16;
#ifdef PE_F3_SAR16_G_IMAD
    asm("mul.hi.s32 %0, %1, 65536;" : "=r"(g) : "r"(yy + (int)tu.x + (int)tv.y));   // one of the three shifts on the FMA-heavy pipe
#else
    g = (yy + (int)tu.x + (int)tv.y) >> 16;
#endif
    b = (yy + (int)tu.y) >> 16;
  };

  // gamma LUT of the three blended channels (16-bit values, the byte to look up is byte 1) -> RGBA word
  auto lut_pack = [&](uint32_t vr, uint32_t vg, uint32_t vb) -> uint32_t {
    if (HAS_LUT) {
      const uint32_t e0 = lds32((vr & 0xFF00u) | lutl);
      const uint32_t e1 = lds32((vg & 0xFF00u) | lutl);
      const uint32_t e2 = lds32((vb & 0xFF00u) | lutl);
      return __byte_perm(__byte_perm(e0, e1, 0x0040u), e2, 0x7410u);
    }
    return __byte_perm(__byte_perm(vr, vg, 0x0051u), vb, 0x0510u) | 0xFF000000u;
  };
  // alpha-over (integer form of compositor.c:120 for alpha = ka / 256) + gamma LUT of one pixel.
  // C16: the filtered channel is byte 2 of its accumulator (byte 3 is zero): one PRMT puts it next to the bg byte and ONE DP2A
  // forms bg * kia + f * ka (no extraction of the filtered bytes, no 16-bit-pair multiplies)
  auto blend_acc = [&](uint32_t bg, uint32_t a0, uint32_t a1, uint32_t a2) -> uint32_t {
    const uint32_t vr = dp2a_hi(kk, __byte_perm(bg, a0, 0x6000u), 0u);
    const uint32_t vg = dp2a_hi(kk, __byte_perm(bg, a1, 0x6100u), 0u);
    const uint32_t vb = dp2a_hi(kk, __byte_perm(bg, a2, 0x6200u), 0u);
    return lut_pack(vr, vg, vb);
  };
  auto blend = [&](uint32_t bg, uint32_t frb, uint32_t fg_) -> uint32_t {   // frb = filtered R | B << 16
    const uint32_t rb = (bg & 0x00FF00FFu) * kia + frb * ka;   // R | B in the 16-bit halves
    const uint32_t gg = __byte_perm(bg, 0u, 0x4441u) * kia + fg_ * ka;
    return lut_pack(rb, gg, rb >> 16);
  };
  auto blend_border = [&](uint32_t bg) -> uint32_t {  // letterbox border: fg = black (blank_pixel, colourspace.c:11169)
    const uint32_t rb = (bg & 0x00FF00FFu) * kia;
    const uint32_t gg = __byte_perm(bg, 0u, 0x4441u) * kia;
    return lut_pack(rb, gg, rb >> 16);
  };

  // ---- the warp's share of the cost sequence (frame, band, strip, row)
  const int cb = P.cost_b, ci = P.cost_i, Hb = P.band_h, nstrips = P.nstrips;
  auto cost_upto = [&](int r) -> int {  // cost of the rows [0, r) of one strip
    const int top = min(r, oy), mid = min(max(r - oy, 0), ih), bot = max(r - oy - ih, 0);
    return cb * (top + bot) + ci * mid;
  };
  auto row_at = [&](int t) -> int {  // smallest r with cost_upto(r) >= t
    if (t <= cb * oy) return (t + cb - 1) / cb;
    t -= cb * oy;
    if (t <= ci * ih) return oy + (t + ci - 1) / ci;
    t -= ci * ih;
    return oy + ih + (t + cb - 1) / cb;
  };
  // static part: [0, static_cost) in equal shares; the rest is handed out in small chunks as warps run dry (P.sched: the
  // chunk counter and the count of finished warps; the last warp to finish zeroes both for the next launch)
  // share number of this warp: consecutive shares to the warps of one CTA (neighbouring strips of a band: the chroma sectors two
  // strips share are fetched once per SM), or (P.spread) shares gridDim.x apart, so that the warps of an SM are in different
  // phases of a frame -- border rows are pure memory traffic, inner rows mostly arithmetic
  const long long nwarps = (long long)gridDim.x * F3_NW;
  const long long gw = P.spread ? (long long)warp * gridDim.x + blockIdx.x : (long long)blockIdx.x * F3_NW + warp;
  uint32_t bg_push_seq = 0u, bg_pop_seq = 0u;  // TMA variant: rows pushed into / popped from the warp's bg ring so far
  long long pos = P.static_cost * min(gw, P.nshares) / P.nshares;
  long long pos_end = P.static_cost * min(gw + 1, P.nshares) / P.nshares;

#ifdef PE_F3_TIMELINE
  int tl_nseg = 0, tl_rows = 0, tl_inner = 0;
#endif
  for (;;) {
  while (pos < pos_end) {
    // ---- locate the unit (frame, band, strip) that holds `pos` and the rows of it that belong to this warp
    const int f = (int)(pos / P.frame_cost);
    int rem = (int)(pos - (long long)f * P.frame_cost);  // (a frame's cost fits 31 bits: checked by the launcher)
    // the band that holds `rem`, without walking the bands (a single-frame launch has ~80 of them per frame and every warp looked its
    // two or three units up by linear search: a quarter of that launch's instructions): cost_upto is strictly increasing, so the row
    // rq with cost_upto(rq) <= rem / nstrips < cost_upto(rq + 1) lies in the band
    int base, bc;
    {
      const int q = rem / nstrips;
      int rq;
      if (q < cb * oy) rq = q / cb;
      else if (q < cb * oy + ci * ih) rq = oy + (q - cb * oy) / ci;
      else rq = oy + ih + (q - cb * oy - ci * ih) / cb;
      const int b = rq / Hb;
      const int r0 = b * Hb, r1 = min(oh, r0 + Hb);
      base = cost_upto(r0);
      bc = cost_upto(r1) - base;
      rem -= base * nstrips;
    }
    const int s = rem / bc;
    const int within = rem - s * bc;
    const long long unit0 = pos - within;
    const int hi = (int)min((long long)bc, pos_end - unit0);
    const int ra = row_at(base + within), rb = row_at(base + hi);
    pos = unit0 + bc;

    const int x = 128 * s + 4 * lane;  // the lane's first output column
    if (x >= P.ow || ra >= rb) continue;
#ifdef PE_F3_TIMELINE
    tl_nseg++; tl_rows += rb - ra; tl_inner += max(0, min(rb, oy + ih) - max(ra, oy));
#endif
    const int xs = x - P.ox;           // ... and its first source column (pillarbox: the inner rectangle starts at ox)
    const Fused3Frame &F = P.fr[f];
    const uint8_t *bgp = F.bg + 4 * (size_t)x;
    uint8_t *outp = F.out + 4 * (size_t)x;
    const uint32_t rs_bg = (uint32_t)P.rs_bg, rs_out = (uint32_t)P.rs_out;

    // ---- border rows above / below the inner rectangle: bg * (1 - alpha) -> gamma
    auto border_rows = [&](int r0, int r1) {
      if (r0 >= r1) return;
      // four rows per iteration, the next four in flight meanwhile (the loads of a row past r1 - 1 re-read row r1 - 1)
      uint4 w[4], wn[4];
#pragma unroll
      for (int j = 0; j < 4; j++) w[j] = ld_stream_u4(bgp + (size_t)rs_bg * (uint32_t)min(r0 + j, r1 - 1));
      for (int r = r0; r < r1; r += 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) wn[j] = ld_stream_u4(bgp + (size_t)rs_bg * (uint32_t)min(r + 4 + j, r1 - 1));
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (r + j < r1) {
            uint4 o;
            o.x = blend_border(w[j].x); o.y = blend_border(w[j].y); o.z = blend_border(w[j].z); o.w = blend_border(w[j].w);
            st_stream_u4(outp + (size_t)rs_out * (uint32_t)(r + j), o);
          }
          w[j] = wn[j];
        }
      }
    };
    if (xs < 0 || xs >= fw) {  // a lane left / right of the inner rectangle: border on every row
      border_rows(ra, rb);
      continue;
    }
    border_rows(ra, min(rb, oy));

    const int ia = max(ra, oy) - oy, ib = min(rb, oy + ih) - oy;  // inner rows [ia, ib)
    if (ia < ib) {
      // ---- per-lane constants
      const int jc0 = xs >> 1;                // first chroma column of the lane
      const int o = jc0 - 1;                  // byte offset of chroma column jc0 - 1
      const int off0 = xs == 0 ? 0 : (o & ~3);  // the second word is the next one (the lane of source column 0 only uses the first)
      L.x = xs;
      L.yp = F.y + xs;
      L.up0 = F.u + off0;
      L.vp0 = F.v + off0;
      L.vfp = F.v;
      L.sel = xs == 0 ? 0x2100u : ((o & 3) == 3 ? 0x6543u : 0x4321u);
      L.selB = xs == 0 ? 0x3254u : 0x3210u;
      const uint32_t rs_y = (uint32_t)P.rs_y, rs_u = (uint32_t)P.rs_u, rs_v = (uint32_t)P.rs_v;
      const int k_fast_max = P.k_fast_max;

      // raw words of a fast step: luma rows 2k-1, 2k, chroma row k.  The step loads the words of step k + 1 through RUNNING
      // pointers (yq, uq, vq, vfq: advanced once per step, whatever kind of step it is) -- per step 4 64-bit additions instead of
      // 7 address computations from the row number (profiles/r02k: a tenth of the step's instructions were address arithmetic)
      auto load_at = [&](const uint8_t *yr, const uint8_t *ur, const uint8_t *vr, const uint8_t *vfr, const uint8_t *ufr, Pre &p) {
#ifdef PE_F3_LUMA_NC
        p.yA = ld_keep_u32(yr);
        p.yB = ld_keep_u32(yr + rs_y);
#else
        p.yA = ld_stream_u32(yr);
        p.yB = ld_stream_u32(yr + rs_y);
#endif
        p.u0 = ld_keep_u32(ur); p.u1 = ld_keep_u32(ur + 4);
        p.v0 = ld_keep_u32(vr); p.v1 = ld_keep_u32(vr + 4);
        if (IS422) {   // ur / vr: chroma row 2k - 1; ufr / vfr: column 0 of chroma row k - 1 (the seed slip, colourspace.c:3600: columns
                       // <= 0 of rows 2k - 1 / 2k take column 0 of chroma rows k - 1 / k; read by the lane of column 0 only)
          p.u2 = ld_keep_u32(ur + rs_u); p.u3 = ld_keep_u32(ur + rs_u + 4);
          p.v2 = ld_keep_u32(vr + rs_v); p.v3 = ld_keep_u32(vr + rs_v + 4);
          if (QUIRKS && L.x == 0) p.sd = ldg_u8(ufr) | (ldg_u8(vfr) << 8) | (ldg_u8(ufr + rs_u) << 16) | (ldg_u8(vfr + rs_v) << 24);
        } else {
          p.vf = ldg_u8(vfr);
        }
      };
      auto init_carry = [&](int r, Carry &c) {  // sums of chroma row r (0 <= r <= ch - 2), as a fast step leaves them
        const uint32_t uo = rs_u * (uint32_t)r, vo = rs_v * (uint32_t)r;
        const RowC U = unpack_row(ld_stream_u32(L.up0 + uo), ld_stream_u32(L.up0 + uo + 4), L.sel);
        const RowC V = unpack_row(ld_stream_u32(L.vp0 + vo), ld_stream_u32(L.vp0 + vo + 4), L.sel);
        const uint32_t RU = U.a + U.c, RV = V.a + V.c, LU = U.a + U.b, LV = V.a + V.b;
        c.DUr = RU * 2u; c.MUr = RU & MSK; c.DVr = RV * 2u; c.MVr = RV & MSK;
        c.DUl = LU * 2u; c.MUl = LU & MSK; c.DVl = LV * 2u; c.MVl = LV & MSK;
        c.QUL = LU * 2u + (LU & MSK) + K3;
        c.aV = V.a;
      };

      // ---- one step: rows A = 2k-1, B = 2k of the lane's 4 columns -> Wc = [B, A, Wp.byte0, Wp.byte1]
      int iy = ia;
      int4 ri = lds128_ro(A_ROWS + 16u * (uint32_t)iy);
      int k = ((ri.x + 4) >> 1) - 2;
      bool carry_ok = false;
      uint32_t order_mask = 0u;   // 0 for the segment's first step: its words were only just requested, the loads of step k + 1 go out with them
      Carry C;
      Pre preA, preB;
      C.DUr = C.MUr = C.DVr = C.MVr = C.DUl = C.MUl = C.DVl = C.MVl = C.QUL = C.aV = 0u;
      preA.yA = preA.yB = preA.u0 = preA.u1 = preA.v0 = preA.v1 = preA.vf = preA.u2 = preA.u3 = preA.v2 = preA.v3 = preA.sd = 0u;
      preB = preA;
      // rows of step k (k may be <= 0 here: the pointers are only dereferenced for fast steps)
      const uint8_t *yq = L.yp + (long long)rs_y * (2 * k - 1);
      // 4:2:0: chroma row k, its first V sample.  4:2:2: chroma row 2k - 1, column 0 of chroma row k - 1 of V and U (the seed samples)
      const uint8_t *uq = L.up0 + (long long)rs_u * (IS422 ? 2 * k - 1 : k);
      const uint8_t *vq = L.vp0 + (long long)rs_v * (IS422 ? 2 * k - 1 : k);
      const uint8_t *vfq = L.vfp + (long long)rs_v * (IS422 ? k - 1 : k);
      const uint8_t *ufq = F.u + (long long)rs_u * (k - 1);
      if (k >= 1 && k <= k_fast_max) {
        load_at(yq, uq, vq, vfq, ufq, preA);
        if (!IS422) {
          init_carry(k - 1, C);
          carry_ok = true;
        }
      }
      // bg rows travel through the warp's cp.async ring, F3_RING - 1 rows ahead of the emit: no registers are tied up and the
      // emit never waits on a load it has just issued.  A lane only ever reads the 16 bytes it copied itself.
      const uint32_t ring_w = A_RING + (uint32_t)warp * (F3_RING * 512);
      const uint32_t ring = ring_w + 16u * (uint32_t)lane;
      constexpr uint32_t RMASK = (uint32_t)(F3_RING - 1) * 512u;   // the slot bits of a ring address (the warp's ring is aligned to its size)
      // TMA variant: rows enter the ring in the order they are emitted; bg_seq counts the rows this warp has pushed / popped since the
      // kernel started (slot = seq % F3_RING, the slot's mbarrier phase = seq / F3_RING); every pushed row is popped before the
      // segment ends
      const uint32_t mbar_w = A_ROWS + 16u * (uint32_t)(ih + 1) + 8u * (uint32_t)(warp * F3_RING);
      const uint8_t *bg_strip = F.bg + 512u * (size_t)s + (size_t)rs_bg * (uint32_t)oy;   // row 0 of the inner rectangle, this strip
      auto bg_push = [&](int row) {   // all lanes call it; lane 0 issues
        if (lane == 0) {
          const uint32_t slot = bg_push_seq & (F3_RING - 1);
          mbar_expect_tx(mbar_w + 8u * slot, 512u);
          bulk_g2s(ring_w + 512u * slot, bg_strip + (size_t)rs_bg * (uint32_t)row, 512u, mbar_w + 8u * slot);
        }
        bg_push_seq++;
      };
      // running state of the emit: the bg row the ring is refilled with when row iy leaves, the output row, the ring slot of row
      // iy and of row iy - 1 (== the slot of row iy + F3_RING - 1), the filter row of iy + 1
      const uint8_t *bg_fill = bgp + (size_t)rs_bg * (uint32_t)(oy + iy + F3_RING - 1);
      uint8_t *out_row = outp + (size_t)rs_out * (uint32_t)(oy + iy);
      uint32_t rd = ring | (((uint32_t)iy * 512u) & RMASK);
      uint32_t rows_sa = A_ROWS + 16u * (uint32_t)(iy + 1);
      if (TMA) {
        __syncwarp();
        for (int j = 0; j < F3_RING - 1 && iy + j < ib; j++) bg_push(iy + j);
      } else {
#pragma unroll
        for (int j = 0; j < F3_RING - 1; j++) {
          if (iy + j < ib) cp_async16(ring | (((uint32_t)(iy + j) * 512u) & RMASK), bgp + (size_t)rs_bg * (uint32_t)(oy + iy + j));
          cp_async_commit();
        }
      }

      // pre: the words of step k (loaded one step ago); nxt: where the words of step k + 1 go.  The two buffers swap roles from
      // step to step (the marching loop is unrolled by two), so nothing is copied.
      // Order matters: the step first WAITS for its own words, then issues the loads of step k + 1 -- those have the whole step and
      // the emit behind it to land.  Issued the other way round, the first use of `pre` sits right behind the new loads and ptxas
      // waits on a scoreboard they share (long-scoreboard stalls per issue 0.41 -> 1.90, 50.3k -> 47.9k fps: profiles/r02l).
      // Nothing in CUDA C orders independent loads behind a use, so the order is a data dependency: the row pointers advance by
      // stride + dep, dep = (words of this step) & P.zero -- a kernel parameter that is 0, which the compiler cannot know.
      auto step = [&](int k, uint32_t(&Wc)[12], const uint32_t(&Wp)[12], const Pre &pre, Pre &nxt) {
        int rA[12], rB[12];
        const bool fast = k >= 1 && k <= k_fast_max;
        const bool next_fast = k + 1 >= 1 && k + 1 <= k_fast_max;
        const uint32_t dep = (IS422 ? ((pre.u0 | pre.v0 | pre.yA) | pre.yB | pre.u2 | pre.v2) : ((pre.u0 | pre.v0 | pre.yA) | pre.yB | pre.vf)) & zero & order_mask;
        order_mask = 0xFFFFFFFFu;
        yq += 2 * (size_t)rs_y + dep; uq += (IS422 ? 2u : 1u) * rs_u + dep; vq += (IS422 ? 2u : 1u) * rs_v + dep; vfq += rs_v + dep;
        if (IS422) ufq += rs_u;
        if (next_fast) load_at(yq, uq, vq, vfq, ufq, nxt);
        if (fast && IS422) {
          // the lane of column 0 under the seed slip stays on the fast path: sample 0 of its chroma words is replaced by the seed sample,
          // which column -1 then replicates through the lane's PRMT selector like any frame edge
          uint32_t u0 = pre.u0, v0 = pre.v0, u2 = pre.u2, v2 = pre.v2;
          if (QUIRKS && L.x == 0) {
            u0 = __byte_perm(u0, pre.sd, 0x3214); v0 = __byte_perm(v0, pre.sd, 0x3215);
            u2 = __byte_perm(u2, pre.sd, 0x3216); v2 = __byte_perm(v2, pre.sd, 0x3217);
          }
          const RowC UA = unpack_row(u0, pre.u1, L.sel), VA = unpack_row(v0, pre.v1, L.sel);
          const RowC UB = unpack_row(u2, pre.u3, L.sel), VB = unpack_row(v2, pre.v3, L.sel);
          const uint32_t LUA = UA.a + UA.b, RUA = UA.a + UA.c, LVA = VA.a + VA.b, RVA = VA.a + VA.c;
          const uint32_t LUB = UB.a + UB.b, RUB = UB.a + UB.c, LVB = VB.a + VB.b, RVB = VB.a + VB.c;
#pragma unroll
          for (int col = 0; col < 4; col++) {
            const bool hi_half = col >> 1, right = col & 1;
            const uint32_t sua = right ? RUA : LUA, sva = right ? RVA : LVA, sub = right ? RUB : LUB, svb = right ? RVB : LVB;
            // table offset 128 * (s >> 1) = (s & ~1) << 6, OR-ed with the region base and the lane's bank pair in the same LOP3
            auto off = [&](uint32_t sm, uint32_t tl) -> uint32_t { return ((hi_half ? (sm >> 10) : (sm << 6)) & 0x7F80u) | tl; };
            rgb(yaddr(pre.yA, col, tyl), off(sua, tul), off(sva, tvl), rA[3 * col], rA[3 * col + 1], rA[3 * col + 2]);
            rgb(yaddr(pre.yB, col, tyl), off(sub, tul), off(svb, tvl), rB[3 * col], rB[3 * col + 1], rB[3 * col + 2]);
          }
#pragma unroll
          for (int i = 0; i < 12; i++) Wc[i] = pack_sat(rA[i], rB[i], Wp[i]);
        } else if (fast) {
          const RowC U = unpack_row(pre.u0, pre.u1, L.sel), V = unpack_row(pre.v0, pre.v1, L.sel);
          // right pixel of both chroma columns: this + next
          const uint32_t RU = U.a + U.c, RV = V.a + V.c;
          const uint32_t DUn = RU * 2u, MUn = RU & MSK, DVn = RV * 2u, MVn = RV & MSK;
          const uint32_t QUR_up = C.DUr + MUn + K3, QUR_lo = C.MUr + DUn + K3;
          const uint32_t QVR_up = C.DVr + MVn + K3, QVR_lo = C.MVr + DVn + K3;
          C.DUr = DUn; C.MUr = MUn; C.DVr = DVn; C.MVr = MVn;
          // left pixel: this + last
          uint32_t QUL_up, QUL_lo, QVL_up, QVL_lo;
          const uint32_t LU = U.a + U.b;
          if (QUIRKS) {
            QUL_up = QUL_lo = C.QUL;                        // u2 = this_u1 + last_u1 (colourspace.c:3461)
            C.QUL = LU * 2u + (LU & MSK) + K3;
            const uint32_t bq = __byte_perm(V.b, C.aV, L.selB);   // last_v1 = this_v2 (:3544), except at column 0
            const uint32_t v1 = C.aV + bq;
            const uint32_t v2 = V.a + pre.vf * 0x10001u;    // last_v2 is never advanced: the row's first V sample
            QVL_up = v1 * 2u + (v2 & MSK) + K3;
            QVL_lo = (v1 & MSK) + v2 * 2u + K3;
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
            const uint32_t mu_up = hi_half ? idx_hi(qu_up, tul) : idx_lo(qu_up, tul), mv_up = hi_half ? idx_hi(qv_up, tvl) : idx_lo(qv_up, tvl);
            const uint32_t mv_lo = hi_half ? idx_hi(qv_lo, tvl) : idx_lo(qv_lo, tvl);
            const uint32_t mu_lo = (QUIRKS && !right) ? mu_up : (hi_half ? idx_hi(qu_lo, tul) : idx_lo(qu_lo, tul));
            rgb(yaddr(pre.yA, col, tyl), mu_up, mv_up, rA[3 * col], rA[3 * col + 1], rA[3 * col + 2]);
            rgb(yaddr(pre.yB, col, tyl), mu_lo, mv_lo, rB[3 * col], rB[3 * col + 1], rB[3 * col + 2]);
          }
#pragma unroll
          for (int i = 0; i < 12; i++) Wc[i] = pack_sat(rA[i], rB[i], Wp[i]);
        } else {
          // ---- slow step (frame edges): out of line; rows beyond the frame replicate the last row (the filter clamps its
          //      source indices)
          if (k <= 0 || 2 * k - 1 <= fh - 1) {
            const SlowRows sr = IS422 ? slow_rows422<QUIRKS>(F.y, F.u, F.v, rs_y, rs_u, rs_v, L.x, k, fh, cw, ch, lane)
                                      : slow_rows<QUIRKS>(F.y, F.u, F.v, rs_y, rs_u, rs_v, L.x, k, fh, cw, ch, lane);
#pragma unroll
            for (int i = 0; i < 12; i++) Wc[i] = __byte_perm(sr.ab[i], Wp[i], 0x5410u);
          } else {
#pragma unroll
            for (int i = 0; i < 12; i++) Wc[i] = __byte_perm(Wp[i], 0u, 0x1000u);   // [b0, b0, b0, b1]: the newest row twice more
          }
          carry_ok = false;
        }
        if (!IS422 && next_fast && !carry_ok) {
          init_carry(k, C);
          carry_ok = true;
        }
      };

      // ---- emit every output row whose window [first, first + 3] is complete after step k
      auto emit = [&](int k, const uint32_t(&Wc)[12], const uint32_t(&Wp)[12]) {
        while (iy < ib && ri.x + 3 <= 2 * k) {
          const int j0 = 2 * k - ri.x - 3;  // 0: the window is Wc; 1: one row older
          const uint32_t CA = (uint32_t)ri.y, CB = (uint32_t)ri.z;
          uint4 bgw;
          if (TMA) {
            __syncwarp();  // every lane has read the slot that is refilled now (row iy - 1's)
            if (iy + F3_RING - 1 < ib) bg_push(iy + F3_RING - 1);
            const uint32_t slot = bg_pop_seq & (F3_RING - 1);
            mbar_wait(mbar_w + 8u * slot, (bg_pop_seq / F3_RING) & 1u);
            bgw = lds128(ring + 512u * slot);
            bg_pop_seq++;
          } else {
            if (iy + F3_RING - 1 < ib) cp_async16(ring | ((rd - 512u) & RMASK), bg_fill);   // the slot row iy - 1 has left
            cp_async_commit();
            cp_async_wait<F3_RING - 1>();  // all but the newest F3_RING - 1 groups have landed: row iy is in its slot
            bgw = lds128(rd);
          }
          const int4 rin = lds128_ro(rows_sa);
          const uint32_t bgv[4] = {bgw.x, bgw.y, bgw.z, bgw.w};
          uint32_t ov[4];
          if (j0 == 0) {
#pragma unroll
            for (int col = 0; col < 4; col++) {
              uint32_t acc[3];
#pragma unroll
              for (int c = 0; c < 3; c++) {
                const uint32_t win = Wc[3 * col + c];
                acc[c] = dp2a_hi(CB, win, dp2a_lo(CA, win, C16 ? 32768u : 2048u));
              }
              ov[col] = C16 ? blend_acc(bgv[col], acc[0], acc[1], acc[2])
                            : blend(bgv[col], __byte_perm(shr12(acc[0]), shr12(acc[2]), 0x5410u), shr12(acc[1]));
            }
          } else {
#pragma unroll
            for (int col = 0; col < 4; col++) {
              uint32_t acc[3];
#pragma unroll
              for (int c = 0; c < 3; c++) {
                const uint32_t win = __byte_perm(Wc[3 * col + c], Wp[3 * col + c], 0x6321u);
                acc[c] = dp2a_hi(CB, win, dp2a_lo(CA, win, C16 ? 32768u : 2048u));
              }
              ov[col] = C16 ? blend_acc(bgv[col], acc[0], acc[1], acc[2])
                            : blend(bgv[col], __byte_perm(shr12(acc[0]), shr12(acc[2]), 0x5410u), shr12(acc[1]));
            }
          }
          st_stream_u4(out_row, make_uint4(ov[0], ov[1], ov[2], ov[3]));
          bg_fill += rs_bg; out_row += rs_out;
          rd = ring | ((rd + 512u) & RMASK);
          rows_sa += 16u;
          ri = rin;
          iy++;
        }
      };

      uint32_t W0[12], W1[12];
#pragma unroll
      for (int i = 0; i < 12; i++) W0[i] = W1[i] = 0u;
      for (;;) {
        step(k, W0, W1, preA, preB);
        emit(k, W0, W1);
        if (iy >= ib) break;
        k++;
        step(k, W1, W0, preB, preA);
        emit(k, W1, W0);
        if (iy >= ib) break;
        k++;
      }
      if (!TMA) cp_async_wait<0>();  // the ring is reused by the warp's next segment
    }
    border_rows(max(ra, oy + ih), rb);
  }
    // ---- next chunk of the dynamic tail
    __syncwarp();
    long long t = 0;
    if (lane == 0) t = (long long)atomicAdd(P.sched, 1u);
    t = __shfl_sync(0xFFFFFFFFu, t, 0);
    pos = P.static_cost + t * P.chunk_cost;
    if (pos >= P.total_cost) break;
    pos_end = min(pos + P.chunk_cost, P.total_cost);
  }
#ifdef PE_F3_TIMELINE
  if (lane == 0) {
    stamp(2 + warp);
    P.tl[(size_t)gridDim.x * (2 + F3_NW) + (size_t)blockIdx.x * F3_NW + warp] =
        (unsigned long long)tl_nseg | ((unsigned long long)tl_rows << 16) | ((unsigned long long)tl_inner << 40);
  }
#endif
  if (lane == 0) {
    const unsigned int done = atomicAdd(P.sched + 1, 1u);
    if (done == (unsigned int)nwarps - 1u) {  // every warp has drawn its last (out of range) chunk: reset for the next launch
      P.sched[0] = 0u;
      P.sched[1] = 0u;
    }
  }
}

}  // namespace

// CLAMP16_240 / CLAMP0_255 before the chroma lookups (colourspace.c:3465-3469) must be no-ops on the tables: flat below
// min_uv and above max_uv.  True for all four variants the reference builds; checked because the kernel relies on it.
bool fused3_tables_ok(const ConvTables &t) {
  const int tabs[4] = {R_CR, G_CB, G_CR, B_CB};
  for (int i = 0; i < 4; i++)
    for (int m = 0; m < 256; m++) {
      const int c = m < t.min_uv ? t.min_uv : (m > t.max_uv ? t.max_uv : m);
      if (t.t[tabs[i]][m] != t.t[tabs[i]][c]) return false;
    }
  return true;
}

// Can the register-resident kernel take this batch?  (everything else: k_fused2 / k_fused)
bool fused3_supported(const FusedArgs *a, int n, int fy_taps) {
  if (n <= 0 || fy_taps > 4) return false;
  const FusedArgs &f0 = a[0];
  if (f0.low_quality) return false;
  if (f0.is_422 && getenv("PE_F3_NO_422")) return false;   // measurement switch: 4:2:2 back to k_fused2
  if (f0.iw != f0.fw || (f0.ox & 3) || (f0.ow & 3) || f0.ox + f0.iw > f0.ow) return false;   // no horizontal scaling; lanes = 4 px
  if ((f0.fw & 3) || f0.fw < 4 || f0.fh < 4 || f0.ih > F3_MAX_IH) return false;
  if (f0.fg.cw != f0.fw / 2 || f0.fg.ch != (f0.is_422 ? f0.fh : (f0.fh + 1) / 2)) return false;
  for (int i = 0; i < n; i++) {
    const FusedArgs &f = a[i];
    if (f.is_422 != f0.is_422 || f.low_quality != f0.low_quality || f.quirks != f0.quirks || f.conv.t != f0.conv.t) return false;
    if (f.fw != f0.fw || f.fh != f0.fh || f.ow != f0.ow || f.oh != f0.oh || f.ih != f0.ih || f.oy != f0.oy || f.ox != f0.ox) return false;
    if (f.fg.rs_y != f0.fg.rs_y || f.fg.rs_u != f0.fg.rs_u || f.fg.rs_v != f0.fg.rs_v || f.bg.rs != f0.bg.rs || f.out.rs != f0.out.rs)
      return false;
    if (((uintptr_t)f.fg.y | (uintptr_t)f.fg.u | (uintptr_t)f.fg.v) & 3) return false;
    if ((f.fg.rs_y | f.fg.rs_u | f.fg.rs_v) & 3) return false;
    if (((uintptr_t)f.bg.p | (uintptr_t)f.out.p) & 15) return false;
    if ((f.bg.rs | f.out.rs) & 15) return false;
  }
  return true;
}

// rows4_dev: int4 per inner output row {first, c3 | c2 << 16, c1 | c0 << 16, 0} (built by the engine from the filter bank)
cudaError_t launch_fused3(const Launch &L, const FusedArgs *frames_host, int nframes, int blend_a, const uint8_t *lut8_dev,
                          const void *rows4_dev, int coef16, unsigned int *sched_dev) {
  static PerDevice attr_set;
  if (!attr_set.cur()) {
    cudaError_t e;
    const int mx = f3_smem_bytes(F3_MAX_IH);
    const void *fns[16] = {(const void *)k_fused3<true, true, true, false>,   (const void *)k_fused3<true, true, false, false>,
                           (const void *)k_fused3<true, false, true, false>,  (const void *)k_fused3<true, false, false, false>,
                           (const void *)k_fused3<false, true, true, false>,  (const void *)k_fused3<false, true, false, false>,
                           (const void *)k_fused3<false, false, true, false>, (const void *)k_fused3<false, false, false, false>,
                           (const void *)k_fused3<true, true, true, true>,    (const void *)k_fused3<true, true, false, true>,
                           (const void *)k_fused3<true, false, true, true>,   (const void *)k_fused3<true, false, false, true>,
                           (const void *)k_fused3<false, true, true, true>,   (const void *)k_fused3<false, true, false, true>,
                           (const void *)k_fused3<false, false, true, true>,  (const void *)k_fused3<false, false, false, true>};
    for (int i = 0; i < 16; i++)
      if ((e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
    const void *fns422[8] = {(const void *)k_fused3<true, true, true, false, true>,   (const void *)k_fused3<true, true, false, false, true>,
                             (const void *)k_fused3<true, false, true, false, true>,  (const void *)k_fused3<true, false, false, false, true>,
                             (const void *)k_fused3<false, true, true, false, true>,  (const void *)k_fused3<false, true, false, false, true>,
                             (const void *)k_fused3<false, false, true, false, true>, (const void *)k_fused3<false, false, false, false, true>};
    for (int i = 0; i < 8; i++)
      if ((e = cudaFuncSetAttribute(fns422[i], cudaFuncAttributeMaxDynamicSharedMemorySize, mx)) != cudaSuccess) return e;
    attr_set.cur() = 1;
  }
  // Defaults measured on B200 (tools/f3_single_variants.sh, profiles/r02zc_f3_split_variants.log): the whole sequence in static shares
  // with border : inner rows costed 2 : 5.  Round 1's split (92 % static + a dynamic tail of 24-row chunks, 2 : 7) is as fast at 32
  // frames per launch (51.3 k fps either way) but every chunk is a segment start-up of ~3 us: a single-frame launch 41 -> 34 us
  static int cost_b = 0, cost_i = 0, static_pct = 100, chunk_rows = 24;
  if (!cost_b) {
    if (getenv("PE_F3_STATIC_PCT")) static_pct = atoi(getenv("PE_F3_STATIC_PCT"));
    if (getenv("PE_F3_CHUNK_ROWS")) chunk_rows = atoi(getenv("PE_F3_CHUNK_ROWS"));
    if (static_pct < 0 || static_pct > 100) static_pct = 100;
    if (chunk_rows < 4) chunk_rows = 4;
    const char *eb = getenv("PE_F3_COST_B"), *ei = getenv("PE_F3_COST_I");
    cost_b = eb ? atoi(eb) : 2;
    cost_i = ei ? atoi(ei) : 5;
    if (cost_b < 1) cost_b = 1;
    if (cost_i < 1) cost_i = 1;
  }
  const FusedArgs &a0 = frames_host[0];
  for (int base = 0; base < nframes; base += F3_MAXF) {
    Fused3Params P;
    P.nframes = nframes - base < F3_MAXF ? nframes - base : F3_MAXF;
    for (int i = 0; i < F3_MAXF; i++) {
      const FusedArgs &a = frames_host[base + (i < P.nframes ? i : 0)];
      P.fr[i] = Fused3Frame{a.fg.y, a.fg.u, a.fg.v, a.bg.p, a.out.p};
    }
    P.fw = a0.fw; P.fh = a0.fh; P.cw = a0.fg.cw; P.ch = a0.fg.ch;
    P.rs_y = a0.fg.rs_y; P.rs_u = a0.fg.rs_u; P.rs_v = a0.fg.rs_v; P.rs_bg = a0.bg.rs; P.rs_out = a0.out.rs;
    P.oh = a0.oh; P.oy = a0.oy; P.ih = a0.ih; P.ow = a0.ow; P.ox = a0.ox;
    P.nstrips = (a0.ow + 127) / 128;
    // the fast step reads whole words behind chroma column cw: on the last chroma row that needs 4 bytes of row padding
    const bool last_row_unsafe = a0.fg.rs_u < a0.fg.cw + 4 || a0.fg.rs_v < a0.fg.cw + 4;
    // (4:2:2: a fast step reads luma and chroma rows 2k - 1 and 2k)
    P.k_fast_max = a0.is_422 ? (a0.fh - 1 - (last_row_unsafe ? 1 : 0)) / 2 : a0.fg.ch - 1 - (last_row_unsafe ? 1 : 0);
    P.cost_b = cost_b; P.cost_i = cost_i;
    const long long strip_cost = (long long)cost_b * (a0.oh - a0.ih) + (long long)cost_i * a0.ih;
    P.frame_cost = strip_cost * P.nstrips;
    P.total_cost = P.frame_cost * P.nframes;
    const int grid = L.sm_count;
    // band height: the rows of an ALL-INNER share, so that inside the inner rectangle a warp owns about one (band, strip) unit and
    // consecutive warps march down neighbouring strips of the same band in step: the chroma sectors two strips share are then still
    // in L2 when the neighbour asks for them.  Measured at 32 frames per launch (tools/f3_band_variants.sh,
    // profiles/r02zf_f3_band_variants.log): 602 us and 1.69 GB of DRAM reads per launch, against 610 us / 1.81 GB with the average
    // rows of a share (border rows included) and 623 - 628 us / 1.88 GB with heights in between
    const long long share = P.total_cost / ((long long)grid * F3_NW);
    long long bh = share / cost_i;
    if (getenv("PE_F3_BAND_MODE")) {   // experiments: 1 = the average rows of a share, >= 16 = that many rows
      const int m = atoi(getenv("PE_F3_BAND_MODE"));
      if (m == 1) bh = share * a0.oh / (strip_cost > 0 ? strip_cost : 1);
      else if (m >= 16) bh = m;
    }
    if (bh < 16) bh = 16;
    if (bh > a0.oh) bh = a0.oh;
    P.band_h = (int)bh;
    P.nshares = (long long)grid * F3_NW;
    // Few frames per launch (the realtime path issues ONE): a share is a few dozen rows, and every unit a share straddles is another
    // segment start-up (~3 us: unit lookup, carried sums, ring prefetch, two steps before the first row leaves --
    // profiles/r02zc_f3_split_variants.log: warps with one segment end at 26 us, with two at 30).  So when the warps divide among the
    // strips almost evenly, the bands become the whole strip and the share count a multiple of the strip count: every share lies inside
    // one strip = one segment per warp; the few warps beyond the share count idle (<= 5 %).
    {
      const long long strips_total = (long long)P.nstrips * P.nframes;
      const long long per_strip = P.nshares / strips_total;
      if (getenv("PE_F3_NO_STRIP_SHARES") == nullptr && per_strip >= 1 && per_strip * strips_total * 100 >= P.nshares * 95) {
        P.nshares = per_strip * strips_total;
        P.band_h = a0.oh;
      }
    }
    // the last part of the sequence is handed out dynamically, in chunks of about F3_CHUNK_ROWS inner rows
    // (a chunk is at most a third of a warp's static share, so that a single-frame launch is not stretched by its last chunks)
    P.chunk_cost = (long long)cost_i * chunk_rows;
    if (P.chunk_cost > share / 3) P.chunk_cost = share / 3;
    if (P.chunk_cost < (long long)cost_i * 4) P.chunk_cost = (long long)cost_i * 4;
    P.static_cost = P.total_cost * static_pct / 100;
    if (P.total_cost - P.static_cost < P.chunk_cost * grid) P.static_cost = P.total_cost;  // small jobs: all static
    P.sched = sched_dev;
    P.ka = (uint32_t)blend_a; P.kia = (uint32_t)(256 - blend_a);
    P.zero = 0u;
    static int spread_pref = -1;
    if (spread_pref < 0) spread_pref = getenv("PE_F3_SPREAD") ? atoi(getenv("PE_F3_SPREAD")) != 0 : 0;
    P.spread = spread_pref;
    P.rows4 = reinterpret_cast<const int4 *>(rows4_dev);
    P.conv = a0.conv.t;
    P.lut8 = lut8_dev;
    if (P.frame_cost >= (1ll << 31)) return cudaErrorInvalidConfiguration;
#ifdef PE_F3_TIMELINE
    static unsigned long long *tl_dev = nullptr;
    const size_t tl_n = (size_t)grid * (2 + F3_NW) + (size_t)grid * F3_NW;
    if (!tl_dev) cudaMalloc(&tl_dev, tl_n * 8);
    cudaMemsetAsync(tl_dev, 0, tl_n * 8, L.stream);
    P.tl = tl_dev;
#endif
    const int smem_bytes = f3_smem_bytes(a0.ih);
    // the TMA variant of the bg ring takes whole 128-column strips of rows that start 16-byte aligned (PE_F3_TMA=0 / 1 overrides the
    // default, which is what measured faster: profiles/r02_k_fused3_tma_vs_ldgsts.txt)
    static int tma_pref = -1;
    if (tma_pref < 0) tma_pref = getenv("PE_F3_TMA") ? atoi(getenv("PE_F3_TMA")) != 0 : PE_F3_TMA_DEFAULT;
    const bool tma = tma_pref && a0.ox == 0 && a0.ow == a0.fw && (a0.ow & 127) == 0;
    auto go = [&](auto q, auto l, auto c) {
      if (a0.is_422) k_fused3<decltype(q)::value, decltype(l)::value, decltype(c)::value, false, true><<<grid, F3_NT, smem_bytes, L.stream>>>(P);
      else if (tma) k_fused3<decltype(q)::value, decltype(l)::value, decltype(c)::value, true><<<grid, F3_NT, smem_bytes, L.stream>>>(P);
      else k_fused3<decltype(q)::value, decltype(l)::value, decltype(c)::value, false><<<grid, F3_NT, smem_bytes, L.stream>>>(P);
    };
    auto pick_c = [&](auto q, auto l) { if (coef16) go(q, l, std::true_type()); else go(q, l, std::false_type()); };
    auto pick_l = [&](auto q) { if (lut8_dev) pick_c(q, std::true_type()); else pick_c(q, std::false_type()); };
    if (a0.quirks) pick_l(std::true_type()); else pick_l(std::false_type());
    PE_COUNT_LAUNCH(L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
#ifdef PE_F3_TIMELINE
    if (getenv("PE_F3_TIMELINE_DUMP")) {   // one line per launch: stamps relative to the earliest CTA start, in ns
      std::vector<unsigned long long> h(tl_n);
      cudaStreamSynchronize(L.stream);
      cudaMemcpy(h.data(), tl_dev, tl_n * 8, cudaMemcpyDeviceToHost);
      unsigned long long t0 = ~0ull, s_max = 0, f_max = 0, e_min = ~0ull, e_max = 0;
      double e_sum = 0;
      for (int b = 0; b < grid; b++) t0 = h[(size_t)b * (2 + F3_NW)] < t0 ? h[(size_t)b * (2 + F3_NW)] : t0;
      for (int b = 0; b < grid; b++) {
        const unsigned long long *r = &h[(size_t)b * (2 + F3_NW)];
        if (r[0] - t0 > s_max) s_max = r[0] - t0;
        if (r[1] - t0 > f_max) f_max = r[1] - t0;
        for (int w = 0; w < F3_NW; w++) {
          const unsigned long long x = r[2 + w] - t0;
          if (x < e_min) e_min = x;
          if (x > e_max) e_max = x;
          e_sum += (double)x;
        }
      }
      fprintf(stderr, "f3 timeline (ns after the first CTA start): last CTA start %llu, last table fill done %llu, warp ends min %llu mean %.0f max %llu, frames %d\n",
              s_max, f_max, e_min, e_sum / ((double)grid * F3_NW), e_max, P.nframes);
      // mean / max end per 5 % bucket of the share number (shares follow the frame from its top border band to its bottom one)
      fprintf(stderr, "f3 timeline buckets (mean/max us):");
      for (int q = 0; q < 20; q++) {
        double sum = 0; unsigned long long mx = 0; int cnt = 0;
        for (int b = 0; b < grid; b++)
          for (int w = 0; w < F3_NW; w++) {
            const long long gw = (long long)b * F3_NW + w;
            if (gw * 20 / ((long long)grid * F3_NW) != q) continue;
            const unsigned long long x = h[(size_t)b * (2 + F3_NW) + 2 + w] - t0;
            sum += (double)x; if (x > mx) mx = x; cnt++;
          }
        fprintf(stderr, " %.1f/%.1f", sum / (cnt ? cnt : 1) / 1e3, (double)mx / 1e3);
      }
      fprintf(stderr, "\n");
      if (P.nframes == 1) {   // end time (us) by number of segments, and the slowest warps
        double sum[8] = {0}; int cnt[8] = {0};
        std::vector<std::pair<unsigned long long, int>> all;
        for (int b = 0; b < grid; b++)
          for (int w = 0; w < F3_NW; w++) {
            const unsigned long long x = h[(size_t)b * (2 + F3_NW) + 2 + w] - t0, info = h[(size_t)grid * (2 + F3_NW) + (size_t)b * F3_NW + w];
            const int ns = (int)(info & 0xFFFF) < 7 ? (int)(info & 0xFFFF) : 7;
            sum[ns] += (double)x; cnt[ns]++;
            all.push_back({x, b * F3_NW + w});
          }
        fprintf(stderr, "f3 timeline by segments:");
        for (int i = 0; i < 8; i++) if (cnt[i]) fprintf(stderr, " nseg %d: %d warps mean %.1f us;", i, cnt[i], sum[i] / cnt[i] / 1e3);
        fprintf(stderr, "\n");
        std::sort(all.begin(), all.end());
        for (int i = 0; i < 6; i++) {
          const auto &a = all[all.size() - 1 - i];
          const unsigned long long info = h[(size_t)grid * (2 + F3_NW) + a.second];
          fprintf(stderr, "  slow warp gw %d (sm %d warp %d): %.1f us, nseg %llu rows %llu inner %llu\n", a.second, a.second / F3_NW, a.second % F3_NW, a.first / 1e3,
                  info & 0xFFFF, (info >> 16) & 0xFFFFFF, info >> 40);
        }
        for (int i = 0; i < 3; i++) {
          const auto &a = all[i];
          const unsigned long long info = h[(size_t)grid * (2 + F3_NW) + a.second];
          fprintf(stderr, "  fast warp gw %d: %.1f us, nseg %llu rows %llu inner %llu\n", a.second, a.first / 1e3, info & 0xFFFF, (info >> 16) & 0xFFFFFF, info >> 40);
        }
      }
    }
#endif
  }
  return cudaSuccess;
}

}  // namespace pe
