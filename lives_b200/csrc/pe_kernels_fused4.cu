// pe_kernels_fused4.cu -- k_cvt_resize: planar 4:2:0 -> RGBA32 / BGRA32 AND the separable resize (both axes scaled, <= 4 non-negative
// taps per axis: the bilinear banks) in ONE kernel.  BASELINE config 2 (1080p YUV420P -> RGBA32 -> 1280 x 720) as resize_layer_full
// runs it (convert, then scale: colourspace.c:14601-14620 / our contract DESIGN.md section 5), without the 8.3 MB RGBA intermediate
// ever reaching HBM: the unfused pair k_yuv_march + k_resize_tile4 moved 2.5 x the algorithmic bytes (profiles/r01s_*, r01t_*).
// Same arithmetic, bit for bit, as the two unfused ops (tests/test_gpu_parity.py::test_cvt_resize_*).
//
// One 512-thread CTA per SM walks output tiles of 128 x 32 pixels (persistent, tile = blockIdx.x + i * gridDim.x); per tile
//   0. (issued one tile ahead with cp.async, landing while the previous tile's passes 2 and 3 run) the raw source words the tile
//      needs: luma rows [2 k0 - 1, 2 k1], chroma rows [k0 - 1, k1] -- whole reference row pairs (colourspace.c:3440-3549);
//   1. conversion, one thread = 4 columns x one row pair, with k_fused3's arithmetic (bank-replicated tables, packed chroma sums,
//      third_round as one multiply-high) but BOTH chroma rows summed locally; the result goes to three byte planes in shared memory
//      (R, G, B: the alpha of every source pixel is 255 and is carried as two per-column / per-row factors instead);
//   2. horizontal pass: the four taps of a column are four consecutive bytes of a plane row = two aligned words + one funnel shift,
//      two DP2A, >> 7, min 32767 -> a 16-bit intermediate row in shared memory (one warp = one source row, a lane = 4 output columns
//      32 apart: every access of the warp is one contiguous wavefront, the coefficients stay in registers across the warp's rows);
//   3. vertical pass: 4 taps x 3 channels of 16-bit loads and multiply-adds, + 2^18 >> 19, saturating pack, one 32-bit store per pixel
//      (a warp = one output row: 128-byte contiguous stores).
// Rows the fast step cannot take (row 0, the last row of an even frame, the last chroma row of a plane without padding) go through
// slow_unit(): scalar code with the reference's edge rules, a few units per frame.
#include <cuda.h>   // CUtensorMap (the 2-D TMA staging variant); the encoder is fetched through cudaGetDriverEntryPoint: no -lcuda

#include <cstdlib>
#include <cstdio>

#include "pe_device.cuh"
#include "pe_kernels.h"
#include "pe_tables.h"

namespace pe {

namespace {

#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

constexpr int F4_NT = 512, F4_NW = F4_NT / 32;
constexpr int F4_TW = 128, F4_TH = 32;
constexpr int F4_TY = 0, F4_TV = 32768, F4_TU = 65536, F4_DYN = 98304;  // u32 [256][32] | uint2 [256][16] | uint2 [256][16] | per-tile part
constexpr int F4_MAXF = 32;
constexpr int F4_SMEM_MAX = 227 * 1024;
// Per-tile maxima, COMPILE-TIME so that every shared-memory address in the three passes is a register plus an immediate: a 128 x 32 tile
// of a <= 4-tap bank (scale factors down to 1 : 1.5 per axis) needs at most 4-column groups / row pairs / source rows as below; a
// geometry that asks for more takes the unfused kernels (launch_cvt_resize returns cudaErrorInvalidConfiguration).
constexpr int F4_GW = 56;                   // words per staged luma row (224 bytes = 14 x 16: 51 groups + 12 bytes of alignment slack)
constexpr int F4_NG = 51;                   // 4-column groups a tile may convert per row pair
constexpr int F4_PR = 28;                   // row pairs
constexpr int F4_TR = 54;                   // source rows the vertical pass reads
constexpr int F4_CWB = 128;                 // bytes per staged chroma row (8 x 16)
constexpr int F4_PLS = (F4_NG + 1) * 4;     // plane row stride
constexpr int F4_PLP = 2 * F4_PR * F4_PLS;  // plane size
constexpr int F4_TMP = (F4_TR + 3) * F4_TW; // 16-bit samples per intermediate plane
constexpr int O_RAWY = F4_DYN, O_RAWU = O_RAWY + 2 * F4_PR * F4_GW * 4, O_RAWV = O_RAWU + (F4_PR + 1) * F4_CWB, O_VF = O_RAWV + (F4_PR + 1) * F4_CWB;
constexpr int O_PL = (O_VF + (F4_PR + 1) * 4 + 15) & ~15, O_TMP = (O_PL + 3 * F4_PLP + 15) & ~15, O_CX = (O_TMP + 3 * F4_TMP * 2 + 15) & ~15;
// filter rows of the tile, double buffered (tile parity): int4 per output column / row {c0 | c1 << 16, c2 | c3 << 16, first, aux};
// tile tables: the tap range [first, last] of every tile column / tile row, read once per launch
constexpr int F4_MAXTX = 128, F4_MAXTY = 160;
constexpr int O_PY = O_CX + 2 * F4_TW * 16, O_TCOL = O_PY + 2 * F4_TH * 16, O_TROW = O_TCOL + F4_MAXTX * 16;
constexpr int O_MBAR = O_TROW + F4_MAXTY * 8;   // one mbarrier: the bulk copies of a tile's raw rows complete on it
constexpr int F4_SMEM = O_MBAR + 16;
static_assert(F4_SMEM <= F4_SMEM_MAX, "k_cvt_resize: shared memory budget");

struct CvtRszParams {
  const uint8_t *y[F4_MAXF], *u[F4_MAXF], *v[F4_MAXF];
  uint8_t *dst[F4_MAXF];
  int nframes;
  int fw, fh, cw, ch, rs_y, rs_u, rs_v;   // source (shared by all frames of the launch)
  int dw, dh, drs;                        // destination
  int tiles_x, tiles_y;
  int vec16;                              // planes and strides 16-byte aligned: the raw words travel as 16-byte cp.async
  int tma;                                // ... or (PE_F4_TMA=1, needs vec16) as one cp.async.bulk per row, see stage()
  int k_fast_max;
  int swap_rb;                            // BGRA32
  const int4 *px, *py;                    // [dw] / [dh]: {c0 | c1 << 16, c2 | c3 << 16, first, aux}; aux = alpha's 15-bit intermediate / the sum of the row's taps
  int fx_taps, fy_taps;
  uint32_t magic_tpf, magic_tx;           // ceil(2^32 / tiles per frame), ceil(2^32 / tiles_x): tile index -> (frame, ty, tx) without a division (0: divide)
  const int32_t *conv;                    // [14][256]
#ifdef PE_F4_TIMELINE
  unsigned long long *tl;                 // [grid][6]: cycles of CTA thread 0 in: wait for the raw words, conversion, staging the next tile, horizontal, vertical, total
#endif
};

// 2-D tensor maps of the launch's planes (PE_F4_TMA=2): box = the largest staged rectangle of a tile, so ONE cp.async.bulk.tensor per plane
// stages a tile (rows outside the plane arrive as zeros: only the slow units, which read global memory themselves, would look at them)
struct alignas(64) CvtRszMaps {
  CUtensorMap y[F4_MAXF], u[F4_MAXF], v[F4_MAXF];
};

__device__ __forceinline__ void f4_tensor_g2s_2d(uint32_t smem_dst, const CUtensorMap *map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t f4_dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t f4_dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t f4_pack_sat(int a, int b, uint32_t c) {  // (c[15:0] << 16) | (sat_u8(a) << 8) | sat_u8(b)
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ void f4_cp_async4(uint32_t smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void f4_cp_async16(uint32_t smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void f4_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void f4_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: the raw rows of a tile.  One instruction moves a whole
// row (<= 224 bytes); as 16-byte cp.async (LDGSTS) the staging of a tile was ~1200 instructions with their 64-bit addresses spread over
// all 16 warps and took 19 % of the kernel (-DPE_F4_TIMELINE); as bulk copies it is ~4 instructions per lane of ONE warp.
__device__ __forceinline__ void f4_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void f4_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f4_bulk_g2s(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void f4_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PE_F4_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PE_F4_MBAR_DONE;\n"
      "bra PE_F4_MBAR_WAIT;\n"
      "PE_F4_MBAR_DONE:\n"
      "}" :: "r"(bar), "r"(parity) : "memory");
}

constexpr uint32_t F4_MSK = 0xFFFEFFFEu, F4_K3 = 0x00030003u;
// third_round of the HIGH / LOW half of a packed Q = 2 n + 3, as the byte offset 128 * m of the chroma table entry (k_fused3: idx_hi / idx_lo)
__device__ __forceinline__ uint32_t f4_idx_hi(uint32_t q) { return __umulhi(q, 10923u * 128u) & 0x7F80u; }
__device__ __forceinline__ uint32_t f4_idx_lo(uint32_t q) { return __umulhi(q << 16, 10923u * 128u) & 0x7F80u; }

struct F4Row {
  uint32_t a, b, c;  // a = [c(jc0), c(jc0+1)], b = [c(jc0-1), c(jc0)], c = [c(jc0+1), c(jc0+2)] as 16-bit halves
};
__device__ __forceinline__ F4Row f4_unpack(uint32_t w0, uint32_t w1, uint32_t sel) {
  const uint32_t cw4 = __byte_perm(w0, w1, sel);
  F4Row r;
  r.a = __byte_perm(cw4, 0u, 0x4241u);
  r.b = __byte_perm(cw4, 0u, 0x4140u);
  r.c = __byte_perm(cw4, 0u, 0x4342u);
  return r;
}

__device__ __forceinline__ int f4_chroma_at(const uint8_t *__restrict__ p, int stride, int r, int c, int cw, int ch) {
  if (c >= cw) c = (cw < stride || r + 1 < ch) ? cw : cw - 1;
  return __ldg(p + (long long)stride * r + c);
}

// (u, v) of source pixel (sx, sy) of a 4:2:0 frame with the reference's edge rules: the per-pixel form of colourspace.c:3391-3642
// (identical to chroma_for_pixel of pe_kernels_fused.cu / k_yuv_planar_to_rgb)
template <bool QUIRKS>
__device__ void f4_chroma_px(const uint8_t *pu, const uint8_t *pv, int rs_u, int rs_v, int cw, int ch, int h, int sx, int sy, int &u, int &v) {
  const int jc = sx >> 1, right = sx & 1;
  if (sy == 0 || (!(h & 1) && sy == h - 1)) {
    const int cr = sy == 0 ? 0 : ch - 1;
    const int ca = jc, cb = right ? jc + 1 : jc - 1;
    const int ua = ca <= 0 ? __ldg(pu + (long long)rs_u * cr) : f4_chroma_at(pu, rs_u, cr, ca, cw, ch);
    const int va = ca <= 0 ? __ldg(pv + (long long)rs_v * cr) : f4_chroma_at(pv, rs_v, cr, ca, cw, ch);
    const int ub = cb <= 0 ? __ldg(pu + (long long)rs_u * cr) : f4_chroma_at(pu, rs_u, cr, cb, cw, ch);
    const int vb = cb <= 0 ? __ldg(pv + (long long)rs_v * cr) : f4_chroma_at(pv, rs_v, cr, cb, cw, ch);
    u = (ua + ub) >> 1;
    v = (va + vb) >> 1;
    return;
  }
  const int k = (sy + 1) >> 1, cr_a = k - 1, cr_b = k, upper = sy & 1;
  const int cn = right ? jc + 1 : max(jc - 1, 0);
  const int u1t = f4_chroma_at(pu, rs_u, cr_a, jc, cw, ch), u1n = f4_chroma_at(pu, rs_u, cr_a, cn, cw, ch);
  const int u2t = f4_chroma_at(pu, rs_u, cr_b, jc, cw, ch), u2n = f4_chroma_at(pu, rs_u, cr_b, cn, cw, ch);
  const int v1t = f4_chroma_at(pv, rs_v, cr_a, jc, cw, ch), v1n = f4_chroma_at(pv, rs_v, cr_a, cn, cw, ch);
  const int v2t = f4_chroma_at(pv, rs_v, cr_b, jc, cw, ch), v2n = f4_chroma_at(pv, rs_v, cr_b, cn, cw, ch);
  int u1 = u1t + u1n, u2 = u2t + u2n, v1 = v1t + v1n, v2 = v2t + v2n;
  if (QUIRKS && !right) {
    u2 = u1;                                             // colourspace.c:3461
    if (jc > 0) v1 = v1t + v2n;                          // :3544
    v2 = v2t + __ldg(pv + (long long)rs_v * cr_b);       // last_v2 never advanced
  }
  u = upper ? third_round(u1 + (u2 >> 1)) : third_round((u1 >> 1) + u2);
  v = upper ? third_round(v1 + (v2 >> 1)) : third_round((v1 >> 1) + v2);
}

template <bool QUIRKS>
__global__ void __launch_bounds__(F4_NT, 1) k_cvt_resize(const __grid_constant__ CvtRszParams P, const __grid_constant__ CvtRszMaps M) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  fill_replicated_yuv_tables<F4_NT>(smem + F4_TY, smem + F4_TV, smem + F4_TU, P.conv, tid);

  uint8_t *const s_rawy = smem + O_RAWY, *const s_rawu = smem + O_RAWU, *const s_rawv = smem + O_RAWV;
  uint32_t *const s_vf = reinterpret_cast<uint32_t *>(smem + O_VF);
  uint8_t *const s_pl = smem + O_PL;
  uint16_t *const s_tmp = reinterpret_cast<uint16_t *>(smem + O_TMP);
  int4 *const s_tcol = reinterpret_cast<int4 *>(smem + O_TCOL);   // per tile column: first / last tap column, ceil(2^32 / groups) of the column
  int2 *const s_trow = reinterpret_cast<int2 *>(smem + O_TROW);
  for (int i = tid; i < P.tiles_x; i += F4_NT) {
    const int vc0 = __ldg(&P.px[i * F4_TW].z), vc1 = __ldg(&P.px[min(i * F4_TW + F4_TW, P.dw) - 1].z) + P.fx_taps - 1;
    const int ng = (((vc1 | 3) + 1) - (vc0 & ~3)) >> 2;
    s_tcol[i] = make_int4(vc0, vc1, (int)(0xFFFFFFFFu / (uint32_t)ng + 1u), 0);   // (the third word: ceil(2^32 / ng) for ng > 1)
  }
  for (int i = tid; i < P.tiles_y; i += F4_NT)
    s_trow[i] = make_int2(__ldg(&P.py[i * F4_TH].z), __ldg(&P.py[min(i * F4_TH + F4_TH, P.dh) - 1].z) + P.fy_taps - 1);
  if (tid == 0) {
    f4_mbar_init(sbase + O_MBAR, 32u);   // the 32 lanes of warp 0 arrive once per tile, each with the bytes it has asked for
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  constexpr int GW = F4_GW, CWB = F4_CWB;
  const int tiles_per_frame = P.tiles_x * P.tiles_y;
  const int total = tiles_per_frame * P.nframes;  // (fits 31 bits: checked by the launcher)
  const uint32_t lane4 = 4u * (uint32_t)lane, lane8 = 8u * (uint32_t)(lane & 15);

  struct Geo {
    int f, x0, y0, ncol, nrow, vr0, nvr, cb, ng, k0, np, ybase, ubase;  // cb: first converted column; ybase / ubase: first staged luma column / chroma byte
    uint32_t ng_magic;  // ceil(2^32 / ng): i / ng = umulhi(i, magic) for the unit indices of a tile (i < 2^16)
  };
  auto geometry = [&](int t) -> Geo {
    Geo g;
    // (the divisions of the tile index were a visible part of the 19 % of the kernel spent between the conversion and the horizontal
    // pass -- every thread computes the next tile's geometry: multiply-high by constants of the launch instead)
    g.f = P.magic_tpf ? (int)__umulhi((uint32_t)t, P.magic_tpf) : t / tiles_per_frame;
    const int r = t - g.f * tiles_per_frame;
    const int ty = P.magic_tx ? (int)__umulhi((uint32_t)r, P.magic_tx) : r / P.tiles_x, tx = r - ty * P.tiles_x;
    g.x0 = tx * F4_TW; g.y0 = ty * F4_TH;
    const int x1 = min(g.x0 + F4_TW, P.dw), y1 = min(g.y0 + F4_TH, P.dh);
    g.ncol = x1 - g.x0; g.nrow = y1 - g.y0;
    const int4 tc = s_tcol[tx];
    const int2 trw = s_trow[ty];
    const int vc0 = tc.x, vc1 = tc.y, vr1 = trw.y;
    g.vr0 = trw.x;
    g.nvr = vr1 - g.vr0 + 1;
    g.cb = vc0 & ~3;
    g.ng = (((vc1 | 3) + 1) - g.cb) >> 2;
    g.k0 = (g.vr0 + 1) >> 1;
    g.np = ((vr1 + 1) >> 1) - g.k0 + 1;
    g.ubase = g.cb == 0 ? 0 : (((g.cb >> 1) - 1) & ~3);
    g.ybase = g.cb;
    if (P.vec16) { g.ybase &= ~15; g.ubase &= ~15; }  // the staged rectangle starts on a 16-byte boundary of the planes
    g.ng_magic = (uint32_t)tc.z;  // = ceil(2^32 / ng) for ng > 1
    return g;
  };
  // cp.async of the raw words of tile t: luma rows 2 k0 - 1 .. 2 k1 (clamped to the frame), chroma rows k0 - 1 .. k1, the first V
  // sample of every chroma row (the reference's never-advanced last_v2, colourspace.c:3544)
  auto box_staged = [&](const Geo &g) -> bool { return P.tma == 2 && g.ubase + 2 * g.ng + 21 <= min(P.rs_u, P.rs_v); };
  auto mbar_staged = [&](const Geo &g) -> bool { return P.tma == 1 || box_staged(g); };   // which completion the tile's raw words signal
  auto stage = [&](const Geo &g, int buf) {
    const uint8_t *Fy = P.y[g.f], *Fu = P.u[g.f], *Fv = P.v[g.f];
    const int rowbase = 2 * g.k0 - 1;
    const long long ulim = (long long)P.rs_u * P.ch, vlim = (long long)P.rs_v * P.ch;
    // A tile whose chroma bytes run past the end of the plane's rows (the reference's one-past-row read, colourspace.c:3508: with a
    // stride equal to the width those bytes are the NEXT row's first samples) cannot come through a 2-D box -- the box would deliver
    // zeros there -- and takes the 16-byte cp.async path (box_staged() below: its completion is the cp.async group, not the mbarrier).
    if (box_staged(g)) {
      // three tensor copies (luma, U, V rectangles of the compile-time box sizes) + two bulk copies (filter rows), all from lane 0 of warp 0
      if (warp == 0) {
        const uint32_t bar = sbase + O_MBAR;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        f4_mbar_expect_tx(bar, lane == 0 ? (uint32_t)(2 * F4_PR * GW * 4 + 2 * (F4_PR + 1) * CWB) + (uint32_t)(g.ncol + g.nrow) * 16u : 0u);
        if (lane == 0) {
          f4_tensor_g2s_2d(sbase + O_RAWY, &M.y[g.f], g.ybase, rowbase, bar);
          f4_tensor_g2s_2d(sbase + O_RAWU, &M.u[g.f], g.ubase, g.k0 - 1, bar);
          f4_tensor_g2s_2d(sbase + O_RAWV, &M.v[g.f], g.ubase, g.k0 - 1, bar);
          f4_bulk_g2s(sbase + O_CX + (buf * F4_TW) * 16, P.px + g.x0, (uint32_t)g.ncol * 16u, bar);
          f4_bulk_g2s(sbase + O_PY + (buf * F4_TH) * 16, P.py + g.y0, (uint32_t)g.nrow * 16u, bar);
        }
      } else if (warp == 1 && lane <= g.np) {
        const int cr = min(max(g.k0 - 1 + lane, 0), P.ch - 1);
        f4_cp_async4(sbase + O_VF + 4 * lane, Fv + (size_t)P.rs_v * cr);
      }
      f4_cp_async_commit();
      return;
    }
    if (P.tma == 1) {
      // warp 0: lane l asks for rows l, l + 32, ... of the tile's 2 np luma + 2 (np + 1) chroma rows with one bulk copy each (a row:
      // <= 224 / 128 bytes, 16-byte aligned at both ends), lanes 0 / 1 for the tile's filter rows as well; every lane arrives on the
      // mbarrier with the byte count it is about to request.  The other warps go straight on to the horizontal pass.
      if (warp == 0) {
        const uint32_t bar = sbase + O_MBAR;
        const int ny = 2 * g.np, nc = g.np + 1, nrows = ny + 2 * nc;
        const uint32_t ybytes = (uint32_t)((g.cb - g.ybase + 4 * g.ng + 15) >> 4) * 16u;
        const uint32_t cbytes = (uint32_t)((2 * g.ng + 21 + 15) >> 4) * 16u;   // (a row pair's last unit reads up to byte 2 ng + 20 of the staged row)
        auto row_copy = [&](int r, const uint8_t *&src, uint32_t &dst, uint32_t &bytes) {
          if (r < ny) {
            const int sr = min(max(rowbase + r, 0), P.fh - 1);
            src = Fy + (size_t)P.rs_y * sr + g.ybase; dst = sbase + O_RAWY + r * (GW * 4); bytes = ybytes;
          } else {
            const bool isv = r - ny >= nc;
            const int ri = r - ny - (isv ? nc : 0);
            const int cr = min(max(g.k0 - 1 + ri, 0), P.ch - 1);
            const long long o = (long long)(isv ? P.rs_v : P.rs_u) * cr + g.ubase, lim = isv ? vlim : ulim;
            src = (isv ? Fv : Fu) + o; dst = sbase + (isv ? O_RAWV : O_RAWU) + ri * CWB;
            bytes = (uint32_t)min((long long)cbytes, lim - o);   // never past the end of the plane (planes are whole 16-byte chunks here)
          }
        };
        uint32_t mine = 0;
        for (int r = lane; r < nrows; r += 32) {
          const uint8_t *src; uint32_t dst, bytes;
          row_copy(r, src, dst, bytes);
          mine += bytes;
        }
        if (lane == 0) mine += (uint32_t)g.ncol * 16u;
        if (lane == 1) mine += (uint32_t)g.nrow * 16u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffers were read through the generic proxy until the barrier before
        f4_mbar_expect_tx(bar, mine);
        for (int r = lane; r < nrows; r += 32) {
          const uint8_t *src; uint32_t dst, bytes;
          row_copy(r, src, dst, bytes);
          f4_bulk_g2s(dst, src, bytes, bar);
        }
        // the filter rows of the tile (16-byte entries of cudaMalloc'ed arrays; x0 / y0 are multiples of the tile size)
        if (lane == 0) f4_bulk_g2s(sbase + O_CX + (buf * F4_TW) * 16, P.px + g.x0, (uint32_t)g.ncol * 16u, bar);
        if (lane == 1) f4_bulk_g2s(sbase + O_PY + (buf * F4_TH) * 16, P.py + g.y0, (uint32_t)g.nrow * 16u, bar);
      } else if (warp == 1 && lane <= g.np) {
        const int cr = min(max(g.k0 - 1 + lane, 0), P.ch - 1);
        f4_cp_async4(sbase + O_VF + 4 * lane, Fv + (size_t)P.rs_v * cr);
      }
      f4_cp_async_commit();
      return;
    }
    if (P.vec16) {
      // a warp = one row at a time, a lane = one 16-byte chunk of it (rows are <= 13 / 8 chunks: few lanes, but no index arithmetic)
      const int nchy = (g.cb - g.ybase + 4 * g.ng + 15) >> 4;
      for (int ri = warp; ri < 2 * g.np; ri += F4_NW) {
        const int sr = min(max(rowbase + ri, 0), P.fh - 1);
        if (lane < nchy) f4_cp_async16(sbase + O_RAWY + ri * (GW * 4) + 16 * lane, Fy + (size_t)P.rs_y * sr + g.ybase + 16 * lane);
      }
      for (int ri = warp; ri <= g.np; ri += F4_NW) {
        const int cr = min(max(g.k0 - 1 + ri, 0), P.ch - 1);
        const int w = lane & 7;
        if (lane < 16 && 16 * w < 2 * g.ng + 21) {  // lanes 0 .. 7: U, 8 .. 15: V (a row pair's last unit reads up to byte 2 ng + 20 of the staged row)
          const bool isv = lane >= 8;
          const long long o = min((long long)(isv ? P.rs_v : P.rs_u) * cr + g.ubase + 16 * w, (isv ? vlim : ulim) - 16);
          f4_cp_async16(sbase + (isv ? O_RAWV : O_RAWU) + ri * CWB + 16 * w, (isv ? Fv : Fu) + o);
        }
        if (lane == 16) f4_cp_async4(sbase + O_VF + 4 * ri, Fv + (size_t)P.rs_v * cr);
      }
    } else {
      for (int i = tid; i < 2 * g.np * g.ng; i += F4_NT) {
        const int ri = g.ng == 1 ? i : (int)__umulhi((uint32_t)i, g.ng_magic), w = i - ri * g.ng;
        const int sr = min(max(rowbase + ri, 0), P.fh - 1);
        f4_cp_async4(sbase + O_RAWY + ri * (GW * 4) + 4 * w, Fy + (size_t)P.rs_y * sr + g.ybase + 4 * w);
      }
      const int cww = g.ng / 2 + 3;
      for (int i = tid; i < (g.np + 1) * cww; i += F4_NT) {
        const int ri = i / cww, w = i - ri * cww;
        const int cr = min(max(g.k0 - 1 + ri, 0), P.ch - 1);
        const long long ou = min((long long)P.rs_u * cr + g.ubase + 4 * w, ulim - 4), ov = min((long long)P.rs_v * cr + g.ubase + 4 * w, vlim - 4);
        f4_cp_async4(sbase + O_RAWU + ri * CWB + 4 * w, Fu + ou);
        f4_cp_async4(sbase + O_RAWV + ri * CWB + 4 * w, Fv + ov);
      }
      for (int i = tid; i <= g.np; i += F4_NT) {
        const int cr = min(max(g.k0 - 1 + i, 0), P.ch - 1);
        f4_cp_async4(sbase + O_VF + 4 * i, Fv + (size_t)P.rs_v * cr);
      }
    }
    // the filter rows of the tile (their 16-byte entries are aligned: cudaMalloc'ed arrays, x0 / y0 multiples of the tile size)
    if (tid < g.ncol) f4_cp_async16(sbase + O_CX + (buf * F4_TW + tid) * 16, P.px + g.x0 + tid);
    else if (tid >= 256 && tid - 256 < g.nrow) f4_cp_async16(sbase + O_PY + (buf * F4_TH + tid - 256) * 16, P.py + g.y0 + tid - 256);
    f4_cp_async_commit();
  };

  // yuv2rgb_int through the replicated tables, UNSATURATED (ou, ov = 128 * m)
  auto rgb = [&](uint32_t y, uint32_t ou, uint32_t ov, int &r, int &g, int &b) {
    const int yy = (int)*reinterpret_cast<const uint32_t *>(smem + F4_TY + (y * 128u + lane4));
    const uint2 tv = *reinterpret_cast<const uint2 *>(smem + F4_TV + (ov | lane8));
    const uint2 tu = *reinterpret_cast<const uint2 *>(smem + F4_TU + (ou | lane8));
    r = (yy + (int)tv.x) >> 16;
    g = (yy + (int)tu.x + (int)tv.y) >> 16;
    b = (yy + (int)tu.y) >> 16;
  };

#ifdef PE_F4_TIMELINE
  long long tl_acc[5] = {0, 0, 0, 0, 0};
  const long long tl_begin = clock64();
  long long tl_mark = tl_begin;
  auto tl_lap = [&](int k) { const long long now = clock64(); tl_acc[k] += now - tl_mark; tl_mark = now; };
#endif
  int t = blockIdx.x;
  if (t >= total) return;
  Geo G = geometry(t);
  int buf = 0;
  uint32_t tile_phase = 0u;   // tiles this CTA has staged through the mbarrier so far: its phase
  stage(G, buf);
  for (; t < total; t += gridDim.x, buf ^= 1) {
    f4_cp_async_wait_all();
    if (mbar_staged(G)) { f4_mbar_wait(sbase + O_MBAR, tile_phase & 1u); tile_phase++; }
    __syncthreads();  // raw words of this tile have landed; the previous tile's passes are done with the planes, the intermediate and the filter rows
#ifdef PE_F4_TIMELINE
    tl_lap(0);
#endif
    const Geo g = G;
    const int rowbase = 2 * g.k0 - 1;
    const int4 *const s_px = reinterpret_cast<const int4 *>(smem + O_CX) + buf * F4_TW, *const s_py = reinterpret_cast<const int4 *>(smem + O_PY) + buf * F4_TH;
    // ---- 1. conversion: unit = (row pair p, 4-column group gi) -> rows 2p, 2p + 1 of the three planes
    for (int i = tid; i < g.np * g.ng; i += F4_NT) {
      const int p = g.ng == 1 ? i : (int)__umulhi((uint32_t)i, g.ng_magic), gi = i - p * g.ng;  // (ceil(2^32 / 1) does not fit the magic)
      const int k = g.k0 + p, x = g.cb + 4 * gi;
      uint32_t oA[3], oB[3];
      if (k >= 1 && k <= P.k_fast_max) {
        const uint32_t yA = *reinterpret_cast<const uint32_t *>(s_rawy + (2 * p) * (GW * 4) + (x - g.ybase));
        const uint32_t yB = *reinterpret_cast<const uint32_t *>(s_rawy + (2 * p + 1) * (GW * 4) + (x - g.ybase));
        const int jc0 = x >> 1, o = jc0 - 1;
        const int off0 = x == 0 ? 0 : (o & ~3);
        const uint32_t sel = x == 0 ? 0x2100u : ((o & 3) == 3 ? 0x6543u : 0x4321u);
        const uint32_t selB = x == 0 ? 0x3254u : 0x3210u;
        const int wo = off0 - g.ubase;
        const uint32_t *up = reinterpret_cast<const uint32_t *>(s_rawu + p * CWB + wo), *uc = reinterpret_cast<const uint32_t *>(s_rawu + (p + 1) * CWB + wo);
        const uint32_t *vp = reinterpret_cast<const uint32_t *>(s_rawv + p * CWB + wo), *vc = reinterpret_cast<const uint32_t *>(s_rawv + (p + 1) * CWB + wo);
        const F4Row Up = f4_unpack(up[0], up[1], sel), Uc = f4_unpack(uc[0], uc[1], sel);
        const F4Row Vp = f4_unpack(vp[0], vp[1], sel), Vc = f4_unpack(vc[0], vc[1], sel);
        // right pixel of both chroma columns: this + next; rows 2k-1 ("up") / 2k ("lo") weigh the chroma rows k-1 / k 2 : 1 and 1 : 2
        const uint32_t RUp = Up.a + Up.c, RUc = Uc.a + Uc.c, RVp = Vp.a + Vp.c, RVc = Vc.a + Vc.c;
        const uint32_t QUR_up = RUp * 2u + (RUc & F4_MSK) + F4_K3, QUR_lo = (RUp & F4_MSK) + RUc * 2u + F4_K3;
        const uint32_t QVR_up = RVp * 2u + (RVc & F4_MSK) + F4_K3, QVR_lo = (RVp & F4_MSK) + RVc * 2u + F4_K3;
        // left pixel: this + last
        uint32_t QUL_up, QUL_lo, QVL_up, QVL_lo;
        const uint32_t LUp = Up.a + Up.b;
        if (QUIRKS) {
          QUL_up = QUL_lo = LUp * 2u + (LUp & F4_MSK) + F4_K3;          // u2 = this_u1 + last_u1 (colourspace.c:3461)
          const uint32_t bq = __byte_perm(Vc.b, Vp.a, selB);           // last_v1 = this_v2 (:3544), except at column 0
          const uint32_t v1 = Vp.a + bq;
          const uint32_t v2 = Vc.a + (s_vf[p + 1] & 0xFFu) * 0x10001u;  // last_v2 is never advanced: the row's first V sample
          QVL_up = v1 * 2u + (v2 & F4_MSK) + F4_K3;
          QVL_lo = (v1 & F4_MSK) + v2 * 2u + F4_K3;
        } else {
          const uint32_t LUc = Uc.a + Uc.b, LVp = Vp.a + Vp.b, LVc = Vc.a + Vc.b;
          QUL_up = LUp * 2u + (LUc & F4_MSK) + F4_K3; QUL_lo = (LUp & F4_MSK) + LUc * 2u + F4_K3;
          QVL_up = LVp * 2u + (LVc & F4_MSK) + F4_K3; QVL_lo = (LVp & F4_MSK) + LVc * 2u + F4_K3;
        }
        int rA[12], rB[12];
#pragma unroll
        for (int col = 0; col < 4; col++) {
          const bool hi_half = col >> 1, right = col & 1;
          const uint32_t qu_up = right ? QUR_up : QUL_up, qu_lo = right ? QUR_lo : QUL_lo;
          const uint32_t qv_up = right ? QVR_up : QVL_up, qv_lo = right ? QVR_lo : QVL_lo;
          const uint32_t mu_up = hi_half ? f4_idx_hi(qu_up) : f4_idx_lo(qu_up), mv_up = hi_half ? f4_idx_hi(qv_up) : f4_idx_lo(qv_up);
          const uint32_t mv_lo = hi_half ? f4_idx_hi(qv_lo) : f4_idx_lo(qv_lo);
          const uint32_t mu_lo = (QUIRKS && !right) ? mu_up : (hi_half ? f4_idx_hi(qu_lo) : f4_idx_lo(qu_lo));
          rgb(byte_of(yA, col), mu_up, mv_up, rA[3 * col], rA[3 * col + 1], rA[3 * col + 2]);
          rgb(byte_of(yB, col), mu_lo, mv_lo, rB[3 * col], rB[3 * col + 1], rB[3 * col + 2]);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {  // bytes = columns 0 .. 3 of channel c
          oA[c] = f4_pack_sat(rA[3 + c], rA[c], f4_pack_sat(rA[9 + c], rA[6 + c], 0u));
          oB[c] = f4_pack_sat(rB[3 + c], rB[c], f4_pack_sat(rB[9 + c], rB[6 + c], 0u));
        }
      } else {
        // ---- slow unit (frame edges): per pixel with the reference's edge rules, straight from global memory
        oA[0] = oA[1] = oA[2] = oB[0] = oB[1] = oB[2] = 0u;
        const uint8_t *Fy = P.y[g.f], *Fu = P.u[g.f], *Fv = P.v[g.f];
#pragma unroll 1
        for (int rr = 0; rr < 2; rr++) {
          const int sy = 2 * k - 1 + rr;
          if (sy < 0 || sy >= P.fh) continue;
          uint32_t o3[3] = {0u, 0u, 0u};
#pragma unroll 1
          for (int col = 0; col < 4; col++) {
            int u, v, r, gg, b;
            f4_chroma_px<QUIRKS>(Fu, Fv, P.rs_u, P.rs_v, P.cw, P.ch, P.fh, x + col, sy, u, v);
            rgb(__ldg(Fy + (size_t)P.rs_y * sy + x + col), (uint32_t)u * 128u, (uint32_t)v * 128u, r, gg, b);
            o3[0] |= (uint32_t)min(max(r, 0), 255) << (8 * col);
            o3[1] |= (uint32_t)min(max(gg, 0), 255) << (8 * col);
            o3[2] |= (uint32_t)min(max(b, 0), 255) << (8 * col);
          }
          if (rr == 0) { oA[0] = o3[0]; oA[1] = o3[1]; oA[2] = o3[2]; }
          else { oB[0] = o3[0]; oB[1] = o3[1]; oB[2] = o3[2]; }
        }
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        uint8_t *pl = s_pl + c * F4_PLP + 4 * gi;
        *reinterpret_cast<uint32_t *>(pl + (2 * p) * F4_PLS) = oA[c];
        *reinterpret_cast<uint32_t *>(pl + (2 * p + 1) * F4_PLS) = oB[c];
      }
    }
    __syncthreads();
#ifdef PE_F4_TIMELINE
    tl_lap(1);
#endif
    // the raw buffers are free again: bring in the next tile's words while passes 2 and 3 run
    const int tn = t + (int)gridDim.x;
    if (tn < total) {
      G = geometry(tn);
      stage(G, buf ^ 1);
    }
#ifdef PE_F4_TIMELINE
    tl_lap(2);
#endif
    // ---- 2. horizontal pass: warp = one source row at a time, lane = output columns lane, lane + 32, lane + 64, lane + 96
    {
      uint2 cf[4];
      int fo[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int xo = lane + 32 * j;
        const bool ok = xo < g.ncol;
        const int4 e = ok ? s_px[xo] : make_int4(0, 0, g.cb, 0);
        cf[j] = make_uint2((uint32_t)e.x, (uint32_t)e.y);
        fo[j] = e.z - g.cb;
      }
      for (int tr = warp; tr < g.nvr; tr += F4_NW) {
        const int pr = g.vr0 + tr - rowbase;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const uint8_t *row = s_pl + c * F4_PLP + pr * F4_PLS;
          uint16_t *trow = s_tmp + c * F4_TMP + tr * F4_TW;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(row) + (fo[j] >> 2);
            const uint32_t win = __funnelshift_r(q[0], q[1], 8u * (uint32_t)(fo[j] & 3));
            const uint32_t v = min(f4_dp2a_hi(cf[j].y, win, f4_dp2a_lo(cf[j].x, win, 0u)) >> 7, 32767u);
            trow[lane + 32 * j] = (uint16_t)v;
          }
        }
      }
    }
    __syncthreads();
#ifdef PE_F4_TIMELINE
    tl_lap(3);
#endif
    // ---- 3. vertical pass: warp = one output row at a time
    for (int yo = warp; yo < g.nrow; yo += F4_NW) {
      const int4 ey = s_py[yo];
      const int4 cy = make_int4(ey.x & 0xFFFF, (int)((uint32_t)ey.x >> 16), ey.y & 0xFFFF, (int)((uint32_t)ey.y >> 16));
      const int tr0 = ey.z - g.vr0, sy = ey.w;
      uint8_t *drow = P.dst[g.f] + (size_t)P.drs * (g.y0 + yo) + (size_t)g.x0 * 4;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int xo = lane + 32 * j;
        if (xo >= g.ncol) continue;
        int acc[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const uint16_t *tp = s_tmp + c * F4_TMP + tr0 * F4_TW + xo;
          acc[c] = ((1 << 18) + cy.x * (int)tp[0] + cy.y * (int)tp[F4_TW] + cy.z * (int)tp[2 * F4_TW] + cy.w * (int)tp[3 * F4_TW]) >> 19;
        }
        const int al = ((1 << 18) + sy * s_px[xo].w) >> 19;
        const int c0 = P.swap_rb ? acc[2] : acc[0], c2 = P.swap_rb ? acc[0] : acc[2];
        // bytes c0, G, c2, A
        const uint32_t px = f4_pack_sat(acc[1], c0, 0u) | (f4_pack_sat(al, c2, 0u) << 16);
        st_stream_u32(drow + 4 * xo, px);
      }
    }
#ifdef PE_F4_TIMELINE
    tl_lap(4);   // (thread 0's own share of the vertical pass; the barrier wait of the next tile lands in lap 0)
#endif
  }
#ifdef PE_F4_TIMELINE
  if (tid == 0) {
    for (int k = 0; k < 5; k++) P.tl[(size_t)blockIdx.x * 6 + k] = (unsigned long long)tl_acc[k];
    P.tl[(size_t)blockIdx.x * 6 + 5] = (unsigned long long)(clock64() - tl_begin);
  }
#endif
}

}  // namespace

// Can k_cvt_resize take this (conversion, resize) pair?  a: the queued planar -> RGB conversion (already yuv_planar_fast_ok), its
// destination the 4-byte packed frame the resize reads.
bool cvt_resize_supported(const YuvToRgbArgs &a, int dw, int dh, int drs, const uint8_t *dst, const ResizeFilter &hx, const ResizeFilter &hy) {
  if (a.is_422 || a.low_quality || a.lut16 || a.blend2) return false;
  if (a.out.psize != 4 || a.out.a != 3 || a.out.g != 1) return false;   // RGBA32 / BGRA32
  if ((a.width & 3) || a.width < 8 || a.height < 4 || (a.height & 1)) return false;
  if (hx.taps > 4 || hy.taps > 4 || !hx.nonneg() || !hy.nonneg()) return false;
  if ((drs & 3) || (reinterpret_cast<uintptr_t>(dst) & 3) || dw < 1 || dh < 1) return false;
  // every tap inside the frame (the libswscale recipes fold the border taps in; the round-1 triangle contract clamps instead)
  for (int i = 0; i < dw; i++)
    if (hx.first[i] < 0 || hx.first[i] + hx.taps > a.width) return false;
  for (int i = 0; i < dh; i++)
    if (hy.first[i] < 0 || hy.first[i] + hy.taps > a.height) return false;
  for (int i = 1; i < dw; i++)
    if (hx.first[i] < hx.first[i - 1]) return false;
  for (int i = 1; i < dh; i++)
    if (hy.first[i] < hy.first[i - 1]) return false;
  return true;
}

// frames: n conversions of the same shape (yuv_planar_same_shape), dsts[i] the resized destination of frames[i]
cudaError_t launch_cvt_resize(const Launch &L, const YuvToRgbArgs *frames, uint8_t *const *dsts, int n, int dw, int dh, int drs, const void *px_dev,
                              const void *py_dev, const ResizeFilter &hx, const ResizeFilter &hy) {
  const YuvToRgbArgs &a0 = frames[0];
  CvtRszParams P;
  memset(&P, 0, sizeof(P));
  P.fw = a0.width; P.fh = a0.height; P.cw = a0.src.cw; P.ch = a0.src.ch;
  P.rs_y = a0.src.rs_y; P.rs_u = a0.src.rs_u; P.rs_v = a0.src.rs_v;
  P.dw = dw; P.dh = dh; P.drs = drs;
  P.tiles_x = (dw + F4_TW - 1) / F4_TW; P.tiles_y = (dh + F4_TH - 1) / F4_TH;
  P.px = reinterpret_cast<const int4 *>(px_dev); P.py = reinterpret_cast<const int4 *>(py_dev);
  P.fx_taps = hx.taps; P.fy_taps = hy.taps;
  if (P.tiles_x > F4_MAXTX || P.tiles_y > F4_MAXTY) return cudaErrorInvalidConfiguration;
  P.conv = a0.conv.t;
  P.swap_rb = a0.out.r == 2;
  const bool last_row_unsafe = a0.src.rs_u < a0.src.cw + 4 || a0.src.rs_v < a0.src.cw + 4;
  P.k_fast_max = a0.src.ch - 1 - (last_row_unsafe ? 1 : 0);
  // per-tile maxima against the kernel's compile-time buffers
  for (int x0 = 0; x0 < dw; x0 += F4_TW) {
    const int x1 = (x0 + F4_TW < dw ? x0 + F4_TW : dw) - 1;
    const int vc0 = hx.first[x0], vc1 = hx.first[x1] + hx.taps - 1;
    if (((((vc1 | 3) + 1) - (vc0 & ~3)) >> 2) > F4_NG) return cudaErrorInvalidConfiguration;  // the caller runs the unfused pair
  }
  for (int y0 = 0; y0 < dh; y0 += F4_TH) {
    const int y1 = (y0 + F4_TH < dh ? y0 + F4_TH : dh) - 1;
    const int vr0 = hy.first[y0], vr1 = hy.first[y1] + hy.taps - 1;
    if (((vr1 + 1) >> 1) - ((vr0 + 1) >> 1) + 1 > F4_PR || vr1 - vr0 + 1 > F4_TR) return cudaErrorInvalidConfiguration;
  }
  if ((long long)P.tiles_x * P.tiles_y * F4_MAXF >= (1ll << 31)) return cudaErrorInvalidConfiguration;
  {
    // t / d == umulhi(t, ceil(2^32 / d)) for every t with t * d < 2^32 (d > 1): true for every tile index of the launch or not used
    const long long tpf = (long long)P.tiles_x * P.tiles_y, tmax = tpf * F4_MAXF;
    P.magic_tpf = (tpf > 1 && tmax * tpf < (1ll << 32)) ? (uint32_t)(0xFFFFFFFFu / (uint32_t)tpf + 1u) : 0u;
    P.magic_tx = (P.tiles_x > 1 && tpf * P.tiles_x < (1ll << 32)) ? (uint32_t)(0xFFFFFFFFu / (uint32_t)P.tiles_x + 1u) : 0u;
  }
  P.vec16 = !((a0.src.rs_y | a0.src.rs_u | a0.src.rs_v) & 15);
  for (int i = 0; i < n && P.vec16; i++)
    P.vec16 = !((reinterpret_cast<uintptr_t>(frames[i].src.y) | reinterpret_cast<uintptr_t>(frames[i].src.u) | reinterpret_cast<uintptr_t>(frames[i].src.v)) & 15);
  if (getenv("PE_F4_NOVEC")) P.vec16 = 0;
  // How a tile's raw rows reach shared memory (PE_F4_TMA; profiles/r02zg_k_cvt_resize_staging.txt, config 2):
  //   2 (default)  three cp.async.bulk.tensor.2d (TMA, one per plane, from 2-D tensor maps of the launch's planes) + two bulk copies
  //                of the filter rows, issued by ONE lane: 105.4 k frames/s.  Tiles at the right edge of a plane whose stride equals
  //                its width keep the cp.async path (box_staged()).
  //   0            16-byte cp.async (LDGSTS) from all warps: 100.2 k
  //   1            one cp.async.bulk per row (<= 224 bytes) from one warp: 67 - 69 k -- far below the size a bulk copy needs to pay off
  P.tma = P.vec16 ? (getenv("PE_F4_TMA") ? atoi(getenv("PE_F4_TMA")) : 2) : 0;
  if (P.tma < 0 || P.tma > 2) P.tma = 0;
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static bool encode_tried = false;
  if (P.tma == 2 && !encode_tried) {
    encode_tried = true;
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) encode = (EncodeFn)fp;
  }
  // (the boxes are the kernel's compile-time maxima: they must fit the planes' rows)
  if (P.tma == 2 && (!encode || a0.src.rs_y < F4_GW * 4 || a0.src.rs_u < F4_CWB || a0.src.rs_v < F4_CWB)) P.tma = 0;   // (narrow planes: the cp.async path)
  CvtRszMaps maps;   // (12 KB of kernel parameters, copied at launch)
  memset(&maps, 0, sizeof(maps));
  auto make_map = [&](CUtensorMap *m, const uint8_t *base, int rs, int rows, int box_w, int box_h) -> bool {
    const cuuint64_t gdim[2] = {(cuuint64_t)rs, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)rs};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, estr[2] = {1u, 1u};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  const int smem_bytes = F4_SMEM;
  static PerDevice attr_set;
  if (!attr_set.cur()) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_cvt_resize<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM_MAX)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_cvt_resize<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM_MAX)) != cudaSuccess) return e;
    attr_set.cur() = 1;
  }
  for (int base = 0; base < n; base += F4_MAXF) {
    P.nframes = n - base < F4_MAXF ? n - base : F4_MAXF;
    for (int i = 0; i < F4_MAXF; i++) {
      const int k = base + (i < P.nframes ? i : 0);
      P.y[i] = frames[k].src.y; P.u[i] = frames[k].src.u; P.v[i] = frames[k].src.v;
      P.dst[i] = dsts[k];
    }
    if (P.tma == 2) {
      bool ok = true;
      for (int i = 0; i < P.nframes && ok; i++)
        ok = make_map(&maps.y[i], P.y[i], P.rs_y, P.fh, F4_GW * 4, 2 * F4_PR) && make_map(&maps.u[i], P.u[i], P.rs_u, P.ch, F4_CWB, F4_PR + 1) &&
             make_map(&maps.v[i], P.v[i], P.rs_v, P.ch, F4_CWB, F4_PR + 1);
      if (!ok) P.tma = 0;
    }
    const long long total = (long long)P.tiles_x * P.tiles_y * P.nframes;
    const int grid = (int)(total < L.sm_count ? total : L.sm_count);
#ifdef PE_F4_TIMELINE
    static unsigned long long *tl_dev = nullptr;
    if (!tl_dev) cudaMalloc(&tl_dev, sizeof(unsigned long long) * 6 * 1024);
    cudaMemsetAsync(tl_dev, 0, sizeof(unsigned long long) * 6 * 1024, L.stream);
    P.tl = tl_dev;
#endif
    if (a0.quirks) k_cvt_resize<true><<<grid, F4_NT, smem_bytes, L.stream>>>(P, maps);
    else k_cvt_resize<false><<<grid, F4_NT, smem_bytes, L.stream>>>(P, maps);
#ifdef PE_F4_TIMELINE
    if (getenv("PE_F4_TIMELINE_DUMP")) {   // where the tiles' time goes (thread 0 of every CTA: its waits at the barriers land in the lap before them)
      static unsigned long long h[6 * 1024];
      cudaStreamSynchronize(L.stream);
      cudaMemcpy(h, tl_dev, sizeof(h), cudaMemcpyDeviceToHost);
      double acc[6] = {0, 0, 0, 0, 0, 0};
      for (int b = 0; b < grid; b++) for (int k = 0; k < 6; k++) acc[k] += (double)h[b * 6 + k];
      fprintf(stderr, "f4 timeline (thread 0 of %d CTAs, cycles per CTA): wait %.0f conversion %.0f staging %.0f horizontal %.0f vertical %.0f total %.0f | tiles %lld\n",
              grid, acc[0] / grid, acc[1] / grid, acc[2] / grid, acc[3] / grid, acc[4] / grid, acc[5] / grid, total);
    }
#endif
    PE_COUNT_LAUNCH(L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace pe
