// pe_kernels_fused2.cu -- the fast fused chain for the case the headline workload is in:
//     planar 4:2:x fg -> RGBA  |  letterbox with NO horizontal scaling (inner_w == fg width), vertical filter of <= 4 taps
//     |  scalar alpha-over a RGBA32 bg  |  optional 8-bit gamma LUT          (everything else: k_fused, pe_kernels_fused.cu)
//
// Same arithmetic, bit for bit, as k_fused and as the unfused ops.  What differs is the data path:
//   * persistent CTAs (256 threads, 2 per SM) walk over output tiles of 128 x tile_h pixels.  The raw Y / U / V bytes a tile
//     needs are staged in shared memory by 16-byte cp.async copies, DOUBLE BUFFERED: the copies for tile i+1 are in flight
//     while tile i is converted, and the bg lines of tile i+1 are prefetched into L2 (frames whose planes are not 16-byte
//     aligned take a synchronous staging path with the same layout);
//   * the chroma halo columns carry the reference's edge semantics (patched in after the copy on frame-edge tiles), so the
//     arithmetic below never looks at a frame edge;
//   * conversion works on 4 x 4 pixel units (two of the reference's row pairs): the chroma sums of colourspace.c:3440-3549
//     are shared inside the unit; the chroma tables are indexed by the UN-divided chroma sum (the /3 rounding and the
//     CLAMP16_240 of the reference are folded into the table); the converted bytes are written PLANAR and COLUMN-MAJOR --
//     one 32-bit word = one channel of 4 vertically adjacent pixels (cvt.pack.sat) -- so that
//   * the vertical filter is two DP2A instructions per channel and pixel (16-bit coefficient pairs x 4 byte taps), fed by a
//     funnel shift over two such words; (sum c12 * pix + 2^11) >> 12 equals the two-pass contract of k_resize_h / k_resize_v
//     exactly when the horizontal pass is the identity (pix * 16384 >> 7 = pix * 128);
//   * alpha = k / 256 is blended in integers ((bg * (256 - k) + fg * k) >> 8 is exactly trunc(bg * (1 - a) + fg * a) in double
//     for such alpha: every product and the sum are exact); any other alpha goes through the 64 KB [bg][fg] table in shared
//     memory, as in k_fused.
// Bank-conflict notes: converted tile column stride is 17 words (odd), stage-1 lanes are laid out 8 column-quads x 4 row-quads.
#include <cstdlib>

#include "pe_device.cuh"
#include "pe_kernels.h"

namespace pe {

namespace {

#define PE_COUNT_LAUNCH(L) do { if ((L).launch_counter) ++*(L).launch_counter; } while (0)

constexpr int F2_NT = 256;
constexpr int F2_NW = F2_NT / 32;
constexpr int F2_TW = 128;                 // output tile width
constexpr int F2_MAXVR = 60;               // max virtual source rows (first .. first + 3 of the last row) per 4:2:0 tile
constexpr int F2_MAXVR_422 = 28;           // ... per 4:2:2 tile (one chroma row per luma row: bounded by F2_CROWS)
constexpr int F2_CW = 17;                  // words per column of the converted tile (68 rows)
constexpr int F2_CCOLS = F2_TW + 4;        // columns of the converted tile (alignment slack of one quad)
constexpr int F2_YROWS = 68;               // raw luma rows
constexpr int F2_YRS = 144;                // raw luma row stride in bytes (9 x 16-byte chunks)
constexpr int F2_CROWS = 36;               // raw chroma rows
constexpr int F2_CRS = 112;                // raw chroma row stride in bytes: 16 bytes of slack (left halo at a frame edge) + 6 chunks
constexpr int F2_MAXTH = 48;               // max output rows per tile
constexpr int F2_NEXT = 768;               // entries of an extended chroma table (index n = u1 + (u2 >> 1) <= 765)
constexpr int F2_MAXF = 16;                // frames per launch (their descriptors travel as kernel parameters)

// Shared-memory tables: RGB_Y[256], then four chroma tables indexed by the UN-divided chroma sum n:
//   ext[t][n] = table_t[clamp(third_round(n), lo, hi)]        (third_round(n) = (int)(n / 3. + .5), colourspace.c:3465)
// A plain chroma sample m (single rows, PB_QUALITY_LOW) is looked up at n = 3 * m (third_round(3 m) == m).
constexpr int OFF_TAB = 0;                                   // int32 [256 + 4 * F2_NEXT]
constexpr int OFF_LUT = OFF_TAB + (256 + 4 * F2_NEXT) * 4;   // u8 [256]
constexpr int OFF_ROW = OFF_LUT + 256;                       // int4 [F2_MAXTH]: pos, a0, a1, -
constexpr int OFF_VF = OFF_ROW + F2_MAXTH * 16;              // u32 [2][F2_CROWS]: first word of every staged V row
constexpr int OFF_Y = OFF_VF + 2 * F2_CROWS * 4;
constexpr int OFF_U = OFF_Y + 2 * F2_YROWS * F2_YRS;
constexpr int OFF_V = OFF_U + 2 * F2_CROWS * F2_CRS;
constexpr int OFF_C = OFF_V + 2 * F2_CROWS * F2_CRS;         // u32 [3][F2_CCOLS][F2_CW]
constexpr int OFF_OVER = OFF_C + 3 * F2_CCOLS * F2_CW * 4;   // u8 [65536] (table blend only)
constexpr int F2_SMEM_ARITH = OFF_OVER;
constexpr int F2_SMEM_TABLE = OFF_OVER + 65536;
static_assert(OFF_ROW % 16 == 0 && OFF_VF % 16 == 0 && OFF_Y % 16 == 0 && OFF_U % 16 == 0 && OFF_V % 16 == 0 && OFF_C % 16 == 0 &&
              OFF_OVER % 16 == 0, "16-byte alignment of the cp.async destinations");

struct Fused2Params {
  FusedArgs fr[F2_MAXF];
  int nframes, tiles_x, tiles_y, tile_h;
  int blend_a, blend_ia;     // arithmetic blend: weights of fg / bg, sum 256
  const uint8_t *lut8;       // optional gamma LUT applied after the blend (nullptr: none)
};

__device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }

// chroma sample with the reference's one-past-row read (see k_yuv_planar_to_rgb, pe_kernels_yuv.cu)
__device__ __forceinline__ uint32_t chroma_edge(const uint8_t *__restrict__ p, int stride, int r, int c, int cw, int ch) {
  if (c >= cw) c = (cw < stride || r + 1 < ch) ? cw : cw - 1;
  return p[(size_t)stride * r + c];
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
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// quad index of a luma row: 4:2:0 groups the reference's row pairs (1,2)(3,4) | (5,6)(7,8) ..., row 0 sits alone in quad 0
__device__ __forceinline__ int quad_of(int row, int is422) { return is422 ? (row >> 2) : ((row + 3) >> 2); }
__device__ __forceinline__ int quad_first_row(int g, int is422) { return is422 ? 4 * g : 4 * g - 3; }

// yuv2rgb_int / xyuv2rgb (colourspace.c:2345-2356) through the shared-memory tables.  nu / nv index the extended chroma
// tables; results are the UNSATURATED (sum >> 16) values, saturated when they are packed (pack_sat).
__device__ __forceinline__ void px_rgb(const int32_t *__restrict__ t, int y, int nu, int nv, int &r, int &g, int &b) {
  const int yy = t[y];
  r = (yy + t[256 + nv]) >> 16;
  g = (yy + t[256 + F2_NEXT + nu] + t[256 + 2 * F2_NEXT + nv]) >> 16;
  b = (yy + t[256 + 3 * F2_NEXT + nu]) >> 16;
}

// geometry of one output tile
struct TileGeo {
  int f, x0, y0, x1, y1, ix0, ix1, iy0, iy1, has_inner;
  int vr0, vr1, base_g, ga, gb, cq0, ncq, yrow0, nyr, crow0, ncr;
  int ybase, cbase;  // column (byte) origins of the staged luma / chroma rows
};

template <bool ASYNC>
__device__ __forceinline__ TileGeo tile_geo(const Fused2Params &P, uint32_t tile) {
  TileGeo G;
  const uint32_t tiles_per_frame = (uint32_t)(P.tiles_x * P.tiles_y);
  const uint32_t fi = tile / tiles_per_frame, t = tile - fi * tiles_per_frame;  // 32-bit: a launch has far fewer than 2^31 tiles
  G.f = (int)fi;
  const FusedArgs &A = P.fr[G.f];
  const int ty = (int)(t / (uint32_t)P.tiles_x), tx = (int)t - ty * P.tiles_x;
  G.x0 = tx * F2_TW; G.y0 = ty * P.tile_h;
  G.x1 = min(G.x0 + F2_TW, A.ow); G.y1 = min(G.y0 + P.tile_h, A.oh);
  // intersection with the inner rectangle, in inner (= source column) coordinates
  G.ix0 = max(G.x0 - A.ox, 0); G.ix1 = min(G.x1 - A.ox, A.iw);
  G.iy0 = max(G.y0 - A.oy, 0); G.iy1 = min(G.y1 - A.oy, A.ih);
  G.has_inner = G.ix0 < G.ix1 && G.iy0 < G.iy1;
  G.vr0 = G.vr1 = G.base_g = G.ga = G.gb = G.cq0 = G.ncq = G.yrow0 = G.nyr = G.crow0 = G.ncr = G.ybase = G.cbase = 0;
  if (G.has_inner) {
    const int is422 = A.is_422, fh = A.fh;
    G.vr0 = A.fy.first[G.iy0]; G.vr1 = A.fy.first[G.iy1 - 1] + 3;      // virtual source rows of the tile
    const int ar0 = min(max(G.vr0, 0), fh - 1), ar1 = min(max(G.vr1, 0), fh - 1);
    G.base_g = quad_of(G.vr0, is422);
    G.ga = quad_of(ar0, is422); G.gb = quad_of(ar1, is422);             // quads that hold real rows
    G.cq0 = G.ix0 >> 2;
    G.ncq = ((G.ix1 - 1) >> 2) - G.cq0 + 1;
    G.yrow0 = max(quad_first_row(G.ga, is422), 0);
    G.nyr = min(quad_first_row(G.gb, is422) + 3, fh - 1) - G.yrow0 + 1;
    G.crow0 = is422 ? G.yrow0 : max(2 * G.ga - 2, 0);
    G.ncr = (is422 ? G.yrow0 + G.nyr - 1 : min(2 * G.gb, A.fg.ch - 1)) - G.crow0 + 1;
    if (ASYNC) {
      G.ybase = (4 * G.cq0) & ~15;
      G.cbase = max(2 * G.cq0 - 1, 0) & ~15;
    } else {
      G.ybase = 4 * G.cq0;
      G.cbase = 2 * G.cq0 - 2;  // left halo at byte 17, interior from byte 18 (even: 16-bit stores)
    }
  }
  return G;
}

// issue the 16-byte async copies of the raw planes of a tile into buffer `buf` (and prefetch its bg lines into L2)
__device__ __forceinline__ void issue_loads(const Fused2Params &P, const TileGeo &G, int buf, uint8_t *smem, int tid) {
  const FusedArgs &A = P.fr[G.f];
  // bg lines of the tile -> L2
  {
    const int th = G.y1 - G.y0;
    const int row = tid >> 2, line = tid & 3;
    if (row < th && G.x0 + 32 * line < A.ow) prefetch_l2(A.bg.p + (size_t)A.bg.rs * (G.y0 + row) + 4 * (size_t)(G.x0 + 32 * line));
  }
  if (!G.has_inner) return;
  {
    uint8_t *sy = smem + OFF_Y + buf * (F2_YROWS * F2_YRS);
    const int nch = (4 * (G.cq0 + G.ncq) - G.ybase + 15) >> 4;  // <= 9
    const int c = tid & 15;
    if (c < nch) {
      const uint8_t *src = A.fg.y + (size_t)A.fg.rs_y * G.yrow0 + G.ybase + 16 * c;
      for (int r = tid >> 4; r < G.nyr; r += F2_NT / 16) cp_async16(sy + r * F2_YRS + 16 * c, src + (size_t)A.fg.rs_y * r);
    }
  }
  {
    // chroma columns 2*cq0-1 .. 2*cq1+2, clipped to the row stride (a column beyond it is patched in afterwards)
    const int clast = 2 * (G.cq0 + G.ncq - 1) + 2;
    const int c = tid & 7;
#pragma unroll
    for (int plane = 0; plane < 2; plane++) {
      const int rs = plane ? A.fg.rs_v : A.fg.rs_u;
      const int cend = min((clast + 16) & ~15, rs);
      const int nch = (cend - G.cbase) >> 4;  // <= 6
      uint8_t *sc = smem + (plane ? OFF_V : OFF_U) + buf * (F2_CROWS * F2_CRS);
      if (c < nch) {
        const uint8_t *src = (plane ? A.fg.v : A.fg.u) + (size_t)rs * G.crow0 + G.cbase + 16 * c;
        for (int r = tid >> 3; r < G.ncr; r += F2_NT / 8) cp_async16(sc + r * F2_CRS + 16 + 16 * c, src + (size_t)rs * r);
      }
    }
    if (tid < G.ncr)
      cp_async4(smem + OFF_VF + (buf * F2_CROWS + tid) * 4, A.fg.v + (size_t)A.fg.rs_v * (G.crow0 + tid));
  }
}

template <int MODE, bool ASYNC>  // MODE 0: arithmetic blend (alpha = k / 256), 1: [bg][fg] table blend
__global__ void __launch_bounds__(F2_NT, MODE == 0 ? 2 : 1) k_fused2(const __grid_constant__ Fused2Params P) {
  extern __shared__ __align__(16) uint8_t smem[];
  int32_t *s_tab = reinterpret_cast<int32_t *>(smem + OFF_TAB);
  uint8_t *s_lut = smem + OFF_LUT;
  int4 *s_row = reinterpret_cast<int4 *>(smem + OFF_ROW);
  uint32_t *s_c = reinterpret_cast<uint32_t *>(smem + OFF_C);
  uint8_t *s_over = smem + OFF_OVER;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t *cur_conv = nullptr;
  int cur_clamped = -1;
  const uint8_t *cur_over = nullptr;
  const bool has_lut = P.lut8 != nullptr;
  if (has_lut) for (int i = tid; i < 256; i += F2_NT) s_lut[i] = P.lut8[i];

  const uint32_t total_tiles = (uint32_t)(P.tiles_x * P.tiles_y * P.nframes);
  uint32_t tile = blockIdx.x;
  if (tile >= total_tiles) return;
  TileGeo G = tile_geo<ASYNC>(P, tile);
  int buf = 0;
  if (ASYNC) {
    issue_loads(P, G, 0, smem, tid);
    cp_async_commit();
  }

  for (; tile < total_tiles; tile += gridDim.x, buf ^= 1) {
    const FusedArgs &A = P.fr[G.f];
    const int is422 = A.is_422, fh = A.fh;
    uint8_t *s_y = smem + OFF_Y + (ASYNC ? buf : 0) * (F2_YROWS * F2_YRS);
    uint8_t *s_u = smem + OFF_U + (ASYNC ? buf : 0) * (F2_CROWS * F2_CRS);
    uint8_t *s_v = smem + OFF_V + (ASYNC ? buf : 0) * (F2_CROWS * F2_CRS);
    uint32_t *s_vf = reinterpret_cast<uint32_t *>(smem + OFF_VF) + (ASYNC ? buf : 0) * F2_CROWS;

    // ---- (re)build the tables when the conversion variant changes (first tile, or frames of different clamping);
    //      the previous tile's readers are past the barrier at the end of the loop body
    bool rebuilt = false;
    if (A.conv.t != cur_conv || A.clamped != cur_clamped) {
      rebuilt = true;
      static_assert(F2_NEXT == kExtN, "extended table size");
      for (int i = tid; i < 256 + 4 * F2_NEXT; i += F2_NT) s_tab[i] = A.conv.ext[i];
      cur_conv = A.conv.t;
      cur_clamped = A.clamped;
    }
    if (MODE == 1 && A.over_table != cur_over) {
      for (int i = tid; i < 4096; i += F2_NT) reinterpret_cast<uint4 *>(s_over)[i] = reinterpret_cast<const uint4 *>(A.over_table)[i];
      cur_over = A.over_table;
      rebuilt = true;
    }
    if (rebuilt) __syncthreads();  // block-uniform; a tile without inner rows goes straight to stage 3

    // ---- prefetch: the copies of the NEXT tile go out before this tile is touched
    TileGeo GN = G;
    const bool have_next = tile + gridDim.x < total_tiles;
    if (have_next) GN = tile_geo<ASYNC>(P, tile + gridDim.x);
    if (ASYNC) {
      if (have_next) issue_loads(P, GN, buf ^ 1, smem, tid);
      cp_async_commit();
      cp_async_wait<1>();  // everything but the group just committed has landed: this tile's planes
    }

    const int base_g = G.base_g, cq0 = G.cq0, ncq = G.ncq, crow0 = G.crow0, ncr = G.ncr, yrow0 = G.yrow0;
    const int iy0 = G.iy0, iy1 = G.iy1;
    if (G.has_inner) {
      const int cw = A.fg.cw, ch = A.fg.ch;
      const int cq1 = cq0 + ncq - 1;
      if (!ASYNC) {
        // ---- synchronous staging (unaligned planes): luma words, chroma 16-bit pairs + halo bytes
        const uint8_t *src = A.fg.y + (size_t)A.fg.rs_y * yrow0 + 4 * cq0;
        for (int w = lane; w < ncq; w += 32)
          for (int r = warp; r < G.nyr; r += F2_NW)
            *reinterpret_cast<uint32_t *>(s_y + r * F2_YRS + 4 * w) = ld_stream_u32(src + (size_t)A.fg.rs_y * r + 4 * w);
        for (int rr = warp; rr < 2 * ncr; rr += F2_NW) {
          const int plane = rr >= ncr, r = plane ? rr - ncr : rr;
          const int rs = plane ? A.fg.rs_v : A.fg.rs_u;
          const uint8_t *srow = (plane ? A.fg.v : A.fg.u) + (size_t)rs * (crow0 + r);
          uint8_t *dst = (plane ? s_v : s_u) + r * F2_CRS + 16 - G.cbase;  // dst[c] = column c
          for (int w = lane; w <= ncq; w += 32) {
            const int c = 2 * (cq0 + w);
            if (w < ncq && c + 1 < cw) *reinterpret_cast<uint16_t *>(dst + c) = *reinterpret_cast<const uint16_t *>(srow + c);
            else if (c < cw) dst[c] = srow[c];  // last column / right halo inside the plane; beyond it: patched below
          }
          if (lane == 0) {
            if (2 * cq0 - 1 >= 0) dst[2 * cq0 - 1] = srow[2 * cq0 - 1];
            if (plane) s_vf[r] = srow[0];
          }
        }
      }
      __syncthreads();  // staged planes visible (cp.async data + the sync path's stores)

      // ---- frame-edge patches of the staged chroma (block-uniform conditions, few tiles):
      //      column -1 replicates column 0 (last = this at the start of a row); 4:2:2 with ref_quirks: columns <= 0 take
      //      column 0 of chroma row (r >> 1) (the seed slip, colourspace.c:3600); a column >= cw reads the byte at
      //      plane[stride * r + cw] -- padding or the first sample of the next row -- except on the last chroma row of a
      //      plane without padding, where it is the replicated edge sample (:3508-3512, DESIGN.md "edge read")
      {
        const bool seed = is422 && A.quirks;
        const bool left_edge = cq0 == 0;
        const bool right_edge = 2 * cq1 + 2 >= cw;
        if (left_edge || right_edge) {
          for (int rr = tid; rr < 2 * ncr; rr += F2_NT) {
            const int plane = rr >= ncr, r = plane ? rr - ncr : rr;
            const int cr = crow0 + r;
            const int rs = plane ? A.fg.rs_v : A.fg.rs_u;
            const uint8_t *pl = plane ? A.fg.v : A.fg.u;
            uint8_t *dst = (plane ? s_v : s_u) + r * F2_CRS + 16 - G.cbase;  // dst[c] = column c
            if (left_edge) {
              const uint8_t c0 = seed ? pl[(size_t)rs * (cr >> 1)] : pl[(size_t)rs * cr];
              dst[0] = c0;
              dst[-1] = c0;
            }
            if (right_edge)
              for (int c = cw; c <= 2 * cq1 + 2; c++) dst[c] = (uint8_t)chroma_edge(pl, rs, cr, c, cw, ch);
          }
          __syncthreads();
        }
      }

      // ---- per-row filter data: window position inside the converted tile, coefficient pairs (read in stage 3)
      for (int i = tid; i < iy1 - iy0; i += F2_NT) {
        const int iy = iy0 + i;
        const int16_t *c = A.fy.coef + iy * A.fy.taps;
        const int nt = A.fy.taps;
        const uint32_t c0 = (uint16_t)c[0], c1 = nt > 1 ? (uint16_t)c[1] : 0u, c2 = nt > 2 ? (uint16_t)c[2] : 0u,
                       c3 = nt > 3 ? (uint16_t)c[3] : 0u;
        s_row[i] = make_int4(A.fy.first[iy] + (is422 ? 0 : 3) - 4 * base_g, (int)(c0 | (c1 << 16)), (int)(c2 | (c3 << 16)), 0);
      }

      // ---- stage 1: convert 4 x 4 units.  Warp lanes: 8 column quads x 4 row quads.
      {
        const int ga = G.ga, ngr = G.gb - G.ga + 1;
        const int cq_groups = (ncq + 7) >> 3, rq_groups = (ngr + 3) >> 2;
        const int quirks = A.quirks;
        const int ycol0 = 4 * cq0 - G.ybase;       // byte offset of source column 4*cq0 in a staged luma row
        const int ccol0 = 16 - G.cbase;            // staged chroma byte index of column c is ccol0 + c
        for (int wt = warp; wt < cq_groups * rq_groups; wt += F2_NW) {
          const int rg = wt / cq_groups, cg = wt - rg * cq_groups;
          const int q = cg * 8 + (lane & 7), gl = rg * 4 + (lane >> 3);
          if (q >= ncq || gl >= ngr) continue;
          const int g = ga + gl;
          const int jc0 = 2 * (cq0 + q);  // absolute chroma column of the unit's first pair
          uint32_t *dst = s_c + (4 * q) * F2_CW + (g - base_g);
          // chroma words: columns jc0-1 .. jc0+2 of one staged chroma row
          const int cidx = ccol0 + jc0 - 1;
          auto cword = [&](const uint8_t *pl, int cr) -> uint32_t {
            const uint8_t *rowp = pl + (cr - crow0) * F2_CRS + (cidx & ~3);
            return __funnelshift_r(*reinterpret_cast<const uint32_t *>(rowp), *reinterpret_cast<const uint32_t *>(rowp + 4),
                                   8 * (cidx & 3));
          };
          const uint8_t *yp = s_y + ycol0 + 4 * q;
          const int row0 = quad_first_row(g, is422);
          const bool full420 = !is422 && !A.low_quality && row0 >= 1 && row0 + 3 <= fh - 1;
          if (full420) {
            // ===== fast path: both row pairs of the quad are interior pairs (colourspace.c:3440-3549)
            const int cA = 2 * g - 2 - crow0;  // staged chroma rows cA, cA + 1, cA + 2
            const uint32_t uw[3] = {cword(s_u, crow0 + cA), cword(s_u, crow0 + cA + 1), cword(s_u, crow0 + cA + 2)};
            const uint32_t vw[3] = {cword(s_v, crow0 + cA), cword(s_v, crow0 + cA + 1), cword(s_v, crow0 + cA + 2)};
            const int vf1 = s_vf[cA + 1] & 0xFF, vf2 = s_vf[cA + 2] & 0xFF;
            uint32_t yw[4];
#pragma unroll
            for (int r = 0; r < 4; r++) yw[r] = *reinterpret_cast<const uint32_t *>(yp + (row0 + r - yrow0) * F2_YRS);
#pragma unroll
            for (int p = 0; p < 2; p++) {
              // chroma samples of columns jc0+p-1, jc0+p, jc0+p+1 in the three rows
              int U[3][3], V[3][3];
#pragma unroll
              for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) { U[r][c] = byte_of(uw[r], p + c); V[r][c] = byte_of(vw[r], p + c); }
              // n indices (u1 + (u2 >> 1) upper, (u1 >> 1) + u2 lower) for the left / right pixel of both pairs
              int nu[2][4], nv[2][4];  // [left/right][row in quad]
#pragma unroll
              for (int pr = 0; pr < 2; pr++) {  // pair: chroma rows (pr, pr + 1)
                {  // right pixel: this + next
                  const int u1 = U[pr][1] + U[pr][2], u2 = U[pr + 1][1] + U[pr + 1][2];
                  const int v1 = V[pr][1] + V[pr][2], v2 = V[pr + 1][1] + V[pr + 1][2];
                  nu[1][2 * pr] = u1 + (u2 >> 1); nu[1][2 * pr + 1] = (u1 >> 1) + u2;
                  nv[1][2 * pr] = v1 + (v2 >> 1); nv[1][2 * pr + 1] = (v1 >> 1) + v2;
                }
                {  // left pixel: this + last, with the reference's slips under `quirks`
                  const int u1 = U[pr][1] + U[pr][0];
                  int u2 = U[pr + 1][1] + U[pr + 1][0];
                  int v1 = V[pr][1] + V[pr][0], v2 = V[pr + 1][1] + V[pr + 1][0];
                  if (quirks) {
                    u2 = u1;                                            // colourspace.c:3461
                    if (jc0 + p > 0) v1 = V[pr][1] + V[pr + 1][0];      // :3544
                    v2 = V[pr + 1][1] + (pr ? vf2 : vf1);               // last_v2 never advanced
                  }
                  nu[0][2 * pr] = u1 + (u2 >> 1); nu[0][2 * pr + 1] = (u1 >> 1) + u2;
                  nv[0][2 * pr] = v1 + (v2 >> 1); nv[0][2 * pr + 1] = (v1 >> 1) + v2;
                }
              }
#pragma unroll
              for (int lr = 0; lr < 2; lr++) {
                const int k = 2 * p + lr;
                int r[4], gg[4], b[4];
#pragma unroll
                for (int rw = 0; rw < 4; rw++) px_rgb(s_tab, byte_of(yw[rw], k), nu[lr][rw], nv[lr][rw], r[rw], gg[rw], b[rw]);
                dst[k * F2_CW] = pack_sat(r[1], r[0], pack_sat(r[3], r[2], 0u));
                dst[F2_CCOLS * F2_CW + k * F2_CW] = pack_sat(gg[1], gg[0], pack_sat(gg[3], gg[2], 0u));
                dst[2 * F2_CCOLS * F2_CW + k * F2_CW] = pack_sat(b[1], b[0], pack_sat(b[3], b[2], 0u));
              }
            }
            continue;
          }
          // ===== general path: 4:2:2, PB_QUALITY_LOW, and the quads that hold row 0 / the last rows of the frame
          uint32_t acc[4][3];
#pragma unroll
          for (int k = 0; k < 4; k++) acc[k][0] = acc[k][1] = acc[k][2] = 0;
          auto put = [&](int k, int bytepos, int r, int gg, int b) {
            acc[k][0] |= (uint32_t)sat8(r) << (8 * bytepos);
            acc[k][1] |= (uint32_t)sat8(gg) << (8 * bytepos);
            acc[k][2] |= (uint32_t)sat8(b) << (8 * bytepos);
          };
          // a single row: horizontal average only (row 0, an even frame's last row, every 4:2:2 row)
          auto do_single = [&](int row, int cr, int bytepos) {
            const uint32_t yw = *reinterpret_cast<const uint32_t *>(yp + (row - yrow0) * F2_YRS);
            const uint32_t uw = cword(s_u, cr), vw = cword(s_v, cr);
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int p = k >> 1;
              const int ua = byte_of(uw, p + 1), va = byte_of(vw, p + 1);
              const int ub = (k & 1) ? byte_of(uw, p + 2) : byte_of(uw, p), vb = (k & 1) ? byte_of(vw, p + 2) : byte_of(vw, p);
              int r, gg, b;
              px_rgb(s_tab, byte_of(yw, k), 3 * ((ua + ub) >> 1), 3 * ((va + vb) >> 1), r, gg, b);
              put(k, bytepos, r, gg, b);
            }
          };
          auto do_pair = [&](int row_a, int cr_a, int bytepos) {
            const int cr_b = cr_a + 1;
            const uint32_t ya = *reinterpret_cast<const uint32_t *>(yp + (row_a - yrow0) * F2_YRS);
            const uint32_t yb = *reinterpret_cast<const uint32_t *>(yp + (row_a + 1 - yrow0) * F2_YRS);
            const uint32_t u1w = cword(s_u, cr_a), u2w = cword(s_u, cr_b), v1w = cword(s_v, cr_a), v2w = cword(s_v, cr_b);
            const int v2_first = s_vf[cr_b - crow0] & 0xFF;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int p = k >> 1;
              int u1, u2, v1, v2;
              if (k & 1) {
                u1 = byte_of(u1w, p + 1) + byte_of(u1w, p + 2); u2 = byte_of(u2w, p + 1) + byte_of(u2w, p + 2);
                v1 = byte_of(v1w, p + 1) + byte_of(v1w, p + 2); v2 = byte_of(v2w, p + 1) + byte_of(v2w, p + 2);
              } else {
                u1 = byte_of(u1w, p + 1) + byte_of(u1w, p); u2 = byte_of(u2w, p + 1) + byte_of(u2w, p);
                v1 = byte_of(v1w, p + 1) + byte_of(v1w, p); v2 = byte_of(v2w, p + 1) + byte_of(v2w, p);
                if (quirks) {
                  u2 = u1;
                  if (jc0 + p > 0) v1 = byte_of(v1w, p + 1) + byte_of(v2w, p);
                  v2 = byte_of(v2w, p + 1) + v2_first;
                }
              }
              int n3u, n4u, n3v, n4v;
              if (!A.low_quality) {
                n3u = u1 + (u2 >> 1); n4u = (u1 >> 1) + u2; n3v = v1 + (v2 >> 1); n4v = (v1 >> 1) + v2;
              } else {  // PB_QUALITY_LOW: u3 = u1 >> 1, u4 = u2 >> 1 (:3470-3474)
                n3u = 3 * (u1 >> 1); n4u = 3 * (u2 >> 1); n3v = 3 * (v1 >> 1); n4v = 3 * (v2 >> 1);
              }
              int r, gg, b;
              px_rgb(s_tab, byte_of(ya, k), n3u, n3v, r, gg, b);
              put(k, bytepos, r, gg, b);
              px_rgb(s_tab, byte_of(yb, k), n4u, n4v, r, gg, b);
              put(k, bytepos + 1, r, gg, b);
            }
          };
          if (is422) {
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
              const int row = 4 * g + rr;
              if (row < fh) do_single(row, row, rr);
            }
          } else {
            // quad g holds rows 4g-3 .. 4g: pairs (4g-3, 4g-2) and (4g-1, 4g); chroma rows (2g-2, 2g-1) and (2g-1, 2g)
#pragma unroll
            for (int pp = 0; pp < 2; pp++) {
              const int row_a = 4 * g - 3 + 2 * pp, cr_a = 2 * g - 2 + pp;
              if (row_a + 1 == 0) do_single(0, 0, 2 * pp + 1);                       // row 0 (lower half of the "pair" -1, 0)
              else if (row_a >= 0 && row_a + 1 < fh) do_pair(row_a, cr_a, 2 * pp);
              else if (row_a == fh - 1 && row_a >= 0) do_single(row_a, ch - 1, 2 * pp);  // even height: last row alone
            }
          }
#pragma unroll
          for (int k = 0; k < 4; k++) {
            dst[k * F2_CW] = acc[k][0];
            dst[F2_CCOLS * F2_CW + k * F2_CW] = acc[k][1];
            dst[2 * F2_CCOLS * F2_CW + k * F2_CW] = acc[k][2];
          }
        }
      }
      __syncthreads();
      // ---- edge fix-up: virtual rows outside the frame replicate row 0 / row fh-1 (source indices are clamped in the contract)
      if (G.vr0 < 0 || G.vr1 > fh - 1) {
        const int shift = is422 ? 0 : 3;
        const int ncols = 4 * ncq;
        uint8_t *cb = reinterpret_cast<uint8_t *>(s_c);
        for (int chn = 0; chn < 3; chn++) {
          for (int col = tid; col < ncols; col += F2_NT) {
            uint8_t *colp = cb + ((chn * F2_CCOLS + col) * F2_CW) * 4;
            if (G.vr0 < 0) {
              const uint8_t v = colp[0 + shift - 4 * base_g];
              for (int r = G.vr0; r < 0; r++) colp[r + shift - 4 * base_g] = v;
            }
            if (G.vr1 > fh - 1) {
              const uint8_t v = colp[fh - 1 + shift - 4 * base_g];
              for (int r = fh; r <= G.vr1; r++) colp[r + shift - 4 * base_g] = v;
            }
          }
        }
        __syncthreads();
      }
    }

    // ---- stage 3: vertical filter + letterbox + alpha-over (+ gamma), one thread = one column x 4 consecutive rows.
    //      The four bg words of the NEXT task are loaded before the current one is computed (register double buffer).
    {
      const int x0 = G.x0, y0 = G.y0, y1 = G.y1;
      const int nquads = (y1 - y0 + 3) >> 2;
      const int ntasks = nquads * (F2_TW / 32);
      const int niy = iy1 - iy0;
      const uint32_t ka = (uint32_t)P.blend_a, kia = (uint32_t)P.blend_ia;
      auto blend_store = [&](uint8_t *outp, uint32_t b, uint32_t fr, uint32_t fg_, uint32_t fb) {
        uint32_t o0, o1, o2;
        if (MODE == 0) {
          // two channels per multiply: (bg * kia + fg * ka) for R | B in the 16-bit halves, G alone
          const uint32_t rb = (b & 0x00FF00FFu) * kia + (fr | (fb << 16)) * ka;
          const uint32_t gg = ((b >> 8) & 0xFFu) * kia + fg_ * ka;
          o0 = (rb >> 8) & 0xFFu; o1 = gg >> 8; o2 = rb >> 24;
          if (has_lut) { o0 = s_lut[o0]; o1 = s_lut[o1]; o2 = s_lut[o2]; }
        } else {  // the table already contains the gamma LUT (launch_over_table)
          o0 = s_over[((b & 0xFFu) << 8) | fr];
          o1 = s_over[(((b >> 8) & 0xFFu) << 8) | fg_];
          o2 = s_over[(((b >> 16) & 0xFFu) << 8) | fb];
        }
        st_stream_u32(outp, o0 | (o1 << 8) | (o2 << 16) | 0xFF000000u);
      };
      // row pointers advance by 32-bit strides (frames are far below 4 GB): no 64-bit multiplies per row.
      // Everything the task loop needs from the frame descriptor is pulled into registers once per tile.
      const uint32_t bg_rs32 = (uint32_t)A.bg.rs, out_rs32 = (uint32_t)A.out.rs;
      const uint8_t *const bg_base = A.bg.p;
      uint8_t *const out_base = A.out.p;
      const int a_ow = A.ow, a_ox = A.ox, a_oy = A.oy;
      const int g_ix0 = G.ix0, g_ix1 = G.has_inner ? G.ix1 : G.ix0;  // empty column range without inner rows
      auto load_bg = [&](int task, uint32_t w[4]) {
        const int qd = task >> 2, cwp = task & 3;
        const int x = x0 + cwp * 32 + lane;
        const int oy0 = y0 + qd * 4;
        const int nrow = (task < ntasks && x < a_ow) ? y1 - oy0 : 0;
        const uint8_t *p0 = bg_base + (bg_rs32 * (uint32_t)oy0 + 4u * (uint32_t)x);
        const uint8_t *p1 = p0 + bg_rs32, *p2 = p1 + bg_rs32, *p3 = p2 + bg_rs32;
        w[0] = nrow > 0 ? ld_stream_u32(p0) : 0u;
        w[1] = nrow > 1 ? ld_stream_u32(p1) : 0u;
        w[2] = nrow > 2 ? ld_stream_u32(p2) : 0u;
        w[3] = nrow > 3 ? ld_stream_u32(p3) : 0u;
      };
      uint32_t bgw[4], bgn[4];
      load_bg(warp, bgw);
      for (int task = warp; task < ntasks; task += F2_NW) {
        load_bg(task + F2_NW, bgn);
        const int qd = task >> 2, cwp = task & 3;
        const int x = x0 + cwp * 32 + lane;
        if (x < a_ow) {
          const int ix = x - a_ox;
          const bool col_in = ix >= g_ix0 && ix < g_ix1;
          const uint32_t *colp = s_c + (col_in ? ix - 4 * cq0 : 0) * F2_CW;
          const int oy0 = y0 + qd * 4;
          const int nrow = min(4, y1 - oy0);
          uint8_t *outp = out_base + (out_rs32 * (uint32_t)oy0 + 4u * (uint32_t)x);
          const int iyl0 = oy0 - a_oy - iy0;  // first row of the quad inside the tile's inner rows
          if (nrow == 4 && col_in && iyl0 >= 0 && iyl0 + 3 < niy) {
            // ===== fast path: four inner rows
#pragma unroll
            for (int r = 0; r < 4; r++) {
              const int4 ri = s_row[iyl0 + r];
              const uint32_t *wp = colp + (ri.x >> 2);
              const int sh = 8 * (ri.x & 3);
              const uint32_t b0 = __funnelshift_r(wp[0], wp[1], sh);
              const uint32_t b1 = __funnelshift_r(wp[F2_CCOLS * F2_CW], wp[F2_CCOLS * F2_CW + 1], sh);
              const uint32_t b2 = __funnelshift_r(wp[2 * F2_CCOLS * F2_CW], wp[2 * F2_CCOLS * F2_CW + 1], sh);
              const uint32_t fr = dp2a_hi((uint32_t)ri.z, b0, dp2a_lo((uint32_t)ri.y, b0, 2048u)) >> 12;
              const uint32_t fg_ = dp2a_hi((uint32_t)ri.z, b1, dp2a_lo((uint32_t)ri.y, b1, 2048u)) >> 12;
              const uint32_t fb = dp2a_hi((uint32_t)ri.z, b2, dp2a_lo((uint32_t)ri.y, b2, 2048u)) >> 12;
              blend_store(outp + out_rs32 * (uint32_t)r, bgw[r], fr, fg_, fb);
            }
          } else {
            // ===== general path: tile borders, letterbox border rows / columns
#pragma unroll
            for (int r = 0; r < 4; r++) {
              if (r < nrow) {
                const int iyl = iyl0 + r;
                uint32_t fr = 0, fg_ = 0, fb = 0;  // letterbox border: black (blank_pixel, colourspace.c:11169)
                if (col_in && iyl >= 0 && iyl < niy) {
                  const int4 ri = s_row[iyl];
                  const uint32_t *wp = colp + (ri.x >> 2);
                  const int sh = 8 * (ri.x & 3);
                  const uint32_t b0 = __funnelshift_r(wp[0], wp[1], sh);
                  const uint32_t b1 = __funnelshift_r(wp[F2_CCOLS * F2_CW], wp[F2_CCOLS * F2_CW + 1], sh);
                  const uint32_t b2 = __funnelshift_r(wp[2 * F2_CCOLS * F2_CW], wp[2 * F2_CCOLS * F2_CW + 1], sh);
                  fr = dp2a_hi((uint32_t)ri.z, b0, dp2a_lo((uint32_t)ri.y, b0, 2048u)) >> 12;
                  fg_ = dp2a_hi((uint32_t)ri.z, b1, dp2a_lo((uint32_t)ri.y, b1, 2048u)) >> 12;
                  fb = dp2a_hi((uint32_t)ri.z, b2, dp2a_lo((uint32_t)ri.y, b2, 2048u)) >> 12;
                }
                blend_store(outp + out_rs32 * (uint32_t)r, bgw[r], fr, fg_, fb);
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; r++) bgw[r] = bgn[r];
      }
    }
    __syncthreads();  // everyone is done with this tile's buffers, tables and s_c
    G = GN;
  }
  if (ASYNC) cp_async_wait<0>();
}

}  // namespace

// Can the fast kernel take this job?  (checked per launch by the engine; everything else goes to k_fused)
bool fused2_supported(const FusedArgs &a, int fy_taps, int unused) {
  (void)unused;
  if (a.iw != a.fw) return false;                        // horizontal pass must be the identity
  if (fy_taps > 4) return false;
  if (a.fw & 3) return false;                            // whole 32-bit luma words
  if (((uintptr_t)a.fg.y & 3) || (a.fg.rs_y & 3)) return false;
  if (((uintptr_t)a.fg.u & 1) || ((uintptr_t)a.fg.v & 1) || (a.fg.rs_u & 1) || (a.fg.rs_v & 1)) return false;
  if (((uintptr_t)a.bg.p & 3) || (a.bg.rs & 3) || ((uintptr_t)a.out.p & 3) || (a.out.rs & 3)) return false;
  if (a.fh < 2) return false;
  return true;
}

// 16-byte aligned planes and strides: the cp.async staging path
static bool fused2_async_ok(const FusedArgs &a) {
  return !(((uintptr_t)a.fg.y | (uintptr_t)a.fg.u | (uintptr_t)a.fg.v) & 15) && !((a.fg.rs_y | a.fg.rs_u | a.fg.rs_v) & 15);
}

int fused2_max_virtual_rows(int is422) { return is422 ? F2_MAXVR_422 : F2_MAXVR; }
int fused2_max_tile_h() { return F2_MAXTH; }

// frames_host: FusedArgs[nframes] in HOST memory (they travel as kernel parameters, F2_MAXF per launch).
// blend_a < 0: table blend (FusedArgs::over_table, gamma folded in); else arithmetic blend with weights blend_a / 256 - blend_a
cudaError_t launch_fused2(const Launch &L, const FusedArgs *frames_host, int nframes, int ow, int oh, int tile_h, int blend_a,
                          const uint8_t *lut8_dev) {
  static PerDevice attr_set;
  if (!attr_set.cur()) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_fused2<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_ARITH)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_fused2<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_ARITH)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_fused2<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_TABLE)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_fused2<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_TABLE)) != cudaSuccess) return e;
    attr_set.cur() = 1;
  }
  for (int base = 0; base < nframes; base += F2_MAXF) {
    Fused2Params P;
    P.nframes = nframes - base < F2_MAXF ? nframes - base : F2_MAXF;
    bool async = getenv("PE_FUSED_SYNC") == nullptr;
    for (int i = 0; i < P.nframes; i++) {
      P.fr[i] = frames_host[base + i];
      async = async && fused2_async_ok(P.fr[i]);
    }
    for (int i = P.nframes; i < F2_MAXF; i++) P.fr[i] = frames_host[base];
    P.tiles_x = (ow + F2_TW - 1) / F2_TW; P.tiles_y = (oh + tile_h - 1) / tile_h; P.tile_h = tile_h;
    P.blend_a = blend_a; P.blend_ia = 256 - blend_a; P.lut8 = blend_a >= 0 ? lut8_dev : nullptr;
    const long long total = (long long)P.tiles_x * P.tiles_y * P.nframes;
    const int per_sm = blend_a >= 0 ? 2 : 1;
    const int grid = (int)(total < (long long)L.sm_count * per_sm ? total : (long long)L.sm_count * per_sm);
    if (blend_a >= 0) {
      if (async) k_fused2<0, true><<<grid, F2_NT, F2_SMEM_ARITH, L.stream>>>(P);
      else k_fused2<0, false><<<grid, F2_NT, F2_SMEM_ARITH, L.stream>>>(P);
    } else {
      if (async) k_fused2<1, true><<<grid, F2_NT, F2_SMEM_TABLE, L.stream>>>(P);
      else k_fused2<1, false><<<grid, F2_NT, F2_SMEM_TABLE, L.stream>>>(P);
    }
    PE_COUNT_LAUNCH(L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace pe
