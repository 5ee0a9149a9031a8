// pe_kernels_yuv3.cu -- the YUV <-> YUV family of convert_layer_palette_full and planar 4:4:4 -> RGB (sm_100a).
//
//   k_yuv444p_to_rgb      convert_yuv_planar_to_{rgb,bgr,argb}_frame     colourspace.c:7200 / 7304 / 7405
//   k_combine_planes      convert_combineplanes_frame                    colourspace.c:7593   (4:4:4 planar -> YUV888 / YUVA8888)
//   k_split_planes        convert_splitplanes_frame                      colourspace.c:9198   (the reverse)
//   k_halve_chroma        convert_halve_chroma                           colourspace.c:10578  (4:2:2 -> 4:2:0 chroma planes)
//   k_double_chroma       convert_double_chroma                          colourspace.c:10612  (4:2:0 -> 4:2:2 chroma planes)
//   k_packed422_unpack    convert_{uyvy,yuyv}_to_{yuv422,yuvp,yuv888}_frame  colourspace.c:8093 / 7800 / 7845
//   k_yuv444p_to_packed422  convert_yuv_planar_to_{uyvy,yuyv}_frame      colourspace.c:7500 / 7548
//   k_yuv444p_to_chroma420  convert_yuvp_to_yuv420_frame                 colourspace.c:7690
//   k_planar42x_to_packed422  convert_yuv420_to_{uyvy,yuyv}_frame / convert_yuv422p_to_{uyvy,yuyv}_frame  colourspace.c:7104 / 6442
//   k_quad_chroma         convert_quad_chroma                            colourspace.c:10642  (4:2:0 -> 4:4:4 chroma planes)
//   k_yuv888_subsample    convert_yuv888_to_{uyvy,yuyv,yuv422,yuv420}_frame  colourspace.c:8184 / 8228 / 8129 / 8035
//   k_packed422_to_yuv420p  convert_{uyvy,yuyv}_to_yuv420_frame          colourspace.c:7887 / 7930
//   k_chroma_upsample_packed  convert_quad_chroma_packed / convert_double_chroma_packed  colourspace.c:10715 / 10811
//   k_swab                convert_swab_frame                             colourspace.c:10517  (UYVY <-> YUYV in place)
//   k_clamp_lut           switch_yuv_clamping_and_subspace               colourspace.c:10929  (every byte through Y_to_Y / U_to_U)
//
// All of it is pure byte shuffling / table lookups at 1 load + 1 store per byte: HBM-bound.  One thread handles 4 pixels (one
// 32-bit word per plane, a 128-bit / 3 x 32-bit packed vector); frames whose pointers or strides are not word aligned, and the
// ragged last group of a row, go byte by byte through the same code.
#include "pe_device.cuh"
#include "pe_kernels.h"
#include "pe_tables.h"

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

// up to 4 bytes of a plane row as one word (missing bytes read as 0)
__device__ __forceinline__ uint32_t ld_px4(const uint8_t *p, int n, bool vec) {
  if (vec && n == 4) return ld_stream_u32(p);
  uint32_t w = 0;
  for (int k = 0; k < n; k++) w |= (uint32_t)p[k] << (8 * k);
  return w;
}
__device__ __forceinline__ void st_px4(uint8_t *p, uint32_t w, int n, bool vec) {
  if (vec && n == 4) { st_stream_u32(p, w); return; }
  for (int k = 0; k < n; k++) p[k] = (uint8_t)(w >> (8 * k));
}
// n (<= 4) packed pixels of psize 3 / 4, px[k] = the pixel's bytes in memory order
__device__ __forceinline__ void st_packed4(uint8_t *d, const uint32_t px[4], int psize, int n, bool vec) {
  if (vec && n == 4) {
    if (psize == 4) { st_stream_u4(d, make_uint4(px[0], px[1], px[2], px[3])); return; }
    st_stream_u32(d, __byte_perm(px[0], px[1], 0x4210));
    st_stream_u32(d + 4, __byte_perm(px[1], px[2], 0x5421));
    st_stream_u32(d + 8, __byte_perm(px[2], px[3], 0x6542));
    return;
  }
  for (int k = 0; k < n; k++)
    for (int b = 0; b < psize; b++) d[k * psize + b] = (uint8_t)(px[k] >> (8 * b));
}
__device__ __forceinline__ void ld_packed4(const uint8_t *s, uint32_t px[4], int psize, int n, bool vec) {
  if (vec && n == 4) {
    if (psize == 4) { const uint4 v = ld_stream_u4(s); px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w; return; }
    const uint32_t a = ld_stream_u32(s), b = ld_stream_u32(s + 4), c = ld_stream_u32(s + 8);
    px[0] = a; px[1] = __byte_perm(a, b, 0x0543); px[2] = __byte_perm(b, c, 0x0432); px[3] = c >> 8;
    return;
  }
  for (int k = 0; k < 4; k++) {
    px[k] = 0;
    if (k < n) for (int b = 0; b < psize; b++) px[k] |= (uint32_t)s[k * psize + b] << (8 * b);
  }
}

__device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }

struct Planes4 {
  const uint8_t *p[4];
  int rs;  // all planes of a 4:4:4 frame share one stride
};
struct OutPlanes4 {
  uint8_t *p[4];
  int rs[4];
};

__global__ void __launch_bounds__(kBlock) k_yuv444p_to_rgb(Planes4 S, uint8_t *dst, int orow, int width, int height, int in_alpha,
                                                          RgbLayout out, DevConv conv, int vec) {
  __shared__ int32_t t[5][256];  // RGB_Y R_Cr G_Cb G_Cr B_Cb
  for (int i = threadIdx.x; i < 5 * 256; i += blockDim.x) t[i >> 8][i & 255] = conv.t[9 * 256 + i];
  __syncthreads();
  const int groups = (width + 3) >> 2;
  const long long total = (long long)groups * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int x = 4 * g, n = min(4, width - x);
    const long long o = (long long)S.rs * row + x;
    const uint32_t yw = ld_px4(S.p[0] + o, n, vec), uw = ld_px4(S.p[1] + o, n, vec), vw = ld_px4(S.p[2] + o, n, vec);
    const uint32_t aw = (in_alpha && out.a >= 0) ? ld_px4(S.p[3] + o, n, vec) : 0xFFFFFFFFu;
    uint32_t px[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int yy = t[0][byte_of(yw, k)], u = byte_of(uw, k), v = byte_of(vw, k);
      const uint32_t r = sat8((yy + t[1][v]) >> 16), gg = sat8((yy + t[2][u] + t[3][v]) >> 16), b = sat8((yy + t[4][u]) >> 16);
      uint32_t w = (r << (8 * out.r)) | (gg << (8 * out.g)) | (b << (8 * out.b));
      if (out.a >= 0) w |= byte_of(aw, k) << (8 * out.a);
      px[k] = w;
    }
    st_packed4(dst + (long long)orow * row + (long long)x * out.psize, px, out.psize, n, vec);
  }
}

__global__ void __launch_bounds__(kBlock) k_combine_planes(Planes4 S, uint8_t *dst, int orow, int width, int height, int in_alpha,
                                                          int out_alpha, int vec) {
  const int groups = (width + 3) >> 2, ops = out_alpha ? 4 : 3;
  const long long total = (long long)groups * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int x = 4 * g, n = min(4, width - x);
    const long long o = (long long)S.rs * row + x;
    const uint32_t yw = ld_px4(S.p[0] + o, n, vec), uw = ld_px4(S.p[1] + o, n, vec), vw = ld_px4(S.p[2] + o, n, vec);
    const uint32_t aw = (in_alpha && out_alpha) ? ld_px4(S.p[3] + o, n, vec) : 0xFFFFFFFFu;
    // 4 x 4 byte transpose: px[k] = y_k | u_k << 8 | v_k << 16 | a_k << 24
    const uint32_t yu_lo = __byte_perm(yw, uw, 0x5140), yu_hi = __byte_perm(yw, uw, 0x7362);
    const uint32_t va_lo = __byte_perm(vw, aw, 0x5140), va_hi = __byte_perm(vw, aw, 0x7362);
    const uint32_t px[4] = {__byte_perm(yu_lo, va_lo, 0x5410), __byte_perm(yu_lo, va_lo, 0x7632), __byte_perm(yu_hi, va_hi, 0x5410),
                            __byte_perm(yu_hi, va_hi, 0x7632)};
    st_packed4(dst + (long long)orow * row + (long long)x * ops, px, ops, n, vec);
  }
}

__global__ void __launch_bounds__(kBlock) k_split_planes(const uint8_t *__restrict__ src, int irow, OutPlanes4 D, int width, int height,
                                                        int src_alpha, int dest_alpha, int vec) {
  const int groups = (width + 3) >> 2, ips = src_alpha ? 4 : 3;
  const long long total = (long long)groups * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int x = 4 * g, n = min(4, width - x);
    uint32_t px[4];
    ld_packed4(src + (long long)irow * row + (long long)x * ips, px, ips, n, vec);
    const uint32_t a01 = __byte_perm(px[0], px[1], 0x5140), b01 = __byte_perm(px[0], px[1], 0x7362);
    const uint32_t a23 = __byte_perm(px[2], px[3], 0x5140), b23 = __byte_perm(px[2], px[3], 0x7362);
    st_px4(D.p[0] + (long long)D.rs[0] * row + x, __byte_perm(a01, a23, 0x5410), n, vec);
    st_px4(D.p[1] + (long long)D.rs[1] * row + x, __byte_perm(a01, a23, 0x7632), n, vec);
    st_px4(D.p[2] + (long long)D.rs[2] * row + x, __byte_perm(b01, b23, 0x5410), n, vec);
    if (dest_alpha) st_px4(D.p[3] + (long long)D.rs[3] * row + x, src_alpha ? __byte_perm(b01, b23, 0x7632) : 0xFFFFFFFFu, n, vec);
  }
}

// avg_chroma(x, y) = cavg[x][y] (colourspace.c:2079); the 64 KB table stays in L1 / L2
__device__ __forceinline__ uint32_t avg4(const uint8_t *__restrict__ cavg, uint32_t a, uint32_t b) {
  uint32_t w = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) w |= (uint32_t)__ldg(cavg + ((byte_of(a, k) << 8) | byte_of(b, k))) << (8 * k);
  return w;
}

// out row k = avg_chroma(src row 2k, src row 2k + 1); a trailing unpaired row is copied.  blockIdx.y = plane (U, V)
__global__ void __launch_bounds__(kBlock) k_halve_chroma(const uint8_t *su, const uint8_t *sv, int irs_u, int irs_v, uint8_t *du, uint8_t *dv,
                                                        int ors_u, int ors_v, int cw, int ch, const uint8_t *__restrict__ cavg, int vec) {
  const uint8_t *s = blockIdx.y ? sv : su;
  uint8_t *d = blockIdx.y ? dv : du;
  const int irs = blockIdx.y ? irs_v : irs_u, ors = blockIdx.y ? ors_v : ors_u;
  const int groups = (cw + 3) >> 2, orows = (ch + 1) >> 1;
  const long long total = (long long)groups * orows;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int x = 4 * g, n = min(4, cw - x);
    uint32_t w = ld_px4(s + (long long)irs * (2 * row) + x, n, vec);
    if (2 * row + 1 < ch) w = avg4(cavg, w, ld_px4(s + (long long)irs * (2 * row + 1) + x, n, vec));
    st_px4(d + (long long)ors * row + x, w, n, vec);
  }
}

// out row 2k = src row k; out row 2k + 1 = avg_chroma(src row k, src row k + 1), the last one a copy of src row ch - 1
__global__ void __launch_bounds__(kBlock) k_double_chroma(const uint8_t *su, const uint8_t *sv, int irs_u, int irs_v, uint8_t *du, uint8_t *dv,
                                                         int ors_u, int ors_v, int cw, int ch, const uint8_t *__restrict__ cavg, int vec) {
  const uint8_t *s = blockIdx.y ? sv : su;
  uint8_t *d = blockIdx.y ? dv : du;
  const int irs = blockIdx.y ? irs_v : irs_u, ors = blockIdx.y ? ors_v : ors_u;
  const int groups = (cw + 3) >> 2;
  const long long total = (long long)groups * ch;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int x = 4 * g, n = min(4, cw - x);
    const uint32_t a = ld_px4(s + (long long)irs * row + x, n, vec);
    uint32_t b = a;
    if (row + 1 < ch) b = avg4(cavg, a, ld_px4(s + (long long)irs * (row + 1) + x, n, vec));
    st_px4(d + (long long)ors * (2 * row) + x, a, n, vec);
    st_px4(d + (long long)ors * (2 * row + 1) + x, b, n, vec);
  }
}

// packed 4:2:2 -> planar 4:2:2 (mode 0), planar 4:4:4 (+ alpha) (mode 1), YUV888 / YUVA8888 (mode 2).  One thread = 2 macropixels
// = 4 pixels.  fmt 0 UYVY, 1 YUYV.  first_only: the reference's never-advanced source pointer (colourspace.c:8103, mode 0 only).
__global__ void __launch_bounds__(kBlock) k_packed422_unpack(int fmt, int mode, const uint8_t *__restrict__ src, int irow, OutPlanes4 D,
                                                            int width_mpx, int height, int add_alpha, int first_only, int vec) {
  const int groups = (width_mpx + 1) >> 1;
  const long long total = (long long)groups * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int m = 2 * g, nm = min(2, width_mpx - m);
    const uint8_t *q = first_only ? src : src + (long long)irow * row + 4LL * m;
    uint32_t m0, m1;
    if (first_only) { m0 = m1 = ld_px4(q, 4, vec); }
    else { m0 = ld_px4(q, 4, vec); m1 = nm == 2 ? ld_px4(q + 4, 4, vec) : 0u; }
    if (fmt == 0) { m0 = __byte_perm(m0, 0u, 0x2301); m1 = __byte_perm(m1, 0u, 0x2301); }  // -> y0 u y1 v
    const uint32_t yw = __byte_perm(m0, m1, 0x6420);                                       // y0 y1 y0' y1'
    if (mode == 0) {
      st_px4(D.p[0] + (long long)D.rs[0] * row + 2 * m, yw, 2 * nm, vec);
      uint8_t *du = D.p[1] + (long long)D.rs[1] * row + m, *dv = D.p[2] + (long long)D.rs[2] * row + m;
      du[0] = (uint8_t)(m0 >> 8); dv[0] = (uint8_t)(m0 >> 24);
      if (nm == 2) { du[1] = (uint8_t)(m1 >> 8); dv[1] = (uint8_t)(m1 >> 24); }
    } else if (mode == 1) {
      const uint32_t uw = __byte_perm(m0, m1, 0x5511), vw = __byte_perm(m0, m1, 0x7733);
      st_px4(D.p[0] + (long long)D.rs[0] * row + 2 * m, yw, 2 * nm, vec);
      st_px4(D.p[1] + (long long)D.rs[1] * row + 2 * m, uw, 2 * nm, vec);
      st_px4(D.p[2] + (long long)D.rs[2] * row + 2 * m, vw, 2 * nm, vec);
      if (add_alpha) st_px4(D.p[3] + (long long)D.rs[3] * row + 2 * m, 0xFFFFFFFFu, 2 * nm, vec);
    } else {
      const int ps = add_alpha ? 4 : 3;
      const uint32_t ff = 0xFF000000u;
      const uint32_t px[4] = {__byte_perm(m0, ff, 0x7310), __byte_perm(m0, ff, 0x7312), __byte_perm(m1, ff, 0x7310),
                              __byte_perm(m1, ff, 0x7312)};
      st_packed4(D.p[0] + (long long)D.rs[0] * row + 2LL * m * ps, px, ps, 2 * nm, vec);
    }
  }
}

// planar 4:4:4 -> UYVY / YUYV (convert_yuv_planar_to_{uyvy,yuyv}_frame, colourspace.c:7500 / 7548): one thread = 4 pixels = 2
// macropixels, chroma = avg_chroma(c[2x], c[2x + 1])
__global__ void __launch_bounds__(kBlock) k_yuv444p_to_packed422(int fmt, Planes4 S, uint8_t *dst, int orow, int width_mpx, int height,
                                                                const uint8_t *__restrict__ cavg, int vec) {
  const int groups = (width_mpx + 1) >> 1;
  const long long total = (long long)groups * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int m = 2 * g, nm = min(2, width_mpx - m);
    const long long o = (long long)S.rs * row + 2 * m;
    const uint32_t yw = ld_px4(S.p[0] + o, 2 * nm, vec), uw = ld_px4(S.p[1] + o, 2 * nm, vec), vw = ld_px4(S.p[2] + o, 2 * nm, vec);
    const uint32_t cu0 = __ldg(cavg + ((byte_of(uw, 0) << 8) | byte_of(uw, 1))), cv0 = __ldg(cavg + ((byte_of(vw, 0) << 8) | byte_of(vw, 1)));
    const uint32_t cu1 = __ldg(cavg + ((byte_of(uw, 2) << 8) | byte_of(uw, 3))), cv1 = __ldg(cavg + ((byte_of(vw, 2) << 8) | byte_of(vw, 3)));
    uint32_t m0 = byte_of(yw, 0) | (cu0 << 8) | (byte_of(yw, 1) << 16) | (cv0 << 24);   // YUYV
    uint32_t m1 = byte_of(yw, 2) | (cu1 << 8) | (byte_of(yw, 3) << 16) | (cv1 << 24);
    if (fmt == 0) { m0 = __byte_perm(m0, 0u, 0x2301); m1 = __byte_perm(m1, 0u, 0x2301); }  // UYVY
    uint8_t *d = dst + (long long)orow * row + 4LL * m;
    if (vec && nm == 2) st_stream_u2(d, make_uint2(m0, m1));
    else { st_px4(d, m0, 4, false); if (nm == 2) st_px4(d + 4, m1, 4, false); }
  }
}

// planar 4:4:4 -> the chroma planes of 4:2:0 (convert_yuvp_to_yuv420_frame, colourspace.c:7690): out[k][j] = avg_chroma(h(2k)[j],
// h(2k + 1)[j]), h(r)[j] = avg_chroma(c[r][2j], c[r][2j + 1]); a trailing unpaired row leaves h(r).  One thread = 4 output samples;
// blockIdx.y = plane (U, V)
__global__ void __launch_bounds__(kBlock) k_yuv444p_to_chroma420(const uint8_t *su, const uint8_t *sv, int irs, uint8_t *du, uint8_t *dv,
                                                                int ors_u, int ors_v, int cw, int height, const uint8_t *__restrict__ cavg,
                                                                int vec) {
  const uint8_t *s = blockIdx.y ? sv : su;
  uint8_t *d = blockIdx.y ? dv : du;
  const int ors = blockIdx.y ? ors_v : ors_u;
  const int groups = (cw + 3) >> 2, orows = (height + 1) >> 1;
  const long long total = (long long)groups * orows;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int x = 4 * g, n = min(4, cw - x);
    auto hrow = [&](int r) -> uint32_t {
      const uint8_t *q = s + (long long)irs * r + 2 * x;
      const uint32_t a = ld_px4(q, min(4, 2 * n), vec), b = n > 2 ? ld_px4(q + 4, 2 * n - 4, vec) : 0u;
      return (uint32_t)__ldg(cavg + ((byte_of(a, 0) << 8) | byte_of(a, 1))) | ((uint32_t)__ldg(cavg + ((byte_of(a, 2) << 8) | byte_of(a, 3))) << 8) |
             ((uint32_t)__ldg(cavg + ((byte_of(b, 0) << 8) | byte_of(b, 1))) << 16) | ((uint32_t)__ldg(cavg + ((byte_of(b, 2) << 8) | byte_of(b, 3))) << 24);
    };
    uint32_t w = hrow(2 * row);
    if (2 * row + 1 < height) w = avg4(cavg, w, hrow(2 * row + 1));
    st_px4(d + (long long)ors * row + x, w, n, vec);
  }
}

// planar 4:2:0 / 4:2:2 -> UYVY / YUYV (convert_yuv420_to_{uyvy,yuyv}_frame colourspace.c:7104 / 7152, convert_yuv422p_to_{uyvy,yuyv}_frame
// :6442 / 6470): pure interleave, 4:2:0 chroma rows used twice (the reference's vertical averaging never runs).  One thread = 4 pixels.
__global__ void __launch_bounds__(kBlock) k_planar42x_to_packed422(int fmt, int is_422, const uint8_t *__restrict__ py, const uint8_t *__restrict__ pu,
                                                                  const uint8_t *__restrict__ pv, int rs_y, int rs_u, int rs_v, uint8_t *dst,
                                                                  int orow, int width_mpx, int height, int vec) {
  const int groups = (width_mpx + 1) >> 1;
  const long long total = (long long)groups * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / groups), g = (int)(it - (long long)row * groups);
    const int m = 2 * g, nm = min(2, width_mpx - m), cr = is_422 ? row : row >> 1;
    const uint32_t yw = ld_px4(py + (long long)rs_y * row + 2 * m, 2 * nm, vec);
    const uint8_t *qu = pu + (long long)rs_u * cr + m, *qv = pv + (long long)rs_v * cr + m;
    const uint32_t u0 = __ldg(qu), v0 = __ldg(qv), u1 = nm == 2 ? __ldg(qu + 1) : 0u, v1 = nm == 2 ? __ldg(qv + 1) : 0u;
    uint32_t m0 = byte_of(yw, 0) | (u0 << 8) | (byte_of(yw, 1) << 16) | (v0 << 24);   // YUYV
    uint32_t m1 = byte_of(yw, 2) | (u1 << 8) | (byte_of(yw, 3) << 16) | (v1 << 24);
    if (fmt == 0) { m0 = __byte_perm(m0, 0u, 0x2301); m1 = __byte_perm(m1, 0u, 0x2301); }  // UYVY
    uint8_t *d = dst + (long long)orow * row + 4LL * m;
    if (vec && nm == 2) st_stream_u2(d, make_uint2(m0, m1));
    else { st_px4(d, m0, 4, false); if (nm == 2) st_px4(d + 4, m1, 4, false); }
  }
}

// avg_chroma in closed form (pe_tables.h AvgForm: both tables of init_average are functions of x + y; checked against the table when
// the engine builds it): two multiplies and a clamp instead of a gather from the 64 KB table.  k_quad_chroma / k_chroma_upsample_packed
// do 3 - 7 of these per output pixel: as 32-address gathers they ran at 5 - 9 % of the HBM roofline (round 1: 79 us per 4K frame).
__device__ __forceinline__ uint32_t avg1(const AvgForm &F, uint32_t x, uint32_t y) {
  const int f = (int)__umulhi((x + y) * F.A + F.B, F.M);
  return (uint32_t)min(max(f, F.lo), F.hi);
}
__device__ __forceinline__ uint32_t avg4(const AvgForm &F, uint32_t a, uint32_t b) {
  // the four byte sums as two 16-bit pairs (no carry between the halves: <= 510)
  const uint32_t se = __byte_perm(a, 0u, 0x4240) + __byte_perm(b, 0u, 0x4240), so = __byte_perm(a, 0u, 0x4341) + __byte_perm(b, 0u, 0x4341);
  auto one = [&](uint32_t s) -> uint32_t {
    const int f = (int)__umulhi(s * F.A + F.B, F.M);
    return (uint32_t)min(max(f, F.lo), F.hi);
  };
  const uint32_t r0 = one(se & 0xFFFFu), r2 = one(se >> 16), r1 = one(so & 0xFFFFu), r3 = one(so >> 16);
  return __byte_perm(r0 | (r1 << 8), r2 | (r3 << 8), 0x5410);
}

// 4:2:0 -> 4:4:4 chroma planes (convert_quad_chroma, colourspace.c:10642).  Even destination row 2k from chroma row k: column 0 =
// s[0], column 2m = f(s[m-1], s[m]), column 2m+1 = g(s[m], s[m+1]) (JPEG sampling: f = g = avg_chroma; otherwise U: f = 3:1, g = 1:3,
// V mirrored); odd row r = avg_chroma(row r+1, row r-1), the last odd row of an odd-height frame with the operands swapped, the last
// row of an even-height frame a copy of the row above.  One thread = 4 destination columns of QC_PAIRS consecutive row pairs: the even
// row of chroma row k + 1 is computed once and serves rows 2k + 1 and 2k + 2; blockIdx.y = plane (U, V).
constexpr int QC_PAIRS = 4;
struct QuadChromaParams {
  const uint8_t *su, *sv;
  uint8_t *du, *dv;
  int irs_u, irs_v, ors, w2, height, cw, ch, jpeg;
  AvgForm avg;
};

__device__ __forceinline__ uint32_t quad_even4(const uint8_t *__restrict__ s, int irs, int k, int m, int cw, int ch, bool is_u, bool jpeg,
                                               const AvgForm &F, int n, bool g_is_f = false) {
  // 4 destination samples 2m .. 2m+3 of even row 2k (n of them valid) from s[m-1 .. m+2]
  const uint8_t *r = s + (long long)irs * k;
  auto at = [&](int c) -> uint32_t {
    if (c < 0) c = 0;
    if (c >= cw) c = (cw < irs || k + 1 < ch) ? cw : cw - 1;
    return __ldg(r + c);
  };
  auto av = [&](uint32_t x, uint32_t y) -> uint32_t { return avg1(F, x, y); };
  auto f = [&](uint32_t x, uint32_t y) -> uint32_t { return jpeg ? av(x, y) : (is_u ? av(x, av(x, y)) : av(av(x, y), y)); };   // even column
  // odd column (convert_double_chroma_packed uses the even column's weights there too, colourspace.c:10853-10858: g_is_f)
  auto g = [&](uint32_t x, uint32_t y) -> uint32_t { return jpeg ? av(x, y) : ((is_u != g_is_f) ? av(av(x, y), y) : av(x, av(x, y))); };
  const uint32_t a = at(m), b = n > 2 ? at(m + 1) : a;
  const uint32_t v0 = m == 0 ? a : f(at(m - 1), a);
  const uint32_t v1 = n > 1 ? g(a, n > 2 ? b : at(m + 1)) : 0u;
  const uint32_t v2 = n > 2 ? f(a, b) : 0u;
  const uint32_t v3 = n > 3 ? g(b, at(m + 2)) : 0u;
  return v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
}

// The same four samples for an interior group (m >= 2, all of s[m-1 .. m+2] inside the row, 4-byte aligned plane): the four bytes come
// from two aligned words, av(c1, c2) is shared between the two middle samples (7 table values instead of 8), no edge tests.
// toward_x: A(x, y) = av(x, av(x, y));  toward_y: B(x, y) = av(av(x, y), y).  U: f = A, g = B; V: f = B, g = A; g_is_f: g = f.
__device__ __forceinline__ bool quad_fast_ok(const uint8_t *s, int irs, int m, int cw, int n) {
  return n == 4 && m >= 2 && (((m - 1) & ~3) + 8) <= cw && (((uintptr_t)s | (uint32_t)irs) & 3) == 0;
}
__device__ __forceinline__ uint32_t quad_load4(const uint8_t *__restrict__ s, int irs, int k, int m) {   // bytes m-1 .. m+2 of chroma row k
  const uint8_t *r = s + (long long)irs * k + ((m - 1) & ~3);
  const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t *>(r)), w1 = __ldg(reinterpret_cast<const uint32_t *>(r + 4));
  return __byte_perm(w0, w1, (m & 2) ? 0x4321u : 0x6543u);
}
__device__ __forceinline__ uint32_t quad_even4_from(uint32_t c, bool is_u, bool jpeg, const AvgForm &F, bool g_is_f = false) {
  const uint32_t c0 = byte_of(c, 0), c1 = byte_of(c, 1), c2 = byte_of(c, 2), c3 = byte_of(c, 3);
  const uint32_t t01 = avg1(F, c0, c1), t12 = avg1(F, c1, c2), t23 = avg1(F, c2, c3);
  uint32_t v0, v1, v2, v3;
  if (jpeg) {
    v0 = t01; v1 = t12; v2 = t12; v3 = t23;
  } else {
    const bool f_x = is_u, g_x = is_u == g_is_f;   // f / g lean toward their first operand
    v0 = f_x ? avg1(F, c0, t01) : avg1(F, t01, c1);
    v2 = f_x ? avg1(F, c1, t12) : avg1(F, t12, c2);
    v1 = g_x ? avg1(F, c1, t12) : avg1(F, t12, c2);
    v3 = g_x ? avg1(F, c2, t23) : avg1(F, t23, c3);
  }
  return __byte_perm(v0 | (v1 << 8), v2 | (v3 << 8), 0x5410);
}
__device__ __forceinline__ uint32_t quad_even4_fast(const uint8_t *__restrict__ s, int irs, int k, int m, bool is_u, bool jpeg, const AvgForm &F,
                                                    bool g_is_f = false) {
  return quad_even4_from(quad_load4(s, irs, k, m), is_u, jpeg, F, g_is_f);
}

__global__ void __launch_bounds__(kBlock) k_quad_chroma(const QuadChromaParams P, int vec) {
  const bool is_u = blockIdx.y == 0;
  const uint8_t *s = is_u ? P.su : P.sv;
  uint8_t *d = is_u ? P.du : P.dv;
  const int irs = is_u ? P.irs_u : P.irs_v;
  const int groups = (P.w2 + 3) >> 2, npairs = (P.height + 1) >> 1, nchunks = (npairs + QC_PAIRS - 1) / QC_PAIRS;
  const long long total = (long long)groups * nchunks;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int chunk = (int)(it / groups), g = (int)(it - (long long)chunk * groups);
    const int x = 4 * g, n = min(4, P.w2 - x), m = x >> 1;
    const int k0 = chunk * QC_PAIRS, k1 = min(k0 + QC_PAIRS, npairs);
    const bool fast = quad_fast_ok(s, irs, m, P.cw, n);
    auto even = [&](int k) -> uint32_t {
      return fast ? quad_even4_fast(s, irs, k, m, is_u, P.jpeg, P.avg) : quad_even4(s, irs, k, m, P.cw, P.ch, is_u, P.jpeg, P.avg, n);
    };
    if (fast && vec && 2 * k0 + 2 * QC_PAIRS <= P.height - 2) {
      // a whole chunk away from the frame's edges: every load is issued before the first use (5 chroma rows in flight per thread)
      uint32_t c[QC_PAIRS + 1];
#pragma unroll
      for (int j = 0; j <= QC_PAIRS; j++) c[j] = quad_load4(s, irs, k0 + j, m);
      uint32_t E = quad_even4_from(c[0], is_u, P.jpeg, P.avg);
      uint8_t *dr = d + (long long)P.ors * (2 * k0) + x;
#pragma unroll
      for (int j = 0; j < QC_PAIRS; j++) {
        const uint32_t En = quad_even4_from(c[j + 1], is_u, P.jpeg, P.avg);
        st_stream_u32(dr, E);
        st_stream_u32(dr + P.ors, avg4(P.avg, En, E));
        dr += 2 * (long long)P.ors;
        E = En;
      }
      continue;
    }
    uint32_t E = even(k0);
    for (int k = k0; k < k1; k++) {
      st_px4(d + (long long)P.ors * (2 * k) + x, E, n, vec);
      const int row = 2 * k + 1;
      if (row >= P.height) break;
      uint32_t w = E;   // the last row of an even-height frame copies the row above
      if (row + 1 <= P.height - 1) {
        const uint32_t En = even(k + 1);
        w = ((P.height & 1) && row == P.height - 2) ? avg4(P.avg, E, En) : avg4(P.avg, En, E);
        E = En;
      }
      st_px4(d + (long long)P.ors * row + x, w, n, vec);
    }
  }
}

// YUV888 / YUVA8888 -> UYVY (mode 0) / YUYV (1) / planar 4:2:2 (2) / planar 4:2:0 (3): convert_yuv888_to_{uyvy,yuyv,yuv422,yuv420}_frame
// (colourspace.c:8184 / 8228 / 8129 / 8035).  One thread = 4 pixels = 2 pairs (mode 3: of two rows); chroma of a pair =
// avg_chroma(first, second), 4:2:0 row k = avg_chroma(row 2k pairs, row 2k + 1 pairs).
__global__ void __launch_bounds__(kBlock) k_yuv888_subsample(int mode, const uint8_t *__restrict__ src, int irow, int src_alpha, OutPlanes4 D,
                                                            int width, int height, const uint8_t *__restrict__ cavg, int vec) {
  const int hw = width >> 1, groups = (hw + 1) >> 1, ips = src_alpha ? 4 : 3;
  const int rows = mode == 3 ? (height + 1) >> 1 : height;
  const long long total = (long long)groups * rows;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int rr = (int)(it / groups), g = (int)(it - (long long)rr * groups);
    const int m = 2 * g, nm = min(2, hw - m);  // pairs m, m + 1
    // one source row -> luma word, (cu0, cu1), (cv0, cv1)
    auto pairs = [&](int row, uint32_t &yw, uint32_t &cu, uint32_t &cv) {
      uint32_t px[4];
      ld_packed4(src + (long long)irow * row + (long long)(2 * m) * ips, px, ips, 2 * nm, vec);
      const uint32_t a01 = __byte_perm(px[0], px[1], 0x5140), b01 = __byte_perm(px[0], px[1], 0x7362);
      const uint32_t a23 = __byte_perm(px[2], px[3], 0x5140), b23 = __byte_perm(px[2], px[3], 0x7362);
      yw = __byte_perm(a01, a23, 0x5410);
      const uint32_t uw = __byte_perm(a01, a23, 0x7632), vw = __byte_perm(b01, b23, 0x5410);
      cu = (uint32_t)__ldg(cavg + ((byte_of(uw, 0) << 8) | byte_of(uw, 1))) | ((uint32_t)__ldg(cavg + ((byte_of(uw, 2) << 8) | byte_of(uw, 3))) << 8);
      cv = (uint32_t)__ldg(cavg + ((byte_of(vw, 0) << 8) | byte_of(vw, 1))) | ((uint32_t)__ldg(cavg + ((byte_of(vw, 2) << 8) | byte_of(vw, 3))) << 8);
    };
    uint32_t yw, cu, cv;
    if (mode <= 1) {
      pairs(rr, yw, cu, cv);
      uint32_t m0 = byte_of(yw, 0) | (byte_of(cu, 0) << 8) | (byte_of(yw, 1) << 16) | (byte_of(cv, 0) << 24);   // YUYV
      uint32_t m1 = byte_of(yw, 2) | (byte_of(cu, 1) << 8) | (byte_of(yw, 3) << 16) | (byte_of(cv, 1) << 24);
      if (mode == 0) { m0 = __byte_perm(m0, 0u, 0x2301); m1 = __byte_perm(m1, 0u, 0x2301); }
      uint8_t *d = D.p[0] + (long long)D.rs[0] * rr + 4LL * m;
      if (vec && nm == 2 && ((D.rs[0] | (int)(uintptr_t)D.p[0]) & 7) == 0) st_stream_u2(d, make_uint2(m0, m1));
      else { st_px4(d, m0, 4, false); if (nm == 2) st_px4(d + 4, m1, 4, false); }
    } else if (mode == 2) {
      pairs(rr, yw, cu, cv);
      st_px4(D.p[0] + (long long)D.rs[0] * rr + 2 * m, yw, 2 * nm, vec);
      st_px4(D.p[1] + (long long)D.rs[1] * rr + m, cu, nm, false);
      st_px4(D.p[2] + (long long)D.rs[2] * rr + m, cv, nm, false);
    } else {
      const int r0 = 2 * rr;
      pairs(r0, yw, cu, cv);
      st_px4(D.p[0] + (long long)D.rs[0] * r0 + 2 * m, yw, 2 * nm, vec);
      if (r0 + 1 < height) {
        uint32_t yw1, cu1, cv1;
        pairs(r0 + 1, yw1, cu1, cv1);
        st_px4(D.p[0] + (long long)D.rs[0] * (r0 + 1) + 2 * m, yw1, 2 * nm, vec);
        cu = avg4(cavg, cu, cu1) & 0xFFFFu; cv = avg4(cavg, cv, cv1) & 0xFFFFu;
      }
      st_px4(D.p[1] + (long long)D.rs[1] * rr + m, cu, nm, false);
      st_px4(D.p[2] + (long long)D.rs[2] * rr + m, cv, nm, false);
    }
  }
}

// UYVY / YUYV -> planar 4:2:0 (convert_{uyvy,yuyv}_to_yuv420_frame, colourspace.c:7887 / 7930): luma split, chroma row k =
// avg_chroma(row 2k, row 2k + 1).  One thread = 2 macropixels of a row pair.
__global__ void __launch_bounds__(kBlock) k_packed422_to_yuv420p(int fmt, const uint8_t *__restrict__ src, int irow, OutPlanes4 D, int width_mpx,
                                                                int height, const uint8_t *__restrict__ cavg, int vec) {
  const int groups = (width_mpx + 1) >> 1, rows = (height + 1) >> 1;
  const long long total = (long long)groups * rows;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int rr = (int)(it / groups), g = (int)(it - (long long)rr * groups);
    const int m = 2 * g, nm = min(2, width_mpx - m);
    auto row2 = [&](int row, uint32_t &yw, uint32_t &cu, uint32_t &cv) {
      const uint8_t *q = src + (long long)irow * row + 4LL * m;
      uint32_t m0 = ld_px4(q, 4, vec), m1 = nm == 2 ? ld_px4(q + 4, 4, vec) : 0u;
      if (fmt == 0) { m0 = __byte_perm(m0, 0u, 0x2301); m1 = __byte_perm(m1, 0u, 0x2301); }  // -> y0 u y1 v
      yw = __byte_perm(m0, m1, 0x6420);
      cu = __byte_perm(m0, m1, 0x4451);   // u, u' in bytes 0, 1 (bytes 2, 3 are not stored)
      cv = __byte_perm(m0, m1, 0x4473);
      st_px4(D.p[0] + (long long)D.rs[0] * row + 2 * m, yw, 2 * nm, vec);
    };
    uint32_t yw, cu, cv;
    row2(2 * rr, yw, cu, cv);
    if (2 * rr + 1 < height) {
      uint32_t yw1, cu1, cv1;
      row2(2 * rr + 1, yw1, cu1, cv1);
      cu = avg4(cavg, cu, cu1); cv = avg4(cavg, cv, cv1);
    }
    st_px4(D.p[1] + (long long)D.rs[1] * rr + m, cu, nm, false);
    st_px4(D.p[2] + (long long)D.rs[2] * rr + m, cv, nm, false);
  }
}

// planar 4:2:0 / 4:2:2 -> YUV888 / YUVA8888 with the chroma up-sampled on the fly (convert_quad_chroma_packed colourspace.c:10715,
// convert_double_chroma_packed :10811): the chroma of k_quad_chroma (4:2:0) or its even-row rule on every row with f on both
// columns (4:2:2), interleaved with the luma.  One thread = 4 pixels of QC_PAIRS consecutive row pairs (4:2:0: the even-row chroma of
// chroma row k + 1 is computed once for rows 2k + 1 and 2k + 2).
struct UpsamplePackedParams {
  const uint8_t *y, *u, *v;
  uint8_t *dst;
  int rs_y, rs_u, rs_v, orow, w2, height, cw, ch, jpeg, is_420, add_alpha;
  AvgForm avg;
};

__global__ void __launch_bounds__(kBlock) k_chroma_upsample_packed(const UpsamplePackedParams P, int vec) {
  const int groups = (P.w2 + 3) >> 2, ps = P.add_alpha ? 4 : 3;
  const int npairs = (P.height + 1) >> 1, nchunks = (npairs + QC_PAIRS - 1) / QC_PAIRS;
  const long long total = (long long)groups * nchunks;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int chunk = (int)(it / groups), g = (int)(it - (long long)chunk * groups);
    const int x = 4 * g, n = min(4, P.w2 - x), m = x >> 1;
    const int k0 = chunk * QC_PAIRS, k1 = min(k0 + QC_PAIRS, npairs);
    auto put = [&](int row, uint32_t uw, uint32_t vw) {
      const uint32_t yw = ld_px4(P.y + (long long)P.rs_y * row + x, n, vec);
      const uint32_t aw = 0xFFFFFFFFu;
      const uint32_t yu_lo = __byte_perm(yw, uw, 0x5140), yu_hi = __byte_perm(yw, uw, 0x7362);
      const uint32_t va_lo = __byte_perm(vw, aw, 0x5140), va_hi = __byte_perm(vw, aw, 0x7362);
      const uint32_t px[4] = {__byte_perm(yu_lo, va_lo, 0x5410), __byte_perm(yu_lo, va_lo, 0x7632), __byte_perm(yu_hi, va_hi, 0x5410),
                              __byte_perm(yu_hi, va_hi, 0x7632)};
      uint8_t *d = P.dst + (long long)P.orow * row + (long long)x * ps;
      if (vec && n == 4 && ps == 4) st_stream_u4(d, make_uint4(px[0], px[1], px[2], px[3]));
      else st_packed4(d, px, ps, n, vec);
    };
    const bool fast = quad_fast_ok(P.u, P.rs_u, m, P.cw, n) && quad_fast_ok(P.v, P.rs_v, m, P.cw, n);
    const bool g_is_f = !P.is_420;
    auto even_u = [&](int k) -> uint32_t {
      return fast ? quad_even4_fast(P.u, P.rs_u, k, m, true, P.jpeg, P.avg, g_is_f) : quad_even4(P.u, P.rs_u, k, m, P.cw, P.ch, true, P.jpeg, P.avg, n, g_is_f);
    };
    auto even_v = [&](int k) -> uint32_t {
      return fast ? quad_even4_fast(P.v, P.rs_v, k, m, false, P.jpeg, P.avg, g_is_f) : quad_even4(P.v, P.rs_v, k, m, P.cw, P.ch, false, P.jpeg, P.avg, n, g_is_f);
    };
    auto put_y = [&](uint8_t *d, uint32_t yw, uint32_t uw, uint32_t vw) {   // vec && n == 4
      const uint32_t yu_lo = __byte_perm(yw, uw, 0x5140), yu_hi = __byte_perm(yw, uw, 0x7362);
      const uint32_t va_lo = __byte_perm(vw, 0xFFFFFFFFu, 0x5140), va_hi = __byte_perm(vw, 0xFFFFFFFFu, 0x7362);
      const uint32_t p0 = __byte_perm(yu_lo, va_lo, 0x5410), p1 = __byte_perm(yu_lo, va_lo, 0x7632), p2 = __byte_perm(yu_hi, va_hi, 0x5410),
                     p3 = __byte_perm(yu_hi, va_hi, 0x7632);
      if (ps == 4) {
        st_stream_u4(d, make_uint4(p0, p1, p2, p3));
      } else {
        st_stream_u32(d, __byte_perm(p0, p1, 0x4210));
        st_stream_u32(d + 4, __byte_perm(p1, p2, 0x5421));
        st_stream_u32(d + 8, __byte_perm(p2, p3, 0x6542));
      }
    };
    if (fast && vec && P.is_420 && 2 * k0 + 2 * QC_PAIRS <= P.height - 2) {
      // a whole chunk away from the frame's edges: every load is issued before the first use
      uint32_t cu[QC_PAIRS + 1], cv[QC_PAIRS + 1], yw[2 * QC_PAIRS];
#pragma unroll
      for (int j = 0; j <= QC_PAIRS; j++) { cu[j] = quad_load4(P.u, P.rs_u, k0 + j, m); cv[j] = quad_load4(P.v, P.rs_v, k0 + j, m); }
#pragma unroll
      for (int j = 0; j < 2 * QC_PAIRS; j++) yw[j] = ld_stream_u32(P.y + (long long)P.rs_y * (2 * k0 + j) + x);
      uint32_t EU = quad_even4_from(cu[0], true, P.jpeg, P.avg), EV = quad_even4_from(cv[0], false, P.jpeg, P.avg);
      uint8_t *d = P.dst + (long long)P.orow * (2 * k0) + (long long)x * ps;
#pragma unroll
      for (int j = 0; j < QC_PAIRS; j++) {
        const uint32_t nu = quad_even4_from(cu[j + 1], true, P.jpeg, P.avg), nv = quad_even4_from(cv[j + 1], false, P.jpeg, P.avg);
        put_y(d, yw[2 * j], EU, EV);
        put_y(d + P.orow, yw[2 * j + 1], avg4(P.avg, nu, EU), avg4(P.avg, nv, EV));
        d += 2 * (long long)P.orow;
        EU = nu; EV = nv;
      }
      continue;
    }
    if (fast && vec && !P.is_420 && 2 * k0 + 2 * QC_PAIRS <= P.height) {
      uint32_t cu[2 * QC_PAIRS], cv[2 * QC_PAIRS], yw[2 * QC_PAIRS];
#pragma unroll
      for (int j = 0; j < 2 * QC_PAIRS; j++) {
        cu[j] = quad_load4(P.u, P.rs_u, 2 * k0 + j, m); cv[j] = quad_load4(P.v, P.rs_v, 2 * k0 + j, m);
        yw[j] = ld_stream_u32(P.y + (long long)P.rs_y * (2 * k0 + j) + x);
      }
      uint8_t *d = P.dst + (long long)P.orow * (2 * k0) + (long long)x * ps;
#pragma unroll
      for (int j = 0; j < 2 * QC_PAIRS; j++) {
        put_y(d, yw[j], quad_even4_from(cu[j], true, P.jpeg, P.avg, true), quad_even4_from(cv[j], false, P.jpeg, P.avg, true));
        d += P.orow;
      }
      continue;
    }
    if (!P.is_420) {
      for (int row = 2 * k0; row < min(2 * k1, P.height); row++) put(row, even_u(row), even_v(row));
      continue;
    }
    uint32_t EU = even_u(k0);
    uint32_t EV = even_v(k0);
    for (int k = k0; k < k1; k++) {
      put(2 * k, EU, EV);
      const int row = 2 * k + 1;
      if (row >= P.height) break;
      uint32_t uw = EU, vw = EV;
      if (row + 1 <= P.height - 1) {
        const uint32_t nu = even_u(k + 1);
        const uint32_t nv = even_v(k + 1);
        const bool swapped = (P.height & 1) && row == P.height - 2;
        uw = swapped ? avg4(P.avg, EU, nu) : avg4(P.avg, nu, EU);
        vw = swapped ? avg4(P.avg, EV, nv) : avg4(P.avg, nv, EV);
        EU = nu; EV = nv;
      }
      put(row, uw, vw);
    }
  }
}

// UYVY <-> YUYV in place: swab() of every row
__global__ void __launch_bounds__(kBlock) k_swab(uint8_t *pix, int rs, int width_mpx, int height, int vec) {
  const long long total = (long long)width_mpx * height;
  for (long long it = global_tid(); it < total; it += global_threads()) {
    const int row = (int)(it / width_mpx), m = (int)(it - (long long)row * width_mpx);
    uint8_t *p = pix + (long long)rs * row + 4LL * m;
    if (vec) *reinterpret_cast<uint32_t *>(p) = __byte_perm(*reinterpret_cast<const uint32_t *>(p), 0u, 0x2301);
    else { const uint8_t a = p[0], b = p[1], c = p[2], d = p[3]; p[0] = b; p[1] = a; p[2] = d; p[3] = c; }
  }
}

// every byte of a plane, walked densely, through the luma or the chroma table.  kind 0 all luma, 1 all chroma, 2 YUV888 (byte i is
// luma when i % 3 == 0), 3 YUVA8888 (i % 4: 0 luma, 1 2 chroma, 3 untouched), 4 UYVY (odd bytes luma), 5 YUYV (even bytes luma)
// rs3 (kind 2 only): 0 = the reference's dense walk (the Y U V phase runs on across the row padding), else the rowstride: the phase
// restarts with every row
__global__ void __launch_bounds__(kBlock) k_clamp_lut(uint8_t *plane, long long nbytes, int kind, int rs3, const uint8_t *__restrict__ ty,
                                                     const uint8_t *__restrict__ tc) {
  __shared__ uint8_t sy[256], sc[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { sy[i] = ty[i]; sc[i] = tc[i]; }
  __syncthreads();
  const long long words = (nbytes + 3) >> 2;  // the plane starts word aligned (checked by the launcher)
  for (long long it = global_tid(); it < words; it += global_threads()) {
    const long long i0 = 4 * it;
    const int n = (int)min(4LL, nbytes - i0);
    uint32_t w = n == 4 ? *reinterpret_cast<const uint32_t *>(plane + i0) : 0u;
    if (n < 4) for (int k = 0; k < n; k++) w |= (uint32_t)plane[i0 + k] << (8 * k);
    const int ph3 = (int)((rs3 ? i0 % rs3 : i0) % 3);  // (rowstrides are multiples of 4: a word never straddles two rows)
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t b = byte_of(w, k);
      int luma;  // 1 luma, 0 chroma, -1 untouched
      switch (kind) {
      case 0: luma = 1; break;
      case 1: luma = 0; break;
      case 2: luma = ((ph3 + k) % 3) == 0; break;
      case 3: luma = k == 3 ? -1 : k == 0; break;
      case 4: luma = k & 1; break;
      default: luma = !(k & 1); break;
      }
      o |= (luma < 0 ? b : luma ? (uint32_t)sy[b] : (uint32_t)sc[b]) << (8 * k);
    }
    if (n == 4) *reinterpret_cast<uint32_t *>(plane + i0) = o;
    else for (int k = 0; k < n; k++) plane[i0 + k] = (uint8_t)(o >> (8 * k));
  }
}

inline bool aligned4(const void *p) { return (((uintptr_t)p) & 3) == 0; }
inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace

cudaError_t launch_yuv444p_to_rgb(const Launch &L, const uint8_t *const planes[4], int irow, Img dst, int width, int height, int in_alpha,
                                  RgbLayout out, DevConv conv) {
  Planes4 S;
  for (int k = 0; k < 4; k++) S.p[k] = planes[k];
  S.rs = irow;
  bool vec = aligned4(planes[0]) && aligned4(planes[1]) && aligned4(planes[2]) && (!in_alpha || aligned4(planes[3])) && !(irow & 3);
  vec = vec && (out.psize == 4 ? aligned16(dst.p) && !(dst.rs & 15) : aligned4(dst.p) && !(dst.rs & 3));
  k_yuv444p_to_rgb<<<grid_for(L, (long long)((width + 3) / 4) * height), kBlock, 0, L.stream>>>(S, dst.p, dst.rs, width, height, in_alpha,
                                                                                                out, conv, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_combine_planes(const Launch &L, const uint8_t *const planes[4], int irow, Img dst, int width, int height, int in_alpha,
                                  int out_alpha) {
  Planes4 S;
  for (int k = 0; k < 4; k++) S.p[k] = planes[k];
  S.rs = irow;
  bool vec = aligned4(planes[0]) && aligned4(planes[1]) && aligned4(planes[2]) && (!in_alpha || aligned4(planes[3])) && !(irow & 3);
  vec = vec && (out_alpha ? aligned16(dst.p) && !(dst.rs & 15) : aligned4(dst.p) && !(dst.rs & 3));
  k_combine_planes<<<grid_for(L, (long long)((width + 3) / 4) * height), kBlock, 0, L.stream>>>(S, dst.p, dst.rs, width, height, in_alpha,
                                                                                                out_alpha, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_split_planes(const Launch &L, CImg src, uint8_t *const planes[4], const int orows[4], int width, int height,
                                int src_alpha, int dest_alpha) {
  OutPlanes4 D;
  bool vec = src_alpha ? aligned16(src.p) && !(src.rs & 15) : aligned4(src.p) && !(src.rs & 3);
  for (int k = 0; k < 4; k++) {
    D.p[k] = planes[k]; D.rs[k] = orows[k];
    if (k < 3 || dest_alpha) vec = vec && aligned4(planes[k]) && !(orows[k] & 3);
  }
  k_split_planes<<<grid_for(L, (long long)((width + 3) / 4) * height), kBlock, 0, L.stream>>>(src.p, src.rs, D, width, height, src_alpha,
                                                                                              dest_alpha, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_resample_chroma_v(const Launch &L, int dbl, const uint8_t *su, const uint8_t *sv, int irs_u, int irs_v, uint8_t *du,
                                     uint8_t *dv, int ors_u, int ors_v, int cw, int ch, const uint8_t *cavg_dev) {
  const bool vec = aligned4(su) && aligned4(sv) && aligned4(du) && aligned4(dv) && !((irs_u | irs_v | ors_u | ors_v) & 3);
  const long long work = (long long)((cw + 3) / 4) * (dbl ? ch : (ch + 1) / 2);
  const dim3 grid(grid_for(L, work), 2);
  if (dbl) k_double_chroma<<<grid, kBlock, 0, L.stream>>>(su, sv, irs_u, irs_v, du, dv, ors_u, ors_v, cw, ch, cavg_dev, vec);
  else k_halve_chroma<<<grid, kBlock, 0, L.stream>>>(su, sv, irs_u, irs_v, du, dv, ors_u, ors_v, cw, ch, cavg_dev, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_packed422_unpack(const Launch &L, int fmt, int mode, CImg src, uint8_t *const planes[4], const int orows[4], int width_mpx,
                                    int height, int add_alpha, int first_only) {
  OutPlanes4 D;
  bool vec = aligned4(src.p) && !(src.rs & 3);
  const int np = mode == 2 ? 1 : mode == 0 ? 3 : (add_alpha ? 4 : 3);
  for (int k = 0; k < 4; k++) {
    D.p[k] = k < np ? planes[k] : nullptr; D.rs[k] = k < np ? orows[k] : 0;
    if (k < np && !(mode == 0 && k > 0)) vec = vec && (mode == 2 && add_alpha ? aligned16(planes[k]) && !(orows[k] & 15) : aligned4(planes[k]) && !(orows[k] & 3));
  }
  k_packed422_unpack<<<grid_for(L, (long long)((width_mpx + 1) / 2) * height), kBlock, 0, L.stream>>>(fmt, mode, src.p, src.rs, D, width_mpx,
                                                                                                      height, add_alpha, first_only, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_yuv444p_to_packed422(const Launch &L, int fmt, const uint8_t *const planes[3], int irow, Img dst, int width_mpx, int height,
                                        const uint8_t *cavg_dev) {
  Planes4 S;
  for (int k = 0; k < 3; k++) S.p[k] = planes[k];
  S.p[3] = nullptr;
  S.rs = irow;
  const bool vec = aligned4(planes[0]) && aligned4(planes[1]) && aligned4(planes[2]) && !(irow & 3) && (((uintptr_t)dst.p | (uint32_t)dst.rs) & 7) == 0;
  k_yuv444p_to_packed422<<<grid_for(L, (long long)((width_mpx + 1) / 2) * height), kBlock, 0, L.stream>>>(fmt, S, dst.p, dst.rs, width_mpx, height,
                                                                                                          cavg_dev, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_yuv444p_to_chroma420(const Launch &L, const uint8_t *su, const uint8_t *sv, int irs, uint8_t *du, uint8_t *dv, int ors_u,
                                        int ors_v, int cw, int height, const uint8_t *cavg_dev) {
  const bool vec = aligned4(su) && aligned4(sv) && aligned4(du) && aligned4(dv) && !((irs | ors_u | ors_v) & 3);
  const dim3 grid(grid_for(L, (long long)((cw + 3) / 4) * ((height + 1) / 2)), 2);
  k_yuv444p_to_chroma420<<<grid, kBlock, 0, L.stream>>>(su, sv, irs, du, dv, ors_u, ors_v, cw, height, cavg_dev, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_planar42x_to_packed422(const Launch &L, int fmt, int is_422, const uint8_t *const planes[3], const int irows[3], Img dst,
                                          int width_mpx, int height) {
  const bool vec = aligned4(planes[0]) && !(irows[0] & 3) && (((uintptr_t)dst.p | (uint32_t)dst.rs) & 7) == 0;
  k_planar42x_to_packed422<<<grid_for(L, (long long)((width_mpx + 1) / 2) * height), kBlock, 0, L.stream>>>(
      fmt, is_422, planes[0], planes[1], planes[2], irows[0], irows[1], irows[2], dst.p, dst.rs, width_mpx, height, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_quad_chroma(const Launch &L, const uint8_t *su, const uint8_t *sv, int irs_u, int irs_v, int ch, uint8_t *du, uint8_t *dv,
                               int ors, int width, int height, int jpeg, int clamped) {
  QuadChromaParams P;
  P.su = su; P.sv = sv; P.du = du; P.dv = dv; P.irs_u = irs_u; P.irs_v = irs_v; P.ors = ors;
  P.w2 = (width >> 1) << 1; P.height = height; P.cw = P.w2 >> 1; P.ch = ch; P.jpeg = jpeg;
  P.avg = avg_form(clamped != 0);
  if (P.w2 < 2 || height < 1) return cudaSuccess;
  const int vec = aligned4(du) && aligned4(dv) && !(ors & 3);
  const dim3 grid(grid_for(L, (long long)((P.w2 + 3) / 4) * ((height + 2 * QC_PAIRS - 1) / (2 * QC_PAIRS))), 2);
  k_quad_chroma<<<grid, kBlock, 0, L.stream>>>(P, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_yuv888_subsample(const Launch &L, int mode, CImg src, int src_alpha, uint8_t *const planes[3], const int orows[3], int width,
                                    int height, const uint8_t *cavg_dev) {
  OutPlanes4 D;
  const int np = mode <= 1 ? 1 : 3;
  bool vec = src_alpha ? aligned16(src.p) && !(src.rs & 15) : aligned4(src.p) && !(src.rs & 3);
  for (int k = 0; k < 4; k++) { D.p[k] = k < np ? planes[k] : nullptr; D.rs[k] = k < np ? orows[k] : 0; }
  vec = vec && aligned4(planes[0]) && !(orows[0] & 3);
  const int hw = width >> 1, rows = mode == 3 ? (height + 1) / 2 : height;
  if (hw < 1 || rows < 1) return cudaSuccess;
  k_yuv888_subsample<<<grid_for(L, (long long)((hw + 1) / 2) * rows), kBlock, 0, L.stream>>>(mode, src.p, src.rs, src_alpha, D, width, height,
                                                                                           cavg_dev, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_packed422_to_yuv420p(const Launch &L, int fmt, CImg src, uint8_t *const planes[3], const int orows[3], int width_mpx,
                                        int height, const uint8_t *cavg_dev) {
  OutPlanes4 D;
  for (int k = 0; k < 4; k++) { D.p[k] = k < 3 ? planes[k] : nullptr; D.rs[k] = k < 3 ? orows[k] : 0; }
  const bool vec = aligned4(src.p) && !(src.rs & 3) && aligned4(planes[0]) && !(orows[0] & 3);
  if (width_mpx < 1 || height < 1) return cudaSuccess;
  k_packed422_to_yuv420p<<<grid_for(L, (long long)((width_mpx + 1) / 2) * ((height + 1) / 2)), kBlock, 0, L.stream>>>(fmt, src.p, src.rs, D, width_mpx,
                                                                                                                    height, cavg_dev, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_chroma_upsample_packed(const Launch &L, int is_420, const uint8_t *const planes[3], const int irows[3], int ch, Img dst,
                                          int width, int height, int add_alpha, int jpeg, int clamped) {
  UpsamplePackedParams P;
  P.y = planes[0]; P.u = planes[1]; P.v = planes[2]; P.dst = dst.p; P.rs_y = irows[0]; P.rs_u = irows[1]; P.rs_v = irows[2]; P.orow = dst.rs;
  P.w2 = (width >> 1) << 1; P.height = height; P.cw = P.w2 >> 1; P.ch = ch; P.jpeg = jpeg; P.is_420 = is_420; P.add_alpha = add_alpha;
  P.avg = avg_form(clamped != 0);
  if (P.w2 < 2 || height < 1) return cudaSuccess;
  const int vec = aligned4(planes[0]) && !(irows[0] & 3) && (add_alpha ? aligned16(dst.p) && !(dst.rs & 15) : aligned4(dst.p) && !(dst.rs & 3));
  k_chroma_upsample_packed<<<grid_for(L, (long long)((P.w2 + 3) / 4) * ((height + 2 * QC_PAIRS - 1) / (2 * QC_PAIRS))), kBlock, 0, L.stream>>>(P, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_swab(const Launch &L, Img img, int width_mpx, int height) {
  const bool vec = aligned4(img.p) && !(img.rs & 3);
  k_swab<<<grid_for(L, (long long)width_mpx * height), kBlock, 0, L.stream>>>(img.p, img.rs, width_mpx, height, vec);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

cudaError_t launch_clamp_lut(const Launch &L, uint8_t *plane, long long nbytes, int kind, int row_phase_stride, const uint8_t *ty_dev,
                             const uint8_t *tc_dev) {
  if (!aligned4(plane) || (row_phase_stride & 3)) return cudaErrorMisalignedAddress;
  if (nbytes <= 0) return cudaSuccess;
  k_clamp_lut<<<grid_for(L, (nbytes + 3) / 4), kBlock, 0, L.stream>>>(plane, nbytes, kind, kind == 2 ? row_phase_stride : 0, ty_dev, tc_dev);
  PE_COUNT_LAUNCH(L);
  return cudaGetLastError();
}

}  // namespace pe
