// pe_engine.cu -- the C ABI of include/pixel_engine.h: engine, device frames, and the host-side logic of the
// reference's layer ops (which converter runs, which tables / LUT it gets, how the layer's metadata changes),
// restated from src/colourspace.c with file:line citations.  All pixel arithmetic lives in pe_kernels_*.cu.
//
// There is deliberately NO CPU fallback: every op either launches a kernel on the engine's stream or fails with
// PE_FALSE / PE_ERR_* and a message in pe_last_error().
#include "pe_engine.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace pe;

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------

static thread_local char g_err[512] = "";

static int set_err(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define PE_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t err__ = (call);                                                                    \
    if (err__ != cudaSuccess)                                                                      \
      return set_err(PE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

// for the boolean-returning layer ops
#define PE_CUDA_B(call)                                                                            \
  do {                                                                                             \
    cudaError_t err__ = (call);                                                                    \
    if (err__ != cudaSuccess) {                                                                    \
      set_err(PE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
      return PE_FALSE;                                                                             \
    }                                                                                              \
  } while (0)

extern "C" const char *pe_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------------------
// palettes (libweed/weed-palettes.h, weed_palette_* helpers of libweed/weed-utils.c)
// ---------------------------------------------------------------------------------------------------------

namespace {

inline bool pal_is_rgb(int p) { return p >= PE_PALETTE_RGB24 && p <= PE_PALETTE_ARGB32; }
inline bool pal_is_yuv(int p) { return p >= 512 && p < 1024; }
inline bool pal_has_alpha(int p) {
  return p == PE_PALETTE_RGBA32 || p == PE_PALETTE_BGRA32 || p == PE_PALETTE_ARGB32 || p == PE_PALETTE_YUVA8888 ||
         p == PE_PALETTE_YUVA4444P;
}
inline bool pal_known(int p) {
  switch (p) {
  case PE_PALETTE_RGB24: case PE_PALETTE_BGR24: case PE_PALETTE_RGBA32: case PE_PALETTE_BGRA32: case PE_PALETTE_ARGB32:
  case PE_PALETTE_YUV420P: case PE_PALETTE_YVU420P: case PE_PALETTE_YUV422P: case PE_PALETTE_YUV444P:
  case PE_PALETTE_YUVA4444P: case PE_PALETTE_UYVY: case PE_PALETTE_YUYV: case PE_PALETTE_YUV888: case PE_PALETTE_YUVA8888:
  case PE_PALETTE_YUV411:
    return true;
  }
  return false;
}
inline int pal_nplanes(int p) {
  switch (p) {
  case PE_PALETTE_YUV420P: case PE_PALETTE_YVU420P: case PE_PALETTE_YUV422P: case PE_PALETTE_YUV444P: return 3;
  case PE_PALETTE_YUVA4444P: return 4;
  default: return 1;
  }
}
inline bool pal_is_planar(int p) { return pal_nplanes(p) > 1; }
// bytes per macropixel of plane 0
inline int pal_psize(int p) {
  switch (p) {
  case PE_PALETTE_RGB24: case PE_PALETTE_BGR24: case PE_PALETTE_YUV888: return 3;
  case PE_PALETTE_RGBA32: case PE_PALETTE_BGRA32: case PE_PALETTE_ARGB32: case PE_PALETTE_YUVA8888:
  case PE_PALETTE_UYVY: case PE_PALETTE_YUYV: return 4;
  case PE_PALETTE_YUV411: return 6;
  default: return 1;
  }
}
inline int pal_ppmp(int p) { return (p == PE_PALETTE_UYVY || p == PE_PALETTE_YUYV) ? 2 : p == PE_PALETTE_YUV411 ? 4 : 1; }

inline RgbLayout rgb_layout(int p) {
  switch (p) {
  case PE_PALETTE_RGB24: return {0, 1, 2, -1, 3};
  case PE_PALETTE_BGR24: return {2, 1, 0, -1, 3};
  case PE_PALETTE_RGBA32: return {0, 1, 2, 3, 4};
  case PE_PALETTE_BGRA32: return {2, 1, 0, 3, 4};
  default: return {1, 2, 3, 0, 4};  // ARGB32
  }
}

inline int align_ceil(int a, int b) { return (a + b - 1) / b * b; }

// can_inline_gamma colourspace.c:12128
bool can_inline_gamma(int inpl, int opal) {
  if (pal_is_rgb(inpl) && pal_is_rgb(opal)) return true;
  if ((inpl == PE_PALETTE_YUV420P || inpl == PE_PALETTE_YVU420P || inpl == PE_PALETTE_YUV422P || inpl == PE_PALETTE_YUV444P) &&
      pal_is_rgb(opal))
    return true;
  if (pal_is_rgb(opal)) return true;
  if (opal == PE_PALETTE_UYVY || opal == PE_PALETTE_YUYV) {
    if (pal_is_rgb(inpl) || inpl == PE_PALETTE_UYVY || inpl == PE_PALETTE_YUYV) return true;
  }
  return false;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// device block pool
// ---------------------------------------------------------------------------------------------------------

void *DevPool::get(size_t bytes, size_t *granted) {
  const size_t kClass = 64 * 1024;
  const size_t want = (bytes + kClass - 1) / kClass * kClass;
  auto it = free_.find(want);
  if (it != free_.end()) {
    void *p = it->second;
    free_.erase(it);
    held_ -= want;
    *granted = want;
    return p;
  }
  void *p = nullptr;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    release_all();  // give cached blocks back to the driver and retry once
    if (cudaMalloc(&p, want) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
  }
  *granted = want;
  return p;
}

bool DevPool::undefer(void *p) {
  for (size_t i = 0; i < deferred_.size(); i++)
    if (deferred_[i].first == p) { deferred_.erase(deferred_.begin() + (long)i); return true; }
  return false;
}

void DevPool::flush_deferred() {
  std::vector<std::pair<void *, size_t>> d;
  d.swap(deferred_);
  for (auto &b : d) put(b.first, b.second);
}

void DevPool::put(void *p, size_t granted) {
  if (!p) return;
  if (defer_) { deferred_.emplace_back(p, granted); return; }
  const size_t kMaxHeld = (size_t)8 << 30;  // 8 GiB of recycled blocks at most (HBM is 180 GB)
  if (held_ + granted > kMaxHeld) {
    cudaFree(p);
    return;
  }
  free_.emplace(granted, p);
  held_ += granted;
}

void DevPool::release_all() {
  for (auto &kv : free_) cudaFree(kv.second);
  free_.clear();
  held_ = 0;
}

// ---------------------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------------------

extern "C" void pe_config_default(pe_config_t *cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->device = 0;
  cfg->pb_quality = PE_QUALITY_HIGH;  // render / transcode force HIGH (colourspace.c:2103-2114)
  cfg->screen_gamma = 1.4;            // DEF_SCREEN_GAMMA
  cfg->apply_gamma = 1;
  cfg->alpha_post = 0;
  cfg->ref_quirks = 1;
  cfg->stream = nullptr;
}

static int engine_init(pe_engine *e);

extern "C" int pe_engine_create(const pe_config_t *cfg, pe_engine_t **out) {
  if (!out) return set_err(PE_ERR_ARG, "pe_engine_create: out is NULL");
  *out = nullptr;
  pe_config_t c;
  if (cfg) c = *cfg; else pe_config_default(&c);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return set_err(PE_ERR_CUDA, "no usable CUDA device (%s); the pixel engine has no CPU fallback",
                   ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
  }
  if (c.device < 0 || c.device >= ndev) return set_err(PE_ERR_ARG, "device %d out of range (0..%d)", c.device, ndev - 1);
  PE_CUDA(cudaSetDevice(c.device));
  pe_engine *e = new pe_engine();
  e->cfg = c;
  e->device = c.device;
  const int rc = engine_init(e);
  if (rc != PE_OK) {  // stream, events and tables created so far go the way of a finished engine
    pe_engine_destroy(e);
    return rc;
  }
  *out = e;
  return PE_OK;
}

static int engine_init(pe_engine *e) {
  const pe_config_t &c = e->cfg;
  cudaDeviceProp prop;
  PE_CUDA(cudaGetDeviceProperties(&prop, c.device));
  e->sm_count = prop.multiProcessorCount;
  e->no_host_staging = getenv("PE_HOST_NO_STAGING") != nullptr;
  if (c.stream) {
    e->stream = (cudaStream_t)c.stream;
  } else {
    PE_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->own_stream = true;
  }
  PE_CUDA(cudaEventCreate(&e->ev0));
  PE_CUDA(cudaEventCreate(&e->ev1));
  PE_CUDA(cudaEventCreateWithFlags(&e->args_ev, cudaEventDisableTiming));
  // init_colour_engine (colourspace.c:1973): every (clamping, subspace) variant up front
  for (int cl = 0; cl < 2; cl++) {
    for (int hd = 0; hd < 2; hd++) {
      build_conv_tables(cl == 0 ? PE_YUV_CLAMPING_CLAMPED : PE_YUV_CLAMPING_UNCLAMPED,
                        hd ? PE_YUV_SUBSPACE_BT709 : PE_YUV_SUBSPACE_YCBCR, &e->conv_host[cl][hd]);
      // [14][256] as the reference has them + the extended planar YUV -> RGB tables (DevConv::ext, pe_kernels.h)
      std::vector<int32_t> host(N_CONVTAB * 256 + 256 + 4 * kExtN);
      const ConvTables &ct = e->conv_host[cl][hd];
      memcpy(host.data(), ct.t, sizeof(int32_t) * N_CONVTAB * 256);
      int32_t *ext = host.data() + N_CONVTAB * 256;
      memcpy(ext, ct.t[RGB_Y], sizeof(int32_t) * 256);
      const int lo = cl == 0 ? 16 : 0, hi = cl == 0 ? 240 : 255;
      for (int n = 0; n < kExtN; n++) {
        int c = (2 * n + 3) / 6;  // (int)(n / 3. + .5)
        c = c < lo ? lo : c > hi ? hi : c;
        ext[256 + 0 * kExtN + n] = ct.t[R_CR][c];
        ext[256 + 1 * kExtN + n] = ct.t[G_CB][c];
        ext[256 + 2 * kExtN + n] = ct.t[G_CR][c];
        ext[256 + 3 * kExtN + n] = ct.t[B_CB][c];
      }
      PE_CUDA(cudaMalloc(&e->conv_dev[cl][hd], sizeof(int32_t) * host.size()));
      // (on the engine stream: a synchronous cudaMemcpy from pageable memory returns once the bytes are staged, the DMA itself is
      // ordered on the legacy stream only, and the engine's stream is non-blocking -- a kernel could read the table before it lands)
      PE_CUDA(cudaMemcpyAsync(e->conv_dev[cl][hd], host.data(), sizeof(int32_t) * host.size(), cudaMemcpyHostToDevice, e->stream));
      PE_CUDA(cudaStreamSynchronize(e->stream));
    }
  }
  {
    int32_t luma[3][256];
    build_plugin_luma_tables(luma[0], luma[1], luma[2]);
    PE_CUDA(cudaMalloc(&e->luma_dev, sizeof(luma)));
    PE_CUDA(cudaMemcpyAsync(e->luma_dev, luma, sizeof(luma), cudaMemcpyHostToDevice, e->stream));
    PE_CUDA(cudaStreamSynchronize(e->stream));  // `luma` is a stack buffer
  }
  PE_CUDA(cudaMalloc(&e->stats_dev, sizeof(DevStats)));
  return PE_OK;
}

extern "C" void pe_engine_destroy(pe_engine_t *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream || !e->own_stream) cudaStreamSynchronize(e->stream);
  for (int cl = 0; cl < 2; cl++)
    for (int hd = 0; hd < 2; hd++) cudaFree(e->conv_dev[cl][hd]);
  for (int k = 0; k < 6; k++) cudaFree(e->premult_dev[k]);
  for (int k = 0; k < 2; k++) cudaFree(e->cavg_dev[k]);
  cudaFree(e->yy_dev);
  cudaFree(e->ftab_dev[0]);
  cudaFree(e->ftab_dev[1]);
  cudaFree(e->luma_dev);
  for (auto &kv : e->lut8) cudaFree(kv.second.dev);
  for (auto &kv : e->lut16) cudaFree(kv.second);
  for (auto &kv : e->over) cudaFree(kv.second.dev);
  for (auto &kv : e->filters) { cudaFree((void *)kv.second.dev.first); cudaFree((void *)kv.second.dev.coef); cudaFree(kv.second.rows4); cudaFree(kv.second.pack4); }
  e->pool.release_all();
  cudaFree(e->stats_dev);
  cudaFree(e->args_dev);
  cudaFree(e->f3_sched);
  for (int k = 0; k < 4; k++) { if (e->fan_stream[k]) cudaStreamDestroy(e->fan_stream[k]); if (e->fan_join[k]) cudaEventDestroy(e->fan_join[k]); }
  if (e->fan_fork) cudaEventDestroy(e->fan_fork);
  if (e->args_pinned) cudaFreeHost(e->args_pinned);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->args_ev) cudaEventDestroy(e->args_ev);
  if (e->h2d_stream) cudaStreamDestroy(e->h2d_stream);
  if (e->d2h_stream) cudaStreamDestroy(e->d2h_stream);
  for (int k = 0; k < 6; k++) {
    if (e->pipe_up[k]) cudaEventDestroy(e->pipe_up[k]);
    if (e->pipe_comp[k]) cudaEventDestroy(e->pipe_comp[k]);
    if (e->pipe_free[k]) cudaEventDestroy(e->pipe_free[k]);
  }
  if (e->egress_stream) cudaStreamDestroy(e->egress_stream);
  for (int k = 0; k < 4; k++) { if (e->egress_ready[k]) cudaEventDestroy(e->egress_ready[k]); if (e->egress_done[k]) cudaEventDestroy(e->egress_done[k]); }
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

extern "C" int pe_engine_sync(pe_engine_t *e) {
  if (!e) return set_err(PE_ERR_ARG, "engine is NULL");
  PE_CUDA(cudaStreamSynchronize(e->stream));
  return PE_OK;
}

extern "C" void *pe_engine_stream(pe_engine_t *e) { return e ? (void *)e->stream : nullptr; }
extern "C" long pe_engine_launch_count(pe_engine_t *e) { return e ? e->launches : 0; }
extern "C" int pe_sm_count(pe_engine_t *e) { return e ? e->sm_count : 0; }
extern "C" int pe_engine_set_sm_limit(pe_engine_t *e, int n) {
  if (!e || n < 0) return set_err(PE_ERR_ARG, "pe_engine_set_sm_limit: engine and n >= 0");
  std::lock_guard<std::mutex> lk(e->mu);
  e->sm_limit = n;
  return PE_OK;
}

// which coefficients the resize / letterbox / fused calls of this engine use from now on: 1 libswscale's recipes (default; one bank
// per LiVESInterpType, pe_tables.cpp build_resize_filter_sws), 0 the round-1 triangle contract (every interpolation type)
extern "C" int pe_engine_set_resize_recipe(pe_engine_t *e, int recipe) {
  if (!e || recipe < 0 || recipe > 1) return set_err(PE_ERR_ARG, "resize recipe: 0 or 1");
  std::lock_guard<std::mutex> lk(e->mu);
  e->resize_recipe = recipe;
  return PE_OK;
}

// the coefficient bank the engine would build, on the host (no GPU involved): first[dst_n], coefs[dst_n * max_taps]; returns the
// tap count or -1
extern "C" int pe_host_parallel_copy2d(void *dst, size_t dst_stride, const void *src, size_t src_stride, size_t row_bytes, size_t rows, int threads) {
  if (!dst || !src || threads < 1 || threads > 64) return set_err(PE_ERR_ARG, "bad argument");
  pe::CopyPool pool(threads);
  pool.copy2d(dst, dst_stride, src, src_stride, row_bytes, rows);
  return PE_OK;
}

extern "C" int pe_avg_closed_form(int clamped, uint32_t out[5]) {
  const AvgForm F = avg_form(clamped != 0);
  if (out) { out[0] = F.A; out[1] = F.B; out[2] = F.M; out[3] = (uint32_t)F.lo; out[4] = (uint32_t)F.hi; }
  std::vector<uint8_t> t(65536);
  build_avg_table(clamped != 0, t.data());
  return avg_form_matches(clamped != 0, t.data()) ? 1 : 0;
}

extern "C" int pe_resize_filter_host(int recipe, int src_n, int dst_n, int shift_bits, int32_t *first, int16_t *coefs, int max_taps) {
  ResizeFilter f;
  if (!first || !coefs) return -1;
  if (recipe < 0 || recipe > 5) return -1;  // 0 triangle, 1 .. 5 = pe::SwsKind (bilinear, bicubic, Lanczos, fast vertical / horizontal)
  if (!(recipe ? build_resize_filter_sws(src_n, dst_n, shift_bits, &f, (SwsKind)recipe) : build_resize_filter(src_n, dst_n, shift_bits, &f))) return -1;
  if (f.taps > max_taps) return -1;
  for (int i = 0; i < dst_n; i++) {
    first[i] = f.first[i];
    for (int k = 0; k < max_taps; k++) coefs[(size_t)i * max_taps + k] = k < f.taps ? f.coef[(size_t)i * f.taps + k] : 0;
  }
  return f.taps;
}

extern "C" int pe_timer_start(pe_engine_t *e) {
  if (!e) return set_err(PE_ERR_ARG, "engine is NULL");
  PE_CUDA(cudaEventRecord(e->ev0, e->stream));
  return PE_OK;
}

extern "C" int pe_timer_stop_ms(pe_engine_t *e, float *ms) {
  if (!e || !ms) return set_err(PE_ERR_ARG, "NULL argument");
  PE_CUDA(cudaEventRecord(e->ev1, e->stream));
  PE_CUDA(cudaEventSynchronize(e->ev1));
  PE_CUDA(cudaEventElapsedTime(ms, e->ev0, e->ev1));
  return PE_OK;
}

// ---- engine-internal helpers ----------------------------------------------------------------------------

namespace {

inline const ConvTables &conv_host(pe_engine *e, int clamping, int subspace) {
  return e->conv_host[clamping == PE_YUV_CLAMPING_CLAMPED ? 0 : 1][subspace == PE_YUV_SUBSPACE_BT709 ? 1 : 0];
}
inline DevConv dev_conv(pe_engine *e, int clamping, int subspace) {
  const int cl = clamping == PE_YUV_CLAMPING_CLAMPED ? 0 : 1, hd = subspace == PE_YUV_SUBSPACE_BT709 ? 1 : 0;
  const ConvTables &h = e->conv_host[cl][hd];
  return DevConv{e->conv_dev[cl][hd], h.min_y, h.max_y, h.min_uv, h.max_uv, e->conv_dev[cl][hd] + N_CONVTAB * 256};
}

// create_gamma_lut8 (colourspace.c:655).  Our cache is keyed by the ARGUMENTS; the reference keys its process-wide
// cache by a value it mutates while building (:701,:721-733) which makes its result depend on call history --
// DESIGN.md quirk table, X.
Lut8Entry *get_lut8(pe_engine *e, double fileg, int from, int to) {
  GammaKey k{fileg, from, to};
  auto it = e->lut8.find(k);
  if (it != e->lut8.end()) return &it->second;
  Lut8Entry ent;
  if (!build_gamma_lut8(fileg, from, to, e->cfg.screen_gamma, ent.host)) return nullptr;
  if (cudaMalloc(&ent.dev, 256) != cudaSuccess) return nullptr;
  auto ins = e->lut8.emplace(k, ent).first;
  // source is the map node (stable address), so the async copy may outlive this call
  if (cudaMemcpyAsync(ins->second.dev, ins->second.host, 256, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) return nullptr;
  return &ins->second;
}

uint16_t *get_lut16(pe_engine *e, double fileg, int from, int to) {
  GammaKey k{fileg, from, to};
  auto it = e->lut16.find(k);
  if (it != e->lut16.end()) return it->second;
  std::vector<uint16_t> host(65536);
  if (!build_gamma_lut16(fileg, from, to, e->cfg.screen_gamma, host.data())) return nullptr;
  uint16_t *dev = nullptr;
  if (cudaMalloc(&dev, 65536 * 2) != cudaSuccess) return nullptr;
  if (cudaMemcpyAsync(dev, host.data(), 65536 * 2, cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
      cudaStreamSynchronize(e->stream) != cudaSuccess) {
    cudaFree(dev);
    return nullptr;
  }
  e->lut16.emplace(k, dev);
  return dev;
}

uint8_t *get_premult(pe_engine *e, int which) {
  if (e->premult_dev[which]) return e->premult_dev[which];
  std::vector<uint8_t> host(65536);
  build_premult_table(which, host.data());
  uint8_t *dev = nullptr;
  if (cudaMalloc(&dev, 65536) != cudaSuccess) return nullptr;
  if (cudaMemcpyAsync(dev, host.data(), 65536, cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
      cudaStreamSynchronize(e->stream) != cudaSuccess) {
    cudaFree(dev);
    return nullptr;
  }
  e->premult_dev[which] = dev;
  return dev;
}

uint8_t *get_cavg(pe_engine *e, bool clamped) {
  const int which = clamped ? 0 : 1;
  if (e->cavg_dev[which]) return e->cavg_dev[which];
  std::vector<uint8_t> host(65536);
  build_avg_table(clamped, host.data());
  if (!avg_form_matches(clamped, host.data())) {   // k_quad_chroma / k_chroma_upsample_packed compute avg_chroma in closed form
    set_err(PE_ERR_ARG, "the averaging table left its closed form");
    return nullptr;
  }
  uint8_t *dev = nullptr;
  if (cudaMalloc(&dev, 65536) != cudaSuccess) return nullptr;
  if (cudaMemcpyAsync(dev, host.data(), 65536, cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
      cudaStreamSynchronize(e->stream) != cudaSuccess) {
    cudaFree(dev);
    return nullptr;
  }
  e->cavg_dev[which] = dev;
  return dev;
}

// 64 KB [bg][fg] table of paint_pixel (compositor.c:120) for one scalar alpha, optionally composed with a gamma LUT
uint8_t *get_over_table(pe_engine *e, double alpha, const uint8_t *lut_dev) {
  OverKey k{alpha, lut_dev};
  auto it = e->over.find(k);
  if (it != e->over.end()) { it->second.tick = ++e->cache_tick; return it->second.dev; }
  uint8_t *dev = nullptr;
  // bounded (an animated non-dyadic alpha would otherwise grow HBM without end): past kMaxOver entries the least recently used
  // table is rebuilt in place for the new key -- stream ordered behind the paints that read it, no cudaMalloc / cudaFree
  const size_t kMaxOver = 64;
  if (e->over.size() >= kMaxOver) {
    auto lru = e->over.begin();
    for (auto j = e->over.begin(); j != e->over.end(); ++j)
      if (j->second.tick < lru->second.tick) lru = j;
    dev = lru->second.dev;
    e->over.erase(lru);
  } else if (cudaMalloc(&dev, 65536) != cudaSuccess) {
    return nullptr;
  }
  if (launch_over_table(e->L(), alpha, lut_dev, dev) != cudaSuccess) {
    cudaFree(dev);
    return nullptr;
  }
  e->over.emplace(k, pe::OverEntry{dev, ++e->cache_tick});
  return dev;
}

// kind: 0 the round-1 triangle bank, else a pe::SwsKind
struct AxisKinds { int x, y; };
// the bank each axis of a resize takes: recipe 0 -> the round-1 triangle whatever the interpolation; recipe 1 (default) -> the
// flag resize_layer_full hands libswscale (:14991-14997): NORMAL SWS_BILINEAR, BEST SWS_LANCZOS when the frame grows in either
// direction else SWS_BICUBIC, FAST SWS_FAST_BILINEAR (a position walk horizontally, a two-tap bank vertically)
AxisKinds filter_kinds(const pe_engine *e, int interp, int sw, int sh, int dw, int dh) {
  if (!e->resize_recipe) return AxisKinds{0, 0};
  int kx = SWS_KIND_BILINEAR, ky = SWS_KIND_BILINEAR;
  if (interp == PE_INTERP_BEST) kx = ky = (dw > sw || dh > sh) ? SWS_KIND_LANCZOS : SWS_KIND_BICUBIC;
  else if (interp == PE_INTERP_FAST) { kx = SWS_KIND_FAST_H; ky = SWS_KIND_FAST_V; }
  if (sw == dw) kx = SWS_KIND_BILINEAR;  // an unscaled axis is the identity under every flag
  if (sh == dh) ky = SWS_KIND_BILINEAR;
  return AxisKinds{kx, ky};
}

DevFilterEntry *get_filter(pe_engine *e, int src_n, int dst_n, int bits, int kind) {
  FilterKey k{src_n, dst_n, bits | (kind << 8)};
  auto it = e->filters.find(k);
  if (it != e->filters.end()) { it->second.tick = ++e->cache_tick; return &it->second; }
  const size_t kMaxFilters = 256;  // bounded: a zoom animates through many geometries; callers hold at most a handful of entries at once
  if (e->filters.size() >= kMaxFilters) {
    auto lru = e->filters.begin();
    for (auto j = e->filters.begin(); j != e->filters.end(); ++j)
      if (j->second.tick < lru->second.tick) lru = j;
    cudaStreamSynchronize(e->stream);  // kernels that read the bank have finished
    cudaFree((void *)lru->second.dev.first); cudaFree((void *)lru->second.dev.coef); cudaFree(lru->second.rows4); cudaFree(lru->second.pack4);
    e->filters.erase(lru);
  }
  DevFilterEntry ent;
  ent.tick = ++e->cache_tick;
  if (!(kind ? build_resize_filter_sws(src_n, dst_n, bits, &ent.host, (SwsKind)kind) : build_resize_filter(src_n, dst_n, bits, &ent.host)))
    return nullptr;
  int32_t *first = nullptr;
  int16_t *coef = nullptr;
  if (cudaMalloc(&first, sizeof(int32_t) * dst_n) != cudaSuccess) return nullptr;
  if (cudaMalloc(&coef, sizeof(int16_t) * (size_t)dst_n * ent.host.taps) != cudaSuccess) { cudaFree(first); return nullptr; }
  auto ins = e->filters.emplace(k, std::move(ent)).first;
  DevFilterEntry &r = ins->second;
  r.dev = DevFilter{first, coef, r.host.taps, r.host.nonneg() ? 1 : 0};
  if (cudaMemcpyAsync(first, r.host.first.data(), sizeof(int32_t) * dst_n, cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
      cudaMemcpyAsync(coef, r.host.coef.data(), sizeof(int16_t) * (size_t)dst_n * r.host.taps, cudaMemcpyHostToDevice,
                      e->stream) != cudaSuccess) {
    cudaStreamSynchronize(e->stream);  // no entry with unwritten device data stays behind
    cudaFree(first); cudaFree(coef);
    e->filters.erase(ins);
    return nullptr;
  }
  return &r;
}

// upload a small per-launch argument array (BlendFrame / FusedArgs) through a pinned staging buffer; on failure the
// reason is left in pe_last_error()
void *upload_args(pe_engine *e, const void *src, size_t bytes) {
  cudaError_t ce;
  if (bytes > e->args_cap) {
    if (e->args_dev) { cudaStreamSynchronize(e->stream); cudaFree(e->args_dev); }
    size_t cap = bytes < 4096 ? 4096 : bytes * 2;
    if ((ce = cudaMalloc(&e->args_dev, cap)) != cudaSuccess) {
      set_err(PE_ERR_CUDA, "argument buffer cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(ce));
      e->args_dev = nullptr; e->args_cap = 0;
      return nullptr;
    }
    e->args_cap = cap;
  }
  if (bytes > e->args_pinned_cap) {
    if (e->args_pinned) { cudaEventSynchronize(e->args_ev); cudaFreeHost(e->args_pinned); }
    size_t cap = bytes < 4096 ? 4096 : bytes * 2;
    if ((ce = cudaMallocHost(&e->args_pinned, cap)) != cudaSuccess) {
      set_err(PE_ERR_CUDA, "argument staging cudaMallocHost(%zu) failed: %s", cap, cudaGetErrorString(ce));
      e->args_pinned = nullptr; e->args_pinned_cap = 0;
      return nullptr;
    }
    e->args_pinned_cap = cap;
  } else if ((ce = cudaEventSynchronize(e->args_ev)) != cudaSuccess) {  // previous upload out of the staging buffer is done
    set_err(PE_ERR_CUDA, "an earlier kernel on the engine stream failed: %s", cudaGetErrorString(ce));
    return nullptr;
  }
  memcpy(e->args_pinned, src, bytes);
  // kernels of earlier launches that read args_dev are ordered before this copy on the same stream
  if ((ce = cudaMemcpyAsync(e->args_dev, e->args_pinned, bytes, cudaMemcpyHostToDevice, e->stream)) != cudaSuccess) {
    set_err(PE_ERR_CUDA, "argument upload failed: %s", cudaGetErrorString(ce));
    return nullptr;
  }
  cudaEventRecord(e->args_ev, e->stream);
  return e->args_dev;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// frames
// ---------------------------------------------------------------------------------------------------------

// calc_rowstrides colourspace.c:11252-11365 with RS_ALIGN_DEF = 32 (colourspace.h:45)
extern "C" size_t pe_frame_layout(int palette, int width, int height, int *nplanes, int rowstrides[PE_MAXPLANES],
                                  int plane_heights[PE_MAXPLANES]) {
  if (!pal_known(palette) || width <= 0 || height <= 0) return 0;
  const int np = pal_nplanes(palette);
  int rs[PE_MAXPLANES] = {0, 0, 0, 0}, ph[PE_MAXPLANES] = {0, 0, 0, 0};
  rs[0] = align_ceil((width / pal_ppmp(palette)) * pal_psize(palette), 32);
  ph[0] = height;
  switch (palette) {
  case PE_PALETTE_YUV420P: case PE_PALETTE_YVU420P:
    rs[1] = rs[2] = rs[0] >> 1; ph[1] = ph[2] = (height + 1) >> 1; break;
  case PE_PALETTE_YUV422P:
    rs[1] = rs[2] = rs[0] >> 1; ph[1] = ph[2] = height; break;
  case PE_PALETTE_YUV444P:
    rs[1] = rs[2] = rs[0]; ph[1] = ph[2] = height; break;
  case PE_PALETTE_YUVA4444P:
    rs[1] = rs[2] = rs[3] = rs[0]; ph[1] = ph[2] = ph[3] = height; break;
  default: break;
  }
  size_t total = 0;
  for (int p = 0; p < np; p++) total += ((size_t)rs[p] * ph[p] + 255) / 256 * 256;  // planes start 256-byte aligned
  if (nplanes) *nplanes = np;
  for (int p = 0; p < PE_MAXPLANES; p++) {
    if (rowstrides) rowstrides[p] = rs[p];
    if (plane_heights) plane_heights[p] = ph[p];
  }
  return total;
}

namespace {

// allocate pixel memory for f->d.{palette,width,height}; fills nplanes / rowstrides / planes
int frame_alloc(pe_engine *e, pe_frame *f) {
  int np = 0;
  const size_t bytes = pe_frame_layout(f->d.palette, f->d.width, f->d.height, &np, f->d.rowstrides, f->plane_heights);
  if (!bytes) return set_err(PE_ERR_ARG, "bad frame geometry: palette %d, %d x %d", f->d.palette, f->d.width, f->d.height);
  // + 16 guard bytes: vector tails and the one-past-row chroma read of the last plane stay inside the block
  f->base = e->pool.get(bytes + 16, &f->granted);
  if (!f->base) return set_err(PE_ERR_MEMORY, "device allocation of %zu bytes failed", bytes);
  f->d.nplanes = np;
  size_t off = 0;
  for (int p = 0; p < PE_MAXPLANES; p++) {
    if (p < np) {
      f->d.planes[p] = (uint8_t *)f->base + off;
      off += ((size_t)f->d.rowstrides[p] * f->plane_heights[p] + 255) / 256 * 256;
    } else {
      f->d.planes[p] = nullptr;
    }
  }
  return PE_OK;
}

void frame_release_pixels(pe_frame *f) {
  if (f->base) f->e->pool.put(f->base, f->granted);
  f->base = nullptr;
  f->granted = 0;
  for (int p = 0; p < PE_MAXPLANES; p++) f->d.planes[p] = nullptr;
}

// black per create_empty_pixel_data (colourspace.c:11448-11449, :11213 blank_frame): RGB 0 with opaque alpha,
// YUV 16 (clamped) or 0 / 128 / 128
int frame_fill_black(pe_engine *e, pe_frame *f) {
  const int pal = f->d.palette;
  const uint8_t y0 = (pal_is_yuv(pal) && f->d.yuv_clamping == PE_YUV_CLAMPING_CLAMPED) ? 16 : 0;
  if (pal_is_planar(pal)) {
    for (int p = 0; p < f->d.nplanes; p++) {
      const uint8_t v = p == 0 ? y0 : p == 3 ? 255 : 128;
      PE_CUDA(cudaMemsetAsync(f->d.planes[p], v, (size_t)f->d.rowstrides[p] * f->plane_heights[p], e->stream));
    }
    return PE_OK;
  }
  uint32_t px;
  switch (pal) {
  case PE_PALETTE_RGB24: case PE_PALETTE_BGR24: px = 0; break;
  case PE_PALETTE_RGBA32: case PE_PALETTE_BGRA32: px = 0xFF000000u; break;
  case PE_PALETTE_ARGB32: px = 0x000000FFu; break;
  case PE_PALETTE_YUV888: px = y0 | (128u << 8) | (128u << 16); break;
  case PE_PALETTE_YUVA8888: px = y0 | (128u << 8) | (128u << 16) | 0xFF000000u; break;
  case PE_PALETTE_UYVY: px = 128u | ((uint32_t)y0 << 8) | (128u << 16) | ((uint32_t)y0 << 24); break;
  default: px = y0 | (128u << 8) | ((uint32_t)y0 << 16) | (128u << 24); break;  // YUYV
  }
  PE_CUDA(launch_fill(e->L(), Img{(uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, f->d.width / pal_ppmp(pal), f->d.height,
                      pal_psize(pal), px));
  return PE_OK;
}

inline int chroma_height(const pe_frame *f) { return f->plane_heights[1]; }

// replace the pixel memory and geometry of `f` by those of `n` (which is consumed)
void frame_take(pe_frame *f, pe_frame *n) {
  frame_release_pixels(f);
  f->base = n->base;
  f->granted = n->granted;
  f->d.palette = n->d.palette;
  f->d.width = n->d.width;
  f->d.height = n->d.height;
  f->d.nplanes = n->d.nplanes;
  for (int p = 0; p < PE_MAXPLANES; p++) {
    f->d.rowstrides[p] = n->d.rowstrides[p];
    f->d.planes[p] = n->d.planes[p];
    f->plane_heights[p] = n->plane_heights[p];
  }
  n->base = nullptr;
}

}  // namespace

// fill: 0 leave the block as it is (internal: every payload byte is about to be overwritten), 1 zero, 2 black
static int frame_create_impl(pe_engine_t *e, int palette, int width, int height, int yuv_clamping, int yuv_sampling,
                             int yuv_subspace, int gamma_type, int fill, pe_frame_t **out) {
  if (!e || !out) return set_err(PE_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (!pal_known(palette)) return set_err(PE_ERR_PALETTE, "palette %d is not handled by this build", palette);
  if (palette == PE_PALETTE_YUV420P || palette == PE_PALETTE_YVU420P) {  // colourspace.c:11603-11604
    width = (width >> 1) << 1;
    height = (height >> 1) << 1;
  }
  if (width <= 0 || height <= 0) return set_err(PE_ERR_ARG, "bad size %d x %d", width, height);
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  pe_frame *f = new pe_frame();
  f->e = e;
  f->d.palette = palette; f->d.width = width; f->d.height = height;
  f->d.yuv_clamping = yuv_clamping; f->d.yuv_sampling = yuv_sampling; f->d.yuv_subspace = yuv_subspace;
  f->d.gamma_type = gamma_type; f->d.flags = 0;
  int rc = frame_alloc(e, f);
  if (rc != PE_OK) { delete f; return rc; }
  if (fill == 2) rc = frame_fill_black(e, f);
  else if (fill == 1 && cudaMemsetAsync(f->base, 0, f->granted, e->stream) != cudaSuccess) rc = set_err(PE_ERR_CUDA, "memset failed");
  if (rc != PE_OK) { frame_release_pixels(f); delete f; return rc; }
  *out = f;
  return PE_OK;
}

extern "C" int pe_frame_create(pe_engine_t *e, int palette, int width, int height, int yuv_clamping, int yuv_sampling,
                               int yuv_subspace, int gamma_type, int black_fill, pe_frame_t **out) {
  return frame_create_impl(e, palette, width, height, yuv_clamping, yuv_sampling, yuv_subspace, gamma_type, black_fill ? 2 : 1, out);
}

extern "C" int pe_frame_wrap(pe_engine_t *e, const pe_frame_desc_t *desc, pe_frame_t **out) {
  if (!e || !desc || !out) return set_err(PE_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (!pal_known(desc->palette)) return set_err(PE_ERR_PALETTE, "palette %d is not handled by this build", desc->palette);
  if (desc->width <= 0 || desc->height <= 0) return set_err(PE_ERR_ARG, "bad size %d x %d", desc->width, desc->height);
  const int np = pal_nplanes(desc->palette);
  for (int p = 0; p < np; p++)
    if (!desc->planes[p] || desc->rowstrides[p] <= 0) return set_err(PE_ERR_ARG, "plane %d missing", p);
  pe_frame *f = new pe_frame();
  f->e = e;
  f->d = *desc;
  f->d.nplanes = np;
  int rs[PE_MAXPLANES];
  pe_frame_layout(desc->palette, desc->width, desc->height, nullptr, rs, f->plane_heights);
  *out = f;
  return PE_OK;
}

extern "C" void pe_frame_destroy(pe_frame_t *f) {
  if (!f) return;
  if (f->e) {
    std::lock_guard<std::mutex> lk(f->e->mu);
    frame_release_pixels(f);
  }
  delete f;
}

extern "C" int pe_frame_get_desc(const pe_frame_t *f, pe_frame_desc_t *out) {
  if (!f || !out) return set_err(PE_ERR_ARG, "NULL argument");
  *out = f->d;
  return PE_OK;
}

extern "C" int pe_frame_set_gamma(pe_frame_t *f, int gamma_type) {
  if (!f) return set_err(PE_ERR_ARG, "NULL argument");
  f->d.gamma_type = gamma_type;
  return PE_OK;
}

extern "C" int pe_frame_set_flags(pe_frame_t *f, int flags) {
  if (!f) return set_err(PE_ERR_ARG, "NULL argument");
  f->d.flags = flags;
  return PE_OK;
}

namespace {
// payload bytes of one row of plane p
inline int plane_row_bytes(const pe_frame_desc_t &d, int p) {
  if (p == 0) return (d.width / pal_ppmp(d.palette)) * pal_psize(d.palette);
  switch (d.palette) {
  case PE_PALETTE_YUV420P: case PE_PALETTE_YVU420P: case PE_PALETTE_YUV422P: return d.width >> 1;
  default: return d.width;
  }
}
}  // namespace

extern "C" int pe_frame_upload(pe_engine_t *e, pe_frame_t *f, const void *const host_planes[PE_MAXPLANES],
                               const int host_rowstrides[PE_MAXPLANES]) {
  if (!e || !f || !host_planes) return set_err(PE_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  for (int p = 0; p < f->d.nplanes; p++) {
    if (!host_planes[p]) return set_err(PE_ERR_ARG, "host plane %d is NULL", p);
    const int hrs = host_rowstrides ? host_rowstrides[p] : f->d.rowstrides[p];
    // whole strides when they agree (row padding travels too: the reference's 4:2:x converters read one byte
    // past the chroma row, colourspace.c:3508), else the payload bytes
    int wbytes = plane_row_bytes(f->d, p);
    if (hrs == f->d.rowstrides[p]) wbytes = hrs;
    // a big pageable plane: copy threads fill a page-locked ring buffer, the DMA runs from there (pe_hoststage.h); anything else -- and
    // any failure of the staged path -- is the plain copy, which the driver stages itself
    if ((size_t)f->d.rowstrides[p] * f->plane_heights[p] >= pe::HostStager::kMinBytes && !e->no_host_staging && pe::HostStager::pageable(host_planes[p]) &&
        e->stager.upload(e->stream, f->d.planes[p], (size_t)f->d.rowstrides[p], host_planes[p], (size_t)hrs, (size_t)wbytes,
                         (size_t)f->plane_heights[p]) == cudaSuccess)
      continue;
    cudaGetLastError();
    PE_CUDA(cudaMemcpy2DAsync(f->d.planes[p], f->d.rowstrides[p], host_planes[p], hrs, wbytes, f->plane_heights[p],
                              cudaMemcpyHostToDevice, e->stream));
  }
  return PE_OK;
}

extern "C" int pe_frame_download(pe_engine_t *e, const pe_frame_t *f, void *const host_planes[PE_MAXPLANES],
                                 const int host_rowstrides[PE_MAXPLANES]) {
  if (!e || !f || !host_planes) return set_err(PE_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  std::vector<pe::PendingOut> pending;
  for (int p = 0; p < f->d.nplanes; p++) {
    if (!host_planes[p]) return set_err(PE_ERR_ARG, "host plane %d is NULL", p);
    const int hrs = host_rowstrides ? host_rowstrides[p] : f->d.rowstrides[p];
    if ((size_t)f->d.rowstrides[p] * f->plane_heights[p] >= pe::HostStager::kMinBytes && !e->no_host_staging && pe::HostStager::pageable(host_planes[p])) {
      pe::PendingOut po;
      if (e->stager.download_begin(e->stream, f->d.planes[p], (size_t)f->d.rowstrides[p], host_planes[p], (size_t)hrs,
                                   (size_t)plane_row_bytes(f->d, p), (size_t)f->plane_heights[p], &po) == cudaSuccess) {
        pending.push_back(po);   // (plane p is copied out by the threads while plane p + 1 travels)
        continue;
      }
      cudaGetLastError();
    }
    PE_CUDA(cudaMemcpy2DAsync(host_planes[p], hrs, f->d.planes[p], f->d.rowstrides[p], plane_row_bytes(f->d, p),
                              f->plane_heights[p], cudaMemcpyDeviceToHost, e->stream));
  }
  cudaError_t pend_err = cudaSuccess;
  for (auto &po : pending) { const cudaError_t ce = e->stager.finish(po); if (ce != cudaSuccess) pend_err = ce; }
  PE_CUDA(pend_err);
  PE_CUDA(cudaStreamSynchronize(e->stream));
  return PE_OK;
}

extern "C" int pe_frame_copy(pe_engine_t *e, const pe_frame_t *src, pe_frame_t **out) {
  if (!e || !src || !out) return set_err(PE_ERR_ARG, "NULL argument");
  *out = nullptr;
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  pe_frame *f = new pe_frame();
  f->e = e;
  f->d = src->d;
  int rc = frame_alloc(e, f);
  if (rc != PE_OK) { delete f; return rc; }
  for (int p = 0; p < f->d.nplanes; p++) {
    const int rb = src->d.rowstrides[p] < f->d.rowstrides[p] ? src->d.rowstrides[p] : f->d.rowstrides[p];
    cudaError_t ce = launch_copy2d(e->L(), (const uint8_t *)src->d.planes[p], src->d.rowstrides[p], (uint8_t *)f->d.planes[p],
                                   f->d.rowstrides[p], rb, f->plane_heights[p], 0, 0);
    if (ce != cudaSuccess) { frame_release_pixels(f); delete f; return set_err(PE_ERR_CUDA, "copy failed: %s", cudaGetErrorString(ce)); }
  }
  *out = f;
  return PE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// frame ingest / egress in device memory.
//   ingest: the decoder plugin's get_frame (src/plugins.h:442, decplugin.h:280: "fill these planes with frame N") with the planes in
//           HBM -- a device decoder (NVDEC / nvJPEG writing YUV420P / NV12) implements pe_device_get_frame_f; pe_clip_cache is the
//           built-in source: a clip that already sits in HBM
//   egress: the render-to-disk tail (src/events.c:4247-4263: convert to the clip's palette, layer_to_pixbuf) -- only the final packed
//           frame crosses PCIe, on its own stream, while the next frame is computed
// ---------------------------------------------------------------------------------------------------------

namespace {
int convert_locked(pe_engine *e, pe_frame *f, int outpl, int oclamping, int osampling, int osubspace, int tgt_gamma);  // below
}

extern "C" int pe_ingest_frame(pe_engine_t *e, const pe_clip_source_t *src, int64_t frame, pe_frame_t **out) {
  if (!e || !src || !src->get_frame || !out) return set_err(PE_ERR_ARG, "NULL argument");
  *out = nullptr;
  pe_frame_t *f = nullptr;
  int rc = frame_create_impl(e, src->palette, src->width, src->height, src->yuv_clamping, src->yuv_sampling, src->yuv_subspace,
                             src->gamma_type, 0, &f);
  if (rc != PE_OK) return rc;
  {
    std::lock_guard<std::mutex> lk(e->mu);
    if (cudaSetDevice(e->device) != cudaSuccess) rc = set_err(PE_ERR_CUDA, "cudaSetDevice failed");
    // boolean get_frame(cdata, frame, rowstrides, height, pixel_data): FALSE = the frame could not be produced
    else if (!src->get_frame(src->clip_data, frame, f->d.rowstrides, f->d.height, f->d.planes, (void *)e->stream))
      rc = set_err(PE_ERR_ARG, "the clip source could not produce frame %lld", (long long)frame);
  }
  if (rc != PE_OK) { pe_frame_destroy(f); return rc; }
  *out = f;
  return PE_OK;
}

struct pe_clip_cache {
  pe_engine *e = nullptr;
  pe_frame_desc_t d{};          // geometry of one frame; planes unused
  int plane_heights[PE_MAXPLANES] = {};
  size_t plane_off[PE_MAXPLANES] = {};
  size_t frame_bytes = 0;
  int nframes = 0;
  uint8_t *base = nullptr;      // nframes * frame_bytes of HBM
};

extern "C" int pe_clip_cache_create(pe_engine_t *e, int palette, int width, int height, int nframes, int yuv_clamping, int yuv_sampling,
                                    int yuv_subspace, int gamma_type, pe_clip_cache_t **out) {
  if (!e || !out || nframes <= 0) return set_err(PE_ERR_ARG, "NULL / empty argument");
  *out = nullptr;
  if (!pal_known(palette)) return set_err(PE_ERR_PALETTE, "palette %d is not handled by this build", palette);
  if (palette == PE_PALETTE_YUV420P || palette == PE_PALETTE_YVU420P) { width &= ~1; height &= ~1; }
  pe_clip_cache *c = new pe_clip_cache();
  c->e = e;
  c->d.palette = palette; c->d.width = width; c->d.height = height;
  c->d.yuv_clamping = yuv_clamping; c->d.yuv_sampling = yuv_sampling; c->d.yuv_subspace = yuv_subspace; c->d.gamma_type = gamma_type;
  const size_t bytes = pe_frame_layout(palette, width, height, &c->d.nplanes, c->d.rowstrides, c->plane_heights);
  if (!bytes) { delete c; return set_err(PE_ERR_ARG, "bad frame geometry"); }
  size_t off = 0;
  for (int p = 0; p < c->d.nplanes; p++) { c->plane_off[p] = off; off += ((size_t)c->d.rowstrides[p] * c->plane_heights[p] + 255) / 256 * 256; }
  c->frame_bytes = off + 256;  // slack: the 4:2:x converters read a few bytes past the last chroma row (colourspace.c:3508)
  c->nframes = nframes;
  std::lock_guard<std::mutex> lk(e->mu);
  if (cudaSetDevice(e->device) != cudaSuccess || cudaMalloc(&c->base, c->frame_bytes * (size_t)nframes) != cudaSuccess) {
    cudaGetLastError();
    delete c;
    return set_err(PE_ERR_MEMORY, "device allocation of %zu bytes failed", c->frame_bytes * (size_t)nframes);
  }
  cudaMemsetAsync(c->base, 0, c->frame_bytes * (size_t)nframes, e->stream);
  *out = c;
  return PE_OK;
}

extern "C" void pe_clip_cache_destroy(pe_clip_cache_t *c) {
  if (!c) return;
  {
    std::lock_guard<std::mutex> lk(c->e->mu);
    cudaSetDevice(c->e->device);
    cudaStreamSynchronize(c->e->stream);
    cudaFree(c->base);
  }
  delete c;
}

// the one-time fill (what a device decoder would have written): host planes of frame `frame` -> HBM
extern "C" int pe_clip_cache_load(pe_clip_cache_t *c, int64_t frame, const void *const host_planes[PE_MAXPLANES],
                                  const int host_rowstrides[PE_MAXPLANES]) {
  if (!c || !host_planes || frame < 0 || frame >= c->nframes) return set_err(PE_ERR_ARG, "bad clip cache argument");
  pe_engine *e = c->e;
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  for (int p = 0; p < c->d.nplanes; p++) {
    if (!host_planes[p]) return set_err(PE_ERR_ARG, "host plane %d is NULL", p);
    const int hrs = host_rowstrides ? host_rowstrides[p] : c->d.rowstrides[p];
    const int wbytes = hrs < c->d.rowstrides[p] ? hrs : c->d.rowstrides[p];
    PE_CUDA(cudaMemcpy2DAsync(c->base + c->frame_bytes * (size_t)frame + c->plane_off[p], c->d.rowstrides[p], host_planes[p], hrs, wbytes,
                              c->plane_heights[p], cudaMemcpyHostToDevice, e->stream));
  }
  PE_CUDA(cudaStreamSynchronize(e->stream));
  return PE_OK;
}

// device pointers of a cached frame (a device decoder / another kernel may write them directly)
extern "C" int pe_clip_cache_frame_desc(pe_clip_cache_t *c, int64_t frame, pe_frame_desc_t *out) {
  if (!c || !out || frame < 0 || frame >= c->nframes) return set_err(PE_ERR_ARG, "bad clip cache argument");
  *out = c->d;
  for (int p = 0; p < c->d.nplanes; p++) out->planes[p] = c->base + c->frame_bytes * (size_t)frame + c->plane_off[p];
  return PE_OK;
}

namespace {
// pe_device_get_frame_f of the cache: device-to-device copies of the planes on the engine stream (frame numbers wrap around)
int clip_cache_get_frame(void *clip_data, int64_t frame, const int *rowstrides, int height, void *const *pixel_data, void *stream) {
  pe_clip_cache *c = (pe_clip_cache *)clip_data;
  if (!c || height != c->d.height) return 0;
  const int64_t k = ((frame % c->nframes) + c->nframes) % c->nframes;
  for (int p = 0; p < c->d.nplanes; p++) {
    const int rb = rowstrides[p] < c->d.rowstrides[p] ? rowstrides[p] : c->d.rowstrides[p];
    if (cudaMemcpy2DAsync(pixel_data[p], (size_t)rowstrides[p], c->base + c->frame_bytes * (size_t)k + c->plane_off[p], (size_t)c->d.rowstrides[p],
                          (size_t)rb, (size_t)c->plane_heights[p], cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
  }
  return 1;
}
}  // namespace

extern "C" int pe_clip_cache_source(pe_clip_cache_t *c, pe_clip_source_t *out) {
  if (!c || !out) return set_err(PE_ERR_ARG, "NULL argument");
  out->clip_data = c;
  out->get_frame = clip_cache_get_frame;
  out->palette = c->d.palette; out->width = c->d.width; out->height = c->d.height;
  out->yuv_clamping = c->d.yuv_clamping; out->yuv_sampling = c->d.yuv_sampling; out->yuv_subspace = c->d.yuv_subspace;
  out->gamma_type = c->d.gamma_type;
  return PE_OK;
}

// zero-copy ingest: the cached frame itself as a (borrowed) layer -- for chains that only read it (the fused kernel's fg / bg)
extern "C" int pe_clip_cache_borrow(pe_clip_cache_t *c, int64_t frame, pe_frame_t **out) {
  pe_frame_desc_t d;
  if (!c || !out) return set_err(PE_ERR_ARG, "NULL argument");
  const int64_t k = ((frame % c->nframes) + c->nframes) % c->nframes;
  int rc = pe_clip_cache_frame_desc(c, k, &d);
  if (rc != PE_OK) return rc;
  return pe_frame_wrap(c->e, &d, out);
}

// egress, asynchronous: convert_layer_palette(layer, out_palette) when needed (the render tail converts to the clip's palette,
// src/events.c:4250), then ONLY that packed frame travels to host memory on the download stream; slot 0 .. 3 names the copy for
// pe_render_out_wait.  The layer must stay alive (and unwritten) until the wait.
extern "C" int pe_render_out_begin(pe_engine_t *e, pe_frame_t *layer, int out_palette, void *host_dst, int host_rowstride, int slot) {
  if (!e || !layer || !host_dst || slot < 0 || slot >= 4) return set_err(PE_ERR_ARG, "bad render-out argument");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  if (out_palette != PE_PALETTE_NONE && layer->d.palette != out_palette) {
    if (!convert_locked(e, layer, out_palette, layer->d.yuv_clamping, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YUV, PE_GAMMA_UNKNOWN))
      return PE_ERR_PALETTE;
  }
  if (pal_is_planar(layer->d.palette)) return set_err(PE_ERR_PALETTE, "render-out takes a packed frame (RGB24 / RGBA32 / ...)");
  if (!e->egress_stream) {
    PE_CUDA(cudaStreamCreateWithFlags(&e->egress_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 4; k++) {
      PE_CUDA(cudaEventCreateWithFlags(&e->egress_ready[k], cudaEventDisableTiming));
      PE_CUDA(cudaEventCreateWithFlags(&e->egress_done[k], cudaEventDisableTiming));
    }
  }
  const int hrs = host_rowstride > 0 ? host_rowstride : layer->d.rowstrides[0];
  PE_CUDA(cudaEventRecord(e->egress_ready[slot], e->stream));
  PE_CUDA(cudaStreamWaitEvent(e->egress_stream, e->egress_ready[slot], 0));
  PE_CUDA(cudaMemcpy2DAsync(host_dst, (size_t)hrs, layer->d.planes[0], (size_t)layer->d.rowstrides[0], (size_t)plane_row_bytes(layer->d, 0),
                            (size_t)layer->d.height, cudaMemcpyDeviceToHost, e->egress_stream));
  PE_CUDA(cudaEventRecord(e->egress_done[slot], e->egress_stream));
  e->egress_busy[slot] = true;
  return PE_OK;
}

extern "C" int pe_render_out_wait(pe_engine_t *e, int slot) {
  if (!e || slot < 0 || slot >= 4) return set_err(PE_ERR_ARG, "bad render-out slot");
  if (!e->egress_busy[slot]) return PE_OK;
  PE_CUDA(cudaEventSynchronize(e->egress_done[slot]));
  e->egress_busy[slot] = false;
  return PE_OK;
}

extern "C" int pe_render_out(pe_engine_t *e, pe_frame_t *layer, int out_palette, void *host_dst, int host_rowstride) {
  int rc = pe_render_out_begin(e, layer, out_palette, host_dst, host_rowstride, 0);
  return rc == PE_OK ? pe_render_out_wait(e, 0) : rc;
}

// ---- process-wide engine + host prefs + host-buffer registration (what the weed_layer_t drop-ins and the effect plugin share) ----

namespace {
std::mutex g_shared_mu;
pe_engine_t *g_shared = nullptr;
pe_config_t g_shared_cfg;
bool g_shared_cfg_set = false;
}  // namespace

// the configuration the shared engine is created with on first use (device, prefs); PE_DEVICE in the environment picks the GPU when
// the host never calls this (a plugin has no other channel)
extern "C" int pe_engine_shared_configure(const pe_config_t *cfg) {
  std::lock_guard<std::mutex> lk(g_shared_mu);
  if (g_shared) return set_err(PE_ERR_ARG, "the shared engine already exists: use pe_engine_set_prefs");
  if (!cfg) return set_err(PE_ERR_ARG, "NULL config");
  g_shared_cfg = *cfg;
  g_shared_cfg_set = true;
  return PE_OK;
}

extern "C" pe_engine_t *pe_engine_shared(void) {
  std::lock_guard<std::mutex> lk(g_shared_mu);
  if (!g_shared) {
    if (!g_shared_cfg_set) {
      pe_config_default(&g_shared_cfg);
      if (const char *d = getenv("PE_DEVICE")) g_shared_cfg.device = atoi(d);
    }
    if (pe_engine_create(&g_shared_cfg, &g_shared) != PE_OK) g_shared = nullptr;  // reason in pe_last_error()
  }
  return g_shared;
}

// prefs->pb_quality / screen_gamma / apply_gamma / alpha_post of the host (src/preferences.h) as they change at run time.  The gamma
// LUT caches depend on screen_gamma (colourspace.c:677): they are dropped when it changes.
extern "C" int pe_engine_set_prefs(pe_engine_t *e, int pb_quality, double screen_gamma, int apply_gamma, int alpha_post) {
  if (!e) return set_err(PE_ERR_ARG, "engine is NULL");
  if (pb_quality < PE_QUALITY_LOW || pb_quality > PE_QUALITY_HIGH) return set_err(PE_ERR_ARG, "pb_quality %d out of range", pb_quality);
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  if (screen_gamma != e->cfg.screen_gamma) {
    PE_CUDA(cudaStreamSynchronize(e->stream));
    for (auto &kv : e->lut8) cudaFree(kv.second.dev);
    for (auto &kv : e->lut16) cudaFree(kv.second);
    for (auto &kv : e->over) cudaFree(kv.second.dev);  // keyed by LUT pointers
    e->lut8.clear(); e->lut16.clear(); e->over.clear();
  }
  e->cfg.pb_quality = pb_quality; e->cfg.screen_gamma = screen_gamma; e->cfg.apply_gamma = apply_gamma; e->cfg.alpha_post = alpha_post;
  return PE_OK;
}

extern "C" int pe_engine_get_config(pe_engine_t *e, pe_config_t *out) {
  if (!e || !out) return set_err(PE_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(e->mu);
  *out = e->cfg;
  return PE_OK;
}

// page-lock caller-owned host memory in place (cudaHostRegister): LiVES recycles a fixed set of big pixel buffers (bigblocks,
// src/memory.c:37-47), so a buffer registered once serves every later frame at pinned-copy speed.  Idempotent per range.
extern "C" int pe_host_register(void *p, size_t bytes) {
  if (!p || !bytes) return set_err(PE_ERR_ARG, "NULL / empty range");
  cudaError_t ce = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (ce == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return PE_OK; }
  if (ce != cudaSuccess) { cudaGetLastError(); return set_err(PE_ERR_CUDA, "cudaHostRegister(%zu bytes) failed: %s", bytes, cudaGetErrorString(ce)); }
  return PE_OK;
}
extern "C" int pe_host_unregister(void *p) {
  if (!p) return PE_OK;
  cudaError_t ce = cudaHostUnregister(p);
  if (ce != cudaSuccess) { cudaGetLastError(); return set_err(PE_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(ce)); }
  return PE_OK;
}

extern "C" void *pe_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void pe_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------------------------------------
// gamma (colourspace.c:14034-14160)
// ---------------------------------------------------------------------------------------------------------

namespace {

int gamma_convert_sub_layer_locked(pe_engine *e, int gamma_type, double fileg, pe_frame *f, int x, int y, int width, int height) {
  if (!e->cfg.apply_gamma) return PE_TRUE;                      // :14071
  if (!pal_is_rgb(f->d.palette)) return PE_FALSE;               // :14076 "dont know how to convert in yuv space"
  const int lgamma = f->d.gamma_type;
  if (gamma_type == lgamma && fileg == 1.0) return PE_TRUE;     // :14080
  Lut8Entry *lut = gamma_type == PE_GAMMA_VARIANT ? get_lut8(e, fileg, lgamma, gamma_type) : get_lut8(e, 1.0, lgamma, gamma_type);
  if (!lut) return PE_TRUE;                                     // :14103 (no-op pairs return NULL)
  // the rectangle: the reference's band arithmetic (:14096-14124) is thread-count dependent; contract = whole rect
  if (x < 0 || y < 0 || width <= 0 || height <= 0 || x + width > f->d.width || y + height > f->d.height) {
    set_err(PE_ERR_ARG, "gamma rectangle %d,%d %dx%d outside the %dx%d frame", x, y, width, height, f->d.width, f->d.height);
    return PE_FALSE;
  }
  PE_CUDA_B(launch_lut8_rect(e->L(), Img{(uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, rgb_layout(f->d.palette), x, y, width,
                             height, lut->dev));
  if (gamma_type != PE_GAMMA_VARIANT) f->d.gamma_type = gamma_type;  // :14139
  return PE_TRUE;
}

int gamma_convert_layer_locked(pe_engine *e, int gamma_type, pe_frame *f) {
  if (!f || !f->d.planes[0] || !f->d.width || !f->d.height) return PE_FALSE;  // :14146-14156
  return gamma_convert_sub_layer_locked(e, gamma_type, 1.0, f, 0, 0, f->d.width, f->d.height);
}

}  // namespace

extern "C" int pe_gamma_convert_sub_layer(pe_engine_t *e, int gamma_type, double fileg, pe_frame_t *layer, int x, int y,
                                          int width, int height, int may_thread) {
  (void)may_thread;  // row bands are the CUDA grid's business
  if (!e || !layer) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA_B(cudaSetDevice(e->device));
  return gamma_convert_sub_layer_locked(e, gamma_type, fileg, layer, x, y, width, height);
}

extern "C" int pe_gamma_convert_layer(pe_engine_t *e, int gamma_type, pe_frame_t *layer) {
  if (!e || !layer) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA_B(cudaSetDevice(e->device));
  return gamma_convert_layer_locked(e, gamma_type, layer);
}

extern "C" int pe_gamma_lut8(pe_engine_t *e, double fileg, int gamma_from, int gamma_to, uint8_t out[256]) {
  if (!e || !out) return set_err(PE_ERR_ARG, "NULL argument");
  if (!build_gamma_lut8(fileg, gamma_from, gamma_to, e->cfg.screen_gamma, out))
    return set_err(PE_ERR_ARG, "no LUT for this pair (create_gamma_lut8 returns NULL, colourspace.c:662)");
  return PE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// alpha premultiply (colourspace.c:11968-12106)
// ---------------------------------------------------------------------------------------------------------

namespace {

void alpha_premult_locked(pe_engine *e, pe_frame *f, int direction) {
  const int pal = f->d.palette;
  int coffs, aoffs;
  if (pal == PE_PALETTE_YUVA4444P) {  // "special case - planar with alpha" (:12001-12049)
    const bool clamped = f->d.yuv_clamping == PE_YUV_CLAMPING_CLAMPED;
    const bool rev = direction == PE_DIRECTION_REVERSE;
    const uint8_t *ty = get_premult(e, clamped ? (rev ? 2 : 3) : (rev ? 0 : 1)), *tc = get_premult(e, clamped ? (rev ? 4 : 5) : (rev ? 0 : 1));
    if (!ty || !tc) { set_err(PE_ERR_MEMORY, "premultiply tables could not be built"); return; }
    uint8_t *pl[4] = {(uint8_t *)f->d.planes[0], (uint8_t *)f->d.planes[1], (uint8_t *)f->d.planes[2], (uint8_t *)f->d.planes[3]};
    cudaError_t ce = launch_premult_planar(e->L(), pl, f->d.rowstrides, f->d.width, f->d.height, ty, tc);
    if (ce != cudaSuccess) { set_err(PE_ERR_CUDA, "premult launch failed: %s", cudaGetErrorString(ce)); return; }
    return;  // (the planar branch returns before the flag update of :12100-12104, as the reference does)
  }
  switch (pal) {
  case PE_PALETTE_RGBA32: case PE_PALETTE_BGRA32: case PE_PALETTE_YUVA8888: coffs = 0; aoffs = 3; break;
  case PE_PALETTE_ARGB32: coffs = 1; aoffs = 0; break;
  default: return;  // other palettes: no-op as in the reference
  }
  const bool clamped_yuva = (pal == PE_PALETTE_YUVA8888 && f->d.yuv_clamping == PE_YUV_CLAMPING_CLAMPED);
  const uint8_t *t0, *t1, *t2;
  int quirk = 0;
  if (!clamped_yuva) {
    t0 = t1 = t2 = get_premult(e, direction == PE_DIRECTION_REVERSE ? 0 : 1);  // REVERSE -> unal, FORWARD -> al (:12058-12074)
  } else if (direction == PE_DIRECTION_REVERSE) {
    t0 = get_premult(e, 2); t1 = t2 = get_premult(e, 4);  // unalcy / unalcuv
  } else {
    t0 = get_premult(e, 3); t1 = t2 = get_premult(e, 5);  // alcy / alcuv
    quirk = 1;  // U and V are looked up with the (already rewritten) Y byte, :12093-12094
  }
  if (!t0 || !t1) { set_err(PE_ERR_MEMORY, "premultiply tables could not be built"); return; }
  cudaError_t ce = launch_premult(e->L(), Img{(uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, f->d.width, f->d.height, coffs, 3,
                                  aoffs, t0, t1, t2, quirk);
  if (ce != cudaSuccess) { set_err(PE_ERR_CUDA, "premult launch failed: %s", cudaGetErrorString(ce)); return; }
  if (direction == PE_DIRECTION_FORWARD) f->d.flags |= PE_LAYER_ALPHA_PREMULT;  // :12100-12104
  else f->d.flags &= ~PE_LAYER_ALPHA_PREMULT;
}

}  // namespace

extern "C" void pe_alpha_premult(pe_engine_t *e, pe_frame_t *layer, int direction) {
  if (!e || !layer || !layer->d.planes[0]) return;
  std::lock_guard<std::mutex> lk(e->mu);
  if (cudaSetDevice(e->device) != cudaSuccess) return;
  alpha_premult_locked(e, layer, direction);
}

// ---------------------------------------------------------------------------------------------------------
// convert_layer_palette_full (colourspace.c:12190-13928)
// ---------------------------------------------------------------------------------------------------------

namespace {

int convert_locked(pe_engine *e, pe_frame *f, int outpl, int oclamping, int osampling, int osubspace, int tgt_gamma);

inline int convert_simple_locked(pe_engine *e, pe_frame *f, int outpl, int op_clamping) {  // convert_layer_palette :13931
  return convert_locked(e, f, outpl, op_clamping, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YUV, PE_GAMMA_UNKNOWN);
}

Planes planes_of(const pe_frame *f, bool swap_uv) {
  Planes P;
  P.y = (const uint8_t *)f->d.planes[0];
  P.u = (const uint8_t *)f->d.planes[swap_uv ? 2 : 1];
  P.v = (const uint8_t *)f->d.planes[swap_uv ? 1 : 2];
  P.rs_y = f->d.rowstrides[0];
  P.rs_u = f->d.rowstrides[swap_uv ? 2 : 1];
  P.rs_v = f->d.rowstrides[swap_uv ? 1 : 2];
  P.cw = f->d.width >> 1;
  P.ch = f->plane_heights[1];
  return P;
}

const uint8_t *get_yy(pe_engine *e) {
  if (e->yy_dev) return e->yy_dev;
  uint8_t host[4 * 256];
  for (int k = 0; k < 4; k++) build_yy_table(k, host + 256 * k);
  uint8_t *dev = nullptr;
  if (cudaMalloc(&dev, sizeof(host)) != cudaSuccess) return nullptr;
  if (cudaMemcpyAsync(dev, host, sizeof(host), cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
      cudaStreamSynchronize(e->stream) != cudaSuccess) {
    cudaFree(dev);
    return nullptr;
  }
  e->yy_dev = dev;
  return dev;
}

// switch_yuv_clamping_and_subspace (colourspace.c:10929): in place, all planes
int switch_clamping_locked(pe_engine *e, pe_frame *f, int oclamping) {
  const uint8_t *yy = get_yy(e);
  if (!yy) return set_err(PE_ERR_MEMORY, "clamping tables could not be built");
  const bool to_unclamped = f->d.yuv_clamping == PE_YUV_CLAMPING_CLAMPED;  // :1177-1185: anything else is treated as unclamped input
  const uint8_t *ty = yy + (to_unclamped ? 0 : 512), *tc = yy + (to_unclamped ? 256 : 768);
  const int pal = f->d.palette;
  cudaError_t ce = cudaSuccess;
  for (int p = 0; p < f->d.nplanes && ce == cudaSuccess; p++) {
    int kind;
    switch (pal) {
    case PE_PALETTE_YUV888: kind = 2; break;
    case PE_PALETTE_YUVA8888: kind = 3; break;
    case PE_PALETTE_UYVY: kind = 4; break;
    case PE_PALETTE_YUYV: kind = 5; break;
    default: kind = p == 0 ? 0 : 1; break;
    }
    if (p == 3) break;  // the alpha plane of YUVA4444P is not touched (:10971-10984)
    // YUV888: the reference walks Y U V Y U V ... densely across the row padding (:10959-10969), so with a rowstride that is not
    // a multiple of 3 the tables are misapplied from row 1 on -- replicated under ref_quirks, per-row phase without
    ce = launch_clamp_lut(e->L(), (uint8_t *)f->d.planes[p], (long long)f->d.rowstrides[p] * f->plane_heights[p], kind,
                          e->cfg.ref_quirks ? 0 : f->d.rowstrides[p], ty, tc);
  }
  if (ce != cudaSuccess) return set_err(PE_ERR_CUDA, "clamping switch launch failed: %s", cudaGetErrorString(ce));
  f->d.yuv_clamping = oclamping;
  return PE_OK;
}

int convert_locked(pe_engine *e, pe_frame *f, int outpl, int oclamping, int osampling, int osubspace, int tgt_gamma) {
  if (!f || !f->d.planes[0]) return PE_FALSE;  // :12206
  if (!pal_known(outpl)) { set_err(PE_ERR_PALETTE, "output palette %d is not handled by this build", outpl); return PE_FALSE; }
  int inpl = f->d.palette;
  int isampling = f->d.yuv_sampling, iclamping = f->d.yuv_clamping, isubspace = f->d.yuv_subspace;

  if (pal_is_yuv(inpl) && pal_is_yuv(outpl) && (iclamping != oclamping || isubspace != osubspace)) {  // :12241
    if (isubspace == osubspace) {
      // switch_yuv_clamping_and_subspace (:10929-11092): every byte of every plane through Y_to_Y / U_to_U (= V_to_V), walked
      // densely over height * rowstride bytes (row padding included); "currently subspace conversions are not performed"
      if (switch_clamping_locked(e, f, oclamping) != PE_OK) return PE_FALSE;
      iclamping = oclamping;
    } else {
      // different subspace: go through RGB(A) first (:12249-12262)
      if (!convert_simple_locked(e, f, pal_has_alpha(inpl) ? PE_PALETTE_RGBA32 : PE_PALETTE_RGB24, 0)) return PE_FALSE;
      inpl = f->d.palette;
      isubspace = osubspace; isampling = osampling; iclamping = oclamping;
    }
  }
  if (inpl == outpl) return PE_TRUE;  // :12265-12288 (sampling switches "not yet written" in the reference either)

  // premultiplied-alpha bookkeeping (:12290-12306)
  int flags = f->d.flags;
  if (e->cfg.alpha_post) {
    if ((flags & PE_LAYER_ALPHA_PREMULT) && pal_has_alpha(inpl) && !pal_has_alpha(outpl)) {
      alpha_premult_locked(e, f, PE_DIRECTION_REVERSE);
      flags = f->d.flags;
    }
  } else if (!pal_has_alpha(inpl) && pal_has_alpha(outpl)) {
    flags |= PE_LAYER_ALPHA_PREMULT;
  }
  if (pal_has_alpha(inpl) && !pal_has_alpha(outpl) && (flags & PE_LAYER_ALPHA_PREMULT)) flags &= ~PE_LAYER_ALPHA_PREMULT;
  f->d.flags = flags;

  // gamma decision (:12311-12332)
  int gamma_type = PE_GAMMA_UNKNOWN, new_gamma_type = PE_GAMMA_UNKNOWN;
  if (e->cfg.apply_gamma) {
    gamma_type = f->d.gamma_type;
    if (gamma_type != PE_GAMMA_UNKNOWN) {
      if (tgt_gamma != PE_GAMMA_UNKNOWN) new_gamma_type = tgt_gamma;
      else if (pal_is_rgb(inpl) && !pal_is_rgb(outpl))
        new_gamma_type = osubspace == PE_YUV_SUBSPACE_BT709 ? PE_GAMMA_BT709 : PE_GAMMA_SRGB;
      else new_gamma_type = gamma_type;
      if (pal_is_rgb(inpl) && !pal_is_rgb(outpl) && !can_inline_gamma(inpl, outpl)) {
        gamma_convert_layer_locked(e, new_gamma_type, f);
        gamma_type = new_gamma_type = f->d.gamma_type;
      }
    }
  }

  const int width = f->d.width, height = f->d.height;
  const Launch L = e->L();

  // destination frame
  pe_frame n;
  n.e = e;
  n.d.palette = outpl; n.d.width = width; n.d.height = height;
  bool inplace = false;
  cudaError_t ce = cudaSuccess;

  if (pal_is_rgb(inpl) && pal_is_rgb(outpl)) {
    // all RGB <-> RGB conversions (:12370-12556): byte permutation + alpha add / drop, optional LUT8
    const uint8_t *lut = nullptr;
    if (gamma_type != new_gamma_type) {
      Lut8Entry *le = get_lut8(e, 1.0, gamma_type, new_gamma_type);
      lut = le ? le->dev : nullptr;
    }
    const RgbLayout li = rgb_layout(inpl), lo = rgb_layout(outpl);
    inplace = (li.psize == lo.psize);  // pconv_can_inplace :12148
    if (!inplace && frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    Img dst = inplace ? Img{(uint8_t *)f->d.planes[0], f->d.rowstrides[0]} : Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]};
    if (e->rgb_defer)  // batch call: launched by flush_rgb_pending
      e->rgb_pending.push_back(pe_engine::RgbJob{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0], dst.p, dst.rs, width, height, li, lo, lut});
    else
      ce = launch_rgb_to_rgb(L, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, dst, width, height, li, lo, lut);
  } else if ((inpl == PE_PALETTE_YUV420P || inpl == PE_PALETTE_YVU420P || inpl == PE_PALETTE_YUV422P) && pal_is_rgb(outpl)) {
    // convert_yuv420p_to_{rgb,bgr,argb}_frame (:13521-13560, :13648-13685)
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    YuvToRgbArgs A;
    A.src = planes_of(f, inpl == PE_PALETTE_YVU420P);  // swap_chroma_planes :12354
    A.dst = Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]};
    A.width = width; A.height = height;
    A.out = rgb_layout(outpl);
    A.is_422 = inpl == PE_PALETTE_YUV422P;
    A.clamped = iclamping == PE_YUV_CLAMPING_CLAMPED;
    // only the RGB-order converter has the PB_QUALITY_LOW chroma shortcut (:3470; none at :4090, :4700)
    A.low_quality = (e->cfg.pb_quality == PE_QUALITY_LOW && (outpl == PE_PALETTE_RGB24 || outpl == PE_PALETTE_RGBA32));
    A.quirks = e->cfg.ref_quirks;
    A.conv = dev_conv(e, iclamping, isubspace);
    A.lut16 = nullptr;
    if (new_gamma_type != PE_GAMMA_UNKNOWN) A.lut16 = get_lut16(e, 1.0, gamma_type, new_gamma_type);  // `if (tgt_gamma)` :3273
    A.blend2 = e->fuse_blend2; A.blend2_rs = e->fuse_blend2_rs; A.blend_bf = e->fuse_blend_bf;
    if (getenv("PE_YUV_SLOW") == nullptr && yuv_planar_fast_ok(A, &conv_host(e, iclamping, isubspace))) {
      if (e->yuv_defer) e->yuv_pending.push_back(A);  // batch call: launched by flush_yuv_pending
      else ce = launch_yuv_planar_to_rgb_fast(L, A);
    } else {
      ce = launch_yuv_planar_to_rgb(L, A);
    }
  } else if ((inpl == PE_PALETTE_UYVY || inpl == PE_PALETTE_YUYV) && pal_is_rgb(outpl)) {
    // convert_{uyvy,yuyv}_to_*_frame (:13147-13190, :13244-13290).  Table choice as the reference makes it:
    // uyvy->RGB24 passes the layer's subspace, uyvy->RGBA32 passes its SAMPLING in that slot (:13160), every other
    // variant selects YCbCr itself.
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    int sub = PE_YUV_SUBSPACE_YCBCR;
    if (inpl == PE_PALETTE_UYVY && outpl == PE_PALETTE_RGB24) sub = isubspace;
    else if (inpl == PE_PALETTE_UYVY && outpl == PE_PALETTE_RGBA32) sub = isampling;
    ce = launch_packed422_to_rgb(L, inpl == PE_PALETTE_UYVY ? 0 : 1, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]},
                                 Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width >> 1, height, rgb_layout(outpl),
                                 dev_conv(e, iclamping, sub));
  } else if ((inpl == PE_PALETTE_YUV888 || inpl == PE_PALETTE_YUVA8888) && pal_is_rgb(outpl)) {
    // convert_yuv888_to_*_frame / convert_yuva8888_to_*_frame: the dispatcher passes isampling where the converter
    // expects the subspace (:13359-13385, :13453-13478) -- replicated
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    ce = launch_yuv888_to_rgb(L, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]},
                              Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width, height, inpl == PE_PALETTE_YUVA8888,
                              rgb_layout(outpl), dev_conv(e, iclamping, isampling));
  } else if (pal_is_rgb(inpl) && (outpl == PE_PALETTE_YUV888 || outpl == PE_PALETTE_YUVA8888)) {
    // convert_{rgb,bgr,argb}_to_yuv_frame: always the YCbCr tables (:5710), clamping = oclamping
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    if (width & 1) {  // the converter drops an odd last column (:5750): leave it defined (black)
      n.d.yuv_clamping = oclamping;
      if (frame_fill_black(e, &n) != PE_OK) { frame_release_pixels(&n); return PE_FALSE; }
    }
    ce = launch_rgb_to_yuv888(L, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]},
                              Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width, height, rgb_layout(inpl),
                              outpl == PE_PALETTE_YUVA8888, dev_conv(e, oclamping, PE_YUV_SUBSPACE_YCBCR));
  } else if (pal_is_rgb(inpl) && (outpl == PE_PALETTE_UYVY || outpl == PE_PALETTE_YUYV)) {
    // convert_{rgb,bgr,argb}_to_{uyvy,yuyv}_frame (:12562-12580 etc.): YCbCr tables, clamping = oclamping, the 16-bit gamma
    // LUT inside the converter when the gamma changes; the layer gets width >> 1 macropixels (an odd last column is cut)
    n.d.width = (width >> 1) << 1;
    if (n.d.width < 2) { set_err(PE_ERR_SIZE, "frame too narrow for a 4:2:2 macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint16_t *lut16 = gamma_type == new_gamma_type ? nullptr : get_lut16(e, 1.0, gamma_type, new_gamma_type);
    ce = launch_rgb_to_packed422(L, outpl == PE_PALETTE_UYVY ? 0 : 1, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]},
                                 Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width, height, rgb_layout(inpl),
                                 dev_conv(e, oclamping, PE_YUV_SUBSPACE_YCBCR), lut16);
  } else if (pal_is_rgb(inpl) && (outpl == PE_PALETTE_YUV444P || outpl == PE_PALETTE_YUVA4444P)) {
    // convert_{rgb,bgr,argb}_to_yuvp_frame (:12613-12626 etc.)
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    if (width & 1) {  // the converter drops an odd last column: leave it defined (black)
      n.d.yuv_clamping = oclamping;
      if (frame_fill_black(e, &n) != PE_OK) { frame_release_pixels(&n); return PE_FALSE; }
    }
    uint8_t *pl[4] = {(uint8_t *)n.d.planes[0], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2],
                      outpl == PE_PALETTE_YUVA4444P ? (uint8_t *)n.d.planes[3] : nullptr};
    ce = launch_rgb_to_yuv444p(L, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, pl, n.d.rowstrides[0], width, height,
                               rgb_layout(inpl), dev_conv(e, oclamping, PE_YUV_SUBSPACE_YCBCR));
  } else if (pal_is_rgb(inpl) && (outpl == PE_PALETTE_YUV420P || outpl == PE_PALETTE_YVU420P || outpl == PE_PALETTE_YUV422P)) {
    // convert_{rgb,bgr}_to_yuv420_frame (:12681-12690, :12754-12763, :12600-12609, :12828-12837): width and height cut to even;
    // 4:2:0 takes the tables of osubspace, 4:2:2 gets WEED_YUV_SAMPLING_DEFAULT in that slot (= YCbCr); the planes are
    // written Cb to plane 1, Cr to plane 2 (the reference's dest[1] / dest[2]; a YVU420P layer gets them swapped at conv_done, :13895).  convert_argb_to_yuv420_frame
    // (:6323) takes G1 / B1 one / two bytes too far (:6357: the next pixel's bytes, past the buffer on the last pair of the frame): ARGB32 is
    // converted like the other orders, each pixel's own R, G, B (X).
    n.d.width = width & ~1; n.d.height = height & ~1;
    if (n.d.width < 2 || n.d.height < 2) { set_err(PE_ERR_SIZE, "frame too small for a 4:2:x macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const bool is422 = outpl == PE_PALETTE_YUV422P;
    const uint8_t *cavg = get_cavg(e, oclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    uint8_t *pl[3] = {(uint8_t *)n.d.planes[0], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2]};
    ce = launch_rgb_to_yuv420p(L, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, pl, n.d.rowstrides, n.d.width, n.d.height,
                               rgb_layout(inpl), is422 ? 1 : 0, dev_conv(e, oclamping, is422 ? PE_YUV_SUBSPACE_YCBCR : osubspace), cavg);
    n.d.yuv_sampling = PE_YUV_SAMPLING_DEFAULT;
  } else if ((inpl == PE_PALETTE_YUV444P || inpl == PE_PALETTE_YUVA4444P) && pal_is_rgb(outpl)) {
    // convert_yuv_planar_to_{rgb,bgr,argb}_frame (:12950-12976, :13055-13080): YCbCr tables of the layer's clamping
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *pl[4] = {(const uint8_t *)f->d.planes[0], (const uint8_t *)f->d.planes[1], (const uint8_t *)f->d.planes[2],
                            inpl == PE_PALETTE_YUVA4444P ? (const uint8_t *)f->d.planes[3] : nullptr};
    ce = launch_yuv444p_to_rgb(L, pl, f->d.rowstrides[0], Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width, height,
                               inpl == PE_PALETTE_YUVA4444P, rgb_layout(outpl), dev_conv(e, iclamping, PE_YUV_SUBSPACE_YCBCR));
  } else if ((inpl == PE_PALETTE_YUV444P || inpl == PE_PALETTE_YUVA4444P) && (outpl == PE_PALETTE_YUV888 || outpl == PE_PALETTE_YUVA8888)) {
    // convert_combineplanes_frame (:12999-13009, :13098-13108)
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *pl[4] = {(const uint8_t *)f->d.planes[0], (const uint8_t *)f->d.planes[1], (const uint8_t *)f->d.planes[2],
                            inpl == PE_PALETTE_YUVA4444P ? (const uint8_t *)f->d.planes[3] : nullptr};
    ce = launch_combine_planes(L, pl, f->d.rowstrides[0], Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width, height,
                               inpl == PE_PALETTE_YUVA4444P, outpl == PE_PALETTE_YUVA8888);
  } else if ((inpl == PE_PALETTE_YUV888 || inpl == PE_PALETTE_YUVA8888) && (outpl == PE_PALETTE_YUV444P || outpl == PE_PALETTE_YUVA4444P)) {
    // convert_splitplanes_frame (:13346-13356, :13438-13448)
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    uint8_t *pl[4] = {(uint8_t *)n.d.planes[0], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2], (uint8_t *)n.d.planes[3]};
    ce = launch_split_planes(L, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, pl, n.d.rowstrides, width, height,
                             inpl == PE_PALETTE_YUVA8888, outpl == PE_PALETTE_YUVA4444P);
  } else if ((inpl == PE_PALETTE_YUV444P && outpl == PE_PALETTE_YUVA4444P) || (inpl == PE_PALETTE_YUVA4444P && outpl == PE_PALETTE_YUV444P)) {
    // convert_yuvp_to_yuvap_frame / convert_yuvap_to_yuvp_frame (:7643-7688): plane copies, alpha = 255 over the whole plane
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    for (int p = 0; p < 3 && ce == cudaSuccess; p++)
      ce = launch_copy2d(L, (const uint8_t *)f->d.planes[p], f->d.rowstrides[p], (uint8_t *)n.d.planes[p], n.d.rowstrides[p], width, height, 0, 0);
    if (ce == cudaSuccess && outpl == PE_PALETTE_YUVA4444P)
      ce = cudaMemsetAsync(n.d.planes[3], 255, (size_t)n.d.rowstrides[3] * height, e->stream);
  } else if ((inpl == PE_PALETTE_YUV420P && outpl == PE_PALETTE_YUV422P) || (inpl == PE_PALETTE_YUV422P && outpl == PE_PALETTE_YUV420P)) {
    // 4:2:0 -> 4:2:2: luma copied, convert_double_chroma(width >> 1, height >> 1) (:13587-13597).
    // 4:2:2 -> 4:2:0: luma copied, convert_halve_chroma (:13700-13710) -- the dispatcher passes height >> 1 where the function
    // expects the SOURCE chroma height, so the reference fills only the top half of the 4:2:0 chroma planes (X); here the whole
    // plane goes through the function's own arithmetic.
    const bool dbl = outpl == PE_PALETTE_YUV422P;
    if (!dbl) n.d.height = height & ~1;
    n.d.width = width & ~1;
    if (n.d.width < 2 || n.d.height < 2) { set_err(PE_ERR_SIZE, "frame too small for a 4:2:x macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    ce = launch_copy2d(L, (const uint8_t *)f->d.planes[0], f->d.rowstrides[0], (uint8_t *)n.d.planes[0], n.d.rowstrides[0], n.d.width,
                       n.d.height, 0, 0);
    // source chroma rows that take part: 4:2:0 -> all (the result has 2 ch rows = the luma height rounded up to even is the
    // caller's business: frames are even, :11603); 4:2:2 -> the first n.d.height rows
    const int ch = dbl ? (n.plane_heights[1] >> 1) : n.d.height;
    if (ce == cudaSuccess)
      ce = launch_resample_chroma_v(L, dbl ? 1 : 0, (const uint8_t *)f->d.planes[1], (const uint8_t *)f->d.planes[2], f->d.rowstrides[1],
                                    f->d.rowstrides[2], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2], n.d.rowstrides[1],
                                    n.d.rowstrides[2], n.d.width >> 1, ch, cavg);
  } else if ((inpl == PE_PALETTE_YUV444P || inpl == PE_PALETTE_YUVA4444P) && (outpl == PE_PALETTE_UYVY || outpl == PE_PALETTE_YUYV)) {
    // convert_yuv_planar_to_{uyvy,yuyv}_frame (:12984-12997, :13083-13096): chroma = avg_chroma of the pixel pair, an odd last
    // column is cut (the layer gets width >> 1 macropixels)
    n.d.width = width & ~1;
    if (n.d.width < 2) { set_err(PE_ERR_SIZE, "frame too narrow for a 4:2:2 macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    const uint8_t *pl[3] = {(const uint8_t *)f->d.planes[0], (const uint8_t *)f->d.planes[1], (const uint8_t *)f->d.planes[2]};
    ce = launch_yuv444p_to_packed422(L, outpl == PE_PALETTE_UYVY ? 0 : 1, pl, f->d.rowstrides[0], Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]},
                                     width >> 1, height, cavg);
  } else if ((inpl == PE_PALETTE_YUV444P || inpl == PE_PALETTE_YUVA4444P) && (outpl == PE_PALETTE_YUV420P || outpl == PE_PALETTE_YVU420P)) {
    // convert_yuvp_to_yuv420_frame (:13016-13022, :13115-13121): luma copied, chroma averaged over 2 x 2; the planes are written
    // Cb to plane 1, Cr to plane 2 (YVU420P: swapped at conv_done, :13895); sampling -> DEFAULT
    n.d.width = width & ~1; n.d.height = height & ~1;
    if (n.d.width < 2 || n.d.height < 2) { set_err(PE_ERR_SIZE, "frame too small for a 4:2:x macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    ce = launch_copy2d(L, (const uint8_t *)f->d.planes[0], f->d.rowstrides[0], (uint8_t *)n.d.planes[0], n.d.rowstrides[0], n.d.width,
                       n.d.height, 0, 0);
    if (ce == cudaSuccess)
      ce = launch_yuv444p_to_chroma420(L, (const uint8_t *)f->d.planes[1], (const uint8_t *)f->d.planes[2], f->d.rowstrides[1],
                                       (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2], n.d.rowstrides[1], n.d.rowstrides[2],
                                       n.d.width >> 1, n.d.height, cavg);
    n.d.yuv_sampling = PE_YUV_SAMPLING_DEFAULT;
  } else if ((inpl == PE_PALETTE_YUV420P || inpl == PE_PALETTE_YUV422P) && (outpl == PE_PALETTE_UYVY || outpl == PE_PALETTE_YUYV)) {
    // convert_yuv420_to_{uyvy,yuyv}_frame (:13561-13574): chroma row k serves luma rows 2k and 2k + 1 (the averaging never runs);
    // convert_yuv422p_to_{uyvy,yuyv}_frame (:13686-13699): interleave (the reference's own row advance is wrong beyond row 0, X)
    n.d.width = width & ~1;
    if (n.d.width < 2) { set_err(PE_ERR_SIZE, "frame too narrow for a 4:2:2 macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *pl[3] = {(const uint8_t *)f->d.planes[0], (const uint8_t *)f->d.planes[1], (const uint8_t *)f->d.planes[2]};
    ce = launch_planar42x_to_packed422(L, outpl == PE_PALETTE_UYVY ? 0 : 1, inpl == PE_PALETTE_YUV422P, pl, f->d.rowstrides,
                                       Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width >> 1, height);
  } else if ((inpl == PE_PALETTE_YUV420P || inpl == PE_PALETTE_YVU420P) && (outpl == PE_PALETTE_YUV444P || outpl == PE_PALETTE_YUVA4444P)) {
    // luma copied, convert_quad_chroma on planes 1 and 2 (:13598-13617; YVU420P with its chroma planes swapped first, :12354), the alpha
    // plane of YUVA4444P = 255 (the reference also asks for it on YUV444P, whose 4th plane pointer is not valid: X)
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    const Planes S = planes_of(f, inpl == PE_PALETTE_YVU420P);
    ce = launch_copy2d(L, S.y, S.rs_y, (uint8_t *)n.d.planes[0], n.d.rowstrides[0], width, height, 0, 0);
    if (ce == cudaSuccess)
      ce = launch_quad_chroma(L, S.u, S.v, S.rs_u, S.rs_v, S.ch, (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2], n.d.rowstrides[1], width,
                              height, isampling == PE_YUV_SAMPLING_JPEG, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (ce == cudaSuccess && outpl == PE_PALETTE_YUVA4444P)
      ce = cudaMemsetAsync(n.d.planes[3], 255, (size_t)n.d.rowstrides[3] * height, e->stream);
  } else if ((inpl == PE_PALETTE_YUV888 || inpl == PE_PALETTE_YUVA8888) &&
             (outpl == PE_PALETTE_UYVY || outpl == PE_PALETTE_YUYV || outpl == PE_PALETTE_YUV422P || outpl == PE_PALETTE_YUV420P ||
              outpl == PE_PALETTE_YVU420P)) {
    // convert_yuv888_to_{uyvy,yuyv,yuv422,yuv420}_frame (:13387-13415, :13479-13507): chroma of a pixel pair = avg_chroma(first,
    // second), 4:2:0 also over the row pair; planes written in layer order; odd last column / row cut
    const bool to420 = outpl == PE_PALETTE_YUV420P || outpl == PE_PALETTE_YVU420P;
    n.d.width = width & ~1;
    if (to420) n.d.height = height & ~1;
    if (n.d.width < 2 || n.d.height < (to420 ? 2 : 1)) { set_err(PE_ERR_SIZE, "frame too small for a 4:2:x macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    uint8_t *pl[3] = {(uint8_t *)n.d.planes[0], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2]};
    const int mode = outpl == PE_PALETTE_UYVY ? 0 : outpl == PE_PALETTE_YUYV ? 1 : outpl == PE_PALETTE_YUV422P ? 2 : 3;
    ce = launch_yuv888_subsample(L, mode, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, inpl == PE_PALETTE_YUVA8888, pl,
                                 n.d.rowstrides, n.d.width, n.d.height, cavg);
  } else if ((inpl == PE_PALETTE_YUV888 && outpl == PE_PALETTE_YUVA8888) || (inpl == PE_PALETTE_YUVA8888 && outpl == PE_PALETTE_YUV888)) {
    // convert_addpost_frame / convert_delpost_frame on the packed YUV bytes (:13337-13342, :13430-13435): the RGB24 <-> RGBA32 kernel
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    ce = launch_rgb_to_rgb(L, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width,
                           height, rgb_layout(inpl == PE_PALETTE_YUV888 ? PE_PALETTE_RGB24 : PE_PALETTE_RGBA32),
                           rgb_layout(outpl == PE_PALETTE_YUV888 ? PE_PALETTE_RGB24 : PE_PALETTE_RGBA32), nullptr);
  } else if ((inpl == PE_PALETTE_UYVY || inpl == PE_PALETTE_YUYV) && (outpl == PE_PALETTE_YUV420P || outpl == PE_PALETTE_YVU420P)) {
    // convert_{uyvy,yuyv}_to_yuv420_frame (:13215-13221, :13315-13321): luma split, chroma averaged over the row pair; planes in
    // layer order; an odd last row is cut
    n.d.height = height & ~1;
    if (n.d.height < 2) { set_err(PE_ERR_SIZE, "frame too small for a 4:2:0 macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    uint8_t *pl[3] = {(uint8_t *)n.d.planes[0], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2]};
    ce = launch_packed422_to_yuv420p(L, inpl == PE_PALETTE_UYVY ? 0 : 1, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, pl,
                                     n.d.rowstrides, width >> 1, n.d.height, cavg);
  } else if ((inpl == PE_PALETTE_YUV420P || inpl == PE_PALETTE_YVU420P || inpl == PE_PALETTE_YUV422P) &&
             (outpl == PE_PALETTE_YUV888 || outpl == PE_PALETTE_YUVA8888)) {
    // convert_quad_chroma_packed (:13624-13635) / convert_double_chroma_packed (:13730-13742): chroma up-sampled on the fly, alpha 255
    // on EVERY pixel (the reference skips the alpha of odd rows / second pixels: X)
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    const Planes S = planes_of(f, inpl == PE_PALETTE_YVU420P);
    const uint8_t *pl[3] = {S.y, S.u, S.v};
    const int irs[3] = {S.rs_y, S.rs_u, S.rs_v};
    ce = launch_chroma_upsample_packed(L, inpl != PE_PALETTE_YUV422P, pl, irs, S.ch, Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width, height,
                                       outpl == PE_PALETTE_YUVA8888, isampling == PE_YUV_SAMPLING_JPEG, iclamping == PE_YUV_CLAMPING_CLAMPED);
  } else if ((inpl == PE_PALETTE_YUV420P && outpl == PE_PALETTE_YVU420P) || (inpl == PE_PALETTE_YVU420P && outpl == PE_PALETTE_YUV420P)) {
    // pconv_can_inplace (:12152-12155): no pixel work (:13618-13623) -- the chroma plane pointers and rowstrides change places, on the
    // way in for a V-first source (:12354), on the way out for a V-first target (:13895, below)
    inplace = true;
    if (inpl == PE_PALETTE_YVU420P) { std::swap(f->d.planes[1], f->d.planes[2]); std::swap(f->d.rowstrides[1], f->d.rowstrides[2]); }
  } else if (outpl == PE_PALETTE_YUV411 &&
             (inpl == PE_PALETTE_UYVY || inpl == PE_PALETTE_YUYV || inpl == PE_PALETTE_YUV420P || inpl == PE_PALETTE_YVU420P ||
              inpl == PE_PALETTE_YUV422P || inpl == PE_PALETTE_YUV888 || inpl == PE_PALETTE_YUVA8888 || inpl == PE_PALETTE_YUV444P ||
              inpl == PE_PALETTE_YUVA4444P)) {
    // convert_{uyvy,yuyv}_to_yuv411_frame (:13227, :13327), convert_yuv420_to_yuv411_frame (:13640, :13747),
    // convert_yuv888_to_yuv411_frame (:13420, :13512; every row here, the reference stops after a third / a quarter of them, X),
    // convert_yuvp_to_yuv411_frame (:13028, :13127; per macropixel, the reference never advances its output pointer, X)
    n.d.width = width & ~3;
    if (n.d.width < 4) { set_err(PE_ERR_SIZE, "frame too narrow for a YUV411 macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    const int mode = inpl == PE_PALETTE_UYVY ? 0 : inpl == PE_PALETTE_YUYV ? 1 : (inpl == PE_PALETTE_YUV420P || inpl == PE_PALETTE_YVU420P) ? 2
                     : inpl == PE_PALETTE_YUV422P ? 3 : inpl == PE_PALETTE_YUV888 ? 4 : inpl == PE_PALETTE_YUVA8888 ? 5 : 6;
    const bool swap = inpl == PE_PALETTE_YVU420P;
    const uint8_t *pl[3] = {(const uint8_t *)f->d.planes[0], (const uint8_t *)f->d.planes[swap ? 2 : 1], (const uint8_t *)f->d.planes[swap ? 1 : 2]};
    const int rs[3] = {f->d.rowstrides[0], f->d.rowstrides[swap ? 2 : 1], f->d.rowstrides[swap ? 1 : 2]};
    ce = launch_to_yuv411(L, mode, pl, rs, width >> 2, height, Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, cavg);
  } else if (pal_is_rgb(inpl) && outpl == PE_PALETTE_YUV411) {
    // convert_{rgb,bgr,argb}_to_yuv411_frame (:12627, :12705, :12779, :12852, :12925): YCbCr tables of oclamping, whole macropixels
    // ("cut the rightmost one, two or three pixels": the layer becomes width >> 2 macropixels wide)
    n.d.width = width & ~3;
    if (n.d.width < 4) { set_err(PE_ERR_SIZE, "frame too narrow for a YUV411 macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    ce = launch_rgb_to_yuv411(L, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]},
                              width >> 2, height, rgb_layout(inpl), dev_conv(e, oclamping, PE_YUV_SUBSPACE_YCBCR));
  } else if (inpl == PE_PALETTE_YUV411 && (pal_is_rgb(outpl) || outpl == PE_PALETTE_YUV888 || outpl == PE_PALETTE_YUVA8888 ||
                                           outpl == PE_PALETTE_YUV444P || outpl == PE_PALETTE_YUVA4444P || outpl == PE_PALETTE_UYVY ||
                                           outpl == PE_PALETTE_YUYV || outpl == PE_PALETTE_YUV422P || outpl == PE_PALETTE_YUV420P ||
                                           outpl == PE_PALETTE_YVU420P)) {
    // convert_yuv411_to_{rgb,bgr,argb,yuv888,yuvp,uyvy,yuyv}_frame (:13755-13826): every variant selects the YCbCr tables of the
    // layer's clamping itself (set_conversion_arrays(clamping, WEED_YUV_SUBSPACE_YCBCR)); the layer's width in macropixels is a
    // quarter of the pixel width.  (The reference walks source and most destinations densely; strides are honoured here, X.)
    if (width & 3) { set_err(PE_ERR_SIZE, "a YUV411 frame is a whole number of 4-pixel macropixels wide"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    const int target = pal_is_rgb(outpl) ? 0 : (outpl == PE_PALETTE_YUV888 || outpl == PE_PALETTE_YUVA8888) ? 1
                       : (outpl == PE_PALETTE_YUV444P || outpl == PE_PALETTE_YUVA4444P) ? 2 : outpl == PE_PALETTE_UYVY ? 3
                       : outpl == PE_PALETTE_YUYV ? 4 : outpl == PE_PALETTE_YUV422P ? 5 : 6;   // (4:2:0: Cb to plane 1, swapped below for YVU)
    uint8_t *pl[4] = {(uint8_t *)n.d.planes[0], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2], (uint8_t *)n.d.planes[3]};
    ce = launch_yuv411_to(L, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, width >> 2, height, pl, n.d.rowstrides, target,
                          pal_has_alpha(outpl), pal_is_rgb(outpl) ? rgb_layout(outpl) : RgbLayout{0, 1, 2, -1, 3},
                          e->cfg.ref_quirks && (outpl == PE_PALETTE_BGR24 || outpl == PE_PALETTE_BGRA32),
                          dev_conv(e, iclamping, PE_YUV_SUBSPACE_YCBCR), cavg);
  } else if ((inpl == PE_PALETTE_UYVY && outpl == PE_PALETTE_YUYV) || (inpl == PE_PALETTE_YUYV && outpl == PE_PALETTE_UYVY)) {
    // convert_swab_frame in place (:13138-13140, :13238-13240)
    inplace = true;
    ce = launch_swab(L, Img{(uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, width >> 1, height);
  } else if ((inpl == PE_PALETTE_UYVY || inpl == PE_PALETTE_YUYV) &&
             (outpl == PE_PALETTE_YUV422P || outpl == PE_PALETTE_YUV444P || outpl == PE_PALETTE_YUVA4444P || outpl == PE_PALETTE_YUV888 ||
              outpl == PE_PALETTE_YUVA8888)) {
    // convert_{uyvy,yuyv}_to_yuv422_frame (:13141-13146, :13241-13246; with ref_quirks the reference's never-advanced source
    // pointer, :8103: the whole frame is its first macropixel), _to_yuvp_frame (:13189-13200), _to_yuv888_frame (:13203-13214)
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const int mode = outpl == PE_PALETTE_YUV422P ? 0 : (outpl == PE_PALETTE_YUV888 || outpl == PE_PALETTE_YUVA8888) ? 2 : 1;
    uint8_t *pl[4] = {(uint8_t *)n.d.planes[0], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2], (uint8_t *)n.d.planes[3]};
    ce = launch_packed422_unpack(L, inpl == PE_PALETTE_UYVY ? 0 : 1, mode, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, pl,
                                 n.d.rowstrides, width >> 1, height, outpl == PE_PALETTE_YUVA4444P || outpl == PE_PALETTE_YUVA8888,
                                 mode == 0 && e->cfg.ref_quirks);
    if (ce == cudaSuccess && outpl == PE_PALETTE_YUVA4444P)  // lives_memset(dest[3], 255, orow[3] * height): padding included
      ce = cudaMemsetAsync(n.d.planes[3], 255, (size_t)n.d.rowstrides[3] * height, e->stream);
  } else if (((inpl == PE_PALETTE_YUV444P || inpl == PE_PALETTE_YUVA4444P) && outpl == PE_PALETTE_YUV422P) ||
             (inpl == PE_PALETTE_YUV422P && (outpl == PE_PALETTE_YUV444P || outpl == PE_PALETTE_YUVA4444P))) {
    // 4:4:4 planar <-> 4:2:2 planar.  The reference's dispatcher hands these to the VERTICAL convert_halve_chroma / convert_double_chroma
    // (:12948, :13719: full-width rows written into half-width planes, half of the target never written) -- nothing defined to
    // replicate (X).  Defined here through the reference's own HORIZONTAL converters of the same two samplings, composed:
    //   4:4:4 planar -> (convert_combineplanes_frame) YUV888 -> (convert_yuv888_to_yuv422_frame: avg_chroma of the pixel pair) 4:2:2 planar
    //   4:2:2 planar -> (convert_double_chroma_packed) YUV888 -> (convert_splitplanes_frame) 4:4:4 planar, alpha plane 255
    const bool down = outpl == PE_PALETTE_YUV422P;
    if (down) n.d.width = width & ~1;
    if (n.d.width < 2) { set_err(PE_ERR_SIZE, "frame too narrow for a 4:2:2 macropixel"); return PE_FALSE; }
    if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
    const uint8_t *cavg = get_cavg(e, iclamping == PE_YUV_CLAMPING_CLAMPED);
    if (!cavg) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "averaging table could not be built"); return PE_FALSE; }
    const int trs = align_ceil(width * 3, 32);
    size_t granted = 0;
    uint8_t *tmp = (uint8_t *)e->pool.get((size_t)trs * height, &granted);
    if (!tmp) { frame_release_pixels(&n); set_err(PE_ERR_MEMORY, "device allocation of %zu bytes failed", (size_t)trs * height); return PE_FALSE; }
    if (down) {
      const uint8_t *pl[4] = {(const uint8_t *)f->d.planes[0], (const uint8_t *)f->d.planes[1], (const uint8_t *)f->d.planes[2], nullptr};
      ce = launch_combine_planes(L, pl, f->d.rowstrides[0], Img{tmp, trs}, width, height, 0, 0);
      uint8_t *opl[3] = {(uint8_t *)n.d.planes[0], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2]};
      if (ce == cudaSuccess) ce = launch_yuv888_subsample(L, 2, CImg{tmp, trs}, 0, opl, n.d.rowstrides, n.d.width, height, cavg);
    } else {
      const uint8_t *pl[3] = {(const uint8_t *)f->d.planes[0], (const uint8_t *)f->d.planes[1], (const uint8_t *)f->d.planes[2]};
      ce = launch_chroma_upsample_packed(L, 0, pl, f->d.rowstrides, height, Img{tmp, trs}, width, height, 0, isampling == PE_YUV_SAMPLING_JPEG, iclamping == PE_YUV_CLAMPING_CLAMPED);
      uint8_t *opl[4] = {(uint8_t *)n.d.planes[0], (uint8_t *)n.d.planes[1], (uint8_t *)n.d.planes[2], nullptr};
      if (ce == cudaSuccess) ce = launch_split_planes(L, CImg{tmp, trs}, opl, n.d.rowstrides, width, height, 0, 0);
      if (ce == cudaSuccess && outpl == PE_PALETTE_YUVA4444P) ce = cudaMemsetAsync(n.d.planes[3], 255, (size_t)n.d.rowstrides[3] * height, e->stream);
    }
    e->pool.put(tmp, granted);  // stream ordered: the next user is enqueued behind the two kernels
  } else {
    set_err(PE_ERR_PALETTE, "palette conversion %d -> %d is not handled by this build", inpl, outpl);
    return PE_FALSE;  // memfail: the layer is left as it was
  }
  if (ce != cudaSuccess) {
    set_err(PE_ERR_CUDA, "conversion kernel launch failed: %s", cudaGetErrorString(ce));
    if (!inplace) frame_release_pixels(&n);
    return PE_FALSE;
  }
  if (inplace) f->d.palette = outpl;
  else frame_take(f, &n);
  // "if V plane is before U, swap the pointers" (:13895): every converter above wrote Cb to plane 1 and Cr to plane 2 (the reference's
  // dest[1] / dest[2]); a V-first palette has them the other way round in the finished layer
  if (outpl == PE_PALETTE_YVU420P) { std::swap(f->d.planes[1], f->d.planes[2]); std::swap(f->d.rowstrides[1], f->d.rowstrides[2]); }
  // the converters that produce 4:2:0 / 4:2:2 planes from full-resolution chroma leave WEED_YUV_SAMPLING_DEFAULT (:12691, :13022)
  if ((outpl == PE_PALETTE_YUV420P || outpl == PE_PALETTE_YVU420P || outpl == PE_PALETTE_YUV422P) &&
      (pal_is_rgb(inpl) || inpl == PE_PALETTE_YUV444P || inpl == PE_PALETTE_YUVA4444P))
    f->d.yuv_sampling = PE_YUV_SAMPLING_DEFAULT;

  // conv_done (:13859-13900)
  if (new_gamma_type != PE_GAMMA_UNKNOWN && can_inline_gamma(inpl, outpl)) {
    f->d.gamma_type = new_gamma_type;
    gamma_type = new_gamma_type;
  }
  if (pal_is_rgb(outpl)) {
    f->d.yuv_clamping = 0; f->d.yuv_subspace = 0; f->d.yuv_sampling = 0;  // leaves deleted
  } else {
    f->d.yuv_clamping = oclamping;
    if (pal_is_rgb(inpl)) f->d.yuv_subspace = gamma_type == PE_GAMMA_BT709 ? PE_YUV_SUBSPACE_BT709 : PE_YUV_SUBSPACE_YCBCR;
  }
  return PE_TRUE;
}

}  // namespace

// ---- the reference's float ("experimental") YUV -> RGB path (colourspace.c:101-172, :592, :2367) ---------------------------------
extern "C" int pe_float_yuv_table(int clamping, int which, float out[256]) {
  if (which < 0 || which > 4 || !out) return set_err(PE_ERR_ARG, "table 0 .. 4 (RGBf_Y, Rf_Cr, Gf_Cb, Gf_Cr, Bf_Cb)");
  float t[5][256];
  build_float_yuv_tables(clamping, t);
  memcpy(out, t[which], sizeof(t[which]));
  return PE_OK;
}

extern "C" int pe_convert_yuv888_to_rgb_float(pe_engine_t *e, pe_frame_t *f, int outpl, int mode, float *sums_host) {
  if (!e || !f || !f->d.planes[0]) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  const int inpl = f->d.palette;
  if ((inpl != PE_PALETTE_YUV888 && inpl != PE_PALETTE_YUVA8888) || !pal_is_rgb(outpl)) {
    set_err(PE_ERR_PALETTE, "float path: YUV888 / YUVA8888 -> an RGB palette (%d -> %d asked)", inpl, outpl);
    return PE_FALSE;
  }
  if (f->d.yuv_subspace != PE_YUV_SUBSPACE_BT709) {
    set_err(PE_ERR_PALETTE, "float path: the reference only has float tables for the BT.709 subspace (colourspace.c:279-310, :316-355)");
    return PE_FALSE;
  }
  if (mode != 0 && mode != 1) { set_err(PE_ERR_ARG, "mode 0 (yuv2rgb_float as written) or 1 (RGBf_Y form)"); return PE_FALSE; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (cudaSetDevice(e->device) != cudaSuccess) { set_err(PE_ERR_CUDA, "cudaSetDevice failed"); return PE_FALSE; }
  const int ci = f->d.yuv_clamping == PE_YUV_CLAMPING_UNCLAMPED ? 1 : 0;
  if (!e->ftab_dev[ci]) {
    float t[5][256];
    build_float_yuv_tables(ci ? PE_YUV_CLAMPING_UNCLAMPED : PE_YUV_CLAMPING_CLAMPED, t);
    float *d = nullptr;
    // stream-ordered upload (see conv_dev in pe_engine_create: a synchronous copy from the stack is not ordered with e->stream)
    if (cudaMalloc(&d, sizeof(t)) != cudaSuccess || cudaMemcpyAsync(d, t, sizeof(t), cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
        cudaStreamSynchronize(e->stream) != cudaSuccess) {
      if (d) cudaFree(d);
      set_err(PE_ERR_CUDA, "float table upload failed");
      return PE_FALSE;
    }
    e->ftab_dev[ci] = d;
  }
  const DevConv conv = dev_conv(e, f->d.yuv_clamping, PE_YUV_SUBSPACE_BT709);
  pe_frame n;
  n.e = e;
  n.d = f->d;
  n.d.palette = outpl;
  if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
  const int width = f->d.width, height = f->d.height;
  float *sums_dev = nullptr;
  size_t granted = 0;
  const size_t sums_bytes = sizeof(float) * 3 * (size_t)width * height;
  if (sums_host && !(sums_dev = (float *)e->pool.get(sums_bytes, &granted))) {
    frame_release_pixels(&n);
    set_err(PE_ERR_MEMORY, "device allocation of %zu bytes failed", sums_bytes);
    return PE_FALSE;
  }
  cudaError_t ce = launch_yuv888_to_rgb_float(e->L(), mode, CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]},
                                              Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, width, height, inpl == PE_PALETTE_YUVA8888,
                                              rgb_layout(outpl), e->ftab_dev[ci], conv.t + 9 * 256, sums_dev);
  if (ce == cudaSuccess && sums_dev) {
    ce = cudaMemcpyAsync(sums_host, sums_dev, sums_bytes, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  }
  if (sums_dev) e->pool.put(sums_dev, granted);
  if (ce != cudaSuccess) {
    frame_release_pixels(&n);
    set_err(PE_ERR_CUDA, "float conversion failed: %s", cudaGetErrorString(ce));
    return PE_FALSE;
  }
  frame_take(f, &n);
  f->d.yuv_clamping = 0; f->d.yuv_subspace = 0; f->d.yuv_sampling = 0;  // conv_done (:13859-13900): the YUV leaves are deleted
  return PE_TRUE;
}

extern "C" int pe_mc_publish(pe_engine_t *e, void *mc_dst, const void *src, size_t bytes, void *cuda_stream, int max_ctas) {
  if (!e || !mc_dst || !src) return set_err(PE_ERR_ARG, "NULL argument");
  if ((bytes & 15) || ((uintptr_t)mc_dst & 15) || ((uintptr_t)src & 15)) return set_err(PE_ERR_ARG, "pe_mc_publish: 16-byte granularity");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  pe::Launch L = e->L();
  if (cuda_stream) L.stream = (cudaStream_t)cuda_stream;
  PE_CUDA(launch_mc_publish(L, src, mc_dst, bytes, max_ctas));
  return PE_OK;
}

extern "C" int pe_convert_layer_palette_full(pe_engine_t *e, pe_frame_t *layer, int outpl, int oclamping, int osampling,
                                             int osubspace, int tgt_gamma) {
  if (!e || !layer) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA_B(cudaSetDevice(e->device));
  return convert_locked(e, layer, outpl, oclamping, osampling, osubspace, tgt_gamma);
}

extern "C" int pe_convert_layer_palette(pe_engine_t *e, pe_frame_t *layer, int outpl, int op_clamping) {
  return pe_convert_layer_palette_full(e, layer, outpl, op_clamping, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YUV,
                                       PE_GAMMA_UNKNOWN);
}

// ---------------------------------------------------------------------------------------------------------
// resize_layer_full (colourspace.c:14759) / letterbox_layer (:15343)
// ---------------------------------------------------------------------------------------------------------

namespace {

// one plane (or packed image) src -> dst through the separable filter bank
int resize_plane(pe_engine *e, const uint8_t *src, int srs, int sw, int sh, uint8_t *dst, int drs, int dw, int dh, int psize, AxisKinds kinds) {
  DevFilterEntry *fx = get_filter(e, sw, dw, 14, kinds.x), *fy = get_filter(e, sh, dh, 12, kinds.y);
  if (!fx || !fy) return set_err(PE_ERR_SIZE, "scale factor out of range (%dx%d -> %dx%d; at most 31x down)", sw, sh, dw, dh);
  if (e->rsz_defer && psize == 4 && fx->host.fast_taps() <= 4 && fy->host.fast_taps() <= 4) {  // batch call: launched by flush_rsz_pending
    e->rsz_pending.push_back(pe_engine::RszJob{src, srs, sw, sh, dst, drs, dw, dh, psize, kinds.x, kinds.y});
    return PE_OK;
  }
  // one tiled kernel (15-bit intermediate stays in shared memory) ...
  cudaError_t te = launch_resize_tile(e->L(), CImg{src, srs}, sw, sh, Img{dst, drs}, dw, dh, psize, fx->dev, fy->dev,
                                      fx->host.first.data(), fy->host.first.data());
  if (te == cudaSuccess) return PE_OK;
  if (te != cudaErrorInvalidConfiguration) return set_err(PE_ERR_CUDA, "resize launch failed: %s", cudaGetErrorString(te));
  // ... or, for scale factors whose source rectangle does not fit there, the two-kernel path through an HBM intermediate
  size_t granted = 0;
  const size_t tmp_bytes = sizeof(int16_t) * (size_t)sh * dw * psize;
  int16_t *tmp = (int16_t *)e->pool.get(tmp_bytes, &granted);
  if (!tmp) return set_err(PE_ERR_MEMORY, "device allocation of %zu bytes failed", tmp_bytes);
  cudaError_t ce = launch_resize_h(e->L(), CImg{src, srs}, sw, sh, tmp, dw, psize, fx->dev);
  if (ce == cudaSuccess) ce = launch_resize_v(e->L(), tmp, sh, Img{dst, drs}, dw, dh, psize, fy->dev);
  e->pool.put(tmp, granted);  // stream ordered: the next user is enqueued behind the two kernels
  if (ce != cudaSuccess) return set_err(PE_ERR_CUDA, "resize launch failed: %s", cudaGetErrorString(ce));
  return PE_OK;
}

int resize_locked(pe_engine *e, pe_frame *f, int width, int height, int interp, int opal_hint, int oclamp_hint, int osamp_hint,
                  int osubs_hint, int tgt_gamma) {
  int palette = f->d.palette;
  if (opal_hint == PE_PALETTE_NONE) opal_hint = palette;  // WEED_PALETTE_ANY: keep
  if (!f->d.planes[0]) {  // :14820-14832
    f->d.width = width; f->d.height = height;
    if (pal_known(opal_hint)) f->d.palette = opal_hint;
    f->d.yuv_clamping = oclamp_hint;
    return PE_FALSE;
  }
  if (width <= 0 || height <= 0) return PE_FALSE;  // :14834
  int iwidth = (f->d.width >> 1) << 1, iheight = (f->d.height >> 1) << 1;  // :14850-14851
  if (width < 4) width = 4;
  if (height < 4) height = 4;
  if (iwidth != width || iheight != height) height = (height >> 1) << 1;  // :14856-14859
  if (iwidth == width && iheight == height) return PE_TRUE;

  // tgt_gamma resolution (:14884-14893)
  if (tgt_gamma == PE_GAMMA_UNKNOWN && pal_is_yuv(opal_hint) && osubs_hint == PE_YUV_SUBSPACE_BT709) tgt_gamma = PE_GAMMA_BT709;
  if (tgt_gamma == PE_GAMMA_UNKNOWN) tgt_gamma = f->d.gamma_type;
  if (tgt_gamma == PE_GAMMA_BT709 && pal_is_yuv(opal_hint)) osubs_hint = PE_YUV_SUBSPACE_BT709;

  // get_resizable (:14577): what gets scaled.  Packed 3/4-byte palettes and planar YUV scale as they are; packed
  // 4:2:2 is converted first.  A YUV source with an RGB target is converted BEFORE scaling, with the reference's own
  // converter arithmetic (the reference lets swscale do both at once, :14601-14620 -- see DESIGN.md "resize").
  int resolved = palette;
  if (pal_is_yuv(palette) && pal_is_rgb(opal_hint)) resolved = opal_hint;
  else if (palette == PE_PALETTE_UYVY || palette == PE_PALETTE_YUYV) resolved = pal_is_rgb(opal_hint) ? opal_hint : PE_PALETTE_RGB24;
  if (resolved != palette) {
    if (!convert_locked(e, f, resolved, oclamp_hint, osamp_hint, osubs_hint, tgt_gamma)) return PE_FALSE;
    if (f->d.palette != resolved) return PE_FALSE;
    palette = resolved;
    iwidth = (f->d.width >> 1) << 1; iheight = (f->d.height >> 1) << 1;
    if (iwidth == width && iheight == height) return PE_TRUE;
  }

  pe_frame n;
  n.e = e;
  n.d.palette = palette; n.d.width = width; n.d.height = height;
  if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
  int rc = PE_OK;
  const AxisKinds kinds = filter_kinds(e, interp, iwidth, iheight, width, height);  // one flag per call (:14991-14997), every plane
  if (!pal_is_planar(palette)) {
    rc = resize_plane(e, (const uint8_t *)f->d.planes[0], f->d.rowstrides[0], f->d.width, f->d.height, (uint8_t *)n.d.planes[0],
                      n.d.rowstrides[0], width, height, pal_psize(palette), kinds);
  } else {
    for (int p = 0; p < n.d.nplanes && rc == PE_OK; p++) {
      const bool sub_h = p > 0 && p < 3 && palette != PE_PALETTE_YUV444P && palette != PE_PALETTE_YUVA4444P;
      const int sw = sub_h ? f->d.width >> 1 : f->d.width, dw = sub_h ? width >> 1 : width;
      rc = resize_plane(e, (const uint8_t *)f->d.planes[p], f->d.rowstrides[p], sw, f->plane_heights[p], (uint8_t *)n.d.planes[p],
                        n.d.rowstrides[p], dw, n.plane_heights[p], 1, kinds);
    }
  }
  if (rc != PE_OK) { frame_release_pixels(&n); return PE_FALSE; }
  frame_take(f, &n);
  if (opal_hint != palette && pal_known(opal_hint)) {
    if (!convert_locked(e, f, opal_hint, oclamp_hint, osamp_hint, osubs_hint, tgt_gamma)) return PE_FALSE;
  }
  return PE_TRUE;
}

int letterbox_locked(pe_engine *e, pe_frame *f, int nwidth, int nheight, int width, int height, int interp, int tpal, int tclamp) {
  if (!width || !height || !nwidth || !nheight) return PE_TRUE;  // :15375
  if (nwidth < width) nwidth = width;
  if (nheight < height) nheight = height;
  if (nheight == height && nwidth == width) {  // :15380-15383
    resize_locked(e, f, width, height, interp, tpal, tclamp, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YCBCR, PE_GAMMA_UNKNOWN);
    return PE_TRUE;
  }
  if (f->d.width != width || f->d.height != height) {  // :15388-15393
    if (!resize_locked(e, f, width, height, interp, tpal, tclamp, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YCBCR, PE_GAMMA_UNKNOWN))
      return PE_FALSE;
  }
  width = f->d.width; height = f->d.height;
  const int pal = f->d.palette;
  pe_frame n;
  n.e = e;
  n.d = f->d;
  n.d.width = nwidth; n.d.height = nheight;
  if (pal == PE_PALETTE_YUV420P || pal == PE_PALETTE_YVU420P) { n.d.width &= ~1; n.d.height &= ~1; }
  if (n.d.width < width || n.d.height < height) return PE_FALSE;  // :15503
  if (frame_alloc(e, &n) != PE_OK) return PE_FALSE;
  nwidth = n.d.width; nheight = n.d.height;
  const int offs_x = (nwidth - width + 1) >> 1, offs_y = (nheight - height + 1) >> 1;  // :15522-15523
  cudaError_t ce = cudaSuccess;
  if (!pal_is_planar(pal) && pal_ppmp(pal) == 1) {
    uint32_t black;
    const uint32_t y0 = (pal_is_yuv(pal) && f->d.yuv_clamping == PE_YUV_CLAMPING_CLAMPED) ? 16 : 0;
    switch (pal) {
    case PE_PALETTE_RGB24: case PE_PALETTE_BGR24: black = 0; break;
    case PE_PALETTE_RGBA32: case PE_PALETTE_BGRA32: black = 0xFF000000u; break;
    case PE_PALETTE_ARGB32: black = 0x000000FFu; break;
    case PE_PALETTE_YUV888: black = y0 | (128u << 8) | (128u << 16); break;
    default: black = y0 | (128u << 8) | (128u << 16) | 0xFF000000u; break;
    }
    ce = launch_letterbox(e->L(), CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, width, height,
                          Img{(uint8_t *)n.d.planes[0], n.d.rowstrides[0]}, nwidth, nheight, pal_psize(pal), black);
  } else {
    // planar / macropixel palettes: black frame, then one 2-D copy per plane with the offsets scaled by the plane's
    // subsampling ratios (:15538-15549)
    if (frame_fill_black(e, &n) != PE_OK) { frame_release_pixels(&n); return PE_FALSE; }
    for (int p = 0; p < n.d.nplanes && ce == cudaSuccess; p++) {
      const int rb = plane_row_bytes(f->d, p);
      const int hr = p == 0 ? 1 : f->d.width / (rb ? rb : 1);                     // horizontal ratio denominator
      const int vr = p == 0 ? 1 : f->d.height / (f->plane_heights[p] ? f->plane_heights[p] : 1);
      const int xo = p == 0 ? (offs_x / pal_ppmp(pal)) * pal_psize(pal) : offs_x / (hr ? hr : 1);
      const int yo = offs_y / (vr ? vr : 1);
      ce = launch_copy2d(e->L(), (const uint8_t *)f->d.planes[p], f->d.rowstrides[p],
                         (uint8_t *)n.d.planes[p] + (size_t)yo * n.d.rowstrides[p] + xo, n.d.rowstrides[p], rb,
                         f->plane_heights[p], 0, 0);
    }
  }
  if (ce != cudaSuccess) {
    set_err(PE_ERR_CUDA, "letterbox launch failed: %s", cudaGetErrorString(ce));
    frame_release_pixels(&n);
    return PE_FALSE;
  }
  frame_take(f, &n);
  return PE_TRUE;
}

}  // namespace

extern "C" int pe_resize_layer_full(pe_engine_t *e, pe_frame_t *layer, int width, int height, int interp, int opal_hint,
                                    int oclamp_hint, int osamp_hint, int osubs_hint, int tgt_gamma) {
  if (!e || !layer) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA_B(cudaSetDevice(e->device));
  return resize_locked(e, layer, width, height, interp, opal_hint, oclamp_hint, osamp_hint, osubs_hint, tgt_gamma);
}

namespace {

// launch the planar YUV -> RGB conversions a batch call has queued: runs of same-shaped frames leave as one launch per 32
int flush_yuv_pending(pe_engine *e) {
  std::vector<YuvToRgbArgs> q;
  q.swap(e->yuv_pending);
  size_t i = 0;
  while (i < q.size()) {
    size_t j = i + 1;
    while (j < q.size() && yuv_planar_same_shape(q[i], q[j])) j++;
    cudaError_t ce = j - i == 1 ? launch_yuv_planar_to_rgb_fast(e->L(), q[i]) : launch_yuv_planar_to_rgb_batch(e->L(), &q[i], (int)(j - i));
    if (ce != cudaSuccess) return set_err(PE_ERR_CUDA, "conversion kernel launch failed: %s", cudaGetErrorString(ce));
    i = j;
  }
  return PE_OK;
}

// launch the RGB <-> RGB permutations a batch call has queued: runs of same-shaped frames leave as one launch per 128
int flush_rgb_pending(pe_engine *e) {
  std::vector<pe_engine::RgbJob> q;
  q.swap(e->rgb_pending);
  e->rgb_defer = false;
  auto same_lay = [](const RgbLayout &a, const RgbLayout &b) { return a.r == b.r && a.g == b.g && a.b == b.b && a.a == b.a && a.psize == b.psize; };
  size_t i = 0;
  while (i < q.size()) {
    size_t j = i + 1;
    while (j < q.size() && q[j].irow == q[i].irow && q[j].orow == q[i].orow && q[j].width == q[i].width && q[j].height == q[i].height &&
           q[j].lut == q[i].lut && same_lay(q[j].in, q[i].in) && same_lay(q[j].out, q[i].out))
      j++;
    std::vector<const uint8_t *> srcs;
    std::vector<uint8_t *> dsts;
    for (size_t k = i; k < j; k++) { srcs.push_back(q[k].src); dsts.push_back(q[k].dst); }
    cudaError_t ce = launch_rgb_to_rgb_batch(e->L(), srcs.data(), q[i].irow, dsts.data(), q[i].orow, (int)(j - i), q[i].width, q[i].height,
                                             q[i].in, q[i].out, q[i].lut);
    if (ce != cudaSuccess) return set_err(PE_ERR_CUDA, "conversion kernel launch failed: %s", cudaGetErrorString(ce));
    i = j;
  }
  return PE_OK;
}

// launch the packed resizes a batch call has queued: runs of same-shaped frames leave as one launch per 32
int flush_rsz_pending(pe_engine *e) {
  std::vector<pe_engine::RszJob> q;
  q.swap(e->rsz_pending);
  e->rsz_defer = false;  // (resize_plane below must launch)
  size_t i = 0;
  int rc = PE_OK;
  while (i < q.size() && rc == PE_OK) {
    size_t j = i + 1;
    auto same = [&](const pe_engine::RszJob &a, const pe_engine::RszJob &b) {
      return a.srs == b.srs && a.sw == b.sw && a.sh == b.sh && a.drs == b.drs && a.dw == b.dw && a.dh == b.dh && a.psize == b.psize && a.kx == b.kx && a.ky == b.ky;
    };
    while (j < q.size() && same(q[i], q[j])) j++;
    bool done = false;
    if (j - i > 1) {
      DevFilterEntry *fx = get_filter(e, q[i].sw, q[i].dw, 14, q[i].kx), *fy = get_filter(e, q[i].sh, q[i].dh, 12, q[i].ky);
      if (fx && fy) {
        std::vector<const uint8_t *> srcs;
        std::vector<uint8_t *> dsts;
        for (size_t k = i; k < j; k++) { srcs.push_back(q[k].src); dsts.push_back(q[k].dst); }
        cudaError_t te = launch_resize_tile_batch(e->L(), srcs.data(), q[i].srs, q[i].sw, q[i].sh, dsts.data(), q[i].drs, q[i].dw, q[i].dh,
                                                  q[i].psize, fx->dev, fy->dev, fx->host.first.data(), fy->host.first.data(), (int)(j - i));
        if (te == cudaSuccess) done = true;
        else if (te != cudaErrorInvalidConfiguration) rc = set_err(PE_ERR_CUDA, "resize launch failed: %s", cudaGetErrorString(te));
      }
    }
    if (!done && rc == PE_OK)
      for (size_t k = i; k < j && rc == PE_OK; k++)
        rc = resize_plane(e, q[k].src, q[k].srs, q[k].sw, q[k].sh, q[k].dst, q[k].drs, q[k].dw, q[k].dh, q[k].psize, AxisKinds{q[k].kx, q[k].ky});
    i = j;
  }
  return rc;
}

// A batch call queued both the planar -> RGB conversions (yuv_pending) and the resizes of their results (rsz_pending).  When every
// pair (conversion i, resize i) meets through an intermediate frame that nothing else reads, the pairs leave as ONE k_cvt_resize
// launch per 32 (pe_kernels_fused4.cu) and the intermediate frames are never written; otherwise the queues are flushed in order.
// k_cvt_resize's view of a <= 4-tap non-negative bank, built once per cached bank: per output sample the four coefficients as 16-bit
// pairs, the first source index, and what the pass does to an all-255 alpha channel -- horizontal (14 bit): min((sum c * 255) >> 7,
// 32767), the 15-bit intermediate of alpha; vertical (12 bit): sum c, its multiplier
const void *get_pack4(pe_engine *e, DevFilterEntry *f, int bits) {
  if (f->pack4) return f->pack4;
  const int n = (int)f->host.first.size(), taps = f->host.taps;
  if (taps > 4) return nullptr;
  std::vector<int32_t> h((size_t)n * 4);
  for (int i = 0; i < n; i++) {
    uint32_t c[4] = {0, 0, 0, 0};
    int sum = 0;
    for (int k = 0; k < taps; k++) { c[k] = (uint16_t)f->host.coef[(size_t)i * taps + k]; sum += f->host.coef[(size_t)i * taps + k]; }
    h[4 * (size_t)i] = (int32_t)(c[0] | (c[1] << 16));
    h[4 * (size_t)i + 1] = (int32_t)(c[2] | (c[3] << 16));
    h[4 * (size_t)i + 2] = f->host.first[(size_t)i];
    h[4 * (size_t)i + 3] = bits == 14 ? std::min((sum * 255) >> 7, 32767) : sum;
  }
  void *d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(int32_t)) != cudaSuccess) return nullptr;
  if (cudaMemcpyAsync(d, h.data(), h.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
      cudaStreamSynchronize(e->stream) != cudaSuccess) {  // (h is pageable and dies with this call)
    cudaFree(d);
    return nullptr;
  }
  f->pack4 = d;
  return d;
}

int flush_cvt_rsz_pending(pe_engine *e) {
  std::vector<YuvToRgbArgs> &Y = e->yuv_pending;
  std::vector<pe_engine::RszJob> &R = e->rsz_pending;
  bool fuse = getenv("PE_NO_CVT_RESIZE") == nullptr && !Y.empty() && Y.size() == R.size();
  for (size_t i = 0; i < Y.size() && fuse; i++)
    fuse = R[i].src == Y[i].dst.p && R[i].srs == Y[i].dst.rs && R[i].sw == Y[i].width && R[i].sh == Y[i].height && R[i].psize == 4 &&
           yuv_planar_same_shape(Y[0], Y[i]) && R[i].drs == R[0].drs && R[i].dw == R[0].dw && R[i].dh == R[0].dh && R[i].kx == R[0].kx &&
           R[i].ky == R[0].ky;
  if (fuse) {
    DevFilterEntry *fx = get_filter(e, R[0].sw, R[0].dw, 14, R[0].kx), *fy = get_filter(e, R[0].sh, R[0].dh, 12, R[0].ky);
    fuse = fx && fy;
    for (size_t i = 0; i < Y.size() && fuse; i++) fuse = cvt_resize_supported(Y[i], R[0].dw, R[0].dh, R[0].drs, R[i].dst, fx->host, fy->host);
    const void *px = fuse ? get_pack4(e, fx, 14) : nullptr, *py = fuse ? get_pack4(e, fy, 12) : nullptr;
    fuse = fuse && px && py;
    if (fuse) {
      std::vector<uint8_t *> dsts;
      for (size_t i = 0; i < R.size(); i++) dsts.push_back(R[i].dst);
      cudaError_t ce = launch_cvt_resize(e->L(), Y.data(), dsts.data(), (int)Y.size(), R[0].dw, R[0].dh, R[0].drs, px, py, fx->host, fy->host);
      if (ce == cudaSuccess) {
        Y.clear(); R.clear();
        e->rsz_defer = false;
        return PE_OK;
      }
      if (ce != cudaErrorInvalidConfiguration) {
        Y.clear(); R.clear();
        e->rsz_defer = false;
        return set_err(PE_ERR_CUDA, "conversion + resize launch failed: %s", cudaGetErrorString(ce));
      }
    }
  }
  int rc = flush_yuv_pending(e);
  if (rc == PE_OK) rc = flush_rsz_pending(e);
  else { R.clear(); e->rsz_defer = false; }
  return rc;
}

// Fan a batch of independent per-layer calls out over four side streams.  Layer 0 runs on the engine stream first (it
// creates whatever cached tables the batch needs: filter banks, LUTs -- their uploads are ordered before the fork); the other
// layers of the same geometry run on the side streams, so that the small kernels of different layers overlap instead of
// queueing behind each other's tails; blocks freed meanwhile are parked until the streams have joined.
struct FanOut {
  pe_engine *e;
  cudaStream_t main;
  bool active = false;
  explicit FanOut(pe_engine *e_) : e(e_), main(e_->stream) {}
  bool begin() {
    if (!e->fan_fork) {
      if (cudaEventCreateWithFlags(&e->fan_fork, cudaEventDisableTiming) != cudaSuccess) return false;
      for (int k = 0; k < 4; k++)
        if (cudaStreamCreateWithFlags(&e->fan_stream[k], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&e->fan_join[k], cudaEventDisableTiming) != cudaSuccess)
          return false;
    }
    if (cudaEventRecord(e->fan_fork, main) != cudaSuccess) return false;
    for (int k = 0; k < 4; k++)
      if (cudaStreamWaitEvent(e->fan_stream[k], e->fan_fork, 0) != cudaSuccess) return false;
    e->pool.defer(true);
    active = true;
    return true;
  }
  void use(int i) { e->stream = active ? e->fan_stream[i & 3] : main; }
  void use_main() { e->stream = main; }
  ~FanOut() {
    e->stream = main;
    if (!active) return;
    for (int k = 0; k < 4; k++) {
      cudaEventRecord(e->fan_join[k], e->fan_stream[k]);
      cudaStreamWaitEvent(main, e->fan_join[k], 0);
    }
    e->pool.defer(false);
    e->pool.flush_deferred();
  }
};

// "Left untouched on failure" (colourspace.c:13906-13927) for the batch calls whose kernels leave after the layers have been given
// their new pixel blocks: the old blocks are parked in the pool until the flush (DevPool::defer), so a failed flush can hand them back.
// (An in-place conversion has no old block: its descriptor is restored, its bytes are whatever the failed launch left.)
struct BatchSnapshot {
  std::vector<pe_frame> before;
  BatchSnapshot(int n, pe_frame_t *const *layers) : before((size_t)n) {
    for (int i = 0; i < n; i++)
      if (layers[i]) before[(size_t)i] = *layers[i];
  }
  void restore(pe_engine *e, int n, pe_frame_t *const *layers) {
    for (int i = 0; i < n; i++) {
      pe_frame *f = layers[i];
      if (!f) continue;
      const pe_frame &b = before[(size_t)i];
      if (f->base == b.base) { *f = b; continue; }      // untouched or in place
      if (b.base && !e->pool.undefer(b.base)) continue;  // the old block is gone: keep the new state
      pe_frame cur = *f;
      *f = b;
      frame_release_pixels(&cur);
    }
  }
};

inline bool same_geometry(const pe_frame *a, const pe_frame *b) {
  return a->d.palette == b->d.palette && a->d.width == b->d.width && a->d.height == b->d.height &&
         a->d.yuv_clamping == b->d.yuv_clamping && a->d.yuv_subspace == b->d.yuv_subspace && a->d.gamma_type == b->d.gamma_type;
}

}  // namespace

// A batch of independent layers through resize_layer_full (the render-to-disk loop of src/events.c:4239-4253 issues them one
// by one): one lock, no host work between the launches.  Returns the number of layers that were resized (TRUE results).
extern "C" int pe_resize_layer_batch(pe_engine_t *e, int n, pe_frame_t *const *layers, int width, int height, int interp,
                                     int opal_hint, int oclamp_hint) {
  if (!e || !layers || n <= 0) { set_err(PE_ERR_ARG, "NULL / empty argument"); return 0; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (cudaSetDevice(e->device) != cudaSuccess) { set_err(PE_ERR_CUDA, "cudaSetDevice failed"); return 0; }
  int done = 0;
  // BASELINE config 2 as a batch (and as a single layer): same-shaped 4:2:0 layers with a 4-byte RGB target whose resize takes the
  // <= 4-tap kernels.  Conversion and resize of every layer are queued by the ordinary resize_locked (all metadata rules stay its own)
  // and leave together as one fused launch per 32 frames -- the converted frame never exists in HBM.
  if ((opal_hint == PE_PALETTE_RGBA32 || opal_hint == PE_PALETTE_BGRA32) && width > 0 && height > 0 && layers[0] && layers[0]->d.planes[0]) {
    bool ok = true;
    for (int i = 0; i < n && ok; i++)
      ok = layers[i] && layers[i]->d.planes[0] && (layers[i]->d.palette == PE_PALETTE_YUV420P || layers[i]->d.palette == PE_PALETTE_YVU420P) &&
           same_geometry(layers[0], layers[i]);
    if (ok) {  // the resize must be one flush_rsz_pending would batch (else it would launch ahead of the queued conversions)
      const int iw = (layers[0]->d.width >> 1) << 1, ih = (layers[0]->d.height >> 1) << 1, w2 = width < 4 ? 4 : width;
      int h2 = height < 4 ? 4 : height;
      if (iw != w2 || ih != h2) h2 = (h2 >> 1) << 1;
      ok = !(iw == w2 && ih == h2);
      if (ok) {
        const AxisKinds kinds = filter_kinds(e, interp, iw, ih, w2, h2);
        DevFilterEntry *fx = get_filter(e, layers[0]->d.width, w2, 14, kinds.x), *fy = get_filter(e, layers[0]->d.height, h2, 12, kinds.y);
        ok = fx && fy && fx->host.fast_taps() <= 4 && fy->host.fast_taps() <= 4;
      }
    }
    if (ok) {
      BatchSnapshot snap(n, layers);
      e->yuv_defer = true;
      e->rsz_defer = true;
      e->pool.defer(true);
      for (int i = 0; i < n; i++)
        if (resize_locked(e, layers[i], width, height, interp, opal_hint, oclamp_hint, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YCBCR,
                          PE_GAMMA_UNKNOWN) == PE_TRUE)
          done++;
      e->yuv_defer = false;
      const int frc = flush_cvt_rsz_pending(e);
      e->rsz_defer = false;
      if (frc != PE_OK) snap.restore(e, n, layers);
      e->pool.defer(false);
      e->pool.flush_deferred();
      return frc == PE_OK ? done : 0;
    }
  }
  // phase 1: planar YUV layers with an RGB target are converted first (resize_layer_full converts before it scales, :14601);
  // the conversions of the whole batch are queued and leave as one launch per 32 same-shaped frames
  if (pal_is_rgb(opal_hint)) {
    BatchSnapshot snap(n, layers);
    e->yuv_defer = true;
    e->pool.defer(true);
    for (int i = 0; i < n; i++) {
      pe_frame *f = layers[i];
      if (!f || !f->d.planes[0] || width <= 0 || height <= 0) continue;
      const int pal = f->d.palette;
      if (pal != PE_PALETTE_YUV420P && pal != PE_PALETTE_YVU420P && pal != PE_PALETTE_YUV422P) continue;
      {  // resize_locked returns before converting when the (evened) sizes already match (:14850-14859)
        const int iw = (f->d.width >> 1) << 1, ih = (f->d.height >> 1) << 1, w2 = width < 4 ? 4 : width;
        int h2 = height < 4 ? 4 : height;
        if (iw != w2 || ih != h2) h2 = (h2 >> 1) << 1;
        if (iw == w2 && ih == h2) continue;
      }
      convert_locked(e, f, opal_hint, oclamp_hint, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YCBCR, PE_GAMMA_UNKNOWN);
    }
    e->yuv_defer = false;
    const int frc = flush_yuv_pending(e);
    if (frc != PE_OK) snap.restore(e, n, layers);
    e->pool.defer(false);
    e->pool.flush_deferred();
    if (frc != PE_OK) return 0;
  }
  // phase 2: the resizes.  4-byte packed frames of one geometry are queued and leave as one launch per 32; anything else is
  // fanned out over the side streams
  {
    bool uniform = n > 1 && layers[0] != nullptr;
    for (int i = 0; i < n && uniform; i++)
      uniform = layers[i] && layers[i]->d.planes[0] && pal_psize(layers[i]->d.palette) == 4 && !pal_is_planar(layers[i]->d.palette) &&
                same_geometry(layers[0], layers[i]) && (opal_hint == PE_PALETTE_NONE || opal_hint == layers[i]->d.palette);
    if (uniform) {
      BatchSnapshot snap(n, layers);
      e->rsz_defer = true;
      e->pool.defer(true);
      for (int i = 0; i < n; i++)
        if (resize_locked(e, layers[i], width, height, interp, opal_hint, oclamp_hint, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YCBCR,
                          PE_GAMMA_UNKNOWN) == PE_TRUE)
          done++;
      const int frc = flush_rsz_pending(e);
      if (frc != PE_OK) snap.restore(e, n, layers);
      e->pool.defer(false);
      e->pool.flush_deferred();
      return frc == PE_OK ? done : 0;
    }
  }
  const pe_frame ref0 = layers[0] ? *layers[0] : pe_frame();
  FanOut fan(e);
  for (int i = 0; i < n; i++) {
    if (!layers[i]) continue;
    if (i == 1 && n > 2 && layers[0]) fan.begin();
    if (i >= 1 && same_geometry(&ref0, layers[i])) fan.use(i); else fan.use_main();
    if (resize_locked(e, layers[i], width, height, interp, opal_hint, oclamp_hint, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YCBCR,
                      PE_GAMMA_UNKNOWN) == PE_TRUE)
      done++;
  }
  return done;
}

// ... and through convert_layer_palette
extern "C" int pe_convert_layer_palette_batch(pe_engine_t *e, int n, pe_frame_t *const *layers, int outpl, int op_clamping) {
  if (!e || !layers || n <= 0) { set_err(PE_ERR_ARG, "NULL / empty argument"); return 0; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (cudaSetDevice(e->device) != cudaSuccess) { set_err(PE_ERR_CUDA, "cudaSetDevice failed"); return 0; }
  int done = 0;
  bool all_planar = pal_is_rgb(outpl);
  for (int i = 0; i < n && all_planar; i++)
    all_planar = layers[i] && (layers[i]->d.palette == PE_PALETTE_YUV420P || layers[i]->d.palette == PE_PALETTE_YVU420P ||
                               layers[i]->d.palette == PE_PALETTE_YUV422P);
  bool all_rgb = pal_is_rgb(outpl);
  for (int i = 0; i < n && all_rgb; i++) all_rgb = layers[i] && pal_is_rgb(layers[i]->d.palette);
  if (all_rgb) {
    // RGB <-> RGB: the permutations are queued and leave as one launch per 128 same-shaped frames
    BatchSnapshot snap(n, layers);
    e->rgb_defer = true;
    e->pool.defer(true);
    for (int i = 0; i < n; i++)
      if (convert_locked(e, layers[i], outpl, op_clamping, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YUV, PE_GAMMA_UNKNOWN) == PE_TRUE) done++;
    const int frc = flush_rgb_pending(e);
    if (frc != PE_OK) snap.restore(e, n, layers);
    e->pool.defer(false);
    e->pool.flush_deferred();
    return frc == PE_OK ? done : 0;
  }
  if (all_planar) {
    // planar YUV -> RGB: the conversions are queued and leave as one launch per 32 same-shaped frames
    BatchSnapshot snap(n, layers);
    e->yuv_defer = true;
    e->pool.defer(true);
    for (int i = 0; i < n; i++)
      if (convert_locked(e, layers[i], outpl, op_clamping, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YUV, PE_GAMMA_UNKNOWN) == PE_TRUE) done++;
    e->yuv_defer = false;
    const int frc = flush_yuv_pending(e);
    if (frc != PE_OK) snap.restore(e, n, layers);
    e->pool.defer(false);
    e->pool.flush_deferred();
    return frc == PE_OK ? done : 0;
  }
  const pe_frame ref0 = layers[0] ? *layers[0] : pe_frame();
  FanOut fan(e);
  for (int i = 0; i < n; i++) {
    if (!layers[i]) continue;
    if (i == 1 && n > 2 && layers[0]) fan.begin();
    if (i >= 1 && same_geometry(&ref0, layers[i])) fan.use(i); else fan.use_main();
    if (convert_locked(e, layers[i], outpl, op_clamping, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YUV, PE_GAMMA_UNKNOWN) == PE_TRUE)
      done++;
  }
  return done;
}

extern "C" int pe_resize_layer(pe_engine_t *e, pe_frame_t *layer, int width, int height, int interp, int opal_hint,
                               int oclamp_hint) {  // :15331
  // a 4:2:0 layer with a 4-byte RGB target: conversion + resize as one kernel (the batch call with one layer)
  if (e && layer && layer->d.planes[0] && (opal_hint == PE_PALETTE_RGBA32 || opal_hint == PE_PALETTE_BGRA32) &&
      (layer->d.palette == PE_PALETTE_YUV420P || layer->d.palette == PE_PALETTE_YVU420P)) {
    const pe_frame before = *layer;
    pe_frame_t *one[1] = {layer};
    if (pe_resize_layer_batch(e, 1, one, width, height, interp, opal_hint, oclamp_hint) == 1) return PE_TRUE;
    if (layer->base != before.base || layer->d.palette != before.d.palette) return PE_FALSE;  // changed and failed: nothing to retry
  }
  return pe_resize_layer_full(e, layer, width, height, interp, opal_hint, oclamp_hint, PE_YUV_SAMPLING_DEFAULT,
                              PE_YUV_SUBSPACE_YCBCR, PE_GAMMA_UNKNOWN);
}

extern "C" int pe_letterbox_layer(pe_engine_t *e, pe_frame_t *layer, int nwidth, int nheight, int width, int height,
                                  int interp, int tpal, int tclamp) {
  if (!e || !layer) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA_B(cudaSetDevice(e->device));
  return letterbox_locked(e, layer, nwidth, nheight, width, height, interp, tpal, tclamp);
}

// ---------------------------------------------------------------------------------------------------------
// effects (boundary B1 arithmetic)
// ---------------------------------------------------------------------------------------------------------

namespace {

int check_blend_frames(const pe_frame *in1, const pe_frame *in2, const pe_frame *out, bool rgb24_only) {
  if (!in1 || !in2 || !out || !in1->d.planes[0] || !in2->d.planes[0] || !out->d.planes[0])
    return set_err(PE_ERR_ARG, "NULL frame");
  const int pal = in1->d.palette;
  if (!pal_is_rgb(pal) || (rgb24_only && pal_psize(pal) != 3))
    return set_err(PE_ERR_PALETTE, "palette %d is not in this filter's palette list", pal);
  if (in2->d.palette != pal || out->d.palette != pal) return set_err(PE_ERR_PALETTE, "channel palettes differ");
  if (in2->d.width != in1->d.width || in2->d.height != in1->d.height || out->d.width != in1->d.width ||
      out->d.height != in1->d.height)
    return set_err(PE_ERR_SIZE, "channel sizes differ");
  return PE_OK;
}

inline BlendFrame blend_frame(const pe_frame *in1, const pe_frame *in2, const pe_frame *out) {
  BlendFrame b;
  b.s1 = (const uint8_t *)in1->d.planes[0]; b.s2 = (const uint8_t *)in2->d.planes[0]; b.d = (uint8_t *)out->d.planes[0];
  b.rs1 = in1->d.rowstrides[0]; b.rs2 = in2->d.rowstrides[0]; b.rsd = out->d.rowstrides[0];
  b.s2_bytes = (long long)in2->d.rowstrides[0] * (in2->d.height - 1) + (long long)in2->d.width * pal_psize(in2->d.palette);
  return b;
}

}  // namespace

extern "C" int pe_fx_simple_blend_batch(pe_engine_t *e, int type, int n, const pe_frame_t *const *in1,
                                        const pe_frame_t *const *in2, pe_frame_t *const *out, int blend_factor) {
  if (!e || n <= 0 || !in1 || !in2 || !out) return set_err(PE_ERR_ARG, "NULL / empty argument");
  if (type < 0 || type > 4) return set_err(PE_ERR_ARG, "simple_blend type %d out of range", type);
  // "averaged luma overlay" (type 4, simple_blend.c:153-169): its 3 x 3 average is guarded by `row > 0`, and `row` is only advanced
  // inside that guard -- it stays 0, every pixel falls through to `case 1`: the filter IS the luma overlay (checked against the
  // compiled plugin, tests/test_weed_plugin.py)
  if (type == 4) type = 1;
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  std::vector<BlendFrame> frames(n);
  for (int i = 0; i < n; i++) {
    int rc = check_blend_frames(in1[i], in2[i], out[i], false);
    if (rc != PE_OK) return rc;
    if (in1[i]->d.palette != in1[0]->d.palette || in1[i]->d.width != in1[0]->d.width || in1[i]->d.height != in1[0]->d.height)
      return set_err(PE_ERR_SIZE, "frames of a batch must share palette and size");
    frames[i] = blend_frame(in1[i], in2[i], out[i]);
  }
  const BlendFrame *dev = (const BlendFrame *)upload_args(e, frames.data(), sizeof(BlendFrame) * n);
  if (!dev) return PE_ERR_CUDA;
  PE_CUDA(launch_simple_blend(e->L(), type, dev, n, in1[0]->d.width, in1[0]->d.height, rgb_layout(in1[0]->d.palette), blend_factor,
                              e->luma_dev));
  return PE_OK;
}

extern "C" int pe_fx_simple_blend(pe_engine_t *e, int type, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out,
                                  int blend_factor) {
  const pe_frame_t *a[1] = {in1}, *b[1] = {in2};
  pe_frame_t *o[1] = {out};
  return pe_fx_simple_blend_batch(e, type, 1, a, b, o, blend_factor);
}

// convert_layer_palette(clip, outpl) followed by the 'chroma blend' of simple_blend.c with in1 = the converted clip and
// in2 = operand, written back to the clip -- the multitrack crossfade of BASELINE config 5 -- in ONE kernel: the converted
// frame never travels to HBM and back.  clip: YUV420P / YVU420P / YUV422P; outpl: RGB24 / BGR24 (= the operand's palette).
extern "C" int pe_fx_convert_crossfade(pe_engine_t *e, pe_frame_t *clip, const pe_frame_t *operand, int outpl, int op_clamping,
                                       int blend_factor) {
  if (!e || !clip || !operand || !clip->d.planes[0] || !operand->d.planes[0]) return set_err(PE_ERR_ARG, "NULL argument");
  const int ip = clip->d.palette;
  if (ip != PE_PALETTE_YUV420P && ip != PE_PALETTE_YVU420P && ip != PE_PALETTE_YUV422P)
    return set_err(PE_ERR_PALETTE, "convert_crossfade: the clip must be YUV420P / YVU420P / YUV422P");
  if ((outpl != PE_PALETTE_RGB24 && outpl != PE_PALETTE_BGR24) || operand->d.palette != outpl)
    return set_err(PE_ERR_PALETTE, "convert_crossfade: output and operand must both be RGB24 or both BGR24");
  if (operand->d.width != clip->d.width || operand->d.height != clip->d.height)
    return set_err(PE_ERR_SIZE, "convert_crossfade: clip and operand differ in size");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  e->fuse_blend2 = (const uint8_t *)operand->d.planes[0];
  e->fuse_blend2_rs = operand->d.rowstrides[0];
  e->fuse_blend_bf = blend_factor;
  const int ok = convert_locked(e, clip, outpl, op_clamping, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YUV, PE_GAMMA_UNKNOWN);
  e->fuse_blend2 = nullptr;
  return ok == PE_TRUE ? PE_OK : PE_ERR_PALETTE;
}

// n clips, clip i against operands[i] (the frames of ONE clip against the successive frames of the other track, or the clips of a
// stack against one shared operand): the conversions are queued and leave as one k_yuv_march launch per 32 same-shaped clips
extern "C" int pe_fx_convert_crossfade_batchv(pe_engine_t *e, int n, pe_frame_t *const *clips, const pe_frame_t *const *operands,
                                              int outpl, int op_clamping, int blend_factor) {
  if (!e || !clips || !operands || n < 0) { set_err(PE_ERR_ARG, "NULL argument"); return 0; }
  if (outpl != PE_PALETTE_RGB24 && outpl != PE_PALETTE_BGR24) {
    set_err(PE_ERR_PALETTE, "convert_crossfade: output and operand must both be RGB24 or both BGR24");
    return 0;
  }
  std::lock_guard<std::mutex> lk(e->mu);
  if (cudaSetDevice(e->device) != cudaSuccess) { set_err(PE_ERR_CUDA, "cudaSetDevice failed"); return 0; }
  e->fuse_blend_bf = blend_factor;
  BatchSnapshot snap(n, clips);
  e->yuv_defer = true;
  e->pool.defer(true);
  int done = 0;
  for (int i = 0; i < n; i++) {
    pe_frame_t *c = clips[i];
    const pe_frame_t *op = operands[i];
    if (!c || !c->d.planes[0] || !op || !op->d.planes[0]) continue;
    const int ip = c->d.palette;
    if (ip != PE_PALETTE_YUV420P && ip != PE_PALETTE_YVU420P && ip != PE_PALETTE_YUV422P) {
      set_err(PE_ERR_PALETTE, "convert_crossfade: the clip must be YUV420P / YVU420P / YUV422P");
      continue;
    }
    if (op->d.palette != outpl) { set_err(PE_ERR_PALETTE, "convert_crossfade: output and operand must both be RGB24 or both BGR24"); continue; }
    if (op->d.width != c->d.width || op->d.height != c->d.height) {
      set_err(PE_ERR_SIZE, "convert_crossfade: clip and operand differ in size");
      continue;
    }
    e->fuse_blend2 = (const uint8_t *)op->d.planes[0];
    e->fuse_blend2_rs = op->d.rowstrides[0];
    if (convert_locked(e, c, outpl, op_clamping, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YUV, PE_GAMMA_UNKNOWN) == PE_TRUE) done++;
  }
  e->yuv_defer = false;
  e->fuse_blend2 = nullptr;
  const int frc = flush_yuv_pending(e);
  if (frc != PE_OK) snap.restore(e, n, clips);
  e->pool.defer(false);
  e->pool.flush_deferred();
  return frc == PE_OK ? done : 0;
}

extern "C" int pe_fx_convert_crossfade_batch(pe_engine_t *e, int n, pe_frame_t *const *clips, const pe_frame_t *operand, int outpl,
                                             int op_clamping, int blend_factor) {
  if (!e || !clips || n < 0 || !operand || !operand->d.planes[0]) { set_err(PE_ERR_ARG, "NULL argument"); return 0; }
  std::vector<const pe_frame_t *> ops((size_t)n, operand);
  return pe_fx_convert_crossfade_batchv(e, n, clips, ops.data(), outpl, op_clamping, blend_factor);
}

extern "C" int pe_fx_multi_blend(pe_engine_t *e, int type, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out,
                                 int blend_factor) {
  if (!e) return set_err(PE_ERR_ARG, "NULL engine");
  if (type < 0 || type > 6) return set_err(PE_ERR_ARG, "multi_blend type %d out of range", type);
  int rc = check_blend_frames(in1, in2, out, true);
  if (rc != PE_OK) return rc;
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  PE_CUDA(launch_multi_blend(e->L(), type, blend_frame(in1, in2, out), in1->d.width, in1->d.height,
                             in1->d.palette == PE_PALETTE_BGR24, blend_factor, e->luma_dev));
  return PE_OK;
}

// slide_over.c:94,109,124,135 as the reference's build computes it: -ffast-math (lives-plugins/weed-plugins/Makefile.am:49) turns
// `/ 255.` into `* (1 / 255.)` and regroups `(float)dim * (transval / 255.)` as (dim * (1 / 255.)) * transval.  Host arithmetic
// (this file is compiled without fast-math: the grouping below is what runs); checked on the CPU against the compiled plugin.
extern "C" int pe_fx_slide_over_bound(int direction, int transval, int width, int height) {
  const double r255 = 1. / 255.;
  const double dim = (double)(float)(direction <= 2 ? width : height);
  if (direction == 1 || direction == 3) return (int)((1. - (double)transval * r255) * dim);
  if (direction == 2 || direction == 4) return (int)((dim * r255) * (double)transval);
  return 0;
}

extern "C" int pe_fx_slide_over(pe_engine_t *e, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out, int transval,
                                int direction, int mvlower, int mvupper) {
  if (!e) return set_err(PE_ERR_ARG, "NULL engine");
  if (!in1 || !in2 || !out || !in1->d.planes[0] || !in2->d.planes[0] || !out->d.planes[0]) return set_err(PE_ERR_ARG, "NULL frame");
  const int pal = in1->d.palette;
  if (pal_is_planar(pal)) return set_err(PE_ERR_PALETTE, "palette %d is not in this filter's palette list", pal);  // ALL_PACKED_PALETTES_PLUS :158
  if (in2->d.palette != pal || out->d.palette != pal) return set_err(PE_ERR_PALETTE, "channel palettes differ");
  if (in2->d.width != in1->d.width || in2->d.height != in1->d.height || out->d.width != in1->d.width || out->d.height != in1->d.height)
    return set_err(PE_ERR_SIZE, "channel sizes differ");
  if (out->d.planes[0] == in1->d.planes[0] || out->d.planes[0] == in2->d.planes[0])
    return set_err(PE_ERR_ARG, "slide over is not an in-place filter (out channel flags 0, slide_over.c:162)");
  if (direction < 1 || direction > 4) return set_err(PE_ERR_ARG, "direction %d: 1 .. 4 (plugin_direction, slide_over.c:38-52)", direction);
  if (transval < 0 || transval > 255) return set_err(PE_ERR_ARG, "transition value %d out of 0 .. 255", transval);
  const int ps = pal_psize(pal), width = in1->d.width / pal_ppmp(pal), height = in1->d.height;  // width in macropixels, as the plugin sees it
  const int bound = pe_fx_slide_over_bound(direction, transval, width, height);
  const bool swapped = direction == 2 || direction == 4, along_y = direction >= 3;
  const pe_frame_t *fa = swapped ? in2 : in1, *fb = swapped ? in1 : in2;
  const bool mv_a = swapped ? mvlower : mvupper, mv_b = swapped ? mvupper : mvlower;
  SlideArgs a;
  a.first = (const uint8_t *)fa->d.planes[0]; a.second = (const uint8_t *)fb->d.planes[0]; a.d = (uint8_t *)out->d.planes[0];
  a.rs_first = fa->d.rowstrides[0]; a.rs_second = fb->d.rowstrides[0]; a.rsd = out->d.rowstrides[0];
  a.row_bytes = width * ps; a.height = height; a.along_y = along_y;
  a.bound = along_y ? bound : bound * ps;
  // a moving clip is read shifted: its far edge sits on the line (:95-97, :110-112, :126-127, :137-138)
  a.off_first = !mv_a ? 0 : along_y ? (long long)a.rs_first * (height - bound) : (long long)(width - bound) * ps;
  a.off_second = !mv_b ? 0 : along_y ? -(long long)a.rs_second * bound : -(long long)bound * ps;
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  PE_CUDA(launch_slide_over(e->L(), a));
  return PE_OK;
}

// ---- softlight.c / layout_blends.c / multi_transitions.c ---------------------------------------------------

extern "C" int pe_fx_softlight(pe_engine_t *e, const pe_frame_t *in, pe_frame_t *out) {
  if (!e) return set_err(PE_ERR_ARG, "NULL engine");
  if (!in || !out || !in->d.planes[0] || !out->d.planes[0]) return set_err(PE_ERR_ARG, "NULL frame");
  const int pal = in->d.palette;
  if (pal != PE_PALETTE_YUV444P && pal != PE_PALETTE_YUVA4444P && pal != PE_PALETTE_YUV422P && pal != PE_PALETTE_YUV420P &&
      pal != PE_PALETTE_YVU420P)
    return set_err(PE_ERR_PALETTE, "palette %d is not in this filter's palette list (softlight.c:169)", pal);
  if (out->d.palette != pal) return set_err(PE_ERR_PALETTE, "channel palettes differ");
  if (out->d.width != in->d.width || out->d.height != in->d.height) return set_err(PE_ERR_SIZE, "channel sizes differ");
  if (out->d.planes[0] == in->d.planes[0]) return set_err(PE_ERR_ARG, "softlight is not an in-place filter (out channel flags 0, softlight.c:173)");
  const bool unclamped = in->d.yuv_clamping == PE_YUV_CLAMPING_UNCLAMPED;  // :100-106
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  PE_CUDA(launch_softlight(e->L(), CImg{(const uint8_t *)in->d.planes[0], in->d.rowstrides[0]}, Img{(uint8_t *)out->d.planes[0], out->d.rowstrides[0]},
                           in->d.width, in->d.height, unclamped ? 0 : 16, unclamped ? 255 : 235));
  // the other planes are copied (:144-154: chroma width / height by palette, alpha plane at full size)
  for (int p = 1; p < in->d.nplanes; p++) {
    const bool sub = p < 3 && pal != PE_PALETTE_YUV444P && pal != PE_PALETTE_YUVA4444P;
    const int w = sub ? in->d.width >> 1 : in->d.width;
    const int h = (p < 3 && (pal == PE_PALETTE_YUV420P || pal == PE_PALETTE_YVU420P)) ? in->d.height >> 1 : in->d.height;
    PE_CUDA(launch_copy2d(e->L(), (const uint8_t *)in->d.planes[p], in->d.rowstrides[p], (uint8_t *)out->d.planes[p], out->d.rowstrides[p], w, h, 0, 0));
  }
  return PE_OK;
}

namespace {
// the channel checks the selector filters share: packed palettes of 3 / 4 bytes per (macro)pixel, equal sizes, out may be in1 only
int check_select_frames(const pe_frame *in1, const pe_frame *in2, const pe_frame *out, bool rgb24_only, bool inplace_ok, SelectArgs *a) {
  if (!in1 || !in2 || !out || !in1->d.planes[0] || !in2->d.planes[0] || !out->d.planes[0]) return set_err(PE_ERR_ARG, "NULL frame");
  const int pal = in1->d.palette;
  if (pal_is_planar(pal) || !pal_known(pal) || (rgb24_only && pal != PE_PALETTE_RGB24 && pal != PE_PALETTE_BGR24))
    return set_err(PE_ERR_PALETTE, "palette %d is not in this filter's palette list", pal);
  if (in2->d.palette != pal || out->d.palette != pal) return set_err(PE_ERR_PALETTE, "channel palettes differ");
  if (in2->d.width != in1->d.width || in2->d.height != in1->d.height || out->d.width != in1->d.width || out->d.height != in1->d.height)
    return set_err(PE_ERR_SIZE, "channel sizes differ");
  if (out->d.planes[0] == in2->d.planes[0] || (!inplace_ok && out->d.planes[0] == in1->d.planes[0]))
    return set_err(PE_ERR_ARG, "the out channel may only share pixels with in channel 0 of a CAN_DO_INPLACE filter");
  memset(a, 0, sizeof(*a));
  a->s1 = (const uint8_t *)in1->d.planes[0]; a->s2 = (const uint8_t *)in2->d.planes[0]; a->d = (uint8_t *)out->d.planes[0];
  a->rs1 = in1->d.rowstrides[0]; a->rs2 = in2->d.rowstrides[0]; a->rsd = out->d.rowstrides[0];
  a->psize = pal_psize(pal);
  a->width = in1->d.width / pal_ppmp(pal);  // macropixels, as the plugins see the channel
  a->height = in1->d.height;
  return PE_OK;
}
}  // namespace

// layout_blends.c:43-99: the two double comparisons per column and per row, evaluated once per column / row on the host exactly as
// the reference writes them (bit 0 "outside", bit 1 "inside")
extern "C" void pe_fx_triple_split_classes(int width, int height, double xstart, int sym, double xend, int vert, double bw, uint8_t *colclass,
                                           uint8_t *rowclass) {
  const int wb = width * 3;
  int tbs = height, tbe = height, bbs = height, bbe = height;  // tbs = tbe = bbs = bbe = end (:72)
  if (sym) { xstart /= 2.; xend = 1. - xstart; }
  if (xstart > xend) { const double t = xend; xend = xstart; xstart = t; }
  if (vert) {  // :74-80
    tbs = (int)(height * (xstart - bw) + .5); tbe = (int)(height * (xstart + bw) + .5);
    bbs = (int)(height * (xend - bw) + .5); bbe = (int)(height * (xend + bw) + .5);
    xstart = xend = -bw;
  }
  const double lo_out = wb * (xstart - bw), hi_out = wb * (xend + bw), lo_in = wb * (xstart + bw), hi_in = wb * (xend - bw);
  for (int x = 0; x < width; x++) {
    const int j = 3 * x;
    colclass[x] = (uint8_t)(((j < lo_out || j >= hi_out) ? 1 : 0) | ((j > lo_in && j < hi_in) ? 2 : 0));
  }
  for (int y = 0; y < height; y++) rowclass[y] = (uint8_t)(((y <= tbs || y >= bbe) ? 1 : 0) | ((y > tbe && y < bbs) ? 2 : 0));
}

extern "C" int pe_fx_triple_split(pe_engine_t *e, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out, double start, int sym,
                                  double end, int vert, double borderw, const int bordercol[3]) {
  if (!e) return set_err(PE_ERR_ARG, "NULL engine");
  if (!bordercol) return set_err(PE_ERR_ARG, "NULL border colour");
  SelectArgs a;
  int rc = check_select_frames(in1, in2, out, true, true, &a);
  if (rc != PE_OK) return rc;
  std::vector<uint8_t> cls((size_t)a.width + (size_t)a.height);
  pe_fx_triple_split_classes(a.width, a.height, start, sym, end, vert, borderw, cls.data(), cls.data() + a.width);
  const bool bgr = in1->d.palette == PE_PALETTE_BGR24;  // :66-70
  a.colour[0] = bordercol[bgr ? 2 : 0]; a.colour[1] = bordercol[1]; a.colour[2] = bordercol[bgr ? 0 : 2];
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  const uint8_t *dev = (const uint8_t *)upload_args(e, cls.data(), cls.size());
  if (!dev) return PE_ERR_CUDA;
  a.colclass = dev; a.rowclass = dev + a.width;
  PE_CUDA(launch_select(e->L(), 0, a));
  return PE_OK;
}

struct pe_dissolve_mask {
  pe_engine *e;
  float *dev;
  int width, height;
};

// multi_transitions.c dissolve_init :42-70: width x height floats from the xorshift64 stream of the host's random seed
// (libweed/weed-plugin-utils.c:666,686-704; `val / divd / divd * 1.` is ONE multiplication in the plugins' -ffast-math build)
extern "C" int pe_fx_dissolve_mask_create(pe_engine_t *e, int width, int height, int64_t random_seed, pe_dissolve_mask_t **out) {
  if (!e || !out || width <= 0 || height <= 0) return set_err(PE_ERR_ARG, "NULL / empty argument");
  const size_t n = (size_t)width * (size_t)height;
  std::vector<float> host(n);
  uint64_t x = (uint64_t)random_seed;
  union { uint64_t u; double d; } k;
  k.u = 0x3BF0000000200000ull;  // 1 / (double)0xFFFFFFFF squared, as the compiled plugin holds it
  for (size_t i = 0; i < n; i++) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    host[i] = (float)((double)x * k.d);
  }
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  float *dev = nullptr;
  PE_CUDA(cudaMalloc(&dev, n * sizeof(float)));
  cudaError_t ce = cudaMemcpyAsync(dev, host.data(), n * sizeof(float), cudaMemcpyHostToDevice, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) { cudaFree(dev); return set_err(PE_ERR_CUDA, "mask upload failed: %s", cudaGetErrorString(ce)); }
  *out = new pe_dissolve_mask{e, dev, width, height};
  return PE_OK;
}

extern "C" void pe_fx_dissolve_mask_destroy(pe_dissolve_mask_t *m) {
  if (!m) return;
  {
    std::lock_guard<std::mutex> lk(m->e->mu);
    cudaSetDevice(m->e->device);
    cudaStreamSynchronize(m->e->stream);
    cudaFree(m->dev);
  }
  delete m;
}

extern "C" int pe_fx_multi_transition(pe_engine_t *e, int type, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out, double amount,
                                      const pe_dissolve_mask_t *mask) {
  if (!e) return set_err(PE_ERR_ARG, "NULL engine");
  if (type < 0 || type > 3) return set_err(PE_ERR_ARG, "transition type %d: 0 iris rectangle, 1 iris circle, 2 4 way split, 3 dissolve", type);
  SelectArgs a;
  int rc = check_select_frames(in1, in2, out, false, type != 2, &a);
  if (rc != PE_OK) return rc;
  if (type == 3 && (!mask || mask->width != a.width || mask->height != a.height))
    return set_err(PE_ERR_ARG, "dissolve needs the mask of this channel size (REINIT_ON_SIZE_CHANGE, multi_transitions.c:281)");
  // the per-frame constants of :118-143, in float as the plugin computes them (no -ffast-math here: every step is written out in the
  // form the reference's build evaluates)
  const int psize = a.psize, wb = a.width * psize;
  const float bf = (float)amount, bfneg = 1.f - bf;
  const float hheight = (float)a.height * 0.5f, hwidth_px = (float)a.width * 0.5f, hwidth = (float)wb * 0.5f;
  a.bf = bf;
  a.ihwidth = wb >> 1; a.ihheight = a.height >> 1;
  a.hheight = hheight; a.hwidth = hwidth;
  a.inv_psize = 1.f / (float)psize;
  a.inv_maxradsq = 1.f / (hheight * hheight + hwidth_px * hwidth_px);
  a.inv_hh = 1.f / hheight; a.inv_hw = 1.f / hwidth;
  if (type == 0) {
    a.xx = (int)((double)((float)(int)hwidth * bfneg) + .5);
    a.yy = (int)((double)((float)(int)hheight * bfneg) + .5);
  } else if (type == 2) {
    a.xx = (int)((double)(hheight * bf) + .5);  // rows
    a.yy = (int)((double)((bf * (psize == 3 ? 0.333333343267440796f : 0.25f)) * hwidth) + .5) * psize;  // bytes
  }
  a.mask = mask ? mask->dev : nullptr;
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  PE_CUDA(launch_select(e->L(), type + 1, a));
  return PE_OK;
}

namespace {

// compositor_process (gdk/compositor.c:127) at scale 1 / offset 0, optionally followed by gamma_convert_layer(gamma_to, out)
// folded into the last paint.  Work avoided without changing a byte of the result:
//   * the background fill is skipped when an opaque (alpha == 1) layer covers it -- the fill is dead;
//   * painting that opaque layer is not a pass of its own: the next paint reads it as its background;
//   * alpha = k / 256 uses the integer blend (exact, see k_alpha_over_arith), other alphas the 64 KB [bg][fg] table.
// launch the integer alpha-over paints a compositor batch has queued.  Jobs are taken in order; a run of jobs with the same shape
// and parameters leaves as one launch as long as none of them reads what an earlier job of the run writes (the second paint of
// a frame reads the first one's output: it starts a new run)
int flush_over_pending(pe_engine *e) {
  std::vector<pe_engine::OverJob> q;
  q.swap(e->over_pending);
  e->over_defer = false;
  size_t i = 0;
  while (i < q.size()) {
    size_t j = i + 1;
    auto same = [&](const pe_engine::OverJob &a, const pe_engine::OverJob &b) {
      return a.rs_bg == b.rs_bg && a.rs_fg == b.rs_fg && a.rs_d == b.rs_d && a.w == b.w && a.h == b.h && a.psize == b.psize && a.k256 == b.k256 &&
             a.lut == b.lut;
    };
    auto depends = [&](const pe_engine::OverJob &b) {
      for (size_t k = i; k < j; k++)
        if (q[k].dst == b.bg || q[k].dst == b.fg || q[k].dst == b.dst ||  // reads / rewrites what an earlier job of the run writes
            b.dst == q[k].bg || (b.dst == q[k].fg))                        // ... or writes what an earlier job of the run still reads
          return true;
      return false;
    };
    while (j < q.size() && same(q[i], q[j]) && !depends(q[j])) j++;
    std::vector<const uint8_t *> bgs, fgs;
    std::vector<uint8_t *> dsts;
    for (size_t k = i; k < j; k++) { bgs.push_back(q[k].bg); fgs.push_back(q[k].fg); dsts.push_back(q[k].dst); }
    cudaError_t ce = launch_alpha_over_arith_batch(e->L(), bgs.data(), fgs.data(), dsts.data(), (int)(j - i), q[i].rs_bg, q[i].rs_fg, q[i].rs_d,
                                                   q[i].w, q[i].h, q[i].psize, q[i].k256, q[i].lut, 1);
    if (ce != cudaSuccess) return set_err(PE_ERR_CUDA, "alpha-over launch failed: %s", cudaGetErrorString(ce));
    i = j;
  }
  return PE_OK;
}

int compositor_locked(pe_engine_t *e, pe_frame_t *out, const pe_frame_t *const *layers, const double *alpha, int nlayers,
                      const int bgcol[3], int gamma_to) {
  const int pal = out->d.palette;
  if (!pal_is_rgb(pal) || pal == PE_PALETTE_ARGB32) return set_err(PE_ERR_PALETTE, "compositor palettes: RGB24 BGR24 RGBA32 BGRA32");
  const int psize = pal_psize(pal), w = out->d.width, h = out->d.height;
  const Img dst{(uint8_t *)out->d.planes[0], out->d.rowstrides[0]};
  // the gamma step that follows (gamma_convert_sub_layer :14069): which LUT, if any
  const uint8_t *lut = nullptr;
  bool set_gamma = false;
  if (gamma_to != PE_GAMMA_UNKNOWN && e->cfg.apply_gamma && !(gamma_to == out->d.gamma_type)) {
    Lut8Entry *le = get_lut8(e, 1.0, out->d.gamma_type, gamma_to);
    lut = le ? le->dev : nullptr;
    set_gamma = le != nullptr;
  }
  // layers painted last first (revz == WEED_FALSE, :189-197); a layer with alpha 0 is skipped (:232)
  std::vector<int> order;
  for (int z = nlayers - 1; z >= 0; z--) {
    const pe_frame *l = layers[z];
    if (!l || !l->d.planes[0]) continue;
    if (l->d.palette != pal) return set_err(PE_ERR_PALETTE, "layer %d palette differs from the output's", z);
    if (l->d.width != w || l->d.height != h)
      return set_err(PE_ERR_SIZE, "layer %d: only scale 1 / offset 0 layers are handled by this build", z);
    if (alpha[z] <= 0.) continue;
    order.push_back(z);
  }
  // everything below the last fully opaque layer is dead
  size_t first = 0;
  for (size_t i = 0; i < order.size(); i++)
    if (alpha[order[i]] >= 1.0) first = i;
  const bool opaque_base = !order.empty() && alpha[order[first]] >= 1.0;
  CImg cur{dst.p, dst.rs};  // what the next paint reads as background
  if (opaque_base) {
    const pe_frame *l = layers[order[first]];
    cur = CImg{(const uint8_t *)l->d.planes[0], l->d.rowstrides[0]};
    first++;
  } else {
    // background fill (compositor.c:172-186); alpha byte 0xFF
    const int r = bgcol ? bgcol[0] : 0, g = bgcol ? bgcol[1] : 0, b = bgcol ? bgcol[2] : 0;
    const bool swap = (pal == PE_PALETTE_BGR24 || pal == PE_PALETTE_BGRA32);
    const uint32_t px = (uint32_t)((swap ? b : r) & 255) | ((uint32_t)(g & 255) << 8) | ((uint32_t)((swap ? r : b) & 255) << 16) | 0xFF000000u;
    PE_CUDA(launch_fill(e->L(), dst, w, h, psize, px));
    first = 0;
  }
  bool lut_done = lut == nullptr;
  auto paint = [&](CImg bg, CImg fg, double a, const uint8_t *l8) -> int {
    const double k256 = a * 256.;
    if (k256 >= 0. && k256 <= 256. && k256 == (double)(int)k256) {
      if (e->over_defer)  // batch call: launched by flush_over_pending
        e->over_pending.push_back(pe_engine::OverJob{bg.p, fg.p, dst.p, bg.rs, fg.rs, dst.rs, w, h, psize, (int)k256, l8});
      else
        PE_CUDA(launch_alpha_over_arith(e->L(), bg, fg, dst, w, h, psize, (int)k256, l8, 1));
    } else {
      uint8_t *tab = get_over_table(e, a, l8);
      if (!tab) return set_err(PE_ERR_MEMORY, "alpha-over table could not be built");
      PE_CUDA(launch_alpha_over(e->L(), bg, fg, dst, w, h, psize, tab, 1));
    }
    return PE_OK;
  };
  for (size_t i = first; i < order.size(); i++) {
    const pe_frame *l = layers[order[i]];
    const bool last = i + 1 == order.size();
    int rc = paint(cur, CImg{(const uint8_t *)l->d.planes[0], l->d.rowstrides[0]}, alpha[order[i]], last ? lut : nullptr);
    if (rc != PE_OK) return rc;
    if (last) lut_done = true;
    cur = CImg{dst.p, dst.rs};
  }
  if (cur.p != dst.p) {
    // only the opaque base was painted: out = that layer with the alpha byte forced to 0xFF (+ the LUT)
    int rc = paint(cur, cur, 1.0, lut);
    if (rc != PE_OK) return rc;
    lut_done = true;
  }
  if (!lut_done)  // nothing was painted after the fill: plain gamma pass over the background colour
    PE_CUDA(launch_lut8_rect(e->L(), dst, rgb_layout(pal), 0, 0, w, h, lut));
  if (set_gamma && gamma_to != PE_GAMMA_VARIANT) out->d.gamma_type = gamma_to;
  return PE_OK;
}

}  // namespace

extern "C" int pe_fx_compositor(pe_engine_t *e, pe_frame_t *out, const pe_frame_t *const *layers, const double *alpha,
                                int nlayers, const int bgcol[3]) {
  if (!e || !out || !out->d.planes[0] || nlayers < 0 || (nlayers > 0 && (!layers || !alpha)))
    return set_err(PE_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  return compositor_locked(e, out, layers, alpha, nlayers, bgcol, PE_GAMMA_UNKNOWN);
}

// n independent output frames through compositor_process (+ gamma), each with its own `nlayers` layers (layers[i * nlayers + z])
// and the same per-layer alphas: the render-to-disk loop issues them one by one.  The integer alpha-over paints (alpha = k / 256)
// of the batch are queued and leave as one launch per 32 same-shaped frames and paint pass; fills and table paints are ordered
// before them only when no integer paint is pending, so a batch that mixes the two falls back to frame-by-frame order.
// Returns the number of frames composited.
extern "C" int pe_fx_compositor_gamma_batch(pe_engine_t *e, int n, pe_frame_t *const *outs, const pe_frame_t *const *layers,
                                            const double *alpha, int nlayers, const int bgcol[3], int gamma_to) {
  if (!e || n <= 0 || !outs || nlayers < 0 || (nlayers > 0 && (!layers || !alpha))) { set_err(PE_ERR_ARG, "NULL / empty argument"); return 0; }
  std::lock_guard<std::mutex> lk(e->mu);
  if (cudaSetDevice(e->device) != cudaSuccess) { set_err(PE_ERR_CUDA, "cudaSetDevice failed"); return 0; }
  // deferral keeps stream order only if every launch of the batch is a deferred one: all alphas k / 256 and an opaque base layer
  // (no background fill, no table paint, no trailing LUT pass)
  bool deferrable = nlayers >= 1;
  bool has_opaque = false;
  for (int z = 0; z < nlayers; z++) {
    const double k256 = alpha[z] * 256.;
    if (alpha[z] > 0. && !(k256 <= 256. && k256 == (double)(int)k256)) deferrable = false;
    if (alpha[z] >= 1.0) has_opaque = true;
  }
  deferrable = deferrable && has_opaque;
  if (deferrable)
    for (int i = 0; i < n && deferrable; i++)
      for (int z = 0; z < nlayers; z++)
        if (!layers[(size_t)i * nlayers + z] || !layers[(size_t)i * nlayers + z]->d.planes[0]) deferrable = false;
  int done = 0;
  e->over_defer = deferrable;
  for (int i = 0; i < n; i++) {
    if (!outs[i] || !outs[i]->d.planes[0]) continue;
    if (compositor_locked(e, outs[i], layers + (size_t)i * nlayers, alpha, nlayers, bgcol, gamma_to) == PE_OK) done++;
  }
  if (deferrable && flush_over_pending(e) != PE_OK) return 0;
  e->over_defer = false;
  return done;
}

extern "C" int pe_fx_compositor_gamma(pe_engine_t *e, pe_frame_t *out, const pe_frame_t *const *layers, const double *alpha,
                                      int nlayers, const int bgcol[3], int gamma_to) {
  if (!e || !out || !out->d.planes[0] || nlayers < 0 || (nlayers > 0 && (!layers || !alpha)))
    return set_err(PE_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  return compositor_locked(e, out, layers, alpha, nlayers, bgcol, gamma_to);
}

// ---------------------------------------------------------------------------------------------------------
// fused chain
// ---------------------------------------------------------------------------------------------------------

static int fused_locked(pe_engine_t *e, int n, const pe_frame_t *const *fg, const pe_frame_t *const *bg, pe_frame_t *const *out,
                        int inner_w, int inner_h, double alpha, int gamma_from, int gamma_to) {
  const pe_frame *f0 = fg[0], *b0 = bg[0];
  if (!f0 || !b0) return set_err(PE_ERR_ARG, "NULL frame");
  const int ow = b0->d.width, oh = b0->d.height;
  if (inner_w <= 0 || inner_h <= 0 || inner_w > ow || inner_h > oh) return set_err(PE_ERR_SIZE, "inner rectangle does not fit");
  // gamma LUT folded into the [bg][fg] table
  const uint8_t *lut = nullptr;
  if (e->cfg.apply_gamma) {
    Lut8Entry *le = get_lut8(e, 1.0, gamma_from, gamma_to);
    lut = le ? le->dev : nullptr;
  }
  uint8_t *tab = get_over_table(e, alpha, lut);
  if (!tab) return set_err(PE_ERR_MEMORY, "alpha-over table could not be built");
  const AxisKinds kinds = filter_kinds(e, PE_INTERP_NORMAL, f0->d.width, f0->d.height, inner_w, inner_h);
  DevFilterEntry *fx = get_filter(e, f0->d.width, inner_w, 14, kinds.x), *fy = get_filter(e, f0->d.height, inner_h, 12, kinds.y);
  if (!fx || !fy) return set_err(PE_ERR_SIZE, "scale factor out of range");
  // source extent one output tile can touch
  const int tw = fused_tile_w(), th = fused_tile_h();
  auto span = [](const ResizeFilter &F, int dst_n, int src_n, int tile) {
    int worst = 1;
    for (int i0 = 0; i0 < dst_n; i0 += 1) {
      const int i1 = i0 + tile - 1 < dst_n - 1 ? i0 + tile - 1 : dst_n - 1;
      int lo = F.first[i0], hi = F.first[i1] + F.taps - 1;
      lo = lo < 0 ? 0 : lo; hi = hi > src_n - 1 ? src_n - 1 : hi;
      if (hi - lo + 1 > worst) worst = hi - lo + 1;
    }
    return worst;
  };
  const int max_cols = span(fx->host, inner_w, f0->d.width, tw), max_rows = span(fy->host, inner_h, f0->d.height, th);
  std::vector<FusedArgs> args(n);
  for (int i = 0; i < n; i++) {
    const pe_frame *f = fg[i], *b = bg[i];
    pe_frame *o = out[i];
    if (!f || !b || !o || !f->d.planes[0] || !b->d.planes[0] || !o->d.planes[0]) return set_err(PE_ERR_ARG, "NULL frame");
    if (f->d.palette != PE_PALETTE_YUV420P && f->d.palette != PE_PALETTE_YVU420P && f->d.palette != PE_PALETTE_YUV422P)
      return set_err(PE_ERR_PALETTE, "fused chain: fg must be YUV420P / YVU420P / YUV422P");
    if (b->d.palette != PE_PALETTE_RGBA32 || o->d.palette != PE_PALETTE_RGBA32)
      return set_err(PE_ERR_PALETTE, "fused chain: bg and out must be RGBA32");
    if (b->d.width != ow || b->d.height != oh || o->d.width != ow || o->d.height != oh || f->d.width != f0->d.width ||
        f->d.height != f0->d.height)
      return set_err(PE_ERR_SIZE, "frames of a batch must share sizes");
    FusedArgs &A = args[i];
    A.fg = planes_of(f, f->d.palette == PE_PALETTE_YVU420P);
    A.fw = f->d.width; A.fh = f->d.height;
    A.is_422 = f->d.palette == PE_PALETTE_YUV422P;
    A.clamped = f->d.yuv_clamping == PE_YUV_CLAMPING_CLAMPED;
    A.low_quality = e->cfg.pb_quality == PE_QUALITY_LOW;
    A.quirks = e->cfg.ref_quirks;
    A.conv = dev_conv(e, f->d.yuv_clamping, f->d.yuv_subspace);
    A.bg = CImg{(const uint8_t *)b->d.planes[0], b->d.rowstrides[0]};
    A.out = Img{(uint8_t *)o->d.planes[0], o->d.rowstrides[0]};
    A.ow = ow; A.oh = oh; A.iw = inner_w; A.ih = inner_h;
    A.ox = (ow - inner_w + 1) >> 1; A.oy = (oh - inner_h + 1) >> 1;
    A.fx = fx->dev; A.fy = fy->dev;
    A.over_table = tab;
  }
  // fast path: no horizontal scaling, <= 4 vertical taps, aligned planes (pe_kernels_fused2.cu)
  bool fast = getenv("PE_FUSED_GENERIC") == nullptr;
  for (int i = 0; i < n && fast; i++) fast = fused2_supported(args[i], fy->host.fast_taps(), 0) && args[i].is_422 == args[0].is_422;
  int tile_h = 0;
  if (fast) {
    // tallest tile whose virtual source rows (first .. first + 3 of its last row) fit the shared-memory tile
    const ResizeFilter &F = fy->host;
    for (tile_h = fused2_max_tile_h(); tile_h >= 4; tile_h -= 4) {
      int worst = 0;
      for (int i0 = 0; i0 < inner_h; i0++) {
        const int i1 = i0 + tile_h - 1 < inner_h - 1 ? i0 + tile_h - 1 : inner_h - 1;
        const int spanr = F.first[i1] + 3 - F.first[i0] + 1;
        if (spanr > worst) worst = spanr;
      }
      if (worst <= fused2_max_virtual_rows(args[0].is_422)) break;
    }
    if (tile_h < 4) fast = false;
  }
  const double k256 = alpha * 256.;
  const bool dyadic = k256 >= 0. && k256 <= 256. && k256 == (double)(int)k256;
  // register-resident path (pe_kernels_fused3.cu): 4:2:0, full-width letterbox, alpha = k / 256, one conversion variant
  const bool regs = dyadic && getenv("PE_FUSED_GENERIC") == nullptr && getenv("PE_FUSED3_OFF") == nullptr &&
                    fused3_supported(args.data(), n, fy->host.fast_taps()) &&
                    fused3_tables_ok(conv_host(e, f0->d.yuv_clamping, f0->d.yuv_subspace));
  if (regs) {
    if (!fy->rows4) {
      const ResizeFilter &F = fy->host;
      std::vector<int32_t> r4((size_t)inner_h * 4);
      // scaled by 16 (sum 65536) when every coefficient still fits 16 bits, i.e. no row has the single tap 4096: the kernel
      // then reads the filtered value as byte 2 of its accumulator instead of shifting by 12 (PE_F3_NO_C16 turns it off)
      int cmax = 0;
      for (size_t i = 0; i < (size_t)inner_h * F.taps; i++) cmax = F.coef[i] > cmax ? F.coef[i] : cmax;
      fy->rows4_x16 = cmax < 4096 && getenv("PE_F3_NO_C16") == nullptr;
      const uint32_t scale = fy->rows4_x16 ? 16u : 1u;
      for (int i = 0; i < inner_h; i++) {
        uint32_t c[4] = {0, 0, 0, 0};
        for (int t = 0; t < F.taps; t++) c[t] = (uint32_t)(uint16_t)F.coef[(size_t)i * F.taps + t] * scale;
        r4[4 * i] = F.first[i];
        r4[4 * i + 1] = (int32_t)(c[3] | (c[2] << 16));
        r4[4 * i + 2] = (int32_t)(c[1] | (c[0] << 16));
        r4[4 * i + 3] = 0;
      }
      PE_CUDA(cudaMalloc(&fy->rows4, r4.size() * sizeof(int32_t)));
      PE_CUDA(cudaMemcpyAsync(fy->rows4, r4.data(), r4.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
      PE_CUDA(cudaStreamSynchronize(e->stream));  // r4 is a local
    }
    if (!e->f3_sched) {
      PE_CUDA(cudaMalloc(&e->f3_sched, 2 * sizeof(unsigned int)));
      PE_CUDA(cudaMemsetAsync(e->f3_sched, 0, 2 * sizeof(unsigned int), e->stream));
    }
    PE_CUDA(launch_fused3(e->L(), args.data(), n, (int)k256, lut, fy->rows4, fy->rows4_x16, e->f3_sched));
  } else if (fast) {
    PE_CUDA(launch_fused2(e->L(), args.data(), n, ow, oh, tile_h, dyadic ? (int)k256 : -1, lut));
  } else {
    const FusedArgs *dev = (const FusedArgs *)upload_args(e, args.data(), sizeof(FusedArgs) * n);
    if (!dev) return PE_ERR_CUDA;
    PE_CUDA(launch_fused_dev(e->L(), dev, n, ow, oh, max_rows, max_cols));
  }
  for (int i = 0; i < n; i++) {
    out[i]->d.gamma_type = (lut ? gamma_to : gamma_from);
    out[i]->d.flags = bg[i]->d.flags;
  }
  return PE_OK;
}

extern "C" int pe_fused_convert_letterbox_over_gamma_batch(pe_engine_t *e, int n, const pe_frame_t *const *fg,
                                                           const pe_frame_t *const *bg, pe_frame_t *const *out, int inner_w,
                                                           int inner_h, double alpha, int gamma_from, int gamma_to) {
  if (!e || n <= 0 || !fg || !bg || !out) return set_err(PE_ERR_ARG, "NULL / empty argument");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  return fused_locked(e, n, fg, bg, out, inner_w, inner_h, alpha, gamma_from, gamma_to);
}

extern "C" int pe_fused_convert_letterbox_over_gamma(pe_engine_t *e, const pe_frame_t *fg, const pe_frame_t *bg,
                                                     pe_frame_t *out, int inner_w, int inner_h, double alpha, int gamma_from,
                                                     int gamma_to) {
  const pe_frame_t *f[1] = {fg}, *b[1] = {bg};
  pe_frame_t *o[1] = {out};
  return pe_fused_convert_letterbox_over_gamma_batch(e, 1, f, b, o, inner_w, inner_h, alpha, gamma_from, gamma_to);
}

// ---------------------------------------------------------------------------------------------------------
// the CONVERT step of the node model as one descriptor (src/nodemodel.c:161 get_op_order -> :1065-1282)
// ---------------------------------------------------------------------------------------------------------

static thread_local int g_last_plan_path = -1;
extern "C" int pe_last_plan_path(void) { return g_last_plan_path; }

// the substeps in get_op_order's order, one op at a time: what run_plan's CONVERT step does with the reference's own functions
extern "C" int pe_run_convert_plan(pe_engine_t *e, pe_frame_t *layer, const pe_convert_plan_t *plan) {
  if (!e || !layer || !plan) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  int maxord = 0;
  for (int i = 0; i < PE_N_OP_TYPES; i++) maxord = plan->op_order[i] > maxord ? plan->op_order[i] : maxord;
  const int lbox = plan->op_order[PE_OP_LETTERBOX];
  for (int ord = 1; ord <= maxord; ord++) {
    const bool rsz = plan->op_order[PE_OP_RESIZE] == ord, pcv = plan->op_order[PE_OP_PCONV] == ord, gam = plan->op_order[PE_OP_GAMMA] == ord;
    if (lbox == ord) {
      // letterbox_layer resizes the image to the inner rectangle itself (:15388-15393); a plan that also names OP_RESIZE earlier has
      // already brought the layer to that size
      if (!pe_letterbox_layer(e, layer, plan->lb_width, plan->lb_height, plan->width, plan->height, plan->interp,
                              pcv ? plan->out_palette : PE_PALETTE_NONE, plan->out_clamping))
        return PE_FALSE;
      continue;
    }
    if (rsz) {
      // "resize can do resize OR resize + palconv OR resize + gamma OR resize + palconv + gamma" (:167-169)
      if (lbox == ord + 1 && !pcv && !gam) continue;  // the letterbox that follows does this resize
      if (!pe_resize_layer_full(e, layer, plan->width, plan->height, plan->interp, pcv ? plan->out_palette : PE_PALETTE_NONE,
                                plan->out_clamping, plan->out_sampling, plan->out_subspace, gam ? plan->out_gamma : PE_GAMMA_UNKNOWN))
        return PE_FALSE;
      if (pcv && layer->d.palette != plan->out_palette &&
          !pe_convert_layer_palette_full(e, layer, plan->out_palette, plan->out_clamping, plan->out_sampling, plan->out_subspace,
                                         gam ? plan->out_gamma : PE_GAMMA_UNKNOWN))
        return PE_FALSE;
      if (gam && layer->d.gamma_type != plan->out_gamma && !pe_gamma_convert_layer(e, plan->out_gamma, layer)) return PE_FALSE;
    } else if (pcv) {
      if (!pe_convert_layer_palette_full(e, layer, plan->out_palette, plan->out_clamping, plan->out_sampling, plan->out_subspace,
                                         gam ? plan->out_gamma : PE_GAMMA_UNKNOWN))
        return PE_FALSE;
      if (gam && layer->d.gamma_type != plan->out_gamma && !pe_gamma_convert_layer(e, plan->out_gamma, layer)) return PE_FALSE;
    } else if (gam) {
      if (!pe_gamma_convert_layer(e, plan->out_gamma, layer)) return PE_FALSE;
    }
  }
  return PE_TRUE;
}

// CONVERT (the plan) followed by the APPLY_INST + gamma substeps (:1119-1333) when the instance is the compositor painting the converted
// layer over ONE background layer: the whole chain is one fused kernel launch when the plan is "planar YUV -> RGBA32, letterboxed,
// bilinear" -- which is what the fused kernels cover -- and the reference's op-by-op sequence otherwise.  Same bytes either way.
// plan->no_fuse forces the op-by-op path (tests); pe_last_plan_path(): 1 fused, 0 op by op.
extern "C" int pe_run_convert_plan_over(pe_engine_t *e, const pe_frame_t *fg, const pe_convert_plan_t *plan, const pe_frame_t *bg,
                                        pe_frame_t *out, double alpha, int gamma_to) {
  if (!e || !fg || !plan || !bg || !out) return set_err(PE_ERR_ARG, "NULL argument");
  const int ip = fg->d.palette;
  const bool planar_src = ip == PE_PALETTE_YUV420P || ip == PE_PALETTE_YVU420P || ip == PE_PALETTE_YUV422P;
  const bool wants_gamma_before = plan->op_order[PE_OP_GAMMA] != 0;
  const bool fusable = !plan->no_fuse && planar_src && plan->op_order[PE_OP_PCONV] != 0 && plan->out_palette == PE_PALETTE_RGBA32 &&
                       plan->op_order[PE_OP_LETTERBOX] != 0 && !wants_gamma_before && plan->interp == PE_INTERP_NORMAL &&
                       bg->d.palette == PE_PALETTE_RGBA32 && out->d.palette == PE_PALETTE_RGBA32 && bg->d.width == plan->lb_width &&
                       bg->d.height == plan->lb_height && out->d.width == plan->lb_width && out->d.height == plan->lb_height &&
                       plan->width <= plan->lb_width && plan->height <= plan->lb_height;
  if (fusable) {
    g_last_plan_path = 1;
    return pe_fused_convert_letterbox_over_gamma(e, fg, bg, out, plan->width, plan->height, alpha, bg->d.gamma_type, gamma_to);
  }
  g_last_plan_path = 0;
  pe_frame_t *work = nullptr;
  int rc = pe_frame_copy(e, fg, &work);
  if (rc != PE_OK) return rc;
  if (!pe_run_convert_plan(e, work, plan)) { pe_frame_destroy(work); return PE_ERR_PALETTE; }
  if (work->d.palette != out->d.palette || work->d.width != out->d.width || work->d.height != out->d.height) {
    pe_frame_destroy(work);
    return set_err(PE_ERR_SIZE, "the plan does not end in the out layer's palette / size");
  }
  // compositor: bg opaque underneath, the converted layer over it with `alpha`; then gamma
  const pe_frame_t *layers[2] = {work, bg};
  const double alphas[2] = {alpha, 1.0};
  out->d.gamma_type = bg->d.gamma_type;
  rc = pe_fx_compositor_gamma(e, out, layers, alphas, 2, nullptr, gamma_to);
  pe_frame_destroy(work);
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------------------------

extern "C" int pe_frame_stats(pe_engine_t *e, const pe_frame_t *f, pe_frame_stats_t *out) {
  if (!e || !f || !out || !f->d.planes[0]) return set_err(PE_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  DevStats init;
  memset(&init, 0, sizeof(init));
  for (int k = 0; k < 4; k++) { init.minv[k] = 255; init.maxv[k] = 0; }
  PE_CUDA(cudaMemcpyAsync(e->stats_dev, &init, sizeof(init), cudaMemcpyHostToDevice, e->stream));
  const int pal = f->d.palette;
  const int psize = pal_is_planar(pal) ? 1 : pal_psize(pal);
  int a_off = -1;
  if (pal == PE_PALETTE_RGBA32 || pal == PE_PALETTE_BGRA32 || pal == PE_PALETTE_YUVA8888) a_off = 3;
  else if (pal == PE_PALETTE_ARGB32) a_off = 0;
  PE_CUDA(launch_stats(e->L(), CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, f->d.width / pal_ppmp(pal), f->d.height,
                       psize, a_off, e->stats_dev));
  DevStats res;
  PE_CUDA(cudaMemcpyAsync(&res, e->stats_dev, sizeof(res), cudaMemcpyDeviceToHost, e->stream));
  PE_CUDA(cudaStreamSynchronize(e->stream));
  for (int k = 0; k < 4; k++) { out->min[k] = (uint8_t)res.minv[k]; out->max[k] = (uint8_t)res.maxv[k]; }
  memcpy(out->hist, res.hist, sizeof(out->hist));
  out->sum = res.sum;
  // is_all_black_ish is defined on packed RGB pixels (bytes 0 .. 2 of every pixel, :2575-2577); elsewhere: -1
  const bool rgbish = !pal_is_planar(pal) && (psize == 3 || psize == 4);
  out->all_black = rgbish ? (res.not_black ? 0 : 1) : -1;
  out->all_black_ish = rgbish ? (res.not_black_ish ? 0 : 1) : -1;
  return PE_OK;
}

// hash_cmp_layer (colourspace.c:16044-16075): one 64-bit hash per row of plane 0 (minimd5 of the row's first `nbytes` bytes; nbytes
// <= 0: what the reference hashes -- `width` bytes with width the layer's width leaf in MACROpixels, i.e. the first third / quarter of
// a packed RGB row) and their XOR ("parity", :16068).  hashes: host array of `height` entries.
extern "C" int pe_frame_row_hashes(pe_engine_t *e, const pe_frame_t *f, int nbytes, uint64_t *hashes, uint64_t *parity) {
  if (!e || !f || !hashes || !f->d.planes[0]) return set_err(PE_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  const int pal = f->d.palette, h = f->d.height;
  if (nbytes <= 0) nbytes = f->d.width / pal_ppmp(pal);
  if (nbytes > f->d.rowstrides[0]) return set_err(PE_ERR_ARG, "row hash: %d bytes asked of a %d-byte row", nbytes, f->d.rowstrides[0]);
  size_t granted = 0;
  unsigned long long *dev = (unsigned long long *)e->pool.get(sizeof(unsigned long long) * (size_t)h, &granted);
  if (!dev) return set_err(PE_ERR_MEMORY, "device allocation failed");
  cudaError_t ce = launch_row_hash(e->L(), CImg{(const uint8_t *)f->d.planes[0], f->d.rowstrides[0]}, nbytes, h, dev);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(hashes, dev, sizeof(uint64_t) * (size_t)h, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  e->pool.put(dev, granted);
  if (ce != cudaSuccess) return set_err(PE_ERR_CUDA, "row hash failed: %s", cudaGetErrorString(ce));
  if (parity) { uint64_t x = 0; for (int i = 0; i < h; i++) x ^= hashes[i]; *parity = x; }
  return PE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// host-frame drop-ins: H2D -> device op -> D2H
// ---------------------------------------------------------------------------------------------------------

namespace {

struct HostFrame {
  pe_engine *e;
  pe_frame *f = nullptr;
  explicit HostFrame(pe_engine *eng) : e(eng) {}
  ~HostFrame() { if (f) pe_frame_destroy(f); }
  int upload(const pe_frame_desc_t *d) {
    int rc = frame_create_impl(e, d->palette, d->width, d->height, d->yuv_clamping, d->yuv_sampling, d->yuv_subspace,
                               d->gamma_type, 0, &f);
    if (rc != PE_OK) return rc;
    f->d.flags = d->flags;
    return pe_frame_upload(e, f, (const void *const *)d->planes, d->rowstrides);
  }
  int create_like(const pe_frame_desc_t *d) {
    return frame_create_impl(e, d->palette, d->width, d->height, d->yuv_clamping, d->yuv_sampling, d->yuv_subspace, d->gamma_type,
                             0, &f);
  }
};

void *default_alloc(size_t bytes, void *) { return malloc(bytes); }
void default_free(void *p, void *) { free(p); }

// write the device frame back into the host layer: same geometry -> into the existing buffers, otherwise new
// buffers from the caller's allocator (planes laid out contiguously, as create_empty_pixel_data does with may_contig)
int download_into(pe_engine *e, pe_frame *f, pe_frame_desc_t *layer, const pe_host_allocator_t *alloc) {
  pe_frame_desc_t d = f->d;
  bool same = layer->palette == d.palette && layer->width == d.width && layer->height == d.height && layer->planes[0];
  if (same) {
    int rc = pe_frame_download(e, f, layer->planes, layer->rowstrides);
    if (rc != PE_OK) return rc;
  } else {
    pe_host_allocator_t a = alloc && alloc->alloc && alloc->free ? *alloc : pe_host_allocator_t{default_alloc, default_free, nullptr};
    int np, rs[PE_MAXPLANES], ph[PE_MAXPLANES];
    pe_frame_layout(d.palette, d.width, d.height, &np, rs, ph);
    size_t total = 0;
    for (int p = 0; p < np; p++) total += (size_t)rs[p] * ph[p];
    uint8_t *blk = (uint8_t *)a.alloc(total + 64, a.user);
    if (!blk) return set_err(PE_ERR_MEMORY, "host allocator returned NULL for %zu bytes", total);
    void *planes[PE_MAXPLANES] = {nullptr, nullptr, nullptr, nullptr};
    size_t off = 0;
    for (int p = 0; p < np; p++) { planes[p] = blk + off; off += (size_t)rs[p] * ph[p]; }
    int rc = pe_frame_download(e, f, planes, rs);
    if (rc != PE_OK) { a.free(blk, a.user); return rc; }
    if (layer->planes[0]) a.free(layer->planes[0], a.user);  // the old contiguous block
    for (int p = 0; p < PE_MAXPLANES; p++) { layer->planes[p] = planes[p]; layer->rowstrides[p] = p < np ? rs[p] : 0; }
    layer->nplanes = np;
  }
  layer->palette = d.palette; layer->width = d.width; layer->height = d.height;
  layer->yuv_clamping = d.yuv_clamping; layer->yuv_sampling = d.yuv_sampling; layer->yuv_subspace = d.yuv_subspace;
  layer->gamma_type = d.gamma_type; layer->flags = d.flags;
  return PE_OK;
}

}  // namespace

extern "C" int pe_host_convert_layer_palette_full(pe_engine_t *e, pe_frame_desc_t *layer, int outpl, int oclamping, int osampling,
                                                  int osubspace, int tgt_gamma, const pe_host_allocator_t *alloc) {
  if (!e || !layer || !layer->planes[0]) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  HostFrame h(e);
  if (h.upload(layer) != PE_OK) return PE_FALSE;
  if (!pe_convert_layer_palette_full(e, h.f, outpl, oclamping, osampling, osubspace, tgt_gamma)) return PE_FALSE;
  return download_into(e, h.f, layer, alloc) == PE_OK ? PE_TRUE : PE_FALSE;
}

extern "C" int pe_host_resize_layer(pe_engine_t *e, pe_frame_desc_t *layer, int width, int height, int interp, int opal_hint,
                                    int oclamp_hint, const pe_host_allocator_t *alloc) {
  if (!e || !layer || !layer->planes[0]) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  HostFrame h(e);
  if (h.upload(layer) != PE_OK) return PE_FALSE;
  if (!pe_resize_layer(e, h.f, width, height, interp, opal_hint, oclamp_hint)) return PE_FALSE;
  return download_into(e, h.f, layer, alloc) == PE_OK ? PE_TRUE : PE_FALSE;
}

extern "C" int pe_host_letterbox_layer(pe_engine_t *e, pe_frame_desc_t *layer, int nwidth, int nheight, int width, int height,
                                       int interp, int tpal, int tclamp, const pe_host_allocator_t *alloc) {
  if (!e || !layer || !layer->planes[0]) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  HostFrame h(e);
  if (h.upload(layer) != PE_OK) return PE_FALSE;
  if (!pe_letterbox_layer(e, h.f, nwidth, nheight, width, height, interp, tpal, tclamp)) return PE_FALSE;
  return download_into(e, h.f, layer, alloc) == PE_OK ? PE_TRUE : PE_FALSE;
}

extern "C" int pe_host_gamma_convert_layer(pe_engine_t *e, int gamma_type, pe_frame_desc_t *layer) {
  if (!e || !layer || !layer->planes[0]) { set_err(PE_ERR_ARG, "NULL argument"); return PE_FALSE; }
  HostFrame h(e);
  if (h.upload(layer) != PE_OK) return PE_FALSE;
  if (!pe_gamma_convert_layer(e, gamma_type, h.f)) return PE_FALSE;
  return download_into(e, h.f, layer, nullptr) == PE_OK ? PE_TRUE : PE_FALSE;
}

namespace {
int host_blend(pe_engine_t *e, int which, int type, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2, pe_frame_desc_t *out,
               int bf) {
  if (!e || !in1 || !in2 || !out || !in1->planes[0] || !in2->planes[0] || !out->planes[0]) return set_err(PE_ERR_ARG, "NULL argument");
  HostFrame a(e), b(e), o(e);
  int rc;
  if ((rc = a.upload(in1)) != PE_OK || (rc = b.upload(in2)) != PE_OK) return rc;
  const bool inplace = out->planes[0] == in1->planes[0];  // effects-weed.c:2304-2314
  // the 4-byte chroma blend never writes the alpha byte: a separate out channel keeps what it held
  if (!inplace && (rc = (pal_psize(out->palette) == 4 ? o.upload(out) : o.create_like(out))) != PE_OK) return rc;
  pe_frame *of = inplace ? a.f : o.f;
  rc = which == 0 ? pe_fx_simple_blend(e, type, a.f, b.f, of, bf) : pe_fx_multi_blend(e, type, a.f, b.f, of, bf);
  if (rc != PE_OK) return rc;
  return pe_frame_download(e, of, out->planes, out->rowstrides);
}
}  // namespace

extern "C" int pe_host_simple_blend(pe_engine_t *e, int type, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2,
                                    pe_frame_desc_t *out, int blend_factor) {
  return host_blend(e, 0, type, in1, in2, out, blend_factor);
}

extern "C" int pe_host_multi_blend(pe_engine_t *e, int type, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2,
                                   pe_frame_desc_t *out, int blend_factor) {
  return host_blend(e, 1, type, in1, in2, out, blend_factor);
}

extern "C" int pe_host_slide_over(pe_engine_t *e, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2, pe_frame_desc_t *out,
                                  int transval, int direction, int mvlower, int mvupper) {
  if (!e || !in1 || !in2 || !out || !in1->planes[0] || !in2->planes[0] || !out->planes[0]) return set_err(PE_ERR_ARG, "NULL argument");
  HostFrame a(e), b(e), o(e);
  int rc;
  if ((rc = a.upload(in1)) != PE_OK || (rc = b.upload(in2)) != PE_OK || (rc = o.create_like(out)) != PE_OK) return rc;
  if ((rc = pe_fx_slide_over(e, a.f, b.f, o.f, transval, direction, mvlower, mvupper)) != PE_OK) return rc;
  return pe_frame_download(e, o.f, out->planes, out->rowstrides);
}

extern "C" int pe_host_softlight(pe_engine_t *e, const pe_frame_desc_t *in, pe_frame_desc_t *out) {
  if (!e || !in || !out || !in->planes[0] || !out->planes[0]) return set_err(PE_ERR_ARG, "NULL argument");
  HostFrame a(e), o(e);
  int rc;
  if ((rc = a.upload(in)) != PE_OK || (rc = o.create_like(out)) != PE_OK) return rc;
  if ((rc = pe_fx_softlight(e, a.f, o.f)) != PE_OK) return rc;
  return pe_frame_download(e, o.f, out->planes, out->rowstrides);
}

namespace {
// two host in channels, one host out channel that may be in channel 0 (CAN_DO_INPLACE, effects-weed.c:2304-2314)
template <class F>
int host_select(pe_engine_t *e, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2, pe_frame_desc_t *out, F run) {
  if (!e || !in1 || !in2 || !out || !in1->planes[0] || !in2->planes[0] || !out->planes[0]) return set_err(PE_ERR_ARG, "NULL argument");
  HostFrame a(e), b(e), o(e);
  int rc;
  if ((rc = a.upload(in1)) != PE_OK || (rc = b.upload(in2)) != PE_OK) return rc;
  const bool inplace = out->planes[0] == in1->planes[0];
  if (!inplace && (rc = o.create_like(out)) != PE_OK) return rc;
  pe_frame *of = inplace ? a.f : o.f;
  if ((rc = run(a.f, b.f, of)) != PE_OK) return rc;
  return pe_frame_download(e, of, out->planes, out->rowstrides);
}
}  // namespace

extern "C" int pe_host_triple_split(pe_engine_t *e, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2, pe_frame_desc_t *out, double start,
                                    int sym, double end, int vert, double borderw, const int bordercol[3]) {
  return host_select(e, in1, in2, out, [&](pe_frame *a, pe_frame *b, pe_frame *o) {
    return pe_fx_triple_split(e, a, b, o, start, sym, end, vert, borderw, bordercol);
  });
}

extern "C" int pe_host_multi_transition(pe_engine_t *e, int type, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2, pe_frame_desc_t *out,
                                        double amount, const pe_dissolve_mask_t *mask) {
  return host_select(e, in1, in2, out, [&](pe_frame *a, pe_frame *b, pe_frame *o) { return pe_fx_multi_transition(e, type, a, b, o, amount, mask); });
}

// gdk/compositor.c compositor_process :127 on host channels: every enabled in channel travels up once, the paints run on the device,
// the out channel travels back.  layers[z] == NULL (or without pixel data): the host disabled that channel (:195-199).
extern "C" int pe_host_compositor(pe_engine_t *e, pe_frame_desc_t *out, const pe_frame_desc_t *const *layers, const double *alpha,
                                  int nlayers, const int bgcol[3]) {
  if (!e || !out || !out->planes[0] || nlayers < 0 || (nlayers && (!layers || !alpha))) return set_err(PE_ERR_ARG, "NULL argument");
  std::vector<HostFrame> up;
  up.reserve((size_t)nlayers);
  std::vector<const pe_frame_t *> dev((size_t)nlayers, nullptr);
  int rc;
  for (int z = 0; z < nlayers; z++) {
    up.emplace_back(e);
    if (!layers[z] || !layers[z]->planes[0] || alpha[z] <= 0.) continue;
    if ((rc = up.back().upload(layers[z])) != PE_OK) return rc;
    dev[(size_t)z] = up.back().f;
  }
  HostFrame o(e);
  if ((rc = o.create_like(out)) != PE_OK) return rc;
  if ((rc = pe_fx_compositor(e, o.f, dev.data(), alpha, nlayers, bgcol)) != PE_OK) return rc;
  return pe_frame_download(e, o.f, out->planes, out->rowstrides);
}

// A batch of host frames through the fused chain with the PCIe copies overlapped: three device slots rotate through
// upload (h2d stream) -> kernel (engine stream) -> download (d2h stream), chained by events, so that frame i+1 travels to the
// device and frame i-1 travels back while frame i is computed.  Host buffers should be pinned (pe_host_alloc) for the copies to
// be asynchronous; pageable memory still gives correct results.
extern "C" int pe_host_fused_convert_letterbox_over_gamma_batch(pe_engine_t *e, int n, const pe_frame_desc_t *const *fg,
                                                                const pe_frame_desc_t *const *bg, pe_frame_desc_t *const *out,
                                                                int inner_w, int inner_h, double alpha, int gamma_from,
                                                                int gamma_to) {
  if (!e || n <= 0 || !fg || !bg || !out) return set_err(PE_ERR_ARG, "NULL / empty argument");
  for (int i = 0; i < n; i++)
    if (!fg[i] || !bg[i] || !out[i] || !fg[i]->planes[0] || !bg[i]->planes[0] || !out[i]->planes[0]) return set_err(PE_ERR_ARG, "NULL frame");
  std::lock_guard<std::mutex> lk(e->mu);
  PE_CUDA(cudaSetDevice(e->device));
  constexpr int NS_MAX = 6;
  int NS = 3;  // frames in flight: one uploading, one computing, one downloading
  if (const char *sv = getenv("PE_PIPE_SLOTS")) { NS = atoi(sv); if (NS < 2) NS = 2; if (NS > NS_MAX) NS = NS_MAX; }
  if (!e->h2d_stream) {
    PE_CUDA(cudaStreamCreateWithFlags(&e->h2d_stream, cudaStreamNonBlocking));
    PE_CUDA(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
    for (int k = 0; k < NS_MAX; k++) {
      PE_CUDA(cudaEventCreateWithFlags(&e->pipe_up[k], cudaEventDisableTiming));
      PE_CUDA(cudaEventCreateWithFlags(&e->pipe_comp[k], cudaEventDisableTiming));
      PE_CUDA(cudaEventCreateWithFlags(&e->pipe_free[k], cudaEventDisableTiming));
    }
  }
  // device slots
  pe_frame slots[NS_MAX][3];
  const int ns = n < NS ? n : NS;
  int rc = PE_OK;
  auto release = [&]() { for (int k = 0; k < ns; k++) for (int j = 0; j < 3; j++) frame_release_pixels(&slots[k][j]); };
  for (int k = 0; k < ns && rc == PE_OK; k++) {
    const pe_frame_desc_t *d3[3] = {fg[0], bg[0], out[0]};
    for (int j = 0; j < 3 && rc == PE_OK; j++) {
      slots[k][j].e = e;
      slots[k][j].d = *d3[j];
      rc = frame_alloc(e, &slots[k][j]);
    }
  }
  if (rc != PE_OK) { release(); return rc; }
  // blocks from the pool may still be in use by earlier work on the engine stream
  PE_CUDA(cudaEventRecord(e->pipe_free[0], e->stream));
  PE_CUDA(cudaStreamWaitEvent(e->h2d_stream, e->pipe_free[0], 0));
  const bool copy2d_only = getenv("PE_HOST_COPY2D") != nullptr;  // A/B switch for the measurement only
  // pageable host planes go through the stager (pe_hoststage.h): uploads are staged by the copy threads before their DMA is queued,
  // downloads land in a ring buffer and are copied out when the slot comes round again (pend[k]) or at the end
  std::vector<pe::PendingOut> pend[NS_MAX];
  auto flush_pending = [&](int k) -> cudaError_t {
    cudaError_t r = cudaSuccess;
    for (auto &po : pend[k]) { const cudaError_t c = e->stager.finish(po); if (c != cudaSuccess) r = c; }
    pend[k].clear();
    return r;
  };
  auto copy_planes = [&](cudaStream_t st, pe_frame &dev, const pe_frame_desc_t &host, bool to_device, int slot) -> cudaError_t {
    for (int p = 0; p < dev.d.nplanes; p++) {
      const int wbytes = plane_row_bytes(dev.d, p);
      if ((size_t)dev.d.rowstrides[p] * dev.plane_heights[p] >= pe::HostStager::kMinBytes && !e->no_host_staging && pe::HostStager::pageable(host.planes[p])) {
        if (to_device) {
          const size_t wb = host.rowstrides[p] == dev.d.rowstrides[p] ? (size_t)dev.d.rowstrides[p] : (size_t)wbytes;
          if (e->stager.upload(st, dev.d.planes[p], (size_t)dev.d.rowstrides[p], host.planes[p], (size_t)host.rowstrides[p], wb,
                               (size_t)dev.plane_heights[p]) == cudaSuccess)
            continue;
        } else {
          pe::PendingOut po;
          if (e->stager.download_begin(st, dev.d.planes[p], (size_t)dev.d.rowstrides[p], host.planes[p], (size_t)host.rowstrides[p], (size_t)wbytes,
                                       (size_t)dev.plane_heights[p], &po) == cudaSuccess) {
            pend[slot].push_back(po);
            continue;
          }
        }
        cudaGetLastError();   // the staged path could not take it: the plain copy below
      }
      if (host.rowstrides[p] == dev.d.rowstrides[p] && dev.plane_heights[p] > 0 && !copy2d_only) {  // same pitch on both sides: one linear copy
        size_t nbytes = (size_t)dev.d.rowstrides[p] * (dev.plane_heights[p] - 1) + wbytes;
        // planes that follow each other without a gap on BOTH sides (LiVES allocates planar frames contiguously,
        // WEED_LEAF_HOST_PIXEL_DATA_CONTIGUOUS; ours are one pool block) travel in the same copy
        int q = p;
        while (q + 1 < dev.d.nplanes && host.rowstrides[q + 1] == dev.d.rowstrides[q + 1] && dev.plane_heights[q + 1] > 0) {
          const size_t span = (size_t)dev.d.rowstrides[q] * dev.plane_heights[q];
          if ((const uint8_t *)host.planes[q + 1] != (const uint8_t *)host.planes[q] + span ||
              (const uint8_t *)dev.d.planes[q + 1] != (const uint8_t *)dev.d.planes[q] + span)
            break;
          q++;
        }
        if (q > p) {
          nbytes = (size_t)((const uint8_t *)dev.d.planes[q] - (const uint8_t *)dev.d.planes[p]) +
                   (size_t)dev.d.rowstrides[q] * (dev.plane_heights[q] - 1) + plane_row_bytes(dev.d, q);
        }
        cudaError_t ce1 = to_device ? cudaMemcpyAsync(dev.d.planes[p], host.planes[p], nbytes, cudaMemcpyHostToDevice, st)
                                    : cudaMemcpyAsync(host.planes[p], dev.d.planes[p], nbytes, cudaMemcpyDeviceToHost, st);
        if (ce1 != cudaSuccess) return ce1;
        p = q;
        continue;
      }
      cudaError_t ce = to_device ? cudaMemcpy2DAsync(dev.d.planes[p], dev.d.rowstrides[p], host.planes[p], host.rowstrides[p], wbytes,
                                                     dev.plane_heights[p], cudaMemcpyHostToDevice, st)
                                 : cudaMemcpy2DAsync(host.planes[p], host.rowstrides[p], dev.d.planes[p], dev.d.rowstrides[p], wbytes,
                                                     dev.plane_heights[p], cudaMemcpyDeviceToHost, st);
      if (ce != cudaSuccess) return ce;
    }
    return cudaSuccess;
  };
  for (int i = 0; i < n; i++) {
    const int k = i % NS;
    if (fg[i]->palette != fg[0]->palette || fg[i]->width != fg[0]->width || fg[i]->height != fg[0]->height ||
        bg[i]->width != bg[0]->width || bg[i]->height != bg[0]->height || bg[i]->palette != bg[0]->palette ||
        out[i]->width != out[0]->width || out[i]->height != out[0]->height || out[i]->palette != out[0]->palette) {
      rc = set_err(PE_ERR_SIZE, "frames of a batch must share palette and size");
      break;
    }
    cudaError_t ce = cudaSuccess;
    // upload into slot k once its previous occupant has been downloaded (and, from a ring buffer, copied out to its pageable owner)
    if (i >= NS) ce = cudaStreamWaitEvent(e->h2d_stream, e->pipe_free[k], 0);
    if (ce == cudaSuccess && !pend[k].empty()) ce = flush_pending(k);
    slots[k][0].d.yuv_clamping = fg[i]->yuv_clamping; slots[k][0].d.yuv_subspace = fg[i]->yuv_subspace;
    slots[k][1].d.gamma_type = bg[i]->gamma_type; slots[k][1].d.flags = bg[i]->flags;
    if (ce == cudaSuccess) ce = copy_planes(e->h2d_stream, slots[k][0], *fg[i], true, k);
    if (ce == cudaSuccess) ce = copy_planes(e->h2d_stream, slots[k][1], *bg[i], true, k);
    if (ce == cudaSuccess) ce = cudaEventRecord(e->pipe_up[k], e->h2d_stream);
    // compute
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->stream, e->pipe_up[k], 0);
    if (ce != cudaSuccess) { rc = set_err(PE_ERR_CUDA, "pipeline upload failed: %s", cudaGetErrorString(ce)); break; }
    const pe_frame_t *f1[1] = {&slots[k][0]}, *b1[1] = {&slots[k][1]};
    pe_frame_t *o1[1] = {&slots[k][2]};
    rc = fused_locked(e, 1, f1, b1, o1, inner_w, inner_h, alpha, gamma_from, gamma_to);
    if (rc != PE_OK) break;
    ce = cudaEventRecord(e->pipe_comp[k], e->stream);
    // download
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->d2h_stream, e->pipe_comp[k], 0);
    if (ce == cudaSuccess) ce = copy_planes(e->d2h_stream, slots[k][2], *out[i], false, k);
    if (ce == cudaSuccess) ce = cudaEventRecord(e->pipe_free[k], e->d2h_stream);
    if (ce != cudaSuccess) { rc = set_err(PE_ERR_CUDA, "pipeline download failed: %s", cudaGetErrorString(ce)); break; }
    out[i]->gamma_type = slots[k][2].d.gamma_type;
    out[i]->flags = slots[k][2].d.flags;
  }
  cudaError_t ce = cudaSuccess;
  for (int k = 0; k < NS_MAX; k++) { const cudaError_t c = flush_pending(k); if (c != cudaSuccess) ce = c; }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->d2h_stream);
  cudaError_t ce2 = cudaStreamSynchronize(e->h2d_stream);
  cudaError_t ce3 = cudaStreamSynchronize(e->stream);
  release();
  if (rc == PE_OK && (ce != cudaSuccess || ce2 != cudaSuccess || ce3 != cudaSuccess))
    rc = set_err(PE_ERR_CUDA, "pipeline failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ce2 != cudaSuccess ? ce2 : ce3));
  return rc;
}

extern "C" int pe_host_fused_convert_letterbox_over_gamma(pe_engine_t *e, const pe_frame_desc_t *fg, const pe_frame_desc_t *bg,
                                                          pe_frame_desc_t *out, int inner_w, int inner_h, double alpha,
                                                          int gamma_from, int gamma_to) {
  if (!e || !fg || !bg || !out || !fg->planes[0] || !bg->planes[0] || !out->planes[0]) return set_err(PE_ERR_ARG, "NULL argument");
  HostFrame a(e), b(e), o(e);
  int rc;
  if ((rc = a.upload(fg)) != PE_OK || (rc = b.upload(bg)) != PE_OK || (rc = o.create_like(out)) != PE_OK) return rc;
  rc = pe_fused_convert_letterbox_over_gamma(e, a.f, b.f, o.f, inner_w, inner_h, alpha, gamma_from, gamma_to);
  if (rc != PE_OK) return rc;
  rc = pe_frame_download(e, o.f, out->planes, out->rowstrides);
  if (rc == PE_OK) { out->gamma_type = o.f->d.gamma_type; out->flags = o.f->d.flags; }
  return rc;
}
