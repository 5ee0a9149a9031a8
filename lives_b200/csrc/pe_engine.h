// pe_engine.h -- internal state behind the opaque handles of include/pixel_engine.h.
#pragma once
#include "pe_hoststage.h"
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/pixel_engine.h"
#include "pe_kernels.h"
#include "pe_tables.h"

namespace pe {

// Device block pool (the role of LiVES' bigblock allocator, src/memory.c:37-47): frames are replaced on every
// palette conversion / resize, so blocks are recycled by size class instead of going through cudaMalloc / cudaFree
// (both synchronise the device).  Stream-ordered: a recycled block is only ever reused on the engine's own stream.
class DevPool {
 public:
  void *get(size_t bytes, size_t *granted);
  void put(void *p, size_t granted);
  void release_all();
  size_t bytes_held() const { return held_; }
  // While a batch call fans its layers out over side streams, freed blocks must not be handed out again before the streams
  // have joined: put() parks them, flush_deferred() returns them to the free lists.
  void defer(bool on) { defer_ = on; }
  void flush_deferred();
  bool undefer(void *p);  // take a parked block back (a failed batch restores its layers)

 private:
  std::multimap<size_t, void *> free_;
  std::vector<std::pair<void *, size_t>> deferred_;
  bool defer_ = false;
  size_t held_ = 0;
};

struct GammaKey {
  double fileg;
  int from, to;
  bool operator<(const GammaKey &o) const { return std::tie(fileg, from, to) < std::tie(o.fileg, o.from, o.to); }
};
struct OverKey {
  double alpha;
  const uint8_t *lut;
  bool operator<(const OverKey &o) const { return std::tie(alpha, lut) < std::tie(o.alpha, o.lut); }
};
struct FilterKey {
  int src_n, dst_n, bits;
  bool operator<(const FilterKey &o) const { return std::tie(src_n, dst_n, bits) < std::tie(o.src_n, o.dst_n, o.bits); }
};
struct DevFilterEntry {
  DevFilter dev;
  ResizeFilter host;
  void *rows4 = nullptr;  // int4 per output row {first, c3 | c2 << 16, c1 | c0 << 16, 0}: k_fused3's view of a <= 4-tap bank
  int rows4_x16 = 0;      // its coefficients are scaled by 16 (no tap of the bank is 4096)
  void *pack4 = nullptr;  // int4 per output sample {c0 | c1 << 16, c2 | c3 << 16, first, aux}: k_cvt_resize's view of a <= 4-tap non-negative bank
  unsigned long tick = 0; // last use (LRU bound of the cache)
};
struct OverEntry {
  uint8_t *dev;
  unsigned long tick;
};
struct Lut8Entry {
  uint8_t host[256];
  uint8_t *dev;
};

}  // namespace pe

struct pe_engine {
  pe_config_t cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0;
  int sm_limit = 0;   // > 0: persistent kernels use at most this many SMs (pe_engine_set_sm_limit)
  int resize_recipe = 1;  // 1 libswscale's coefficient recipes (default), 0 the round-1 triangle contract (pe_engine_set_resize_recipe)
  long launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host-frame batch pipeline (pe_host_*_batch): copies run on their own streams, overlapped with the kernels
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  bool no_host_staging = false;      // PE_HOST_NO_STAGING=1: measurement switch (every host copy is the plain cudaMemcpy*Async)
  pe::HostStager stager;             // pageable host planes travel through page-locked ring buffers filled / drained by copy threads
  cudaEvent_t pipe_up[6] = {}, pipe_comp[6] = {}, pipe_free[6] = {};
  cudaStream_t egress_stream = nullptr;  // pe_render_out_*: the final packed frame travels on its own stream
  cudaEvent_t egress_ready[4] = {}, egress_done[4] = {};
  bool egress_busy[4] = {};
  std::mutex mu;  // the reference calls these entry points from several proc-threads (different layers)

  pe::ConvTables conv_host[2][2];    // [clamping][bt709]
  int32_t *conv_dev[2][2] = {};      // 14 x 256 int32 each, followed by the extended planar tables (DevConv::ext)
  uint8_t *cavg_dev[2] = {};         // chroma averaging tables (init_average :190): [0] clamped, [1] unclamped
  float *ftab_dev[2] = {};           // the BT.709 float tables of the reference's experimental float path (pe_convert_yuv888_to_rgb_float)
  uint8_t *yy_dev = nullptr;         // the four 256-byte clamped <-> unclamped tables (init_YUV_to_YUV_tables :1108)
  uint8_t *premult_dev[6] = {};      // built on first use (init_unal is lazy in the reference too, :11985)
  int32_t *luma_dev = nullptr;       // plugin-side calc_luma tables [3][256]
  std::map<pe::GammaKey, pe::Lut8Entry> lut8;
  std::map<pe::GammaKey, uint16_t *> lut16;
  std::map<pe::OverKey, pe::OverEntry> over;
  unsigned long cache_tick = 0;
  std::map<pe::FilterKey, pe::DevFilterEntry> filters;
  pe::DevPool pool;
  pe::DevStats *stats_dev = nullptr;
  // small device scratch for per-launch argument arrays (BlendFrame / FusedArgs), grown on demand
  void *args_dev = nullptr;
  // set around convert_locked by pe_fx_convert_crossfade: the planar YUV -> RGB converter blends with this frame on the fly
  const uint8_t *fuse_blend2 = nullptr;
  int fuse_blend2_rs = 0, fuse_blend_bf = 0;
  // batch calls: planar YUV -> RGB conversions are queued and leave as ONE launch per 32 same-shaped frames (flush_yuv_pending)
  struct RgbJob { const uint8_t *src; int irow; uint8_t *dst; int orow, width, height; pe::RgbLayout in, out; const uint8_t *lut; };
  bool rgb_defer = false;            // ... and the RGB <-> RGB permutations (flush_rgb_pending)
  std::vector<RgbJob> rgb_pending;
  struct RszJob { const uint8_t *src; int srs, sw, sh; uint8_t *dst; int drs, dw, dh, psize, kx, ky; };
  struct OverJob { const uint8_t *bg, *fg; uint8_t *dst; int rs_bg, rs_fg, rs_d, w, h, psize, k256; const uint8_t *lut; };
  bool over_defer = false;           // ... and the integer alpha-over paints of a compositor batch (flush_over_pending)
  std::vector<OverJob> over_pending;
  bool rsz_defer = false;            // ... and so are the resizes of 4-byte packed frames (flush_rsz_pending)
  std::vector<RszJob> rsz_pending;
  bool yuv_defer = false;
  std::vector<pe::YuvToRgbArgs> yuv_pending;
  // batch calls: layers 1 .. n-1 run on four side streams (forked from / joined to the engine stream by events)
  cudaStream_t fan_stream[4] = {};
  cudaEvent_t fan_fork = nullptr, fan_join[4] = {};
  unsigned int *f3_sched = nullptr;  // k_fused3's two work counters (zero between launches)
  size_t args_cap = 0;
  void *args_pinned = nullptr;
  size_t args_pinned_cap = 0;
  cudaEvent_t args_ev = nullptr;  // args_pinned may be rewritten once the previous upload has completed

  pe::Launch L() { return pe::Launch{stream, sm_limit > 0 && sm_limit < sm_count ? sm_limit : sm_count, &launches}; }
};

struct pe_frame {
  pe_engine *e = nullptr;
  pe_frame_desc_t d{};
  void *base = nullptr;  // pool block holding all planes (nullptr for wrapped frames)
  size_t granted = 0;
  int plane_heights[PE_MAXPLANES] = {};
};
