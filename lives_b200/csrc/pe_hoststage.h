// pe_hoststage.h -- PAGEABLE host buffers on the way to / from the device.
// LiVES allocates pixel memory with lives_calloc_safety (pageable).  cudaMemcpyAsync from / to such memory is staged by the driver
// through its own pinned buffer by ONE thread, synchronously: ~10 GB/s, a fifth of the PCIe 5 link (the headline chain on pageable
// buffers: 132 frames/s against 1 046 on page-locked ones, bench.py e2e_pageable).  The stager does the same staging with several host
// threads and keeps the DMA asynchronous: a ring of page-locked buffers, a small pool of copy threads (a plane is cut into one piece per
// thread), the transfer itself from / to the ring.  Hosts that page-lock their pixel blocks once (pe_host_register) never come here.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace pe {

class CopyPool {
 public:
  explicit CopyPool(int nthreads);
  ~CopyPool();
  // dst / src rows of wbytes bytes at the given strides; blocks until done
  void copy2d(void *dst, size_t dst_stride, const void *src, size_t src_stride, size_t wbytes, size_t rows);
  int threads() const { return n_; }

 private:
  void worker(int idx);
  struct Job { uint8_t *dst; const uint8_t *src; size_t ds, ss, wbytes, rows; } job_{};
  int n_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_work_, cv_done_;
  unsigned long gen_ = 0;
  int remaining_ = 0;
  bool stop_ = false;
};

struct StageBuf {
  void *ptr = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
  bool busy = false;     // a DMA recorded on ev may still be using it
  bool pending = false;  // holds a downloaded plane that finish() has not copied out yet
};

// a plane that is on its way back: device -> ring buffer (DMA in flight behind sb->ev) -> caller's rows (finish())
struct PendingOut {
  StageBuf *sb;
  void *host;
  size_t host_stride, dev_stride, wbytes, rows;
};

class HostStager {
 public:
  static constexpr int kRing = 4;   // per direction
  static constexpr size_t kMinBytes = 1u << 20;   // smaller planes: the driver's own staging is as good
  ~HostStager();
  // page-locked (allocated by CUDA or registered) memory needs no staging
  static bool pageable(const void *p);
  // rows of wbytes bytes of a pageable plane into the device plane: threads fill a ring buffer in the DEVICE layout, one asynchronous
  // copy moves it.  Returns cudaSuccess, or an error (the caller falls back to the plain copy on cudaErrorMemoryAllocation).
  cudaError_t upload(cudaStream_t st, void *dev, size_t dev_stride, const void *host, size_t host_stride, size_t wbytes, size_t rows);
  // the way back, first half: device plane -> ring buffer, asynchronous
  cudaError_t download_begin(cudaStream_t st, const void *dev, size_t dev_stride, void *host, size_t host_stride, size_t wbytes, size_t rows,
                             PendingOut *out);
  // second half: wait for the transfer, threads copy the rows out
  cudaError_t finish(PendingOut &p);

 private:
  StageBuf *acquire(StageBuf *ring, int *next, size_t bytes, cudaError_t *err);
  CopyPool *pool();
  StageBuf up_[kRing], down_[kRing];
  int next_up_ = 0, next_down_ = 0;
  CopyPool *pool_ = nullptr;
};

}  // namespace pe
