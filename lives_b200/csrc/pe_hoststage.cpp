// pe_hoststage.cpp -- see pe_hoststage.h
#include "pe_hoststage.h"

#include <cstdlib>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace pe {

// A staging copy writes a buffer it will not read again (the DMA or the caller does): non-temporal stores skip the read-for-ownership
// of the destination lines, a third of the memory traffic of a plain memcpy -- and host memory bandwidth is what bounds the pageable
// path (profiles/r02zj_pageable.log).  glibc's memcpy only switches to them far above the piece a copy thread gets.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void copy_stream_avx2(uint8_t *dst, const uint8_t *src, size_t n) {
  size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
  if (head > n) head = n;
  memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
    _mm256_stream_si256((__m256i *)(dst + i), a); _mm256_stream_si256((__m256i *)(dst + i + 32), b);
    _mm256_stream_si256((__m256i *)(dst + i + 64), c); _mm256_stream_si256((__m256i *)(dst + i + 96), d);
  }
  _mm_sfence();
  memcpy(dst + i, src + i, n - i);
}
#endif

static void copy_bytes(uint8_t *dst, const uint8_t *src, size_t n) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2") && getenv("PE_HOST_COPY_STREAM") != nullptr;   // opt-in, see profiles/r02zj_pageable.log
  if (avx2 && n >= 4096) { copy_stream_avx2(dst, src, n); return; }
#endif
  memcpy(dst, src, n);
}

CopyPool::CopyPool(int nthreads) : n_(nthreads < 1 ? 1 : nthreads) {
  for (int i = 0; i < n_; i++) th_.emplace_back([this, i] { worker(i); });
}

CopyPool::~CopyPool() {
  {
    std::lock_guard<std::mutex> lk(m_);
    stop_ = true;
    gen_++;
  }
  cv_work_.notify_all();
  for (auto &t : th_) t.join();
}

void CopyPool::worker(int idx) {
  unsigned long seen = 0;
  for (;;) {
    Job j;
    {
      std::unique_lock<std::mutex> lk(m_);
      cv_work_.wait(lk, [&] { return gen_ != seen; });
      seen = gen_;
      if (stop_) return;
      j = job_;
    }
    if (j.ds == j.ss && j.ds == j.wbytes) {   // one dense block: pieces of whole pages
      const size_t total = j.wbytes * j.rows;
      size_t piece = (total / (size_t)n_ + 4095) & ~(size_t)4095;
      if (piece == 0) piece = 4096;
      const size_t a = piece * (size_t)idx, b = a + piece < total ? a + piece : total;
      if (a < total) copy_bytes(j.dst + a, j.src + a, b - a);
    } else {
      const size_t per = (j.rows + (size_t)n_ - 1) / (size_t)n_;
      const size_t r0 = per * (size_t)idx, r1 = r0 + per < j.rows ? r0 + per : j.rows;
      for (size_t r = r0; r < r1; r++) copy_bytes(j.dst + r * j.ds, j.src + r * j.ss, j.wbytes);
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      if (--remaining_ == 0) cv_done_.notify_all();
    }
  }
}

void CopyPool::copy2d(void *dst, size_t dst_stride, const void *src, size_t src_stride, size_t wbytes, size_t rows) {
  if (!wbytes || !rows) return;
  std::unique_lock<std::mutex> lk(m_);
  job_ = Job{(uint8_t *)dst, (const uint8_t *)src, dst_stride, src_stride, wbytes, rows};
  remaining_ = n_;
  gen_++;
  cv_work_.notify_all();
  cv_done_.wait(lk, [&] { return remaining_ == 0; });
}

HostStager::~HostStager() {
  for (StageBuf *ring : {up_, down_})
    for (int i = 0; i < kRing; i++) {
      if (ring[i].ev) cudaEventDestroy(ring[i].ev);
      if (ring[i].ptr) cudaFreeHost(ring[i].ptr);
    }
  delete pool_;
}

bool HostStager::pageable(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

CopyPool *HostStager::pool() {
  if (!pool_) {
    int n = (int)std::thread::hardware_concurrency() / 2;
    if (n > 8) n = 8;
    if (n < 1) n = 1;
    if (const char *s = getenv("PE_HOST_COPY_THREADS")) n = atoi(s);
    pool_ = new CopyPool(n);
  }
  return pool_;
}

StageBuf *HostStager::acquire(StageBuf *ring, int *next, size_t bytes, cudaError_t *err) {
  // round robin, skipping buffers whose downloaded plane has not been copied out yet
  int pick = -1;
  for (int t = 0; t < kRing; t++) {
    const int i = (*next + t) % kRing;
    if (!ring[i].pending) { pick = i; break; }
  }
  if (pick < 0) { *err = cudaErrorMemoryAllocation; return nullptr; }   // (the caller falls back to the plain copy)
  StageBuf &b = ring[pick];
  *next = (pick + 1) % kRing;
  *err = cudaSuccess;
  if (b.busy) {   // the transfer that used this buffer kRing planes ago
    if ((*err = cudaEventSynchronize(b.ev)) != cudaSuccess) return nullptr;
    b.busy = false;
  }
  if (!b.ev && (*err = cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming)) != cudaSuccess) return nullptr;
  if (b.cap < bytes) {
    if (b.ptr) cudaFreeHost(b.ptr);
    b.ptr = nullptr;
    b.cap = 0;
    const size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    if ((*err = cudaHostAlloc(&b.ptr, want, cudaHostAllocPortable)) != cudaSuccess) { b.ptr = nullptr; return nullptr; }
    b.cap = want;
  }
  return &b;
}

cudaError_t HostStager::upload(cudaStream_t st, void *dev, size_t dev_stride, const void *host, size_t host_stride, size_t wbytes, size_t rows) {
  cudaError_t err;
  const size_t bytes = dev_stride * (rows - 1) + wbytes;
  StageBuf *b = acquire(up_, &next_up_, bytes, &err);
  if (!b) return err;
  // the ring buffer mirrors the device plane (same stride): the transfer is one linear copy
  if (host_stride == dev_stride) pool()->copy2d(b->ptr, bytes, host, bytes, bytes, 1);
  else pool()->copy2d(b->ptr, dev_stride, host, host_stride, wbytes, rows);
  if ((err = cudaMemcpyAsync(dev, b->ptr, bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) return err;
  if ((err = cudaEventRecord(b->ev, st)) != cudaSuccess) return err;
  b->busy = true;
  return cudaSuccess;
}

cudaError_t HostStager::download_begin(cudaStream_t st, const void *dev, size_t dev_stride, void *host, size_t host_stride, size_t wbytes,
                                       size_t rows, PendingOut *out) {
  cudaError_t err;
  const size_t bytes = dev_stride * (rows - 1) + wbytes;
  StageBuf *b = acquire(down_, &next_down_, bytes, &err);
  if (!b) return err;
  if ((err = cudaMemcpyAsync(b->ptr, dev, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return err;
  if ((err = cudaEventRecord(b->ev, st)) != cudaSuccess) return err;
  b->busy = true;
  b->pending = true;
  *out = PendingOut{b, host, host_stride, dev_stride, wbytes, rows};
  return cudaSuccess;
}

cudaError_t HostStager::finish(PendingOut &p) {
  cudaError_t err = cudaEventSynchronize(p.sb->ev);
  p.sb->pending = false;
  if (err != cudaSuccess) return err;
  p.sb->busy = false;
  if (p.host_stride == p.dev_stride && p.wbytes == p.dev_stride) {   // (otherwise only the payload bytes of a row are the caller's to lose)
    const size_t bytes = p.dev_stride * (p.rows - 1) + p.wbytes;
    pool()->copy2d(p.host, bytes, p.sb->ptr, bytes, bytes, 1);
  } else {
    pool()->copy2d(p.host, p.host_stride, p.sb->ptr, p.dev_stride, p.wbytes, p.rows);
  }
  return cudaSuccess;
}

}  // namespace pe
