// Pageable-memory staging for the SQY_* host entry points: see staging.hpp.
#include "staging.hpp"

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace sqyb {
namespace {

constexpr size_t kChunk = size_t(32) << 20;   // 0.6 ms of DMA at 54 GB/s: long enough to hide the hand-over, short enough to start early
constexpr int kSlots = 3;
constexpr size_t kMinStaged = size_t(8) << 20;
constexpr int kMaxThreads = 16;

// a few persistent threads that copy slices of one buffer; the calling thread copies slice 0 itself
class CopyPool {
 public:
  explicit CopyPool(int workers) {
    for (int i = 0; i < workers; ++i) th_.emplace_back([this, i] { run(i); });
    for (auto& t : th_) t.detach();   // never joined: they sleep on the condition variable until the process ends
  }
  int workers() const { return (int)th_.size(); }
  void copy(void* dst, const void* src, size_t bytes, int parts) {
    parts = std::max(1, std::min(parts, workers() + 1));
    const size_t slice = ((bytes + parts - 1) / parts + 4095) & ~size_t(4095);
    if (parts > 1) {
      std::lock_guard<std::mutex> lk(m_);
      dst_ = static_cast<char*>(dst);
      src_ = static_cast<const char*>(src);
      bytes_ = bytes;
      slice_ = slice;
      parts_ = parts;
      pending_ = workers();
      gen_++;
    }
    if (parts > 1) cv_.notify_all();
    std::memcpy(dst, src, std::min(slice, bytes));
    if (parts > 1) {
      std::unique_lock<std::mutex> lk(m_);
      done_.wait(lk, [this] { return pending_ == 0; });
    }
  }

 private:
  void run(int id) {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(m_);
      cv_.wait(lk, [&] { return gen_ != seen; });
      seen = gen_;
      char* d = dst_;
      const char* s = src_;
      const size_t bytes = bytes_, slice = slice_;
      const int parts = parts_;
      lk.unlock();
      const size_t lo = slice * (size_t)(id + 1);
      if (id + 1 < parts && lo < bytes) std::memcpy(d + lo, s + lo, std::min(slice, bytes - lo));
      lk.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  uint64_t gen_ = 0;
  int pending_ = 0, parts_ = 1;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
};

struct Ring {
  void* slot[kSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[kSlots] = {nullptr, nullptr, nullptr};
  bool ok = false;
  int init() {
    if (ok) return 0;
    for (int i = 0; i < kSlots; ++i) {
      cudaError_t e = cudaHostAlloc(&slot[i], kChunk, cudaHostAllocDefault);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
      if (e != cudaSuccess) { release(); return (int)e; }
    }
    ok = true;
    return 0;
  }
  void release() {
    for (int i = 0; i < kSlots; ++i) {
      if (slot[i]) cudaFreeHost(slot[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
      slot[i] = nullptr;
      ev[i] = nullptr;
    }
    ok = false;
  }
};

Ring g_ring;          // callers hold the library lock (api.cu: g_mu)
CopyPool* g_pool = nullptr;

bool pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

CopyPool& pool() {
  if (!g_pool) g_pool = new CopyPool(std::min<int>(kMaxThreads, std::max(1u, std::thread::hardware_concurrency())) - 1);
  return *g_pool;
}

}  // namespace

int staging_threads(int nthreads) {
  const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
  int t = (nthreads <= 0 || nthreads > hw) ? hw : nthreads;
  return std::min(t, kMaxThreads);
}

int staged_h2d(void* d_dst, const void* h_src, size_t bytes, int nthreads, cudaStream_t st) {
  if (!bytes) return 0;
  const int T = staging_threads(nthreads);
  if (T <= 1 || bytes < kMinStaged || !pageable(h_src))
    return (int)cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
  if (int e = g_ring.init()) return e;
  CopyPool& P = pool();
  const char* src = static_cast<const char*>(h_src);
  char* dst = static_cast<char*>(d_dst);
  size_t c = 0;
  for (size_t off = 0; off < bytes; off += kChunk, ++c) {
    const int s = (int)(c % kSlots);
    const size_t len = std::min(kChunk, bytes - off);
    if (c >= (size_t)kSlots) {
      if (cudaError_t e = cudaEventSynchronize(g_ring.ev[s])) return (int)e;   // the DMA that last read this slot
    }
    P.copy(g_ring.slot[s], src + off, len, T);
    if (cudaError_t e = cudaMemcpyAsync(dst + off, g_ring.slot[s], len, cudaMemcpyHostToDevice, st)) return (int)e;
    if (cudaError_t e = cudaEventRecord(g_ring.ev[s], st)) return (int)e;
  }
  return 0;
}

int staged_d2h(void* h_dst, const void* d_src, size_t bytes, int nthreads, cudaStream_t st) {
  if (!bytes) return 0;
  const int T = staging_threads(nthreads);
  if (T <= 1 || bytes < kMinStaged || !pageable(h_dst))
    return (int)cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st);
  if (int e = g_ring.init()) return e;
  CopyPool& P = pool();
  char* dst = static_cast<char*>(h_dst);
  const char* src = static_cast<const char*>(d_src);
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  auto issue = [&](size_t c) -> cudaError_t {
    const int s = (int)(c % kSlots);
    const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
    if (cudaError_t e = cudaMemcpyAsync(g_ring.slot[s], src + off, len, cudaMemcpyDeviceToHost, st)) return e;
    return cudaEventRecord(g_ring.ev[s], st);
  };
  // kSlots - 1 DMAs are in flight while the host threads empty the oldest chunk
  for (size_t c = 0; c < std::min(nchunks, (size_t)kSlots - 1); ++c)
    if (cudaError_t e = issue(c)) return (int)e;
  for (size_t c = 0; c < nchunks; ++c) {
    const int s = (int)(c % kSlots);
    const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
    if (c + kSlots - 1 < nchunks)
      if (cudaError_t e = issue(c + kSlots - 1)) return (int)e;   // its slot was emptied in the previous iteration
    if (cudaError_t e = cudaEventSynchronize(g_ring.ev[s])) return (int)e;
    P.copy(dst + off, g_ring.slot[s], len, T);
  }
  return 0;
}

void staging_release() { g_ring.release(); }

}  // namespace sqyb
