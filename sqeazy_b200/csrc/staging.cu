// Pageable-memory staging for the SQY_* host entry points: see staging.hpp.
#include "staging.hpp"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace sqyb {
namespace {

constexpr size_t kChunk = size_t(32) << 20;   // 0.6 ms of DMA at 54 GB/s: long enough to hide the hand-over, short enough to start early
constexpr int kSlots = 3;
constexpr size_t kMinStaged = size_t(8) << 20;
constexpr int kMaxThreads = 16;
constexpr int kMaxDevices = 16;

// A few persistent threads that copy slices of buffers. Re-entrant: several callers (one per GPU of a sharded call) may
// copy at the same time; every caller copies slice 0 itself and helps with queued slices while it waits for its own.
class CopyPool {
 public:
  explicit CopyPool(int workers) {
    for (int i = 0; i < workers; ++i) th_.emplace_back([this] { run(); });
    for (auto& t : th_) t.detach();   // never joined: they sleep on the condition variable until the process ends
  }
  int workers() const { return (int)th_.size(); }
  void copy(void* dst, const void* src, size_t bytes, int parts) {
    parts = std::max(1, std::min(parts, workers() + 1));
    const size_t slice = ((bytes + parts - 1) / parts + 4095) & ~size_t(4095);
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    std::atomic<int> left{0};
    if (parts > 1) {
      std::lock_guard<std::mutex> lk(m_);
      for (int p = 1; p < parts; ++p) {
        const size_t lo = slice * (size_t)p;
        if (lo >= bytes) break;
        left.fetch_add(1, std::memory_order_relaxed);
        q_.push_back(Task{d + lo, s + lo, std::min(slice, bytes - lo), &left});
      }
    }
    if (left.load(std::memory_order_relaxed) > 0) cv_.notify_all();
    std::memcpy(d, s, std::min(slice, bytes));
    std::unique_lock<std::mutex> lk(m_);
    while (left.load(std::memory_order_acquire) > 0) {
      if (!q_.empty()) {
        Task t = q_.front();
        q_.pop_front();
        lk.unlock();
        exec(t);
        lk.lock();
      } else {
        done_.wait(lk);
      }
    }
  }

 private:
  struct Task {
    char* d;
    const char* s;
    size_t n;
    std::atomic<int>* left;
  };
  void exec(const Task& t) {
    std::memcpy(t.d, t.s, t.n);
    if (t.left->fetch_sub(1, std::memory_order_acq_rel) == 1) {
      std::lock_guard<std::mutex> lk(m_);   // the waiter checks `left` under this lock: no lost wake-up
      done_.notify_all();
    }
  }
  void run() {
    std::unique_lock<std::mutex> lk(m_);
    for (;;) {
      cv_.wait(lk, [&] { return !q_.empty(); });
      Task t = q_.front();
      q_.pop_front();
      lk.unlock();
      exec(t);
      lk.lock();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::deque<Task> q_;
};

// One ring of page-locked chunks per device (its events belong to that device). The state of a slot outlives a call: `busy`
// says that a DMA which reads or writes the slot has been recorded in `ev` and not been waited for yet, so the next user of
// the slot — in this call or in a later one, on this stream or on another — waits for it first. (Round 1 restarted the slot
// bookkeeping with every call: the first chunks of slab k+1 were copied into slots whose DMAs of slab k were still running.)
struct Ring {
  void* slot[kSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[kSlots] = {nullptr, nullptr, nullptr};
  bool busy[kSlots] = {false, false, false};
  size_t next = 0;     // chunks handed out so far: slots rotate across calls
  bool ok = false;
  int init() {
    if (ok) return 0;
    for (int i = 0; i < kSlots; ++i) {
      cudaError_t e = cudaHostAlloc(&slot[i], kChunk, cudaHostAllocPortable);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
      if (e != cudaSuccess) { release(); return (int)e; }
      busy[i] = false;
    }
    ok = true;
    return 0;
  }
  // the host is about to write the slot: every DMA that touches it must be over
  cudaError_t host_acquire(int s) {
    if (!busy[s]) return cudaSuccess;
    busy[s] = false;
    return cudaEventSynchronize(ev[s]);
  }
  // a DMA on `st` is about to write the slot: it must come after the DMAs recorded on other streams
  cudaError_t stream_acquire(int s, cudaStream_t st) {
    if (!busy[s]) return cudaSuccess;
    return cudaStreamWaitEvent(st, ev[s], 0);
  }
  cudaError_t dma_recorded(int s, cudaStream_t st) {
    busy[s] = true;
    return cudaEventRecord(ev[s], st);
  }
  void release() {
    for (int i = 0; i < kSlots; ++i) {
      if (busy[i] && ev[i]) cudaEventSynchronize(ev[i]);
      if (slot[i]) cudaFreeHost(slot[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
      slot[i] = nullptr;
      ev[i] = nullptr;
      busy[i] = false;
    }
    ok = false;
  }
};

Ring g_ring[kMaxDevices];   // callers hold the lock of the device they work on (api.cu)
std::once_flag g_pool_once;
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
  std::call_once(g_pool_once, [] {
    g_pool = new CopyPool(std::min<int>(kMaxThreads, std::max(1u, std::thread::hardware_concurrency())) - 1);
  });
  return *g_pool;
}

Ring* current_ring() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  return &g_ring[dev];
}

}  // namespace

int staging_threads(int nthreads) {
  const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
  int t = (nthreads <= 0 || nthreads > hw) ? hw : nthreads;
  return std::min(t, kMaxThreads);
}

bool host_buffer_is_pageable(const void* p) { return pageable(p); }

int staged_h2d(void* d_dst, const void* h_src, size_t bytes, int nthreads, cudaStream_t st) {
  if (!bytes) return 0;
  const int T = staging_threads(nthreads);
  if (T <= 1 || bytes < kMinStaged || !pageable(h_src))
    return (int)cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
  Ring* R = current_ring();
  if (!R) return (int)cudaErrorInvalidDevice;
  if (int e = R->init()) return e;
  CopyPool& P = pool();
  const char* src = static_cast<const char*>(h_src);
  char* dst = static_cast<char*>(d_dst);
  for (size_t off = 0; off < bytes; off += kChunk) {
    const int s = (int)(R->next++ % kSlots);
    const size_t len = std::min(kChunk, bytes - off);
    if (cudaError_t e = R->host_acquire(s)) return (int)e;   // the DMA that last used this slot, whichever call issued it
    P.copy(R->slot[s], src + off, len, T);
    if (cudaError_t e = cudaMemcpyAsync(dst + off, R->slot[s], len, cudaMemcpyHostToDevice, st)) return (int)e;
    if (cudaError_t e = R->dma_recorded(s, st)) return (int)e;
  }
  return 0;
}

int staged_d2h(void* h_dst, const void* d_src, size_t bytes, int nthreads, cudaStream_t st) {
  if (!bytes) return 0;
  const int T = staging_threads(nthreads);
  if (T <= 1 || bytes < kMinStaged || !pageable(h_dst))
    return (int)cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st);
  Ring* R = current_ring();
  if (!R) return (int)cudaErrorInvalidDevice;
  if (int e = R->init()) return e;
  CopyPool& P = pool();
  char* dst = static_cast<char*>(h_dst);
  const char* src = static_cast<const char*>(d_src);
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  const size_t base = R->next;
  R->next += nchunks;
  auto issue = [&](size_t c) -> cudaError_t {
    const int s = (int)((base + c) % kSlots);
    const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
    if (cudaError_t e = R->stream_acquire(s, st)) return e;   // e.g. an H2D of an earlier call still reading the slot
    if (cudaError_t e = cudaMemcpyAsync(R->slot[s], src + off, len, cudaMemcpyDeviceToHost, st)) return e;
    return R->dma_recorded(s, st);
  };
  // kSlots - 1 DMAs are in flight while the host threads empty the oldest chunk
  for (size_t c = 0; c < std::min(nchunks, (size_t)kSlots - 1); ++c)
    if (cudaError_t e = issue(c)) return (int)e;
  for (size_t c = 0; c < nchunks; ++c) {
    const int s = (int)((base + c) % kSlots);
    const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
    if (c + kSlots - 1 < nchunks)
      if (cudaError_t e = issue(c + kSlots - 1)) return (int)e;   // its slot was emptied in the previous iteration
    if (cudaError_t e = R->host_acquire(s)) return (int)e;        // chunk c has landed
    P.copy(dst + off, R->slot[s], len, T);
  }
  return 0;
}

void staging_release() {
  if (Ring* R = current_ring()) R->release();
}

}  // namespace sqyb
