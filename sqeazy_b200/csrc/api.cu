// C-ABI layer: SQY_* (include/sqeazy.h, host buffers, the reference's boundary) and sqyx_*
// (include/sqeazy_b200.h, device buffers). Sequences the sm_100a kernels the way
// dynamic_pipeline::detail_encode / detail_decode sequence the reference's stages
// (dynamic_pipeline.hpp:619-690, 772-846). No CPU compute path exists here: without a CUDA device
// every compute entry point fails with 1.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost nothing unless a profiler injects its library

#include <algorithm>
#include <atomic>
#include <climits>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sqeazy.h"
#include "../../include/sqeazy_b200.h"
#include "device/kernels.h"
#include "device/lz4_format.h"
#include "host/numerics.hpp"
#include "host/pipeline.hpp"
#include "host/text.hpp"
#include "staging.hpp"

using namespace sqyb;

namespace sqyb {
std::atomic<long> g_kernel_launches{0};
}

namespace {

constexpr int kVersionMajor = 0, kVersionMinor = 7, kVersionPatch = 2;

#define CK(expr)                                                                                          \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      std::fprintf(stderr, "[sqeazy_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                                           \
    }                                                                                                     \
  } while (0)

#define CKK(expr)                                                                             \
  do {                                                                                        \
    int _r = (expr);                                                                          \
    if (_r != 0) {                                                                            \
      std::fprintf(stderr, "[sqeazy_b200] kernel launch failed (%d) at %s:%d\n", _r, __FILE__, __LINE__); \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

// ---- cached device scratch, one arena per device, guarded by one lock (calls are serialised) ----
enum Slot { kSlotA = 0, kSlotB, kSlotIn, kSlotOut, kSlotWs, kSlotSmall, kSlotOrigins, kNumSlots };

struct Arena {
  void* ptr[kNumSlots] = {nullptr};
  size_t cap[kNumSlots] = {0};
  long last_stats[4] = {0, 0, 0, 0};
  int get(int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 256;
    if (cap[slot] < bytes) {
      if (ptr[slot]) cudaFree(ptr[slot]);
      ptr[slot] = nullptr;
      cap[slot] = 0;
      const size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
      cudaError_t e = cudaMalloc(&ptr[slot], want);
      if (e != cudaSuccess) {
        std::fprintf(stderr, "[sqeazy_b200] cudaMalloc(%zu) failed: %s\n", want, cudaGetErrorString(e));
        return 1;
      }
      cap[slot] = want;
    }
    *out = ptr[slot];
    return 0;
  }
  void release() {
    for (int i = 0; i < kNumSlots; ++i) {
      if (ptr[i]) cudaFree(ptr[i]);
      ptr[i] = nullptr;
      cap[i] = 0;
    }
  }
};

// One context per device: scratch arena, copy stream, batch lanes, and the lock that serialises the calls that use this
// device. Calls on different devices run concurrently (round 1 had one process-wide lock).
// batch entry points (sqyx_*_batch_device_*): every lane of a batch has its own stream and scratch, so the kernels of
// different stacks share the SMs instead of queueing behind each other's tails
constexpr int kBatchLanes = 8;
constexpr int kMaxDevices = 16;
struct Dev {
  std::mutex mu;
  Arena arena;
  cudaStream_t copy_stream = nullptr;
  Arena batch_arena[kBatchLanes];
  cudaStream_t batch_stream[kBatchLanes] = {};
};
Dev g_dev[kMaxDevices];
std::mutex g_timer_mu;   // g_stage_ms

// ---- optional per-stage CUDA-event timing (sqyx_enable_stage_timing) ----
enum StageTimer { kTFilterSwap = 0, kTLz4Enc, kTHist, kTLutApply, kTLz4Dec, kTLutDec, kTSwapDec, kNumTimers };
std::atomic<int> g_timing{0};
float g_stage_ms[kNumTimers] = {0};

// every stage of a call is also an NVTX range (the reference's tracing is wall-clock timers around the stage
// calls, verbs/bench.hpp:181-194): `nsys` / `ncu --nvtx` show filter+transpose, LZ4, histogram, LUT per call
const char* const kStageNames[kNumTimers] = {"sqy:filter+bitswap encode", "sqy:lz4 encode", "sqy:histogram", "sqy:lut apply",
                                             "sqy:lz4 decode", "sqy:lut decode", "sqy:bitswap decode"};
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct ScopedStageTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  int slot;
  bool on;
  NvtxRange range;
  ScopedStageTimer(int slot_, cudaStream_t st_) : st(st_), slot(slot_), on(g_timing.load() != 0), range(kStageNames[slot_]) {
    if (on) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, st);
    }
  }
  void stop() {
    if (on && a) {
      cudaEventRecord(b, st);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      {
        std::lock_guard<std::mutex> lk(g_timer_mu);
        g_stage_ms[slot] += ms;
      }
      cudaEventDestroy(a);
      cudaEventDestroy(b);
      a = b = nullptr;
    }
  }
  ~ScopedStageTimer() { stop(); }
};

std::mutex g_set_mu;
std::vector<int> g_set;          // set by sqyx_set_devices / sqyx_set_device; empty = not pinned by the caller

// the device a single-device call works on: SQY_CUDA_DEVICE, else a device set of one (sqyx_set_devices), else the calling
// thread's current device
int pick_device(int* out) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    std::fprintf(stderr, "[sqeazy_b200] no CUDA device available (this build has no CPU path)\n");
    return 1;
  }
  int dev = 0;
  if (const char* env = std::getenv("SQY_CUDA_DEVICE")) {
    dev = std::atoi(env);
    if (dev < 0 || dev >= n) return 1;
    if (cudaSetDevice(dev) != cudaSuccess) return 1;
  } else {
    int pinned = -1;
    {
      std::lock_guard<std::mutex> lk(g_set_mu);
      if (g_set.size() == 1) pinned = g_set[0];
    }
    if (pinned >= 0) {
      dev = pinned;
      if (dev >= n || cudaSetDevice(dev) != cudaSuccess) return 1;
    } else if (cudaGetDevice(&dev) != cudaSuccess) {
      return 1;
    }
  }
  if (dev >= kMaxDevices) return 1;
  *out = dev;
  return 0;
}

// picks the call's device and holds its lock for the lifetime of the object
struct DevLock {
  Dev* dev = nullptr;
  int id = -1;
  std::unique_lock<std::mutex> lk;
  int acquire() {
    if (pick_device(&id)) return 1;
    dev = &g_dev[id];
    lk = std::unique_lock<std::mutex>(dev->mu);
    return 0;
  }
  int acquire(int device) {   // a given device; made current for the calling thread
    if (device < 0 || device >= kMaxDevices || cudaSetDevice(device) != cudaSuccess) return 1;
    id = device;
    dev = &g_dev[id];
    lk = std::unique_lock<std::mutex>(dev->mu);
    return 0;
  }
};

// LZ4 pitch hint (kernels.h): bytes between vertically adjacent voxels in the stream the `lz4` stage sees — a function of
// pipeline and shape only, so every route (device, streamed, sharded) writes the same bytes
uint32_t lz4_pitch_hint(const Pipeline& pl, const std::vector<uint64_t>& shape) {
  if (shape.size() < 2 || !pl.has_sink) return 0;
  const uint64_t X = shape.back();
  uint64_t pitch = 0;
  if (pl.sink.kind == StageKind::Quantiser && pl.has_tail && pl.head.empty()) pitch = X;                 // rows of 8-bit codes
  else if (pl.sink.kind == StageKind::Lz4 && !pl.head.empty() && pl.head.back().kind == StageKind::Bitswap) {
    const uint64_t bits = X * (uint64_t)pl.head.back().w;                                                 // bits of one row in one plane
    if (bits % 8 == 0) pitch = bits / 8;
  } else if (pl.sink.kind == StageKind::Lz4 && pl.head.empty()) pitch = X * (uint64_t)pl.elem;           // raw voxels
  uint32_t hint = pitch <= 8192 ? (uint32_t)pitch : 0u;
  // bit planes behind a background removal hold no noise planes: every warp of the encoder starts at once (kernels.h)
  if (pl.sink.kind == StageKind::Lz4)
    for (const Stage& s : pl.head)
      if (s.kind == StageKind::RemoveBackground || s.kind == StageKind::RmEstBkrd) hint |= kLz4HintNoNoise;
  return hint;
}

uint64_t shape_product(const std::vector<uint64_t>& shape) {
  uint64_t n = 1;
  for (uint64_t d : shape) n *= d;
  return shape.empty() ? 0 : n;
}

// ---- rmestbkrd threshold: GPU histograms of the sampled faces/rows + host support formula ----
// what the estimate reads, as device pointers: the first `portion` voxels of the z = 0 and z = Z-1 faces and the rows y = 0 /
// y = Y-1 at z in {1, Z/2, Z-2} (nullptr: that z does not exist). The callers whose stack is not resident as a whole (streamed
// and sharded host paths) send just these pieces.
struct EstimateSamples {
  const uint16_t* face[2];
  const uint16_t* row[2][3];
  uint64_t portion, X;
};

EstimateSamples estimate_samples_of(const uint16_t* d_src, uint64_t Z, uint64_t Y, uint64_t X, size_t l2) {
  EstimateSamples S;
  const uint64_t frame = Y * X;
  S.portion = rmest_frame_portion(frame, l2);
  S.X = X;
  S.face[0] = d_src;
  S.face[1] = d_src + (Z - 1) * frame;
  // rows y = 0 and y = Y-1 at z in {1, Z/2, Z-2}, background_scheme_utils.hpp:79-103 (z indices are used as given, like the reference)
  const uint64_t zs[3] = {1, Z / 2, Z - 2};
  const uint64_t ys[2] = {0, Y - 1};
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 3; ++k)   // Z < 3: the reference would read out of bounds; we skip those rows
      S.row[i][k] = zs[k] >= Z ? nullptr : d_src + zs[k] * frame + ys[i] * X;
  return S;
}

int estimate_background_samples(Arena& A, const EstimateSamples& S, float* supports, int* threshold, cudaStream_t st) {
  void* p = nullptr;
  if (A.get(kSlotSmall, 4 * 65536 * sizeof(uint32_t) + 4096, &p)) return 1;
  uint32_t* d_h = static_cast<uint32_t*>(p);
  CK(cudaMemsetAsync(d_h, 0, 4 * 65536 * sizeof(uint32_t), st));
  // z = 0 and z = Z-1 faces (first `portion` elements), background_scheme_utils.hpp:57-77
  CKK(k_histogram_u16(S.face[0], S.portion, d_h, st));
  CKK(k_histogram_u16(S.face[1], S.portion, d_h + 65536, st));
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 3; ++k)
      if (S.row[i][k]) CKK(k_histogram_u16(S.row[i][k], S.X, d_h + (2 + i) * 65536, st));
  // support index of the four histograms on the device; 64 bytes come back instead of 1 MiB of bins
  uint32_t* d_idx = d_h + 4 * 65536;
  CKK(k_support_index(d_h, 4, 0.99f, d_idx, st));
  uint32_t idx[16];
  CK(cudaMemcpyAsync(idx, d_idx, sizeof(idx), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  float mn = 0.f;
  for (int i = 0; i < 4; ++i) {
    supports[i] = support_from_index(idx[4 * i], idx[4 * i + 1], idx[4 * i + 2]);
    if (i == 0 || supports[i] < mn) mn = supports[i];
  }
  *threshold = (int)(uint16_t)mn;  // remove_background_scheme(raw_type) ctor truncates
  return 0;
}

int estimate_background(Arena& A, const uint16_t* d_src, uint64_t Z, uint64_t Y, uint64_t X, long l2_bytes, float* supports,
                        int* threshold, cudaStream_t st) {
  if (Z == 0 || Y == 0 || X == 0) return 1;
  const size_t l2 = l2_bytes >= 0 ? (size_t)l2_bytes : host_l2_cache_bytes();
  return estimate_background_samples(A, estimate_samples_of(d_src, Z, Y, X, l2), supports, threshold, st);
}

// the same from a stack in HOST memory: only the sampled pieces cross the bus, packed into `d_scratch` (>= 4*portion + 12*X bytes)
int estimate_background_host(Arena& A, const uint16_t* h_src, uint64_t Z, uint64_t Y, uint64_t X, uint16_t* d_scratch, int* threshold,
                             cudaStream_t st) {
  if (Z == 0 || Y == 0 || X == 0) return 1;
  const EstimateSamples H = estimate_samples_of(h_src, Z, Y, X, host_l2_cache_bytes());
  EstimateSamples D = H;
  uint16_t* q = d_scratch;
  for (int f = 0; f < 2; ++f) {
    CK(cudaMemcpyAsync(q, H.face[f], 2 * H.portion, cudaMemcpyHostToDevice, st));
    D.face[f] = q;
    q += H.portion;
  }
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 3; ++k) {
      if (!H.row[i][k]) continue;
      CK(cudaMemcpyAsync(q, H.row[i][k], 2 * X, cudaMemcpyHostToDevice, st));
      D.row[i][k] = q;
      q += X;
    }
  float sup[4];
  return estimate_background_samples(A, D, sup, threshold, st);
}

// ---- encode ----
int encode_device_impl(Arena& A, const Pipeline& pl_in, const void* d_src_any, const std::vector<uint64_t>& shape, uint8_t* d_dst,
                       uint64_t dst_cap, uint64_t* dst_bytes, const uint32_t* d_global_hist, cudaStream_t st) {
  Pipeline pl = pl_in;
  const uint64_t N = shape_product(shape);
  const bool u8 = pl.elem == 1;                       // dypeline<uint8_t>: bitswap / remove_background / lz4 / pass_through only
  const uint64_t raw_bytes = (uint64_t)pl.elem * N;
  const uint16_t* d_src = static_cast<const uint16_t*>(d_src_any);   // (typed per stage below)
  auto b8 = [](const uint16_t* q) { return reinterpret_cast<const uint8_t*>(q); };
  auto m8 = [](uint16_t* q) { return reinterpret_cast<uint8_t*>(q); };
  const size_t reserve = header_reserve_bytes(pl, shape);
  uint8_t* payload = d_dst + reserve;
  if (dst_cap < reserve) return 1;
  const uint64_t payload_cap = dst_cap - reserve;

  const uint16_t* cur = d_src;
  uint16_t* bufs[2] = {nullptr, nullptr};
  int next = 0;
  auto next_buf = [&](uint16_t** out) -> int {
    if (!bufs[next]) {
      void* p = nullptr;
      if (A.get(next == 0 ? kSlotA : kSlotB, raw_bytes, &p)) return 1;
      bufs[next] = static_cast<uint16_t*>(p);
    }
    *out = bufs[next];
    next ^= 1;
    return 0;
  };

  for (size_t i = 0; i < pl.head.size(); ++i) {
    const Stage& s = pl.head[i];
    uint16_t* out = nullptr;
    ScopedStageTimer tm(kTFilterSwap, st);
    if (s.kind == StageKind::RemoveBackground || s.kind == StageKind::RmEstBkrd) {
      int t = s.threshold;
      if (s.kind == StageKind::RmEstBkrd) {
        if (shape.size() != 3) {
          std::fprintf(stderr, "[sqeazy_b200] rmestbkrd needs a rank-3 shape\n");
          return 1;
        }
        float sup[4];
        if (estimate_background(A, cur, shape[0], shape[1], shape[2], -1, sup, &t, st)) return 1;
      }
      if (next_buf(&out)) return 1;
      if (i + 1 < pl.head.size() && pl.head[i + 1].kind == StageKind::Bitswap && t > 0) {
        // filter fused into the transpose
        if (u8) CKK(k_bitswap8_encode(pl.head[i + 1].w, b8(cur), m8(out), N, t, st));
        else CKK(k_bitswap_encode(pl.head[i + 1].w, cur, out, N, t, st));
        ++i;
      } else if (u8) {
        CKK(k_remove_background8(b8(cur), m8(out), N, t, st));
      } else {
        CKK(k_remove_background(cur, out, N, t, st));
      }
    } else if (s.kind == StageKind::Bitshuffle) {
      if (next_buf(&out)) return 1;
      if (u8 ? k_bitshuffle8_encode(b8(cur), m8(out), N, s.block_size, st) : k_bitshuffle16_encode(cur, out, N, s.block_size, st)) {
        std::fprintf(stderr, "[sqeazy_b200] bitshuffle failed (block_size=%u must be a multiple of 8)\n", s.block_size);
        return 1;
      }
    } else if (s.kind == StageKind::Diff) {
      if (next_buf(&out)) return 1;
      if (shape.size() != 3 || k_diff_encode(pl.elem, cur, out, shape[0], shape[1], shape[2], st)) {
        // diff_scheme_impl.hpp:84-87 needs rank 3; shapes on which the reference's own loops leave the plane are refused
        std::fprintf(stderr, "[sqeazy_b200] diff3x3x1: shape not supported\n");
        return 1;
      }
    } else {  // Bitswap
      if (next_buf(&out)) return 1;
      if (u8) CKK(k_bitswap8_encode(s.w, b8(cur), m8(out), N, 0, st));
      else CKK(k_bitswap_encode(s.w, cur, out, N, 0, st));
    }
    cur = out;
  }

  uint64_t payload_bytes = 0;
  auto lz4_into_payload = [&](const uint8_t* src, uint64_t nbytes) -> int {
    if (payload_cap < lz4_payload_bound(nbytes)) {
      std::fprintf(stderr, "[sqeazy_b200] destination too small for lz4 (%llu < %llu)\n", (unsigned long long)payload_cap,
                   (unsigned long long)lz4_payload_bound(nbytes));
      return 1;
    }
    void* ws = nullptr;
    if (A.get(kSlotWs, k_lz4_encode_workspace_bytes(nbytes), &ws)) return 1;
    {
      ScopedStageTimer tm(kTLz4Enc, st);
      CKK(k_lz4_encode(src, nbytes, payload, ws, lz4_pitch_hint(pl, shape), st));
    }
    unsigned long long hres[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(hres, ws, 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    payload_bytes = hres[0];
    const uint32_t* stats = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(hres) + 16);
    A.last_stats[0] = stats[0]; A.last_stats[1] = stats[1]; A.last_stats[2] = stats[2]; A.last_stats[3] = (long)payload_bytes;
    return 0;
  };

  if (!pl.has_sink) {
    if (payload_cap < raw_bytes) return 1;
    if (raw_bytes) CK(cudaMemcpyAsync(payload, cur, raw_bytes, cudaMemcpyDeviceToDevice, st));
    payload_bytes = raw_bytes;
  } else if (pl.sink.kind == StageKind::Lz4) {
    if (lz4_into_payload(reinterpret_cast<const uint8_t*>(cur), raw_bytes)) return 1;
  } else if (pl.sink.kind == StageKind::PassThrough) {
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(cur);
    if (pl.has_tail) {
      if (lz4_into_payload(bytes, raw_bytes)) return 1;
    } else {
      if (payload_cap < raw_bytes) return 1;
      if (raw_bytes) CK(cudaMemcpyAsync(payload, bytes, raw_bytes, cudaMemcpyDeviceToDevice, st));
      payload_bytes = raw_bytes;
    }
  } else {  // Quantiser
    if (u8) return 1;   // (the planner never builds it for uint8)
    void* sp = nullptr;
    if (A.get(kSlotSmall, 4 * 65536 * sizeof(uint32_t) + 4096, &sp)) return 1;
    uint32_t* d_hist = static_cast<uint32_t*>(sp);
    uint8_t* d_lut = reinterpret_cast<uint8_t*>(d_hist + 65536);
    std::vector<uint32_t> h(65536);
    if (N > 0 || d_global_hist) {
      if (d_global_hist) {
        CK(cudaMemcpyAsync(h.data(), d_global_hist, 65536 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      } else {
        CK(cudaMemsetAsync(d_hist, 0, 65536 * sizeof(uint32_t), st));
        ScopedStageTimer tm(kTHist, st);
        CKK(k_histogram_u16(cur, N, d_hist, st));
        tm.stop();
        CK(cudaMemcpyAsync(h.data(), d_hist, 65536 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      }
      CK(cudaStreamSynchronize(st));
    }
    std::vector<uint8_t> enc(65536);
    uint16_t dec[256];
    quantiser_luts_from_histogram(h.data(), enc.data(), dec);
    // the decode LUT travels in the header (quantiser_scheme_impl.hpp:200-204)
    pl.sink.kv["decode_lut_string"] = std::string(kVerbatimOpen) + base64_encode(dec, sizeof(dec)) + kVerbatimClose;
    CK(cudaMemcpyAsync(d_lut, enc.data(), 65536, cudaMemcpyHostToDevice, st));
    if (pl.has_tail) {
      uint16_t* codes16 = nullptr;
      if (next_buf(&codes16)) return 1;
      uint8_t* codes = reinterpret_cast<uint8_t*>(codes16);
      {
        ScopedStageTimer tm(kTLutApply, st);
        CKK(k_lut_apply(cur, codes, N, d_lut, st));
      }
      if (lz4_into_payload(codes, N)) return 1;
    } else {
      if (payload_cap < N) return 1;
      CKK(k_lut_apply(cur, payload, N, d_lut, st));
      CK(cudaStreamSynchronize(st));  // enc (host vector) must outlive the H2D copy
      payload_bytes = N;
    }
  }

  // header, right-aligned in its slot (leading blanks are what header::pack itself pads with)
  const std::string hdr = pack_header(pl.type_name(), pl.elem, shape, pl.canonical(), payload_bytes);
  if (hdr.size() > reserve) {
    std::fprintf(stderr, "[sqeazy_b200] header (%zu B) exceeds its slot (%zu B)\n", hdr.size(), reserve);
    return 1;
  }
  std::string slot(reserve - hdr.size(), ' ');
  slot += hdr;
  CK(cudaMemcpyAsync(d_dst, slot.data(), slot.size(), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  *dst_bytes = reserve + payload_bytes;
  return 0;
}

// ---- shared by the streamed host paths (host_encode_streamed below, the host sink of decode_device_impl) ----
constexpr uint64_t kStreamSlabBytes = uint64_t(256) << 20;
constexpr uint64_t kStreamMinBytes = uint64_t(512) << 20;
constexpr uint64_t kStreamGrainVoxels = 131072;   // a 16 KiB block of a 1-bit plane covers 131072 voxels (8192 * P for wider atoms)
int copy_stream(cudaStream_t* cs) {   // of the current device; its lock is held
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev >= kMaxDevices) return 1;
  if (!g_dev[dev].copy_stream) CK(cudaStreamCreateWithFlags(&g_dev[dev].copy_stream, cudaStreamNonBlocking));
  *cs = g_dev[dev].copy_stream;
  return 0;
}

struct EventList {
  std::vector<cudaEvent_t> ev;
  ~EventList() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
  int next(cudaEvent_t* out) {
    cudaEvent_t e = nullptr;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ev.push_back(e);
    *out = e;
    return 0;
  }
};

// where decode_device_impl may deliver the volume slab by slab while the last inverse transpose is still running
struct HostSink {
  char* dst;
  int nthreads;
  bool done = false;
};

// ---- decode ----
int lz4_decode_checked(Arena& A, const uint8_t* src, uint64_t nbytes, uint8_t* dst, uint64_t dst_bytes, uint64_t* decoded,
                       cudaStream_t st) {
  void* ws = nullptr;
  if (A.get(kSlotWs, k_lz4_decode_workspace_bytes(dst_bytes), &ws)) return 1;
  for (int attempt = 0; attempt < 2; ++attempt) {
    {
      ScopedStageTimer tm(kTLz4Dec, st);
      CKK(k_lz4_decode(src, nbytes, dst, dst_bytes, ws, attempt, 1, st));
    }
    uint32_t err = 0, deferred = 0;
    uint64_t total = 0;
    if (k_lz4_decode_status(ws, &err, &total, &deferred, st)) return 1;
    if (err == 0 && deferred) {
      // block-linked frames (reference serial mode / sqy CLI default): deferred cross-block references, lz4_decode.cu
      void* origins = nullptr;
      if (A.get(kSlotOrigins, k_lz4_decode_linked_workspace_bytes(dst_bytes), &origins)) return 1;
      {
        ScopedStageTimer tm(kTLz4Dec, st);
        CKK(k_lz4_decode_linked(src, nbytes, dst, dst_bytes, ws, origins, st));
      }
      uint32_t none = 0;
      if (k_lz4_decode_status(ws, &err, &total, &none, st)) return 1;
    }
    if (err == 0) {
      if (decoded) *decoded = total;
      return 0;
    }
    // 4 = bad block, 5 = size mismatch: foreign frames with short inner blocks -> measure every block and retry
    if (!(attempt == 0 && (err == 4 || err == 5))) {
      std::fprintf(stderr, "[sqeazy_b200] lz4 decode failed (code %u)\n", err);
      return 1;
    }
  }
  return 1;
}

// Host decode of `[filters ->] bitswapN -> lz4` blobs of this library, >= 512 MiB: the LZ4 blocks are decoded slab by slab (the
// blocks of all 16/N bit planes that make up a z-slab: one block-set launch), each slab is transposed back and leaves for the
// host while the next one is decoded; only the first slab's kernels run in front of the bus. Returns -1 when the blob is not
// of that kind (the caller takes the general route), 0 on success, an error code otherwise.
int decode_streamed(Arena& A, const Pipeline& pl, const uint8_t* d_payload, uint64_t payload_bytes, uint16_t* d_dst, uint64_t N,
                    cudaStream_t st, HostSink* host, uint64_t slab_voxels = kStreamSlabBytes / 2, uint64_t min_bytes = kStreamMinBytes) {
  const uint64_t raw_bytes = 2 * N;
  if (pl.elem != 2 || !pl.has_sink || pl.sink.kind != StageKind::Lz4 || pl.has_tail || g_timing.load()) return -1;
  if (raw_bytes < min_bytes || N % kStreamGrainVoxels) return -1;
  int w = 0, nswaps = 0;
  for (const Stage& s : pl.head) {
    if (s.kind == StageKind::Bitswap) { w = s.w; nswaps++; }
    else if (s.kind == StageKind::Bitshuffle || s.kind == StageKind::Diff) return -1;
  }
  if (nswaps != 1) return -1;
  const int P = 16 / w;
  void *ws = nullptr, *planes_v = nullptr;
  if (A.get(kSlotWs, k_lz4_decode_workspace_bytes(raw_bytes), &ws) || A.get(kSlotA, raw_bytes, &planes_v)) return 1;
  uint8_t* planes = static_cast<uint8_t*>(planes_v);
  if (((((uintptr_t)planes) | ((uintptr_t)d_dst)) & 31) != 0) return -1;
  CKK(k_lz4_decode_tables(d_payload, payload_bytes, raw_bytes, ws, 0, 0, st));
  uint32_t err = 0, own = 0, nblk = 0, bb = 0;
  if (k_lz4_decode_peek(ws, &err, &own, &nblk, &bb, st)) return 1;
  if (err || !own || bb != kLz4BlockBytes || (uint64_t)nblk * bb != raw_bytes) return -1;
  cudaStream_t cs = nullptr;
  if (copy_stream(&cs)) return 1;
  // at most 32 slabs (one work counter each)
  uint64_t slab = slab_voxels;
  while ((N + slab - 1) / slab > 32) slab *= 2;
  const uint32_t blocks_per_plane = (uint32_t)(raw_bytes / P / kLz4BlockBytes);
  EventList events;
  uint32_t slot = 0;
  for (uint64_t first = 0; first < N; first += slab, ++slot) {
    const uint64_t count = std::min(slab, N - first);
    CKK(k_lz4_decode_run(d_payload, payload_bytes, planes, raw_bytes, ws, (uint32_t)(2 * first / P / kLz4BlockBytes),
                         (uint32_t)(2 * count / P / kLz4BlockBytes), (uint32_t)P, blocks_per_plane, slot, st));
    if (k_bitswap_decode_range(w, reinterpret_cast<const uint16_t*>(planes), d_dst, N, first, count, st)) return 101;
    cudaEvent_t e = nullptr;
    if (events.next(&e)) return 1;
    CK(cudaEventRecord(e, st));
  }
  size_t i = 0;
  for (uint64_t first = 0; first < N; first += slab, ++i) {
    CK(cudaStreamWaitEvent(cs, events.ev[i], 0));
    CK((cudaError_t)staged_d2h(host->dst + 2 * first, d_dst + first, 2 * std::min(slab, N - first), host->nthreads, cs));
  }
  uint64_t total = 0;
  uint32_t none = 0;
  if (k_lz4_decode_status(ws, &err, &total, &none, st)) return 1;
  CK(cudaStreamSynchronize(cs));
  if (err || total != raw_bytes) {
    std::fprintf(stderr, "[sqeazy_b200] lz4 decode failed (code %u)\n", err);
    return 11;
  }
  host->done = true;
  return 0;
}

int decode_streamed_codes(Arena& A, const Pipeline& pl, const uint8_t* d_payload, uint64_t payload_bytes, uint16_t* d_dst, uint64_t N,
                          cudaStream_t st, HostSink* host, uint64_t slab_voxels, uint64_t min_bytes);   // sharded.inl

int decode_device_impl(Arena& A, const Header& hdr, const Pipeline& pl, const uint8_t* d_payload, uint64_t payload_bytes,
                       void* d_dst_any, uint64_t dst_cap, cudaStream_t st, HostSink* host = nullptr) {
  const uint64_t N = shape_product(hdr.shape);
  const bool u8 = pl.elem == 1;
  const uint64_t raw_bytes = (uint64_t)pl.elem * N;
  uint16_t* d_dst = static_cast<uint16_t*>(d_dst_any);
  if (dst_cap < raw_bytes) return 1;
  if (N == 0) return 0;
  if (host) {
    int rc = decode_streamed(A, pl, d_payload, payload_bytes, d_dst, N, st, host);
    if (rc >= 0) return rc;
    rc = decode_streamed_codes(A, pl, d_payload, payload_bytes, d_dst, N, st, host, kStreamSlabBytes / 2, kStreamMinBytes);
    if (rc >= 0) return rc;
  }

  // data-moving head stages in decode order (reverse); the background filters decode as identity
  std::vector<int> swaps;                 // bitswap: bits per plane; bitshuffle: -1 - block_size; diff3x3x1: kDiffOp
  constexpr int kDiffOp = INT_MIN;
  for (size_t i = pl.head.size(); i-- > 0;) {
    if (pl.head[i].kind == StageKind::Bitswap) swaps.push_back(pl.head[i].w);
    if (pl.head[i].kind == StageKind::Diff) {
      if (hdr.shape.size() != 3 || !diff_shape_supported(hdr.shape[0], hdr.shape[1], hdr.shape[2], pl.elem)) return 100 + 1;
      swaps.push_back(kDiffOp);
    }
    if (pl.head[i].kind == StageKind::Bitshuffle) swaps.push_back(-1 - (int)pl.head[i].block_size);
  }

  uint16_t* bufs[2] = {nullptr, nullptr};
  auto scratch = [&](int which, uint16_t** out) -> int {
    if (!bufs[which]) {
      void* p = nullptr;
      if (A.get(which == 0 ? kSlotA : kSlotB, raw_bytes, &p)) return 1;
      bufs[which] = static_cast<uint16_t*>(p);
    }
    *out = bufs[which];
    return 0;
  };
  int toggle = 0;
  // destination of the op after which `remaining` data-moving ops still follow
  auto out_for = [&](size_t remaining, uint16_t** out) -> int {
    if (remaining == 0) { *out = d_dst; return 0; }
    const int r = scratch(toggle, out);
    toggle ^= 1;
    return r;
  };

  const uint16_t* cur = nullptr;
  size_t remaining = swaps.size();
  if (pl.has_sink) {
    uint16_t* out = nullptr;
    if (out_for(remaining, &out)) return 1;
    if (pl.sink.kind == StageKind::Lz4) {
      uint64_t got = 0;
      if (lz4_decode_checked(A, d_payload, payload_bytes, reinterpret_cast<uint8_t*>(out), raw_bytes, &got, st)) return 11;
    } else if (pl.sink.kind == StageKind::PassThrough) {
      if (pl.has_tail) {
        uint64_t got = 0;
        if (lz4_decode_checked(A, d_payload, payload_bytes, reinterpret_cast<uint8_t*>(out), raw_bytes, &got, st)) return 101;
      } else {
        if (payload_bytes < raw_bytes) return 11;
        CK(cudaMemcpyAsync(out, d_payload, raw_bytes, cudaMemcpyDeviceToDevice, st));
      }
    } else {  // Quantiser
      if (u8) return 11;
      if (!pl.sink.has_decode_lut) {
        std::fprintf(stderr, "[sqeazy_b200] quantiser stage without decode_lut_string in the header\n");
        return 11;
      }
      const uint8_t* codes = d_payload;
      if (pl.has_tail) {
        uint16_t* cbuf = nullptr;
        // codes (N bytes) live in the scratch buffer that is NOT the LUT output
        if (scratch(out == bufs[0] ? 1 : 0, &cbuf)) return 1;
        uint64_t got = 0;
        if (lz4_decode_checked(A, d_payload, payload_bytes, reinterpret_cast<uint8_t*>(cbuf), N, &got, st)) return 101;
        codes = reinterpret_cast<const uint8_t*>(cbuf);
      } else if (payload_bytes < N) {
        return 11;
      }
      void* sp = nullptr;
      if (A.get(kSlotSmall, 4 * 65536 * sizeof(uint32_t) + 4096, &sp)) return 1;
      uint16_t* d_lut = reinterpret_cast<uint16_t*>(static_cast<uint8_t*>(sp) + 4 * 65536 * sizeof(uint32_t));
      CK(cudaMemcpyAsync(d_lut, pl.sink.decode_lut, 512, cudaMemcpyHostToDevice, st));
      {
        ScopedStageTimer tm(kTLutDec, st);
        CKK(k_lut_decode(codes, out, N, d_lut, st));
      }
      CK(cudaStreamSynchronize(st));
    }
    cur = out;
  } else {
    if (payload_bytes < raw_bytes) return 1;
    cur = reinterpret_cast<const uint16_t*>(d_payload);
    if (remaining == 0) {
      CK(cudaMemcpyAsync(d_dst, cur, raw_bytes, cudaMemcpyDeviceToDevice, st));
      cur = d_dst;
    }
  }
  for (size_t k = 0; k < swaps.size(); ++k) {
    uint16_t* out = nullptr;
    remaining = swaps.size() - 1 - k;
    if (remaining == 0) out = d_dst;
    else {
      // pick the scratch buffer that does not hold `cur`
      if (scratch(cur == bufs[0] ? 1 : 0, &out)) return 1;
    }
    if (swaps[k] == kDiffOp) {
      if (k_diff_decode(pl.elem, cur, out, hdr.shape[0], hdr.shape[1], hdr.shape[2], st)) return 100 + 1;
      cur = out;
      continue;
    }
    if (swaps[k] < 0) {
      const uint32_t bsz = (uint32_t)(-1 - swaps[k]);
      if (u8 ? k_bitshuffle8_decode(reinterpret_cast<const uint8_t*>(cur), reinterpret_cast<uint8_t*>(out), N, bsz, st)
             : k_bitshuffle16_decode(cur, out, N, bsz, st))
        return 100 + 1;
      cur = out;
      continue;
    }
    if (remaining == 0 && host && !u8 && raw_bytes >= kStreamMinBytes && N % 128 == 0 && !g_timing.load() &&
        ((((uintptr_t)cur) | ((uintptr_t)out)) & 31) == 0) {
      // last stage of a host call: transpose back slab by slab, every slab leaves for the host as soon as it is done
      cudaStream_t cs = nullptr;
      if (copy_stream(&cs)) return 1;
      EventList events;
      const uint64_t slab = kStreamSlabBytes / 2;
      for (uint64_t first = 0; first < N; first += slab) {
        if (k_bitswap_decode_range(swaps[k], cur, out, N, first, std::min(slab, N - first), st)) return 100 + 1;
        cudaEvent_t e = nullptr;
        if (events.next(&e)) return 1;
        CK(cudaEventRecord(e, st));
      }
      size_t i = 0;
      for (uint64_t first = 0; first < N; first += slab, ++i) {
        CK(cudaStreamWaitEvent(cs, events.ev[i], 0));
        CK((cudaError_t)staged_d2h(host->dst + 2 * first, out + first, 2 * std::min(slab, N - first), host->nthreads, cs));
      }
      CK(cudaStreamSynchronize(cs));
      host->done = true;
      cur = out;
      continue;
    }
    {
      ScopedStageTimer tm(kTSwapDec, st);
      if (u8 ? k_bitswap8_decode(swaps[k], reinterpret_cast<const uint8_t*>(cur), reinterpret_cast<uint8_t*>(out), N, st)
             : k_bitswap_decode(swaps[k], cur, out, N, st))
        return 100 + 1;
    }
    cur = out;
  }
  CK(cudaStreamSynchronize(st));
  return 0;
}

int parse_blob_header_device(const uint8_t* d_blob, uint64_t blob_bytes, Header& hdr, cudaStream_t st) {
  size_t peek = 16384;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const size_t n = blob_bytes < peek ? (size_t)blob_bytes : peek;
    std::vector<char> h(n);
    CK(cudaMemcpyAsync(h.data(), d_blob, n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    hdr = unpack_header(h.data(), n);
    if (hdr.valid || n == blob_bytes) break;
    peek = 1 << 20;
  }
  return hdr.valid ? 0 : 1;
}

std::vector<uint64_t> to_shape(const long* shape, unsigned n) {
  std::vector<uint64_t> v;
  for (unsigned i = 0; i < n; ++i) v.push_back(shape[i] < 0 ? 0 : (uint64_t)shape[i]);
  return v;
}

#include "sharded.inl"

}  // namespace

// =================================================================================================
// sqyx_* — device pointers
// =================================================================================================
extern "C" {

int sqyx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int sqyx_set_device(int device) {
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  std::lock_guard<std::mutex> lk(g_set_mu);
  g_set.assign(1, device);      // a caller that names its device (one process per GPU) is not sharded over the others
  return 0;
}

int sqyx_set_devices(int n, const int* devices) {
  int have = 0;
  if (n < 0 || (n > 0 && !devices) || cudaGetDeviceCount(&have) != cudaSuccess) return 1;
  std::vector<int> v;
  for (int i = 0; i < n; ++i) {
    if (devices[i] < 0 || devices[i] >= have || devices[i] >= kMaxDevices) return 1;
    if (std::find(v.begin(), v.end(), devices[i]) == v.end()) v.push_back(devices[i]);
  }
  std::sort(v.begin(), v.end());
  std::lock_guard<std::mutex> lk(g_set_mu);
  g_set = v;                    // empty: back to the default (SQY_CUDA_DEVICES / every visible device)
  return 0;
}

int sqyx_last_shard_info(long* out3) {
  if (!out3) return 1;
  std::lock_guard<std::mutex> lk(g_shard_stats_mu);
  out3[0] = g_shard_stats.gpus;
  out3[1] = g_shard_stats.nccl;
  out3[2] = g_shard_stats.pieces;
  return 0;
}

long sqyx_nccl_allreduces(void) { return g_nccl_allreduces.load(); }

long sqyx_kernel_launches(void) { return sqyb::g_kernel_launches.load(); }

int sqyx_enable_stage_timing(int on) {
  g_timing.store(on ? 1 : 0);
  return 0;
}

int sqyx_stage_ms(float* out7, int reset) {
  std::lock_guard<std::mutex> lk(g_timer_mu);
  for (int i = 0; i < kNumTimers; ++i) {
    if (out7) out7[i] = g_stage_ms[i];
    if (reset) g_stage_ms[i] = 0;
  }
  return 0;
}

long sqyx_host_l2_bytes(void) { return (long)host_l2_cache_bytes(); }

int sqyx_last_lz4_stats(long* out4) {
  DevLock dl;
  if (dl.acquire()) return 1;
  Arena* A = &dl.dev->arena;
  for (int i = 0; i < 4; ++i) out4[i] = A->last_stats[i];
  return 0;
}

long sqyx_set_lz4_defer_min(long nblocks) { return k_lz4_set_defer_min(nblocks); }

int sqyx_release_scratch(void) {
  DevLock dl;
  if (dl.acquire()) return 1;
  Arena* A = &dl.dev->arena;
  cudaDeviceSynchronize();
  A->release();
  for (int l = 0; l < kBatchLanes; ++l) dl.dev->batch_arena[l].release();
  staging_release();
  return 0;
}

int sqyx_encode_device_ex_UI16(const char* pipeline, const void* d_src, const long* shape, unsigned shape_size, void* d_dst,
                               long dst_capacity, long* dst_bytes, const void* d_global_hist, void* stream) {
  try {
    if (!pipeline || !shape || !d_dst || !dst_bytes || dst_capacity < 0) return 1;
    Pipeline pl;
    if (!build_pipeline_u16(pipeline, pl) || pl.empty()) return 1;
    DevLock dl;
    if (dl.acquire()) return 1;
    Arena* A = &dl.dev->arena;
    uint64_t out = 0;
    const int rc = encode_device_impl(*A, pl, static_cast<const uint16_t*>(d_src), to_shape(shape, shape_size),
                                      static_cast<uint8_t*>(d_dst), (uint64_t)dst_capacity, &out,
                                      static_cast<const uint32_t*>(d_global_hist), static_cast<cudaStream_t>(stream));
    if (rc == 0) *dst_bytes = (long)out;
    return rc;
  } catch (...) {
    return 1;
  }
}

int sqyx_encode_device_UI16(const char* pipeline, const void* d_src, const long* shape, unsigned shape_size, void* d_dst,
                            long dst_capacity, long* dst_bytes, void* stream) {
  return sqyx_encode_device_ex_UI16(pipeline, d_src, shape, shape_size, d_dst, dst_capacity, dst_bytes, nullptr, stream);
}

int sqyx_decode_device_UI16(const void* d_blob, long blob_bytes, void* d_dst, long dst_capacity, void* stream) {
  try {
    if (!d_blob || blob_bytes <= 0 || !d_dst || dst_capacity < 0) return 1;
    DevLock dl;
    if (dl.acquire()) return 1;
    Arena* A = &dl.dev->arena;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Header hdr;
    if (parse_blob_header_device(static_cast<const uint8_t*>(d_blob), (uint64_t)blob_bytes, hdr, st)) return 1;
    if (sizeof_typename(hdr.raw_type) != 2) return 1;
    Pipeline pl;
    if (!build_pipeline_u16(hdr.pipeline, pl) || pl.empty()) {
      std::fprintf(stderr, "[sqeazy]\t%s cannot be build with this version of sqeazy\n", hdr.pipeline.c_str());
      return 1;
    }
    return decode_device_impl(*A, hdr, pl, static_cast<const uint8_t*>(d_blob) + hdr.size, (uint64_t)blob_bytes - hdr.size,
                              static_cast<uint16_t*>(d_dst), (uint64_t)dst_capacity, st);
  } catch (...) {
    return 1;
  }
}

// ---- batches of independent stacks (time-lapse, cfg4/cfg5): lanes of one stack each, run concurrently ----
extern "C++" {
namespace {
template <class Fn>
int run_batch(int n, int* rcs, Fn&& one) {
  if (n < 0) return 1;
  if (n == 0) return 0;
  DevLock dl;
  if (dl.acquire()) return 1;
  const int dev = dl.id;
  Dev& D = *dl.dev;
  const int lanes = n < kBatchLanes ? n : kBatchLanes;
  for (int l = 0; l < lanes; ++l)
    if (!D.batch_stream[l] && cudaStreamCreateWithFlags(&D.batch_stream[l], cudaStreamNonBlocking) != cudaSuccess) return 1;
  const int timing = g_timing.exchange(0);   // the stage timers are per call, not per lane
  std::vector<int> rc((size_t)n, 1);
  std::vector<std::thread> th;
  for (int l = 0; l < lanes; ++l)
    th.emplace_back([&, l] {
      if (cudaSetDevice(dev) != cudaSuccess) return;
      for (int i = l; i < n; i += lanes) {
        try {
          rc[(size_t)i] = one(i, D.batch_arena[l], D.batch_stream[l]);
        } catch (...) {
          rc[(size_t)i] = 1;
        }
      }
    });
  for (auto& t : th) t.join();
  g_timing.store(timing);
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    if (rcs) rcs[i] = rc[(size_t)i];
    bad |= rc[(size_t)i] != 0;
  }
  return bad ? 1 : 0;
}
}  // namespace
}  // extern "C++"

int sqyx_decode_batch_device_UI16(int n, const void* const* d_blobs, const long* blob_bytes, void* const* d_dsts,
                                  const long* dst_capacities, int* rcs) {
  if (n > 0 && (!d_blobs || !blob_bytes || !d_dsts || !dst_capacities)) return 1;
  return run_batch(n, rcs, [&](int i, Arena& A, cudaStream_t st) -> int {
    if (!d_blobs[i] || blob_bytes[i] <= 0 || !d_dsts[i] || dst_capacities[i] < 0) return 1;
    Header hdr;
    if (parse_blob_header_device(static_cast<const uint8_t*>(d_blobs[i]), (uint64_t)blob_bytes[i], hdr, st)) return 1;
    if (sizeof_typename(hdr.raw_type) != 2) return 1;
    Pipeline pl;
    if (!build_pipeline_u16(hdr.pipeline, pl) || pl.empty()) return 1;
    return decode_device_impl(A, hdr, pl, static_cast<const uint8_t*>(d_blobs[i]) + hdr.size, (uint64_t)blob_bytes[i] - hdr.size,
                              d_dsts[i], (uint64_t)dst_capacities[i], st);
  });
}

int sqyx_encode_batch_device_UI16(int n, const char* pipeline, const void* const* d_srcs, const long* shape, unsigned shape_size,
                                  void* const* d_dsts, const long* dst_capacities, long* dst_bytes, int* rcs) {
  if (!pipeline || !shape || (n > 0 && (!d_srcs || !d_dsts || !dst_capacities || !dst_bytes))) return 1;
  Pipeline pl;
  try {
    if (!build_pipeline_u16(pipeline, pl) || pl.empty()) return 1;
  } catch (...) {
    return 1;
  }
  const std::vector<uint64_t> shp = to_shape(shape, shape_size);
  return run_batch(n, rcs, [&](int i, Arena& A, cudaStream_t st) -> int {
    if (!d_dsts[i] || dst_capacities[i] < 0) return 1;
    uint64_t out = 0;
    const int rc = encode_device_impl(A, pl, d_srcs[i], shp, static_cast<uint8_t*>(d_dsts[i]), (uint64_t)dst_capacities[i], &out,
                                      nullptr, st);
    if (rc == 0) dst_bytes[i] = (long)out;
    return rc;
  });
}

int sqyx_encode_device_UI8(const char* pipeline, const void* d_src, const long* shape, unsigned shape_size, void* d_dst,
                           long dst_capacity, long* dst_bytes, void* stream) {
  try {
    if (!pipeline || !shape || !d_dst || !dst_bytes || dst_capacity < 0) return 1;
    Pipeline pl;
    if (!build_pipeline_u8(pipeline, pl) || pl.empty()) return 1;
    DevLock dl;
    if (dl.acquire()) return 1;
    Arena* A = &dl.dev->arena;
    uint64_t out = 0;
    const int rc = encode_device_impl(*A, pl, d_src, to_shape(shape, shape_size), static_cast<uint8_t*>(d_dst), (uint64_t)dst_capacity,
                                      &out, nullptr, static_cast<cudaStream_t>(stream));
    if (rc == 0) *dst_bytes = (long)out;
    return rc;
  } catch (...) {
    return 1;
  }
}

int sqyx_decode_device_UI8(const void* d_blob, long blob_bytes, void* d_dst, long dst_capacity, void* stream) {
  try {
    if (!d_blob || blob_bytes <= 0 || !d_dst || dst_capacity < 0) return 1;
    DevLock dl;
    if (dl.acquire()) return 1;
    Arena* A = &dl.dev->arena;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Header hdr;
    if (parse_blob_header_device(static_cast<const uint8_t*>(d_blob), (uint64_t)blob_bytes, hdr, st)) return 1;
    if (sizeof_typename(hdr.raw_type) != 1) return 1;
    Pipeline pl;
    if (!build_pipeline_u8(hdr.pipeline, pl) || pl.empty()) {
      std::fprintf(stderr, "[sqeazy]\t%s cannot be build with this version of sqeazy\n", hdr.pipeline.c_str());
      return 1;
    }
    return decode_device_impl(*A, hdr, pl, static_cast<const uint8_t*>(d_blob) + hdr.size, (uint64_t)blob_bytes - hdr.size, d_dst,
                              (uint64_t)dst_capacity, st);
  } catch (...) {
    return 1;
  }
}

int sqyx_bitswap_encode_UI8(int w, const void* d_src, void* d_dst, long n, int threshold, void* stream) {
  if (n < 0) return 1;
  return k_bitswap8_encode(w, static_cast<const uint8_t*>(d_src), static_cast<uint8_t*>(d_dst), (uint64_t)n, threshold,
                           static_cast<cudaStream_t>(stream)) ? 1 : 0;
}
int sqyx_bitswap_decode_UI8(int w, const void* d_src, void* d_dst, long n, void* stream) {
  if (n < 0) return 1;
  return k_bitswap8_decode(w, static_cast<const uint8_t*>(d_src), static_cast<uint8_t*>(d_dst), (uint64_t)n,
                           static_cast<cudaStream_t>(stream)) ? 1 : 0;
}
int sqyx_remove_background_UI8(const void* d_src, void* d_dst, long n, int threshold, void* stream) {
  if (n < 0) return 1;
  return k_remove_background8(static_cast<const uint8_t*>(d_src), static_cast<uint8_t*>(d_dst), (uint64_t)n, threshold,
                              static_cast<cudaStream_t>(stream)) ? 1 : 0;
}

int sqyx_bitshuffle_encode_UI16(const void* d_src, void* d_dst, long n, long block_size, void* stream) {
  if (n < 0 || block_size < 0 || block_size > (1l << 30)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k_bitshuffle16_encode(static_cast<const uint16_t*>(d_src), static_cast<uint16_t*>(d_dst), (uint64_t)n, (uint32_t)block_size, st)) return 1;
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}
int sqyx_bitshuffle_decode_UI16(const void* d_src, void* d_dst, long n, long block_size, void* stream) {
  if (n < 0 || block_size < 0 || block_size > (1l << 30)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k_bitshuffle16_decode(static_cast<const uint16_t*>(d_src), static_cast<uint16_t*>(d_dst), (uint64_t)n, (uint32_t)block_size, st)) return 1;
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}

int sqyx_bitshuffle_encode_UI8(const void* d_src, void* d_dst, long n, long block_size, void* stream) {
  if (n < 0 || block_size < 0 || block_size > (1l << 30)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k_bitshuffle8_encode(static_cast<const uint8_t*>(d_src), static_cast<uint8_t*>(d_dst), (uint64_t)n, (uint32_t)block_size, st)) return 1;
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}
int sqyx_bitshuffle_decode_UI8(const void* d_src, void* d_dst, long n, long block_size, void* stream) {
  if (n < 0 || block_size < 0 || block_size > (1l << 30)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k_bitshuffle8_decode(static_cast<const uint8_t*>(d_src), static_cast<uint8_t*>(d_dst), (uint64_t)n, (uint32_t)block_size, st)) return 1;
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}

int sqyx_diff_device(int decode, int sizeof_voxel, const void* d_src, void* d_dst, long z, long y, long x, void* stream) {
  if (z < 0 || y < 0 || x < 0 || (sizeof_voxel != 1 && sizeof_voxel != 2) || d_src == d_dst) return 1;
  if (!diff_shape_supported((uint64_t)z, (uint64_t)y, (uint64_t)x, sizeof_voxel)) return 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (decode ? k_diff_decode(sizeof_voxel, d_src, d_dst, (uint64_t)z, (uint64_t)y, (uint64_t)x, st)
             : k_diff_encode(sizeof_voxel, d_src, d_dst, (uint64_t)z, (uint64_t)y, (uint64_t)x, st))
    return 1;
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}
int sqyx_diff_shape_supported(int sizeof_voxel, long z, long y, long x) {
  return z >= 0 && y >= 0 && x >= 0 && diff_shape_supported((uint64_t)z, (uint64_t)y, (uint64_t)x, sizeof_voxel) ? 1 : 0;
}

int sqyx_bitswap_encode_UI16(int w, const void* d_src, void* d_dst, long n, int threshold, void* stream) {
  if (n < 0) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k_bitswap_encode(w, static_cast<const uint16_t*>(d_src), static_cast<uint16_t*>(d_dst), (uint64_t)n, threshold, st)) return 1;
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}

int sqyx_bitswap_decode_UI16(int w, const void* d_src, void* d_dst, long n, void* stream) {
  if (n < 0) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k_bitswap_decode(w, static_cast<const uint16_t*>(d_src), static_cast<uint16_t*>(d_dst), (uint64_t)n, st)) return 1;
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}

int sqyx_remove_background_UI16(const void* d_src, void* d_dst, long n, int threshold, void* stream) {
  if (n < 0) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k_remove_background(static_cast<const uint16_t*>(d_src), static_cast<uint16_t*>(d_dst), (uint64_t)n, threshold, st)) return 1;
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : 1;
}

int sqyx_estimate_background_UI16(const void* d_src, const long* shape3, long l2_bytes, float* supports4, int* threshold,
                                  void* stream) {
  if (!d_src || !shape3 || !supports4 || !threshold) return 1;
  DevLock dl;
  if (dl.acquire()) return 1;
  Arena* A = &dl.dev->arena;
  return estimate_background(*A, static_cast<const uint16_t*>(d_src), (uint64_t)shape3[0], (uint64_t)shape3[1], (uint64_t)shape3[2],
                             l2_bytes, supports4, threshold, static_cast<cudaStream_t>(stream));
}

int sqyx_histogram_UI16(const void* d_src, long n, void* d_hist, void* stream) {
  if (n < 0 || !d_hist) return 1;
  return k_histogram_u16(static_cast<const uint16_t*>(d_src), (uint64_t)n, static_cast<uint32_t*>(d_hist),
                         static_cast<cudaStream_t>(stream)) ? 1 : 0;
}

float sqyx_histogram_support(const unsigned* hist, float threshold) { return hist ? histogram_support(hist, threshold) : 0.f; }

long sqyx_rmest_frame_portion(long frame_elems, long l2_bytes) {
  const size_t l2 = l2_bytes >= 0 ? (size_t)l2_bytes : host_l2_cache_bytes();
  return (long)rmest_frame_portion((uint64_t)frame_elems, l2);
}

int sqyx_quantiser_luts(const unsigned* hist, unsigned char* enc, unsigned short* dec) {
  if (!hist || !enc || !dec) return 1;
  quantiser_luts_from_histogram(hist, enc, dec);
  return 0;
}

int sqyx_lut_apply_UI16(const void* d_src, void* d_codes, long n, const unsigned char* enc_host, void* stream) {
  if (n < 0 || !enc_host) return 1;
  DevLock dl;
  if (dl.acquire()) return 1;
  Arena* A = &dl.dev->arena;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* sp = nullptr;
  if (A->get(kSlotSmall, 4 * 65536 * sizeof(uint32_t) + 4096, &sp)) return 1;
  CK(cudaMemcpyAsync(sp, enc_host, 65536, cudaMemcpyHostToDevice, st));
  CKK(k_lut_apply(static_cast<const uint16_t*>(d_src), static_cast<uint8_t*>(d_codes), (uint64_t)n, static_cast<uint8_t*>(sp), st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int sqyx_lut_decode_UI16(const void* d_codes, void* d_dst, long n, const unsigned short* dec_host, void* stream) {
  if (n < 0 || !dec_host) return 1;
  DevLock dl;
  if (dl.acquire()) return 1;
  Arena* A = &dl.dev->arena;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* sp = nullptr;
  if (A->get(kSlotSmall, 4 * 65536 * sizeof(uint32_t) + 4096, &sp)) return 1;
  CK(cudaMemcpyAsync(sp, dec_host, 512, cudaMemcpyHostToDevice, st));
  CKK(k_lut_decode(static_cast<const uint8_t*>(d_codes), static_cast<uint16_t*>(d_dst), (uint64_t)n, static_cast<uint16_t*>(sp), st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

long sqyx_lz4_bound(long nbytes) { return nbytes < 0 ? -1 : (long)lz4_payload_bound((uint64_t)nbytes); }

int sqyx_lz4_encode_ex(const void* d_src, long nbytes, void* d_dst, long dst_capacity, long* payload_bytes, long pitch_bytes, void* stream) {
  if (nbytes < 0 || !d_dst || !payload_bytes) return 1;
  if ((uint64_t)dst_capacity < lz4_payload_bound((uint64_t)nbytes)) return 1;
  DevLock dl;
  if (dl.acquire()) return 1;
  Arena* A = &dl.dev->arena;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* ws = nullptr;
  if (A->get(kSlotWs, k_lz4_encode_workspace_bytes((uint64_t)nbytes), &ws)) return 1;
  // (pitch_bytes may carry kLz4HintNoNoise, kernels.h)
  const long pitch = pitch_bytes & ~(long)kLz4HintNoNoise;
  const uint32_t hint = (pitch > 0 && pitch <= 8192 ? (uint32_t)pitch : 0u) | (uint32_t)(pitch_bytes & (long)kLz4HintNoNoise);
  CKK(k_lz4_encode(static_cast<const uint8_t*>(d_src), (uint64_t)nbytes, static_cast<uint8_t*>(d_dst), ws, hint, st));
  unsigned long long hres[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(hres, ws, 32, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *payload_bytes = (long)hres[0];
  const uint32_t* stats = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(hres) + 16);
  A->last_stats[0] = stats[0]; A->last_stats[1] = stats[1]; A->last_stats[2] = stats[2]; A->last_stats[3] = (long)hres[0];
  return 0;
}

int sqyx_lz4_encode(const void* d_src, long nbytes, void* d_dst, long dst_capacity, long* payload_bytes, void* stream) {
  return sqyx_lz4_encode_ex(d_src, nbytes, d_dst, dst_capacity, payload_bytes, 0, stream);
}

int sqyx_lz4_decode(const void* d_src, long nbytes, void* d_dst, long dst_bytes, long* decoded_bytes, void* stream) {
  if (nbytes < 0 || dst_bytes < 0 || !d_src || !d_dst) return 1;
  DevLock dl;
  if (dl.acquire()) return 1;
  Arena* A = &dl.dev->arena;
  uint64_t got = 0;
  const int rc = lz4_decode_checked(*A, static_cast<const uint8_t*>(d_src), (uint64_t)nbytes, static_cast<uint8_t*>(d_dst),
                                    (uint64_t)dst_bytes, &got, static_cast<cudaStream_t>(stream));
  if (decoded_bytes) *decoded_bytes = (long)got;
  return rc;
}

// =================================================================================================
// SQY_* — the reference's boundary (host buffers)
// =================================================================================================

int SQY_Header_Size(const char* src, long* length) {
  if (!src || !length) return 1;
  const Header h = unpack_header(src, (size_t)(*length < 0 ? 0 : *length));
  *length = (long)h.size;  // 0 when no valid header is found (the reference's empty header)
  return 0;
}

int SQY_Decompressed_NDims(const char* src, long* num) {
  if (!src || !num) return 1;
  const Header h = unpack_header(src, (size_t)(*num < 0 ? 0 : *num));
  *num = (long)h.shape.size();
  return 0;
}

int SQY_Decompressed_Shape(const char* src, long* shape) {
  if (!src || !shape) return 1;
  const Header h = unpack_header(src, (size_t)(shape[0] < 0 ? 0 : shape[0]));
  for (size_t i = 0; i < h.shape.size(); ++i) shape[i] = (long)h.shape[i];
  return 0;
}

int SQY_Decompressed_Sizeof(const char* src, long* Sizeof) {
  if (!src || !Sizeof) return 1;
  const Header h = unpack_header(src, (size_t)(*Sizeof < 0 ? 0 : *Sizeof));
  *Sizeof = h.valid ? (long)sizeof_typename(h.raw_type) : 0;
  return 0;
}

int SQY_Decompressed_Length(const char* src, long* length) {
  if (!src || !length) return 1;
  const Header h = unpack_header(src, (size_t)(*length < 0 ? 0 : *length));
  // header::raw_size_byte (sqeazy_header.hpp:467-477); an empty shape multiplies out to 1 element in the reference
  uint64_t n = 1;
  for (uint64_t d : h.shape) n *= d;
  *length = (long)(n * sizeof_typename(h.raw_type));
  return 0;
}

int SQY_Version_Triple(int* version) {
  if (!version) return 1;
  version[0] = kVersionMajor;
  version[1] = kVersionMinor;
  version[2] = kVersionPatch;
  return 0;
}

bool SQY_Pipeline_Possible_UI16(const char* pipeline) {
  try {
    return pipeline ? pipeline_possible_u16(pipeline) : false;
  } catch (...) {
    return false;
  }
}
bool SQY_Pipeline_Possible_UI8(const char* pipeline) {
  try {
    return pipeline ? pipeline_possible_u8(pipeline) : false;
  } catch (...) {
    return false;
  }
}
bool SQY_Pipeline_Possible(const char* pipeline, int sizeofpixel) {   // src/sqeazy.cpp:254-268
  return sizeofpixel == 2 ? SQY_Pipeline_Possible_UI16(pipeline) : (sizeofpixel == 1 ? SQY_Pipeline_Possible_UI8(pipeline) : false);
}

int SQY_Pipeline_Max_Compressed_Length_UI16(const char* pipeline, long pipeline_length, long* length) {
  try {
    if (!pipeline || !length || pipeline_length < 0 || *length < 0) return 1;
    Pipeline pl;
    if (!build_pipeline_u16(std::string(pipeline, (size_t)pipeline_length), pl) || pl.empty()) return 1;
    *length = (long)max_encoded_size(pl, (uint64_t)*length);
    return 0;
  } catch (...) {
    return 1;
  }
}

int SQY_Pipeline_Max_Compressed_Length_3D_UI16(const char* pipeline, long* shape, unsigned shape_size, long* length) {
  try {
    if (!pipeline || !length || !shape || *length < 0) return 1;
    Pipeline pl;
    if (!build_pipeline_u16(std::string(pipeline, (size_t)*length), pl) || pl.empty()) return 1;
    uint64_t n = 1;
    for (unsigned i = 0; i < shape_size; ++i) n *= (uint64_t)shape[i];  // 64-bit, unlike the reference
    *length = (long)max_encoded_size(pl, 2 * n);
    return 0;
  } catch (...) {
    return 1;
  }
}

int SQY_Pipeline_Max_Compressed_Length_UI8(const char* pipeline, long pipeline_length, long* length) {   // src/sqeazy.cpp:144-163
  try {
    if (!pipeline || !length || pipeline_length < 0 || *length < 0) return 1;
    Pipeline pl;
    if (!build_pipeline_u8(std::string(pipeline, (size_t)pipeline_length), pl) || pl.empty()) return 1;
    *length = (long)max_encoded_size(pl, (uint64_t)*length);
    return 0;
  } catch (...) {
    return 1;
  }
}
int SQY_Pipeline_Max_Compressed_Length_3D_UI8(const char* pipeline, long* shape, unsigned shape_size, long* length) {   // :209-231
  try {
    if (!pipeline || !length || !shape || *length < 0) return 1;
    Pipeline pl;
    if (!build_pipeline_u8(std::string(pipeline, (size_t)*length), pl) || pl.empty()) return 1;
    uint64_t n = 1;
    for (unsigned i = 0; i < shape_size; ++i) n *= (uint64_t)shape[i];
    *length = (long)max_encoded_size(pl, n);
    return 0;
  } catch (...) {
    return 1;
  }
}

// ---- streamed host paths: the PCIe hop of a large stack overlaps the kernels ----
// `[remove_background | rmestbkrd ->] bitswapN -> lz4` on uint16 (the headline pipelines): the stack arrives in z-slabs on a
// copy stream; as soon as a slab is there it is filtered + transposed into its place in the plane-major buffer and the
// 16 KiB LZ4 blocks it completes in each of the 16/N bit planes are compressed, while the next slab is still on the bus.
// Only the last slab's kernels, the offset scan and the compaction remain after the last byte has arrived. Decode: the final
// inverse transpose runs slab by slab and every slab leaves for the host as soon as it is done.
bool streamable_encode(const Pipeline& pl, const std::vector<uint64_t>& shape, uint64_t N) {
  if (pl.elem != 2 || !pl.has_sink || pl.sink.kind != StageKind::Lz4 || pl.has_tail) return false;
  if (2 * N < kStreamMinBytes || N % kStreamGrainVoxels) return false;
  if (pl.head.size() == 1) return pl.head[0].kind == StageKind::Bitswap;
  if (pl.head.size() != 2 || pl.head[1].kind != StageKind::Bitswap) return false;
  if (pl.head[0].kind == StageKind::RmEstBkrd) return shape.size() == 3 && shape[0] >= 3;
  return pl.head[0].kind == StageKind::RemoveBackground;
}

// returns 0 and the blob size, or an error; the caller has checked streamable_encode()
int host_encode_streamed(Arena& A, const Pipeline& pl, const char* src, const std::vector<uint64_t>& shape, uint64_t N, char* dst,
                         uint64_t cap, uint64_t* out_bytes, int nthreads) {
  const uint64_t raw_bytes = 2 * N;
  cudaStream_t st = nullptr, cs = nullptr;
  if (copy_stream(&cs)) return 1;
  void *d_in_v = nullptr, *d_out_v = nullptr, *d_planes_v = nullptr, *ws = nullptr;
  if (A.get(kSlotIn, raw_bytes, &d_in_v) || A.get(kSlotOut, cap, &d_out_v) || A.get(kSlotA, raw_bytes, &d_planes_v) ||
      A.get(kSlotWs, k_lz4_encode_workspace_bytes(raw_bytes), &ws))
    return 1;
  uint16_t* d_in = static_cast<uint16_t*>(d_in_v);
  uint16_t* d_planes = static_cast<uint16_t*>(d_planes_v);
  uint8_t* d_out = static_cast<uint8_t*>(d_out_v);
  const size_t reserve = header_reserve_bytes(pl, shape);
  if (cap < reserve || cap - reserve < lz4_payload_bound(raw_bytes)) return 1;
  uint8_t* payload = d_out + reserve;
  const Stage& swap = pl.head.back();
  const int P = 16 / swap.w;

  int thr = pl.head.size() == 2 ? pl.head[0].threshold : 0;
  // rmestbkrd: the estimate samples the two z faces and six border rows. They are sent on their own (behind the first slab,
  // which is already on its way: the bus never waits for the estimate) and the threshold is there before the first kernel.
  auto estimate = [&]() -> int {
    if (pl.head.size() != 2 || pl.head[0].kind != StageKind::RmEstBkrd) return 0;
    const uint64_t Z = shape[0], Y = shape[1], X = shape[2], frame = Y * X;
    const uint64_t portion = rmest_frame_portion(frame, host_l2_cache_bytes());
    const uint16_t* h = reinterpret_cast<const uint16_t*>(src);
    CK(cudaMemcpyAsync(d_in, h, 2 * portion, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_in + (Z - 1) * frame, h + (Z - 1) * frame, 2 * portion, cudaMemcpyHostToDevice, st));
    const uint64_t zs[3] = {1, Z / 2, Z - 2}, ys[2] = {0, Y - 1};
    for (uint64_t z : zs)
      for (uint64_t y : ys) CK(cudaMemcpyAsync(d_in + z * frame + y * X, h + z * frame + y * X, 2 * X, cudaMemcpyHostToDevice, st));
    float sup[4];
    return estimate_background(A, d_in, Z, Y, X, -1, sup, &thr, st);
  };

  CKK(k_lz4_encode_begin(raw_bytes, payload, ws, st));
  const uint64_t slab_voxels = kStreamSlabBytes / 2;          // a multiple of the grain
  const uint32_t blocks_per_plane = (uint32_t)(raw_bytes / P / kLz4BlockBytes);
  EventList events;
  for (uint64_t first = 0; first < N; first += slab_voxels) {
    const uint64_t count = std::min(slab_voxels, N - first);
    CK((cudaError_t)staged_h2d(d_in + first, src + 2 * first, 2 * count, nthreads, cs));
    cudaEvent_t arrived = nullptr;
    if (events.next(&arrived)) return 1;
    CK(cudaEventRecord(arrived, cs));
    if (first == 0 && estimate()) return 1;
    CK(cudaStreamWaitEvent(st, arrived, 0));
    CKK(k_bitswap_encode_range(swap.w, d_in, d_planes, N, first, count, thr, st));
    // plane piece of this slab: bytes [2*first/P, 2*(first+count)/P) of every plane
    CKK(k_lz4_encode_blocks(reinterpret_cast<const uint8_t*>(d_planes), raw_bytes, payload, ws,
                            (uint32_t)(2 * first / P / kLz4BlockBytes), (uint32_t)(2 * count / P / kLz4BlockBytes), (uint32_t)P,
                            blocks_per_plane, lz4_pitch_hint(pl, shape), st));
  }
  CKK(k_lz4_encode_end(reinterpret_cast<const uint8_t*>(d_planes), raw_bytes, payload, ws, st));
  unsigned long long hres[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(hres, ws, 32, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const uint64_t payload_bytes = hres[0];
  const uint32_t* stats = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(hres) + 16);
  A.last_stats[0] = stats[0]; A.last_stats[1] = stats[1]; A.last_stats[2] = stats[2]; A.last_stats[3] = (long)payload_bytes;

  Pipeline named = pl;
  const std::string hdr = pack_header(named.type_name(), named.elem, shape, named.canonical(), payload_bytes);
  if (hdr.size() > reserve) return 1;
  std::string slot(reserve - hdr.size(), ' ');
  slot += hdr;
  CK(cudaMemcpyAsync(d_out, slot.data(), slot.size(), cudaMemcpyHostToDevice, st));
  CK((cudaError_t)staged_d2h(dst, d_out, reserve + payload_bytes, nthreads, st));
  CK(cudaStreamSynchronize(st));
  *out_bytes = reserve + payload_bytes;
  return 0;
}

static int host_encode(int elem, const char* pipeline, const char* src, long* shape, unsigned shape_size, char* dst, long* dstlength,
                       int nthreads) {
  NvtxRange range(elem == 1 ? "SQY_PipelineEncode_UI8" : "SQY_PipelineEncode_UI16");
  try {
    if (!pipeline || !src || !shape || !dst || !dstlength) return 1;
    Pipeline pl;
    if (!(elem == 1 ? build_pipeline_u8(pipeline, pl) : build_pipeline_u16(pipeline, pl))) return 1;
    if (pl.empty()) {
      std::fprintf(stderr, "[sqeazy]\t received pipeline of size 0, cannot encode buffer\n");
      return 1;
    }
    const std::vector<uint64_t> shp = to_shape(shape, shape_size);
    const uint64_t N = shape_product(shp), raw_bytes = (uint64_t)elem * N;
    const uint64_t cap = max_encoded_size(pl, raw_bytes);
    if (elem == 2) {
      // one stack over several GPUs (and the streamed quantiser path on one): sharded.inl
      {
        std::lock_guard<std::mutex> lk(g_shard_stats_mu);
        g_shard_stats = ShardStats();
      }
      uint64_t out = 0;
      const int rc = host_encode_sharded(pl, src, shp, N, dst, cap, &out, nthreads);
      if (rc > 0) return 1;
      if (rc == 0) {
        *dstlength = (long)out;
        return 0;
      }
    }
    DevLock dl;
    if (dl.acquire()) return 1;
    Arena* A = &dl.dev->arena;
    if (streamable_encode(pl, shp, N) && !g_timing.load()) {
      uint64_t out = 0;
      if (host_encode_streamed(*A, pl, src, shp, N, dst, cap, &out, nthreads)) return 1;
      *dstlength = (long)out;
      return 0;
    }
    void *d_in = nullptr, *d_out = nullptr;
    if (A->get(kSlotIn, raw_bytes, &d_in) || A->get(kSlotOut, cap, &d_out)) return 1;
    cudaStream_t st = nullptr;
    CK((cudaError_t)staged_h2d(d_in, src, raw_bytes, nthreads, st));   // pageable buffers: nthreads host threads feed a pinned ring
    uint64_t out = 0;
    const int rc = encode_device_impl(*A, pl, d_in, shp, static_cast<uint8_t*>(d_out), cap, &out,
                                      nullptr, st);
    if (rc) return 1;
    CK((cudaError_t)staged_d2h(dst, d_out, out, nthreads, st));
    CK(cudaStreamSynchronize(st));
    *dstlength = (long)out;
    return 0;
  } catch (...) {
    return 1;
  }
}

static int host_decode(int elem, const char* src, long srclength, char* dst, int nthreads) {
  NvtxRange range(elem == 1 ? "SQY_Decode_UI8" : "SQY_Decode_UI16");
  try {
    if (!src || !dst || srclength <= 0) return 1;
    const Header hdr = unpack_header(src, (size_t)srclength);
    if (!hdr.valid) return 1;
    Pipeline pl;
    if (!(elem == 1 ? build_pipeline_u8(hdr.pipeline, pl) : build_pipeline_u16(hdr.pipeline, pl))) {
      std::fprintf(stderr, "[sqeazy]\t%s cannot be build with this version of sqeazy\n", hdr.pipeline.c_str());
      return 1;
    }
    if (pl.empty()) {
      std::fprintf(stderr, "[sqeazy]\t received pipeline of size 0, no decoding possible\n");
      return 1;
    }
    if (sizeof_typename(hdr.raw_type) != (unsigned)elem) return 1;
    const uint64_t raw_bytes = (uint64_t)elem * shape_product(hdr.shape);
    if (elem == 2) {
      {
        std::lock_guard<std::mutex> lk(g_shard_stats_mu);
        g_shard_stats = ShardStats();
      }
      const int rc = host_decode_sharded(hdr, pl, src, (uint64_t)srclength, dst, nthreads);
      if (rc >= 0) return rc;
    }
    DevLock dl;
    if (dl.acquire()) return 1;
    Arena* A = &dl.dev->arena;
    void *d_in = nullptr, *d_out = nullptr;
    const uint64_t payload_bytes = (uint64_t)srclength - hdr.size;
    // the payload is staged at a 256-byte aligned device address, whatever the header length was
    if (A->get(kSlotIn, payload_bytes + 256, &d_in) || A->get(kSlotOut, raw_bytes, &d_out)) return 1;
    cudaStream_t st = nullptr;
    CK((cudaError_t)staged_h2d(d_in, src + hdr.size, payload_bytes, nthreads, st));
    HostSink sink{dst, nthreads};
    const int rc = decode_device_impl(*A, hdr, pl, static_cast<const uint8_t*>(d_in), payload_bytes, d_out,
                                      raw_bytes, st, &sink);
    if (rc) return rc;
    if (!sink.done) CK((cudaError_t)staged_d2h(dst, d_out, raw_bytes, nthreads, st));
    CK(cudaStreamSynchronize(st));
    return 0;
  } catch (...) {
    return 1;
  }
}

int SQY_PipelineEncode_UI16(const char* pipeline, const char* src, long* shape, unsigned shape_size, char* dst, long* dstlength,
                            int nthreads) {
  // nthreads: host threads that stage a pageable caller buffer (staging.hpp); the kernels have no thread knob
  return host_encode(2, pipeline, src, shape, shape_size, dst, dstlength, nthreads);
}

int SQY_Decode_UI16(const char* src, long srclength, char* dst, int nthreads) {
  return host_decode(2, src, srclength, dst, nthreads);
}

int SQY_PipelineDecode_UI16(const char* src, long srclength, char* dst, int nthreads) {
  return SQY_Decode_UI16(src, srclength, dst, nthreads);
}

// uint8 volumes (src/sqeazy.cpp:72-106, 309-335)
int SQY_PipelineEncode_UI8(const char* pipeline, const char* src, long* shape, unsigned shape_size, char* dst, long* dstlength,
                           int nthreads) {
  return host_encode(1, pipeline, src, shape, shape_size, dst, dstlength, nthreads);
}
int SQY_Decode_UI8(const char* src, long srclength, char* dst, int nthreads) {
  return host_decode(1, src, srclength, dst, nthreads);
}

int SQY_h5_query_sizeof(const char*, const char*, unsigned*) { return 1; }
int SQY_h5_query_dtype(const char*, const char*, unsigned*) { return 1; }
int SQY_h5_query_ndims(const char*, const char*, unsigned*) { return 1; }
int SQY_h5_query_shape(const char*, const char*, unsigned*) { return 1; }
int SQY_h5_read_UI16(const char*, const char*, unsigned short*) { return 1; }
int SQY_h5_write_UI16(const char*, const char*, const unsigned short*, unsigned, const unsigned*, const char*) { return 1; }
int SQY_h5_write(const char*, const char*, const char*, unsigned long) { return 1; }
int SQY_h5_link(const char*, const char*, const char*, const char*, const char*, const char*) { return 1; }

}  // extern "C"
