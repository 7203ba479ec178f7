// Included by api.cu (inside its anonymous namespace): one stack sharded over several GPUs behind the SQY_* host entry points.
//
// SURVEY §8e / BASELINE north_star: "stacks and z-slabs are partitioned over the GPUs; the only collective is an allreduce of
// the quantiser's global histogram". The reference has one address space and OpenMP (dynamic_pipeline.hpp:560-690); here a
// caller's host buffer is cut into contiguous voxel ranges (z-slabs, boundaries at multiples of 131 072 voxels so that no
// 16 KiB LZ4 block of any bit plane straddles two GPUs), every GPU pulls its range over its own PCIe link, runs the SAME
// kernels on it as a stack of its own, and the pieces are merged into ONE blob that is byte-identical to what one GPU
// writes:
//   * `[filter ->] bitswapN -> lz4`: the bit-plane transpose of voxels [a, b) taken alone is exactly the piece
//     [a/P', b/P') of every plane of the whole stack (plane-major layout, §3 of DESIGN.md), so GPU g's local LZ4 blocks of
//     plane p are the global blocks [p*bpp + k0_g, p*bpp + k0_g + bpp_g): a contiguous byte range of the final frame. The
//     block-size words come back (4 B per block), the host takes the exclusive scan over (plane, GPU) pieces and every GPU
//     copies its 16/N pieces straight into the caller's `dst` at their final offsets. No GPU <-> GPU payload traffic.
//   * `quantiser -> lz4`: local 65536-bin histograms, summed over the GPUs by ncclAllReduce (uint32, sum — wraps like the
//     reference's uint32 bins, histogram_utils.hpp:141-150), one LUT on the host, local LUT apply + LZ4 of a contiguous
//     byte range of the code stream, merged the same way (one piece per GPU).
//   * rmestbkrd's threshold samples two faces and six rows of the whole stack: they are sent to the first GPU on their own
//     before the slabs move (estimate_background_host).
// Decode is the mirror image: the index frame of the blob (host memory) gives every piece's byte range; every GPU gets a
// small local blob (index words + pieces), decodes it slab by slab with the single-GPU streamed path and writes its voxel
// range of the caller's buffer.
//
// The same code with one GPU is the streamed host path of `quantiser -> lz4` (histogram while the stack arrives).

// ---------------------------------------------------------------------------------------------------------------------
// device set
// ---------------------------------------------------------------------------------------------------------------------
std::vector<int> parse_device_list(const char* s, int n) {
  std::vector<int> v;
  if (!s) return v;
  if (std::strcmp(s, "all") == 0) {
    for (int i = 0; i < n; ++i) v.push_back(i);
    return v;
  }
  const char* p = s;
  while (*p) {
    char* end = nullptr;
    const long d = std::strtol(p, &end, 10);
    if (end == p) break;
    if (d >= 0 && d < n && d < kMaxDevices && std::find(v.begin(), v.end(), (int)d) == v.end()) v.push_back((int)d);
    p = end;
    while (*p == ',' || *p == ' ') ++p;
  }
  return v;
}

// devices a shardable host call may use, ascending. Priority: sqyx_set_devices / sqyx_set_device, SQY_CUDA_DEVICES
// ("all" or "0,1,.."), SQY_CUDA_DEVICE (one), else every visible device.
std::vector<int> shard_devices() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return {};
  std::vector<int> v;
  {
    std::lock_guard<std::mutex> lk(g_set_mu);
    v = g_set;
  }
  if (v.empty()) {
    if (const char* e = std::getenv("SQY_CUDA_DEVICES")) v = parse_device_list(e, n);
    else if (const char* e1 = std::getenv("SQY_CUDA_DEVICE")) v = parse_device_list(e1, n);
    else
      for (int i = 0; i < n && i < kMaxDevices; ++i) v.push_back(i);
  }
  v.erase(std::remove_if(v.begin(), v.end(), [&](int d) { return d < 0 || d >= n || d >= kMaxDevices; }), v.end());
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
  return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// NCCL, bound at run time (the library does not link against it: a single-GPU caller needs no NCCL installed)
// ---------------------------------------------------------------------------------------------------------------------
struct NcclApi {
  typedef void* comm_t;
  int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  static constexpr int kUint32 = 3, kSum = 0;   // ncclUint32, ncclSum (nccl.h)
  NcclApi() {
    if (const char* off = std::getenv("SQY_NO_NCCL"))
      if (off[0] && off[0] != '0') return;
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return;
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(h, "ncclAllReduce"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(h, "ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    ok = CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
  }
};

NcclApi& nccl() {
  static NcclApi api;
  return api;
}

// communicators of one device set, created once (ncclCommInitAll costs ~100 ms) and kept for the life of the process
struct NcclComms {
  std::vector<int> devs;
  std::vector<NcclApi::comm_t> comm;
};
std::mutex g_nccl_mu;
std::vector<NcclComms*> g_nccl_comms;
std::atomic<long> g_nccl_allreduces{0};

NcclComms* nccl_comms_for(const std::vector<int>& devs) {
  if (!nccl().ok || devs.size() < 2) return nullptr;
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  for (NcclComms* c : g_nccl_comms)
    if (c->devs == devs) return c;
  NcclComms* c = new NcclComms;
  c->devs = devs;
  c->comm.resize(devs.size(), nullptr);
  const int rc = nccl().CommInitAll(c->comm.data(), (int)devs.size(), devs.data());
  if (rc != 0) {
    std::fprintf(stderr, "[sqeazy_b200] ncclCommInitAll failed (%s); histograms are summed through the host\n",
                 nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
    delete c;
    return nullptr;
  }
  g_nccl_comms.push_back(c);
  return c;
}

// ---------------------------------------------------------------------------------------------------------------------
// plumbing: a thread per GPU, phases separated by barriers, first error wins
// ---------------------------------------------------------------------------------------------------------------------
class PhaseBarrier {
 public:
  explicit PhaseBarrier(int n) : n_(n) {}
  void arrive_and_wait() {
    std::unique_lock<std::mutex> lk(m_);
    const uint64_t gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen_ != gen; });
    }
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0;
  uint64_t gen_ = 0;
};

constexpr uint64_t kShardMinStreamBytes = uint64_t(128) << 20;  // >= 8192 LZ4 blocks per GPU: the tiled table builders apply
constexpr uint64_t kShardSubSlabVoxels = uint64_t(32) << 20;    // 64 MiB sub-slabs: 1.2 ms on the bus each, kernels ~0.2 ms

struct ShardPlan {
  std::vector<int> dev;           // ascending
  std::vector<uint64_t> first;    // dev.size() + 1 voxel boundaries, multiples of the grain
  int G() const { return (int)dev.size(); }
  uint64_t count(int g) const { return first[g + 1] - first[g]; }
};

// contiguous voxel ranges, boundaries at multiples of `grain` voxels, at least kShardMinStreamBytes of LZ4 input each
// (stream_bytes = what the whole stack feeds into LZ4: 2 N for bit planes, N for 8-bit codes)
bool plan_shards(uint64_t N, uint64_t stream_bytes, uint64_t grain, const std::vector<int>& devs, ShardPlan& out) {
  if (N == 0 || N % grain) return false;
  const uint64_t units = N / grain;
  uint64_t G = std::min<uint64_t>({(uint64_t)devs.size(), stream_bytes / kShardMinStreamBytes, units});
  if (G < 1) G = 1;
  out.dev.assign(devs.begin(), devs.begin() + (size_t)G);
  out.first.resize((size_t)G + 1);
  for (uint64_t g = 0; g <= G; ++g) out.first[(size_t)g] = (units * g / G) * grain;
  return true;
}

// locks of all devices of a plan (ascending order: no dead-lock with other sharded calls), held by the calling thread
struct DevSetLock {
  std::vector<std::unique_lock<std::mutex>> lk;
  void acquire(const std::vector<int>& devs) {
    for (int d : devs) lk.emplace_back(g_dev[d].mu);
  }
};

enum class ShardKind { None, PlanesLz4, QuantLz4 };

ShardKind shard_kind(const Pipeline& pl, const std::vector<uint64_t>& shape) {
  if (pl.elem != 2 || !pl.has_sink) return ShardKind::None;
  if (pl.sink.kind == StageKind::Lz4 && !pl.has_tail) {
    if (pl.head.size() == 1 && pl.head[0].kind == StageKind::Bitswap) return ShardKind::PlanesLz4;
    if (pl.head.size() == 2 && pl.head[1].kind == StageKind::Bitswap) {
      if (pl.head[0].kind == StageKind::RemoveBackground) return ShardKind::PlanesLz4;
      if (pl.head[0].kind == StageKind::RmEstBkrd && shape.size() == 3 && shape[0] >= 3) return ShardKind::PlanesLz4;
    }
    return ShardKind::None;
  }
  if (pl.sink.kind == StageKind::Quantiser && pl.has_tail && pl.head.empty()) return ShardKind::QuantLz4;
  return ShardKind::None;
}

// host copy of a device range into caller memory of either kind; complete when `st` has been synchronised
int copy_out(void* h_dst, const void* d_src, size_t bytes, int nthreads, cudaStream_t st) {
  if (!bytes) return 0;
  return staged_d2h(h_dst, d_src, bytes, nthreads, st) ? 1 : 0;
}

struct ShardStats {      // what the last sharded call did (sqyx_last_shard_info)
  int gpus = 0, nccl = 0;
  long pieces = 0;
};
std::mutex g_shard_stats_mu;
ShardStats g_shard_stats;

// ---------------------------------------------------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------------------------------------------------
// Returns -1 when the call is not shardable (the caller takes a single-device route), 0 on success, > 0 on failure.
int host_encode_sharded(const Pipeline& pl_in, const char* src, const std::vector<uint64_t>& shape, uint64_t N, char* dst, uint64_t cap,
                        uint64_t* out_bytes, int nthreads) {
  const ShardKind kind = shard_kind(pl_in, shape);
  if (kind == ShardKind::None || g_timing.load()) return -1;
  const uint64_t raw_bytes = 2 * N;
  std::vector<int> devs = shard_devices();
  if (devs.empty()) return -1;
  ShardPlan plan;
  if (!plan_shards(N, kind == ShardKind::PlanesLz4 ? raw_bytes : N, kStreamGrainVoxels, devs, plan)) return -1;
  const int G = plan.G();
  // one GPU: `...bitswapN->lz4` has its own streamed path (host_encode_streamed); quantiser->lz4 streams through this code
  if (G == 1 && (kind == ShardKind::PlanesLz4 || raw_bytes < kStreamMinBytes)) return -1;
  if (G == 1) {   // a one-device call works on the calling thread's device like every other single-device call
    int d = 0;
    if (pick_device(&d)) return 1;
    plan.dev[0] = d;
  }

  Pipeline pl = pl_in;
  const int w = kind == ShardKind::PlanesLz4 ? pl.head.back().w : 0;
  const uint32_t P = kind == ShardKind::PlanesLz4 ? (uint32_t)(16 / w) : 1u;     // LZ4 streams ("planes") per GPU
  const uint64_t stream_bytes = kind == ShardKind::PlanesLz4 ? raw_bytes : N;     // bytes that go through LZ4
  const uint64_t nblocks = stream_bytes / kLz4BlockBytes;
  const uint64_t bpp = nblocks / P;                                                // blocks per plane of the whole stack
  const size_t reserve = header_reserve_bytes(pl, shape);
  if (cap < reserve || cap - reserve < lz4_payload_bound(stream_bytes)) return 1;
  char* payload = dst + reserve;
  const uint64_t prefix = lz4_prefix_bytes(nblocks);
  const int T = std::max(1, staging_threads(nthreads) / G);
  const uint32_t pitch_hint = lz4_pitch_hint(pl_in, shape);

  DevSetLock locks;
  locks.acquire(plan.dev);

  // rmestbkrd: threshold of the WHOLE stack, from its sampled faces and rows, on the first GPU
  int thr = (kind == ShardKind::PlanesLz4 && pl.head.size() == 2) ? pl.head[0].threshold : 0;
  if (kind == ShardKind::PlanesLz4 && pl.head.size() == 2 && pl.head[0].kind == StageKind::RmEstBkrd) {
    if (cudaSetDevice(plan.dev[0]) != cudaSuccess) return 1;
    Arena& A0 = g_dev[plan.dev[0]].arena;
    const uint64_t frame = shape[1] * shape[2];
    void* scratch = nullptr;
    if (A0.get(kSlotB, 4 * rmest_frame_portion(frame, host_l2_cache_bytes()) + 12 * shape[2] + 256, &scratch)) return 1;
    if (estimate_background_host(A0, reinterpret_cast<const uint16_t*>(src), shape[0], shape[1], shape[2],
                                 static_cast<uint16_t*>(scratch), &thr, nullptr))
      return 1;
  }

  struct Share {
    uint64_t a = 0, n = 0, sbytes = 0, nblk = 0, bpp = 0;   // voxel range, LZ4 stream bytes / blocks / blocks per plane of this GPU
    uint8_t* d_out = nullptr;
    uint32_t* d_hist = nullptr;
    uint16_t* d_in = nullptr;
    void* ws = nullptr;
    cudaStream_t cs = nullptr;
    std::vector<uint32_t> idx;                  // block words of this GPU, local order (plane-major)
    std::vector<uint64_t> piece_bytes;          // P pieces
    std::vector<uint64_t> piece_goff;           // final offset of each piece behind the frame header
    std::vector<uint32_t> hist;
    long stats[3] = {0, 0, 0};
  };
  std::vector<Share> S((size_t)G);
  for (int g = 0; g < G; ++g) {
    S[g].a = plan.first[g];
    S[g].n = plan.count(g);
    S[g].sbytes = kind == ShardKind::PlanesLz4 ? 2 * S[g].n : S[g].n;
    S[g].nblk = S[g].sbytes / kLz4BlockBytes;
    S[g].bpp = S[g].nblk / P;
  }
  NcclComms* comms = kind == ShardKind::QuantLz4 ? nccl_comms_for(plan.dev) : nullptr;
  std::vector<uint8_t> lut_enc(65536);
  uint16_t lut_dec[256];
  uint64_t total_block_bytes = 0;
  std::atomic<int> fail{0};
  PhaseBarrier bar(G);

  auto worker = [&](int g) {
    Share& W = S[g];
    Arena& A = g_dev[plan.dev[g]].arena;
    cudaStream_t st = nullptr;    // legacy default stream of this device
    auto phase1 = [&]() -> int {
      if (cudaSetDevice(plan.dev[g]) != cudaSuccess) return 1;
      if (copy_stream(&W.cs)) return 1;
      void *d_in = nullptr, *d_out = nullptr, *d_planes = nullptr;
      if (A.get(kSlotIn, 2 * W.n, &d_in) || A.get(kSlotOut, lz4_payload_bound(W.sbytes), &d_out) || A.get(kSlotA, W.sbytes, &d_planes) ||
          A.get(kSlotWs, k_lz4_encode_workspace_bytes(W.sbytes), &W.ws))
        return 1;
      W.d_in = static_cast<uint16_t*>(d_in);
      W.d_out = static_cast<uint8_t*>(d_out);
      if (kind == ShardKind::QuantLz4) {
        void* sp = nullptr;
        if (A.get(kSlotSmall, 4 * 65536 * sizeof(uint32_t) + 4096, &sp)) return 1;
        W.d_hist = static_cast<uint32_t*>(sp);
        CK(cudaMemsetAsync(W.d_hist, 0, 65536 * sizeof(uint32_t), st));
      } else {
        CKK(k_lz4_encode_begin(W.sbytes, W.d_out, W.ws, st));
      }
      EventList events;
      const uint16_t* h = reinterpret_cast<const uint16_t*>(src) + W.a;
      for (uint64_t first = 0; first < W.n; first += kShardSubSlabVoxels) {
        const uint64_t count = std::min(kShardSubSlabVoxels, W.n - first);
        CK((cudaError_t)staged_h2d(W.d_in + first, h + first, 2 * count, T, W.cs));
        cudaEvent_t arrived = nullptr;
        if (events.next(&arrived)) return 1;
        CK(cudaEventRecord(arrived, W.cs));
        CK(cudaStreamWaitEvent(st, arrived, 0));
        if (kind == ShardKind::QuantLz4) {
          CKK(k_histogram_u16(W.d_in + first, count, W.d_hist, st));
        } else {
          CKK(k_bitswap_encode_range(w, W.d_in, static_cast<uint16_t*>(d_planes), W.n, first, count, thr, st));
          CKK(k_lz4_encode_blocks(static_cast<const uint8_t*>(d_planes), W.sbytes, W.d_out, W.ws, (uint32_t)(2 * first / P / kLz4BlockBytes),
                                  (uint32_t)(2 * count / P / kLz4BlockBytes), P, (uint32_t)W.bpp, pitch_hint, st));
        }
      }
      if (kind != ShardKind::QuantLz4) CKK(k_lz4_encode_end(static_cast<const uint8_t*>(d_planes), W.sbytes, W.d_out, W.ws, st));
      return 0;
    };
    // quantiser: the path's only collective — sum of the 65536-bin histograms over the GPUs (uint32 bins wrap like the
    // reference's). Entered only when every GPU got through phase 1 (a collective with a missing rank never returns).
    auto phase_reduce = [&]() -> int {
      if (comms) {
        const int rc = nccl().AllReduce(W.d_hist, W.d_hist, 65536, NcclApi::kUint32, NcclApi::kSum, comms->comm[g], st);
        if (rc != 0) {
          std::fprintf(stderr, "[sqeazy_b200] ncclAllReduce failed (%d)\n", rc);
          return 1;
        }
        if (g == 0) g_nccl_allreduces.fetch_add(1);
      }
      if (!comms || g == 0) {
        W.hist.resize(65536);
        CK(cudaMemcpyAsync(W.hist.data(), W.d_hist, 65536 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      }
      CK(cudaStreamSynchronize(st));
      return 0;
    };
    // quantiser: LUT apply + LZ4 of this GPU's codes, once the LUT of the whole stack is known
    auto phase_codes = [&]() -> int {
      void *sp = nullptr, *codes = nullptr;
      if (A.get(kSlotSmall, 4 * 65536 * sizeof(uint32_t) + 4096, &sp) || A.get(kSlotA, W.sbytes, &codes)) return 1;
      uint8_t* d_lut = reinterpret_cast<uint8_t*>(static_cast<uint32_t*>(sp) + 65536);
      CK(cudaMemcpyAsync(d_lut, lut_enc.data(), 65536, cudaMemcpyHostToDevice, st));
      CKK(k_lut_apply(W.d_in, static_cast<uint8_t*>(codes), W.n, d_lut, st));
      CKK(k_lz4_encode(static_cast<const uint8_t*>(codes), W.sbytes, W.d_out, W.ws, pitch_hint, st));
      return 0;
    };
    // block words of this GPU -> host, piece sizes
    auto phase_sizes = [&]() -> int {
      W.idx.resize(W.nblk);
      unsigned long long hres[4] = {0, 0, 0, 0};
      CK(cudaMemcpyAsync(W.idx.data(), W.d_out + kSkippableHeaderBytes + sizeof(SqybIndexHeader), 4 * W.nblk, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(hres, W.ws, 32, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      const uint32_t* stats = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(hres) + 16);
      for (int k = 0; k < 3; ++k) W.stats[k] = stats[k];
      W.piece_bytes.assign(P, 0);
      for (uint32_t p = 0; p < P; ++p) {
        uint64_t s = 0;
        const uint32_t* q = W.idx.data() + (uint64_t)p * W.bpp;
        for (uint64_t k = 0; k < W.bpp; ++k) s += 4ull + (q[k] & 0x7FFFFFFFu);
        W.piece_bytes[p] = s;
      }
      return 0;
    };
    // pieces and block words into the caller's buffer at their final places
    auto phase_merge = [&]() -> int {
      uint64_t loc = lz4_prefix_bytes(W.nblk);      // first block header of the local frame
      char* index = payload + kSkippableHeaderBytes + sizeof(SqybIndexHeader);
      const uint64_t k0 = (kind == ShardKind::PlanesLz4 ? 2 * W.a / P : W.a) / kLz4BlockBytes;   // first block of this GPU inside a plane
      for (uint32_t p = 0; p < P; ++p) {
        if (copy_out(payload + prefix + W.piece_goff[p], W.d_out + loc, W.piece_bytes[p], T, st)) return 1;
        loc += W.piece_bytes[p];
        std::memcpy(index + 4 * ((uint64_t)p * bpp + k0), W.idx.data() + (uint64_t)p * W.bpp, 4 * W.bpp);
      }
      CK(cudaStreamSynchronize(st));
      return 0;
    };

    if (phase1()) fail.store(1);
    bar.arrive_and_wait();
    if (kind == ShardKind::QuantLz4) {
      const bool go = !fail.load();      // (the same answer on every thread: nobody writes `fail` between these two barriers)
      bar.arrive_and_wait();
      if (go && phase_reduce()) fail.store(1);
      bar.arrive_and_wait();
      if (g == 0 && !fail.load()) {
        std::vector<uint32_t>& h = S[0].hist;
        if (!comms)
          for (int o = 1; o < G; ++o)
            for (int b = 0; b < 65536; ++b) h[b] += S[o].hist[b];
        quantiser_luts_from_histogram(h.data(), lut_enc.data(), lut_dec);
        pl.sink.kv["decode_lut_string"] = std::string(kVerbatimOpen) + base64_encode(lut_dec, sizeof(lut_dec)) + kVerbatimClose;
      }
      bar.arrive_and_wait();
      if (!fail.load() && phase_codes()) fail.store(1);
    }
    if (!fail.load() && phase_sizes()) fail.store(1);
    bar.arrive_and_wait();
    if (g == 0 && !fail.load()) {
      // exclusive scan over the pieces in their final order: plane 0 of every GPU in slab order, plane 1, ...
      uint64_t off = 0;
      for (int o = 0; o < G; ++o) S[o].piece_goff.assign(P, 0);
      for (uint32_t p = 0; p < P; ++p)
        for (int o = 0; o < G; ++o) {
          S[o].piece_goff[p] = off;
          off += S[o].piece_bytes[p];
        }
      total_block_bytes = off;
    }
    bar.arrive_and_wait();
    if (!fail.load() && phase_merge()) fail.store(1);
  };

  std::vector<std::thread> th;
  for (int g = 1; g < G; ++g) th.emplace_back(worker, g);
  worker(0);
  for (auto& t : th) t.join();
  if (fail.load()) return 1;

  // frame prefix and EndMark (what lz4_write_prefix_kernel / lz4_block_offsets_kernel write on one GPU), header
  {
    uint8_t* q = reinterpret_cast<uint8_t*>(payload);
    auto put32 = [&](uint64_t off, uint32_t v) { for (int k = 0; k < 4; ++k) q[off + k] = (uint8_t)(v >> (8 * k)); };
    auto put64 = [&](uint64_t off, uint64_t v) { for (int k = 0; k < 8; ++k) q[off + k] = (uint8_t)(v >> (8 * k)); };
    put32(0, kSqybSkippableMagic);
    put32(4, (uint32_t)(sizeof(SqybIndexHeader) + 4ull * nblocks));
    put32(8, kSqybIndexMagic);
    put32(12, 1u);
    put32(16, (uint32_t)kLz4BlockBytes);
    put32(20, (uint32_t)nblocks);
    put64(24, stream_bytes);
    put64(32, kLz4FrameHeaderBytes + total_block_bytes + kLz4EndMarkBytes);
    const uint64_t f = prefix - kLz4FrameHeaderBytes;
    put32(f, kLz4FrameMagic);
    q[f + 4] = 0x60;
    q[f + 5] = 0x40;
    q[f + 6] = 0x82;
    put32(prefix + total_block_bytes, 0u);   // EndMark
  }
  const uint64_t payload_bytes = prefix + total_block_bytes + kLz4EndMarkBytes;
  const std::string hdr = pack_header(pl.type_name(), pl.elem, shape, pl.canonical(), payload_bytes);
  if (hdr.size() > reserve) return 1;
  std::memset(dst, ' ', reserve - hdr.size());
  std::memcpy(dst + reserve - hdr.size(), hdr.data(), hdr.size());
  *out_bytes = reserve + payload_bytes;

  long st3[3] = {0, 0, 0};
  for (int g = 0; g < G; ++g)
    for (int k = 0; k < 3; ++k) st3[k] += S[g].stats[k];
  for (int g = 0; g < G; ++g) {
    Arena& A = g_dev[plan.dev[g]].arena;
    for (int k = 0; k < 3; ++k) A.last_stats[k] = st3[k];
    A.last_stats[3] = (long)payload_bytes;
  }
  {
    std::lock_guard<std::mutex> lk(g_shard_stats_mu);
    g_shard_stats.gpus = G;
    g_shard_stats.nccl = comms ? 1 : 0;
    g_shard_stats.pieces = (long)P * G;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------------------------------
// `quantiser -> lz4` blob of this library in device memory -> voxels in caller memory, slab by slab: LZ4 blocks of a slab,
// LUT^-1, D2H while the next slab is decoded. -1: not that kind of blob.
int decode_streamed_codes(Arena& A, const Pipeline& pl, const uint8_t* d_payload, uint64_t payload_bytes, uint16_t* d_dst, uint64_t N,
                          cudaStream_t st, HostSink* host, uint64_t slab_voxels, uint64_t min_bytes) {
  if (pl.elem != 2 || !pl.has_sink || pl.sink.kind != StageKind::Quantiser || !pl.has_tail || !pl.head.empty() || g_timing.load()) return -1;
  if (!pl.sink.has_decode_lut || 2 * N < min_bytes || N % kLz4BlockBytes) return -1;
  void *ws = nullptr, *codes_v = nullptr, *sp = nullptr;
  if (A.get(kSlotWs, k_lz4_decode_workspace_bytes(N), &ws) || A.get(kSlotA, N, &codes_v) ||
      A.get(kSlotSmall, 4 * 65536 * sizeof(uint32_t) + 4096, &sp))
    return 1;
  uint8_t* codes = static_cast<uint8_t*>(codes_v);
  uint16_t* d_lut = reinterpret_cast<uint16_t*>(static_cast<uint8_t*>(sp) + 4 * 65536 * sizeof(uint32_t));
  CK(cudaMemcpyAsync(d_lut, pl.sink.decode_lut, 512, cudaMemcpyHostToDevice, st));
  CKK(k_lz4_decode_tables(d_payload, payload_bytes, N, ws, 0, 0, st));
  uint32_t err = 0, own = 0, nblk = 0, bb = 0;
  if (k_lz4_decode_peek(ws, &err, &own, &nblk, &bb, st)) return 1;
  if (err || !own || bb != kLz4BlockBytes || (uint64_t)nblk * bb != N) return -1;
  cudaStream_t cs = nullptr;
  if (copy_stream(&cs)) return 1;
  uint64_t slab = slab_voxels;
  while ((N + slab - 1) / slab > 32) slab *= 2;   // one work counter per block-set launch
  EventList events;
  uint32_t slot = 0;
  for (uint64_t first = 0; first < N; first += slab, ++slot) {
    const uint64_t count = std::min(slab, N - first);
    CKK(k_lz4_decode_run(d_payload, payload_bytes, codes, N, ws, (uint32_t)(first / kLz4BlockBytes), (uint32_t)(count / kLz4BlockBytes), 1u, 0u,
                         slot, st));
    CKK(k_lut_decode(codes + first, d_dst + first, count, d_lut, st));
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
  if (err || total != N) {
    std::fprintf(stderr, "[sqeazy_b200] lz4 decode failed (code %u)\n", err);
    return 101;
  }
  host->done = true;
  return 0;
}

// -1: not shardable (single-device route), 0 ok, > 0 failure
int host_decode_sharded(const Header& hdr, const Pipeline& pl, const char* src, uint64_t srclength, char* dst, int nthreads) {
  const ShardKind kind = shard_kind(pl, hdr.shape);
  if (kind == ShardKind::None || g_timing.load()) return -1;
  const uint64_t N = shape_product(hdr.shape), raw_bytes = 2 * N;
  if (N == 0 || N % kStreamGrainVoxels) return -1;
  std::vector<int> devs = shard_devices();
  if (devs.empty()) return -1;
  ShardPlan plan;
  if (!plan_shards(N, kind == ShardKind::PlanesLz4 ? raw_bytes : N, kStreamGrainVoxels, devs, plan)) return -1;
  const int G = plan.G();
  if (G < 2) return -1;

  // the blob's own index frame (host memory): block sizes of the whole stream
  const uint8_t* payload = reinterpret_cast<const uint8_t*>(src) + hdr.size;
  const uint64_t payload_bytes = srclength - hdr.size;
  const int w = kind == ShardKind::PlanesLz4 ? pl.head.back().w : 0;
  const uint32_t P = kind == ShardKind::PlanesLz4 ? (uint32_t)(16 / w) : 1u;
  const uint64_t stream_bytes = kind == ShardKind::PlanesLz4 ? raw_bytes : N;
  const uint64_t nblocks = stream_bytes / kLz4BlockBytes, bpp = nblocks / P;
  auto rd32 = [&](uint64_t off) { uint32_t v; std::memcpy(&v, payload + off, 4); return v; };
  auto rd64 = [&](uint64_t off) { uint64_t v; std::memcpy(&v, payload + off, 8); return v; };
  const uint64_t prefix = lz4_prefix_bytes(nblocks);
  if (payload_bytes < prefix + kLz4EndMarkBytes) return -1;
  if (rd32(0) != kSqybSkippableMagic || rd32(4) != sizeof(SqybIndexHeader) + 4 * nblocks || rd32(8) != kSqybIndexMagic || rd32(12) != 1u ||
      rd32(16) != (uint32_t)kLz4BlockBytes || rd32(20) != (uint32_t)nblocks || rd64(24) != stream_bytes)
    return -1;   // a foreign stream (e.g. made by the reference): one device decodes it
  const uint64_t frame_bytes = rd64(32);
  if (rd32(prefix - kLz4FrameHeaderBytes) != kLz4FrameMagic || prefix - kLz4FrameHeaderBytes + frame_bytes != payload_bytes) return -1;
  std::vector<uint32_t> words(nblocks);
  std::memcpy(words.data(), payload + kSkippableHeaderBytes + sizeof(SqybIndexHeader), 4 * nblocks);
  // byte range of every (plane, GPU) piece behind the frame header
  std::vector<uint64_t> piece_off((size_t)P * G + 1);
  {
    uint64_t off = 0;
    size_t q = 0;
    for (uint32_t p = 0; p < P; ++p)
      for (int g = 0; g < G; ++g) {
        piece_off[q++] = off;
        const uint64_t k0 = (kind == ShardKind::PlanesLz4 ? 2 * plan.first[g] / P : plan.first[g]) / kLz4BlockBytes;
        const uint64_t k1 = (kind == ShardKind::PlanesLz4 ? 2 * plan.first[g + 1] / P : plan.first[g + 1]) / kLz4BlockBytes;
        for (uint64_t k = p * bpp + k0; k < p * bpp + k1; ++k) {
          const uint32_t sz = words[k] & 0x7FFFFFFFu;
          if (sz > (uint32_t)kLz4BlockBytes) return 11;   // untrusted index
          off += 4ull + sz;
        }
      }
    piece_off[q] = off;
    if (kLz4FrameHeaderBytes + off + kLz4EndMarkBytes != frame_bytes) return 11;   // the sizes must add up to the frame
  }

  DevSetLock locks;
  locks.acquire(plan.dev);
  const int T = std::max(1, staging_threads(nthreads) / G);
  std::atomic<int> fail{0};

  auto worker = [&](int g) -> int {
    if (cudaSetDevice(plan.dev[g]) != cudaSuccess) return 1;
    Arena& A = g_dev[plan.dev[g]].arena;
    cudaStream_t st = nullptr;
    const uint64_t a = plan.first[g], n = plan.count(g);
    const uint64_t sbytes = kind == ShardKind::PlanesLz4 ? 2 * n : n, nblk = sbytes / kLz4BlockBytes, lbpp = nblk / P;
    const uint64_t k0 = (kind == ShardKind::PlanesLz4 ? 2 * a / P : a) / kLz4BlockBytes;
    uint64_t body = 0;
    for (uint32_t p = 0; p < P; ++p) body += piece_off[(size_t)p * G + g + 1] - piece_off[(size_t)p * G + g];
    // local blob: [skippable header | index header | this GPU's block words | frame header | this GPU's pieces | EndMark]
    const uint64_t lprefix = lz4_prefix_bytes(nblk), lbytes = lprefix + body + kLz4EndMarkBytes;
    void *d_in = nullptr, *d_out = nullptr;
    if (A.get(kSlotIn, lbytes + 256, &d_in) || A.get(kSlotOut, 2 * n, &d_out)) return 1;
    uint8_t* d_blob = static_cast<uint8_t*>(d_in);
    std::vector<uint8_t> head(lprefix);
    {
      auto put32 = [&](uint64_t off, uint32_t v) { std::memcpy(head.data() + off, &v, 4); };
      auto put64 = [&](uint64_t off, uint64_t v) { std::memcpy(head.data() + off, &v, 8); };
      put32(0, kSqybSkippableMagic);
      put32(4, (uint32_t)(sizeof(SqybIndexHeader) + 4 * nblk));
      put32(8, kSqybIndexMagic);
      put32(12, 1u);
      put32(16, (uint32_t)kLz4BlockBytes);
      put32(20, (uint32_t)nblk);
      put64(24, sbytes);
      put64(32, kLz4FrameHeaderBytes + body + kLz4EndMarkBytes);
      for (uint32_t p = 0; p < P; ++p)
        std::memcpy(head.data() + kSkippableHeaderBytes + sizeof(SqybIndexHeader) + 4 * p * lbpp, words.data() + p * bpp + k0, 4 * lbpp);
      put32(lprefix - kLz4FrameHeaderBytes, kLz4FrameMagic);
      head[lprefix - 3] = 0x60;
      head[lprefix - 2] = 0x40;
      head[lprefix - 1] = 0x82;
    }
    CK(cudaMemcpyAsync(d_blob, head.data(), lprefix, cudaMemcpyHostToDevice, st));
    uint64_t loc = lprefix;
    for (uint32_t p = 0; p < P; ++p) {
      const uint64_t o = piece_off[(size_t)p * G + g], len = piece_off[(size_t)p * G + g + 1] - o;
      CK((cudaError_t)staged_h2d(d_blob + loc, payload + prefix + o, len, T, st));
      loc += len;
    }
    CK(cudaMemsetAsync(d_blob + loc, 0, kLz4EndMarkBytes, st));
    CK(cudaStreamSynchronize(st));   // `head` (pageable) has been consumed
    HostSink sink{dst + 2 * a, T};
    int rc;
    if (kind == ShardKind::PlanesLz4)
      rc = decode_streamed(A, pl, d_blob, lbytes, static_cast<uint16_t*>(d_out), n, st, &sink, kShardSubSlabVoxels, 0);
    else
      rc = decode_streamed_codes(A, pl, d_blob, lbytes, static_cast<uint16_t*>(d_out), n, st, &sink, kShardSubSlabVoxels, 0);
    if (rc < 0) {   // not streamable after all: this GPU's stream as a whole
      Header local = hdr;
      local.shape.assign(1, n);
      rc = decode_device_impl(A, local, pl, d_blob, lbytes, d_out, 2 * n, st, nullptr);
      if (rc) return rc;
      CK((cudaError_t)staged_d2h(dst + 2 * a, d_out, 2 * n, T, st));
      CK(cudaStreamSynchronize(st));
    }
    return rc;
  };
  std::vector<int> rcs((size_t)G, 0);
  std::vector<std::thread> th;
  for (int g = 1; g < G; ++g) th.emplace_back([&, g] { rcs[g] = worker(g); });
  rcs[0] = worker(0);
  for (auto& t : th) t.join();
  for (int g = 0; g < G; ++g)
    if (rcs[g]) return rcs[g];
  {
    std::lock_guard<std::mutex> lk(g_shard_stats_mu);
    g_shard_stats.gpus = G;
    g_shard_stats.nccl = 0;
    g_shard_stats.pieces = (long)P * G;
  }
  return 0;
}
