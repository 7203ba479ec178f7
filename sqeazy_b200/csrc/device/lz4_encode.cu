// Chunked LZ4 block encoder for sm_100a — one CTA per 16 KiB block, everything in shared memory.
//
// Replaces the reference's per-chunk liblz4 calls + serial compaction
// (encoders/lz4.hpp:214-242, encoders/lz4_utils.hpp:99-173 encode_serial, :193-274 encode_parallel,
// :175-190 remove_blanks). Output = valid LZ4 frame bytes (see lz4_format.h); compressed-byte parity
// with liblz4 is not defined by the reference — validity, round trip through the
// reference decoder and the compression ratio are.
//
// Per block (16 KiB of input staged in shared memory by 512 threads):
//   load    : 128-bit coalesced loads -> smem; all-equal blocks take a closed-form path
//   phase A : every position finds a match candidate in parallel: 32 rounds of 512 positions against a
//             4096-entry shared-memory hash table (positions of earlier rounds) plus register-only
//             checks of offsets 1..4 for runs inside the current round
//   phase B : 16 warps, one 1 KiB sub-block each, walk their candidates greedily with ballots, extend
//             the selected matches 128 bytes per step, and record (offset,length) in place
//   combine : one thread chains the 16 sub-block summaries (literal carry, output offsets)
//   phase C : warps emit tokens/literals/offsets into a shared-memory output buffer at scanned offsets
//   store   : single-pass decoupled look-back over block sizes gives the final byte offset; the CTA
//             writes its block header + bytes once, coalesced, straight into the frame (no
//             remove_blanks pass, no strided scratch)
#include "common.cuh"
#include "kernels.h"
#include "lz4_format.h"

namespace sqyb {
namespace {

constexpr int kB = kLz4BlockBytes;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kSub = kB / kWarps;        // 1024 positions per warp
constexpr int kWinPerSub = kSub / 32;    // 32 ballot windows per sub-block
constexpr int kPad = 256;
constexpr uint32_t kNone = 0xFFFFu;
constexpr int kHashLog = 12;

static_assert(kWinPerSub == 32, "one selection mask per lane");

struct __align__(16) EncSmem {
  uint32_t data[(kB + kPad) / 4];
  uint32_t out[(kB + 64) / 4];
  uint16_t cand[kB];
  uint16_t htab[1 << kHashLog];
  int sb_nseq[kWarps];
  int sb_first_lit[kWarps];
  int sb_rest[kWarps];
  int sb_tail[kWarps];
  int sb_carry[kWarps];
  int sb_out_off[kWarps];
  int final_off;
  int final_lit;
  int total;
  uint32_t ticket;
  unsigned long long goff;
};

__device__ __forceinline__ int ext_bytes(int v) {
  if (v < 15) return 0;          // by far the common case: keep the division off the hot path
  if (v < 270) return 1;
  return 1 + (v - 15) / 255;
}

__device__ __forceinline__ uint32_t load4(const uint32_t* words, int byte_off) {
  const uint32_t w0 = words[byte_off >> 2], w1 = words[(byte_off >> 2) + 1];
  return __funnelshift_r(w0, w1, (byte_off & 3) * 8);
}

// writes the LZ4 length extension for `v` (nibble already holds 15) at p, returns bytes written
__device__ __forceinline__ int put_ext(uint8_t* p, int v) {
  int r = v - 15, k = 0;
  while (r >= 255) { p[k++] = 255; r -= 255; }
  p[k++] = (uint8_t)r;
  return k;
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPre = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;

// coalesced copy of nbytes from shared memory (word array, arbitrary byte offset 0) to an arbitrarily
// aligned global address
__device__ __forceinline__ void store_bytes(uint8_t* __restrict__ g, const uint32_t* __restrict__ sw, int nbytes, int tid) {
  const uint8_t* sb = reinterpret_cast<const uint8_t*>(sw);
  int head = (int)((16 - ((uintptr_t)g & 15)) & 15);
  if (head > nbytes) head = nbytes;
  if (tid < head) g[tid] = sb[tid];
  const int body = (nbytes - head) >> 4;
  uint4* g4 = reinterpret_cast<uint4*>(g + head);
  for (int v = tid; v < body; v += kThreads) {
    const int so = head + (v << 4);
    const int w = so >> 2, sh = (so & 3) * 8;
    const uint32_t x0 = sw[w], x1 = sw[w + 1], x2 = sw[w + 2], x3 = sw[w + 3], x4 = sw[w + 4];
    uint4 val;
    val.x = __funnelshift_r(x0, x1, sh);
    val.y = __funnelshift_r(x1, x2, sh);
    val.z = __funnelshift_r(x2, x3, sh);
    val.w = __funnelshift_r(x3, x4, sh);
    g4[v] = val;
  }
  const int done = head + (body << 4);
  if (tid < nbytes - done) g[done + tid] = sb[done + tid];
}

__global__ void __launch_bounds__(kThreads, 3)
lz4_encode_kernel(const uint8_t* __restrict__ src, uint64_t raw_bytes, uint8_t* __restrict__ dst, uint32_t nblocks,
                  unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket_counter,
                  unsigned long long* __restrict__ payload_bytes_out, uint32_t* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EncSmem& S = *reinterpret_cast<EncSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) S.ticket = atomicAdd(ticket_counter, 1u);
  __syncthreads();
  const uint32_t b = S.ticket;
  if (b >= nblocks) return;
  const uint64_t boff = (uint64_t)b * kB;
  const int n = (int)((raw_bytes - boff) < (uint64_t)kB ? (raw_bytes - boff) : (uint64_t)kB);
  const uint8_t* bsrc = src + boff;
  uint8_t* data8 = reinterpret_cast<uint8_t*>(S.data);
  uint8_t* out8 = reinterpret_cast<uint8_t*>(S.out);

  // ---------------- load ----------------
  bool same = true;
  if ((((uintptr_t)bsrc) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(bsrc);
    const uint32_t pat = (uint32_t)bsrc[0] * 0x01010101u;
    uint4* d4 = reinterpret_cast<uint4*>(S.data);
#pragma unroll
    for (int k = 0; k < kB / 16 / kThreads; ++k) {
      const int v = tid + k * kThreads;
      uint4 x = make_uint4(0, 0, 0, 0);
      if ((v + 1) * 16 <= n) {
        x = ld_stream(s4 + v);
        same = same && (x.x == pat) && (x.y == pat) && (x.z == pat) && (x.w == pat);
      } else if (v * 16 < n) {
        uint8_t tmp[16];
        for (int q = 0; q < 16; ++q) {
          const bool in = v * 16 + q < n;
          tmp[q] = in ? bsrc[v * 16 + q] : 0;
          same = same && (!in || tmp[q] == (uint8_t)pat);
        }
        x.x = tmp[0] | (tmp[1] << 8) | (tmp[2] << 16) | ((uint32_t)tmp[3] << 24);
        x.y = tmp[4] | (tmp[5] << 8) | (tmp[6] << 16) | ((uint32_t)tmp[7] << 24);
        x.z = tmp[8] | (tmp[9] << 8) | (tmp[10] << 16) | ((uint32_t)tmp[11] << 24);
        x.w = tmp[12] | (tmp[13] << 8) | (tmp[14] << 16) | ((uint32_t)tmp[15] << 24);
      }
      d4[v] = x;
    }
  } else {
    const uint8_t first = bsrc[0];
    for (int i = tid; i < kB; i += kThreads) {
      const uint8_t x = i < n ? bsrc[i] : 0;
      data8[i] = x;
      same = same && (i >= n || x == first);
    }
  }
  if (tid < kPad / 4) S.data[kB / 4 + tid] = 0;
  for (int i = tid; i < (1 << kHashLog) / 2; i += kThreads) reinterpret_cast<uint32_t*>(S.htab)[i] = 0xFFFFFFFFu;
  const int all_same = __syncthreads_and(same ? 1 : 0);

  int csize = 0;        // encoded bytes (without the 4-byte block header)
  bool stored = false;
  int kind = 0;         // stats: 0 general, 1 constant, 2 stored

  if (all_same && n >= 16) {
    // ---------------- closed form: 1 literal, match(offset 1, n-6), 5 literals ----------------
    kind = 1;
    if (warp == 0) {
      const uint8_t v = data8[0];
      const int mlen = n - 6;
      const int e = mlen - 4;
      const int nx = ext_bytes(e);
      if (lane == 0) {
        out8[0] = (uint8_t)((1 << 4) | (e < 15 ? e : 15));
        out8[1] = v;
        out8[2] = 1;
        out8[3] = 0;
      }
      if (e >= 15) {
        const int r = e - 15;
        for (int k = lane; k < nx; k += 32) out8[4 + k] = (k < nx - 1) ? 255 : (uint8_t)(r % 255);
      }
      if (lane < 6) out8[4 + nx + lane] = lane == 0 ? (uint8_t)(5 << 4) : v;
      if (lane == 0) S.total = 4 + nx + 6;
    }
    __syncthreads();
    csize = S.total;
  } else {
    // ---------------- phase A: candidates ----------------
    int nfound = 0;
    for (int r = 0; r < kB / kThreads; ++r) {
      const int i = r * kThreads + tid;
      const int wi = i >> 2, sh = (i & 3) * 8;
      const uint32_t w0 = S.data[wi], w1 = S.data[wi + 1];
      const uint32_t wp = wi > 0 ? S.data[wi - 1] : 0u;
      const uint32_t v = __funnelshift_r(w0, w1, sh);    // bytes i .. i+3
      const uint32_t lo = __funnelshift_r(wp, w0, sh);   // bytes i-4 .. i-1
      const uint32_t h = (v * 2654435761u) >> (32 - kHashLog);
      uint32_t found = kNone;
      if (i + kLz4MFLimit <= n) {
        // short offsets first: in bit-plane data runs and 2/4-byte periods give the long matches, while a
        // table hit from an earlier round is often a stale 4-byte coincidence (tools/lz4_model.c: +20 % ratio
        // on background-removed stacks)
        if (i >= 1 && v == ((lo >> 24) | (v << 8))) found = i - 1;
        else if (i >= 2 && v == ((lo >> 16) | (v << 16))) found = i - 2;
        else if (i >= 4 && v == lo) found = i - 4;
        else if (i >= 3 && v == ((lo >> 8) | (v << 24))) found = i - 3;
        else {
          const uint32_t c = S.htab[h];
          if (c != kNone && load4(S.data, (int)c) == v) found = c;
        }
      }
      S.cand[i] = (uint16_t)found;
      nfound += found != kNone;
      __syncthreads();
      if (i + 4 <= n) S.htab[h] = (uint16_t)i;
      __syncthreads();
    }
    const int any_found = __syncthreads_or(nfound);

    if (!any_found) {
      stored = true;
    } else {
      // ---------------- phase B: greedy selection per 1 KiB sub-block ----------------
      const int sub_lo = warp * kSub;
      const int sub_hi = min(sub_lo + kSub, n);
      uint32_t selmask = 0;  // lane w holds the selection mask of window w
      int nseq = 0, first_lit = 0, rest = 0, anchor = sub_lo;
      if (sub_lo < n) {
        const int match_end_limit = min(sub_hi, n - kLz4LastLiterals);
        int pos = sub_lo;
        for (int win = 0; win < kWinPerSub; ++win) {
          const int base = sub_lo + win * 32;
          if (base >= sub_hi) break;
          const uint32_t c = (base + lane < sub_hi) ? (uint32_t)S.cand[base + lane] : kNone;
          const uint32_t m = __ballot_sync(0xffffffffu, c != kNone);
          while (true) {
            const int rel = pos - base;
            if (rel >= 32) break;
            const uint32_t mm = rel > 0 ? (m & (0xffffffffu << rel)) : m;
            if (!mm) break;
            const int j = __ffs(mm) - 1;
            const int mpos = base + j;
            const int mc = (int)__shfl_sync(0xffffffffu, c, j);
            // cooperative extension, 128 bytes per step
            const int maxlen = match_end_limit - mpos;   // >= 4 is not guaranteed near the sub-block end
            int len = maxlen;
            if (maxlen >= kLz4MinMatch) {
              int done = 4;
              while (done < maxlen) {
                const int k = done + lane * 4;
                const uint32_t x = load4(S.data, mpos + k), y = load4(S.data, mc + k);
                const uint32_t diff = x ^ y;
                const uint32_t bm = __ballot_sync(0xffffffffu, diff != 0);
                if (bm) {
                  const int f = __ffs(bm) - 1;
                  const uint32_t d = __shfl_sync(0xffffffffu, diff, f);
                  len = done + f * 4 + ((__ffs(d) - 1) >> 3);
                  break;
                }
                done += 128;
              }
              if (len > maxlen) len = maxlen;
              const int lit = mpos - anchor;
              if (nseq == 0) {
                first_lit = lit;
                rest += 1 + 2 + ext_bytes(len - 4);
              } else {
                rest += 1 + ext_bytes(lit) + lit + 2 + ext_bytes(len - 4);
              }
              if (lane == 0) {
                S.cand[mpos] = (uint16_t)(mpos - mc);
                S.cand[mpos + 1] = (uint16_t)len;
              }
              if (lane == win) selmask |= 1u << j;
              nseq++;
              anchor = pos = mpos + len;
            } else {
              pos = mpos + 1;  // too close to the sub-block end: leave as literal
            }
          }
        }
      }
      if (lane == 0) {
        S.sb_nseq[warp] = nseq;
        S.sb_first_lit[warp] = first_lit;
        S.sb_rest[warp] = rest;
        S.sb_tail[warp] = (sub_lo < n) ? sub_hi - anchor : 0;
      }
      __syncthreads();
      // ---------------- combine ----------------
      if (tid == 0) {
        int carry = 0, off = 0;
        for (int k = 0; k < kWarps; ++k) {
          S.sb_carry[k] = carry;
          S.sb_out_off[k] = off;
          if (S.sb_nseq[k] == 0) {
            carry += S.sb_tail[k];
          } else {
            const int L = carry + S.sb_first_lit[k];
            off += ext_bytes(L) + L + S.sb_rest[k];
            carry = S.sb_tail[k];
          }
        }
        S.final_off = off;
        S.final_lit = carry;
        S.total = off + 1 + ext_bytes(carry) + carry;
      }
      __syncthreads();
      csize = S.total;
      if (csize >= n) {
        stored = true;
      } else {
        // ---------------- phase C: emission ----------------
        if (nseq > 0) {
          int a_run = sub_lo - S.sb_carry[warp];
          int ooff = S.sb_out_off[warp];
          for (int win = 0; win < kWinPerSub; ++win) {
            const uint32_t mask = __shfl_sync(0xffffffffu, selmask, win);
            if (!mask) continue;
            const int base = sub_lo + win * 32;
            const bool sel = (mask >> lane) & 1u;
            const int mpos = base + lane;
            const int off = sel ? (int)S.cand[mpos] : 0;
            const int len = sel ? (int)S.cand[mpos + 1] : 0;
            const int end = mpos + len;
            const uint32_t prevmask = mask & ((1u << lane) - 1u);
            const int prevlane = prevmask ? 31 - __clz(prevmask) : 0;
            const int prev_end = __shfl_sync(0xffffffffu, end, prevlane);
            const int a = prevmask ? prev_end : a_run;
            const int lit = sel ? mpos - a : 0;
            const int elit = ext_bytes(lit), elen = ext_bytes(len - 4);
            const int size = sel ? 1 + elit + lit + 2 + elen : 0;
            int incl = size;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, incl, d);
              if (lane >= d) incl += t;
            }
            const int my_out = ooff + incl - size;
            if (sel) {
              uint8_t* p = out8 + my_out;
              const int ml = len - 4;
              p[0] = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (ml < 15 ? ml : 15));
              if (lit >= 15) put_ext(p + 1, lit);
              uint8_t* q = p + 1 + elit + lit;
              q[0] = (uint8_t)(off & 0xff);
              q[1] = (uint8_t)(off >> 8);
              if (ml >= 15) put_ext(q + 2, ml);
            }
            // literals: the whole warp copies each selected sequence's run
            uint32_t rem = mask;
            while (rem) {
              const int j = __ffs(rem) - 1;
              rem &= rem - 1;
              const int la = __shfl_sync(0xffffffffu, a, j);
              const int ll = __shfl_sync(0xffffffffu, lit, j);
              const int lo = __shfl_sync(0xffffffffu, my_out, j) + 1 + ext_bytes(ll);
              for (int k = lane; k < ll; k += 32) out8[lo + k] = data8[la + k];
            }
            const int last = 31 - __clz(mask);
            a_run = __shfl_sync(0xffffffffu, end, last);
            ooff += __shfl_sync(0xffffffffu, incl, 31);
          }
        }
        // final literal-only sequence
        {
          const int L = S.final_lit;
          uint8_t* p = out8 + S.final_off;
          const int e = ext_bytes(L);
          if (tid == 0) {
            p[0] = (uint8_t)((L < 15 ? L : 15) << 4);
            if (L >= 15) put_ext(p + 1, L);
          }
          for (int k = tid; k < L; k += kThreads) p[1 + e + k] = data8[n - L + k];
        }
      }
    }
    if (stored) { csize = n; kind = 2; }
  }
  __syncthreads();

  // ---------------- decoupled look-back over (4 + csize) ----------------
  const unsigned long long mine = 4ull + (unsigned long long)csize;
  if (warp == 0) {
    unsigned long long excl = 0;
    if (b == 0) {
      if (lane == 0) st_release_u64(status + 0, kFlagPre | mine);
    } else {
      if (lane == 0) st_release_u64(status + b, kFlagAgg | mine);
      long long look = (long long)b - 1;
      while (true) {
        const long long idx = look - lane;
        unsigned long long st = kFlagPre;  // virtual predecessor of block 0: prefix 0
        if (idx >= 0) {
          st = ld_acquire_u64(status + idx);
          while ((st >> 62) == 0) { __nanosleep(256); st = ld_acquire_u64(status + idx); }
        }
        const uint32_t pm = __ballot_sync(0xffffffffu, (st >> 62) == 2);
        const int first = pm ? __ffs(pm) - 1 : 32;
        unsigned long long contrib = lane <= first ? (st & kValMask) : 0ull;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
        excl += contrib;
        if (pm) break;
        look -= 32;
      }
      if (lane == 0) st_release_u64(status + b, kFlagPre | (excl + mine));
    }
    if (lane == 0) S.goff = excl;
  }
  __syncthreads();

  // ---------------- store ----------------
  const unsigned long long goff = lz4_prefix_bytes(nblocks) + S.goff;
  const uint32_t word = stored ? ((uint32_t)n | kLz4StoredFlag) : (uint32_t)csize;
  uint8_t* g = dst + goff;
  if (tid < 4) g[tid] = (uint8_t)(word >> (8 * tid));
  store_bytes(g + 4, stored ? S.data : S.out, csize, tid);
  // index entry
  if (tid == 0) {
    uint8_t* ip = dst + kSkippableHeaderBytes + sizeof(SqybIndexHeader) + 4ull * b;
    ip[0] = (uint8_t)word; ip[1] = (uint8_t)(word >> 8); ip[2] = (uint8_t)(word >> 16); ip[3] = (uint8_t)(word >> 24);
    if (stats) atomicAdd(stats + kind, 1u);
  }
  if (b == nblocks - 1 && tid == 0) {
    const unsigned long long end = goff + mine;
    for (int k = 0; k < 4; ++k) dst[end + k] = 0;  // EndMark
    *payload_bytes_out = end + 4;
    // frame_bytes field of the index header
    const unsigned long long frame_bytes = end + 4 - (lz4_prefix_bytes(nblocks) - kLz4FrameHeaderBytes);
    uint8_t* fp = dst + kSkippableHeaderBytes + 24;
    for (int k = 0; k < 8; ++k) fp[k] = (uint8_t)(frame_bytes >> (8 * k));
  }
}

// static prefix: skippable header, index header (frame_bytes patched by the last block), frame header
__global__ void lz4_write_prefix_kernel(uint8_t* dst, uint64_t raw_bytes, uint32_t nblocks,
                                        unsigned long long* payload_bytes_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  auto put32 = [&](uint64_t off, uint32_t v) { for (int k = 0; k < 4; ++k) dst[off + k] = (uint8_t)(v >> (8 * k)); };
  auto put64 = [&](uint64_t off, uint64_t v) { for (int k = 0; k < 8; ++k) dst[off + k] = (uint8_t)(v >> (8 * k)); };
  put32(0, kSqybSkippableMagic);
  put32(4, (uint32_t)(sizeof(SqybIndexHeader) + 4ull * nblocks));
  put32(8, kSqybIndexMagic);
  put32(12, 1u);
  put32(16, (uint32_t)kB);
  put32(20, nblocks);
  put64(24, raw_bytes);
  put64(32, kLz4FrameHeaderBytes + kLz4EndMarkBytes);
  const uint64_t f = kSkippableHeaderBytes + sizeof(SqybIndexHeader) + 4ull * nblocks;
  put32(f, kLz4FrameMagic);
  dst[f + 4] = 0x60;  // FLG: version 01, block independent
  dst[f + 5] = 0x40;  // BD: 64 KiB max block
  dst[f + 6] = 0x82;  // HC = (xxh32(FLG BD) >> 8) & 0xff
  if (nblocks == 0) {
    put32(f + 7, 0);
    *payload_bytes_out = f + 7 + 4;
  }
}

}  // namespace

size_t k_lz4_encode_workspace_bytes(uint64_t raw_bytes) {
  return 64 + 8 * lz4_nblocks(raw_bytes) + 64;
}

// workspace layout: [0,8) payload bytes (u64) | [8,12) ticket | [16,32) stats | [64, ...) status[nblocks]
int k_lz4_encode(const uint8_t* src, uint64_t raw_bytes, uint8_t* dst, void* workspace, cudaStream_t st) {
  const uint64_t nb64 = lz4_nblocks(raw_bytes);
  if (nb64 > 0xFFFFFFF0ull) return -2;
  const uint32_t nblocks = (uint32_t)nb64;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  SQYB_CUDA_OK(cudaMemsetAsync(ws, 0, 64 + 8ull * nblocks, st));
  unsigned long long* payload = reinterpret_cast<unsigned long long*>(ws);
  lz4_write_prefix_kernel<<<1, 32, 0, st>>>(dst, raw_bytes, nblocks, payload);
  SQYB_COUNT_LAUNCH(nblocks ? 2 : 1);
  if (nblocks) {
    SQYB_CUDA_OK(cudaFuncSetAttribute(lz4_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem)));
    lz4_encode_kernel<<<nblocks, kThreads, sizeof(EncSmem), st>>>(
        src, raw_bytes, dst, nblocks, reinterpret_cast<unsigned long long*>(ws + 64), reinterpret_cast<uint32_t*>(ws + 8),
        payload, reinterpret_cast<uint32_t*>(ws + 16));
  }
  return (int)cudaGetLastError();
}

}  // namespace sqyb
