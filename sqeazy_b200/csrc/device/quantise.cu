// Quantiser device kernels: the 65536-bin uint32 histogram and the two LUT gathers.
//   histogram : encoders/histogram_utils.hpp:41-55,98-154 (serial/parallel::fill_histogram),
//               called single-threaded by quantiser::computeHistogram (quantiser_utils.hpp:144-151)
//   LUT apply : quantiser_scheme_impl.hpp:206-223 (u16 -> 8-bit code, 65536-entry table)
//   LUT decode: quantiser_scheme_impl.hpp:258-279, quantiser_utils.hpp:26-42 (8-bit -> u16, 256 entries;
//               unsigned index — the reference's signed-char index is UB for codes >= 128)
//
// Histogram layout: 65536 x u32 does not fit one SM's shared memory (256 KiB > 227 KB), so each CTA
// privatises the low kSmemBins = 49152 bins (192 KiB, shared-memory atomics) and sends the rare values
// above that straight to global atomics; one CTA per SM, flushed once at the end. Counts are exact
// uint32 (mod 2^32, like the reference's bins).
#include "common.cuh"
#include "kernels.h"

namespace sqyb {
namespace {

constexpr int kSmemBins = 49152;
static_assert(kSmemBins == 0xC000, "the fast path of histogram_u16_kernel tests the two top bits");
constexpr int kHistThreads = 1024;
constexpr uint32_t kHighDirect = 64;  // voxels above the private bins a thread sends to global atomics before it defers the rest
constexpr uint64_t kHistAhead = 4;   // grid-stride iterations between a load and its L2 prefetch

__global__ void __launch_bounds__(kHistThreads, 1) histogram_u16_kernel(const uint16_t* __restrict__ in, uint64_t n,
                                                                        uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t sh[];
  for (int i = threadIdx.x; i < kSmemBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // head (unaligned) elements are handled by the scalar loop at the end via [0, lead)
  const uintptr_t addr = (uintptr_t)in;
  uint64_t lead = ((16 - (addr & 15)) & 15) / 2;
  if (lead > n) lead = n;
  const uint4* vin = reinterpret_cast<const uint4*>(in + lead);
  const uint64_t nv = (n - lead) / 8;
  auto add = [&](uint32_t v) {
    if (v < (uint32_t)kSmemBins) atomicAdd(&sh[v], 1u);
    else atomicAdd(&hist[v], 1u);
  };
  // Voxels above the private bins go straight to global atomics — fine for the few saturated pixels of a camera stack, a
  // cliff for a bright one (uniform-random voxels: 6.4 ms per 4 GiB against 0.7). A thread that has sent more than
  // kHighDirect of them stops doing so: from its next iteration on it leaves them out, and a second sweep over those
  // iterations counts them in shared memory (the 16384 bins above kSmemBins fit the table once the low bins are flushed).
  // Every voxel is counted once, by whichever path: the sums are exact whatever the threads decide.
  bool defer = false;
  uint32_t nhigh = 0;
  uint64_t i_defer = 0;                // first iteration whose high voxels are left to the second sweep
  auto add_main = [&](uint32_t v) {
    if (v < (uint32_t)kSmemBins) atomicAdd(&sh[v], 1u);
    else if (!defer) { atomicAdd(&hist[v], 1u); ++nhigh; }
  };
  for (uint64_t i = tid; i < nv; i += stride) {
    // One CTA of 1024 threads per SM has 16 KiB in flight — at DRAM latency that is 3.4 TB/s for the whole GPU, which is what
    // this kernel ran at. More loads per thread cost registers the 1024 threads do not have (profiles/README.md); a prefetch
    // into L2 costs none: one lane per 128-byte line asks for the line the CTA reads kHistAhead iterations later, and the
    // loads then see L2 latency (4 GiB of scmos voxels 1.24 -> 0.91 ms = 4.7 TB/s; 2, 4 and 8 iterations ahead alike).
    if ((threadIdx.x & 7) == 0 && i + kHistAhead * stride < nv) asm volatile("prefetch.global.L2 [%0];" ::"l"(vin + i + kHistAhead * stride));
    const uint4 v = ld_stream(vin + i);
    // All eight values in the private bins (below kSmemBins = 0xC000: a value reaches it when its two top bits are set)? Then
    // no per-voxel compare, branch and address multiply: byte addresses from a shift and a mask, eight shared atomics — 50
    // warp instructions per iteration instead of 74, and with the prefetch above the kernel had become bound by those
    // (issue slots 61 % busy at 32 warps per SM): 4 GiB of scmos voxels 0.91 -> 0.69 ms = 6.26 TB/s = 96 % of the measured copy
    // bandwidth. Stacks with voxels >= 49152 take the general path below for the vectors that hold one.
    const uint32_t top = ((v.x & (v.x << 1)) | (v.y & (v.y << 1)) | (v.z & (v.z << 1)) | (v.w & (v.w << 1))) & 0x80008000u;
    if (top == 0u) {
      auto lo = [&](uint32_t w) { atomicAdd(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(sh) + ((w << 2) & 0x3fffcu)), 1u); };
      auto hi = [&](uint32_t w) { atomicAdd(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(sh) + ((w >> 14) & 0x3fffcu)), 1u); };
      lo(v.x); hi(v.x); lo(v.y); hi(v.y); lo(v.z); hi(v.z); lo(v.w); hi(v.w);
      continue;
    }
    add_main(v.x & 0xffff); add_main(v.x >> 16);
    add_main(v.y & 0xffff); add_main(v.y >> 16);
    add_main(v.z & 0xffff); add_main(v.z >> 16);
    add_main(v.w & 0xffff); add_main(v.w >> 16);
    if (!defer && nhigh > kHighDirect) {
      defer = true;
      i_defer = i + stride;
    }
  }
  for (uint64_t i = tid; i < lead; i += stride) add(in[i]);
  for (uint64_t i = lead + nv * 8 + tid; i < n; i += stride) add(in[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < kSmemBins; i += blockDim.x) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&hist[i], c);
    if (i < 65536 - kSmemBins) sh[i] = 0;        // (for the second sweep: every thread clears what it has just read)
  }
  if (!__syncthreads_or(defer)) return;
  // second sweep: the high voxels of the iterations left out above, into shared-memory bins [kSmemBins, 65536)
  if (defer) {
    auto add_high = [&](uint32_t x) {
      if (x >= (uint32_t)kSmemBins) atomicAdd(&sh[x - (uint32_t)kSmemBins], 1u);
    };
    for (uint64_t i = i_defer; i < nv; i += stride) {
      const uint4 v = ld_stream(vin + i);
      const uint32_t top = ((v.x & (v.x << 1)) | (v.y & (v.y << 1)) | (v.z & (v.z << 1)) | (v.w & (v.w << 1))) & 0x80008000u;
      if (top == 0u) continue;
      add_high(v.x & 0xffff); add_high(v.x >> 16);
      add_high(v.y & 0xffff); add_high(v.y >> 16);
      add_high(v.z & 0xffff); add_high(v.z >> 16);
      add_high(v.w & 0xffff); add_high(v.w >> 16);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 65536 - kSmemBins; i += blockDim.x) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&hist[kSmemBins + i], c);
  }
}

// small ranges (rmestbkrd faces/rows): plain global atomics
__global__ void histogram_u16_small_kernel(const uint16_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ hist) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&hist[in[i]], 1u);
}

// calc_support (hist_impl.hpp:359-381) + support_index (:63-84) for one 65536-bin histogram per CTA: the first bin m
// whose cumulative share exceeds `thr`, and the two bin counts the support value is interpolated from. The sums are
// integers below 2^53, so the double-precision running sums of the reference are reproduced exactly in any order; the
// int-typed total wraps like std::accumulate(..., 0) does. out[4*h .. 4*h+3] = {m (65536 = none), bins[m], bins[m-1], 0}.
__global__ void __launch_bounds__(1024) support_index_kernel(const uint32_t* __restrict__ hist, float thr, uint32_t* __restrict__ out) {
  __shared__ unsigned long long ws64[32];
  __shared__ uint32_t ws32[32];
  __shared__ uint32_t best;
  const uint32_t* bins = hist + (size_t)blockIdx.x * 65536;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long s64 = 0;
  uint32_t s32 = 0;
  const uint4* mine = reinterpret_cast<const uint4*>(bins + tid * 64);
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const uint4 v = mine[k];
    s64 += (unsigned long long)v.x + v.y + v.z + v.w;
    s32 += v.x + v.y + v.z + v.w;
  }
  unsigned long long inc = s64;
  uint32_t tot = s32;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
    tot += __shfl_xor_sync(0xffffffffu, tot, d);
  }
  if (lane == 31) ws64[warp] = inc;
  if (lane == 0) ws32[warp] = tot;
  if (tid == 0) best = 65536u;
  __syncthreads();
  unsigned long long running = inc - s64;
  uint32_t isum = 0;
  for (int w = 0; w < 32; ++w) {
    if (w < warp) running += ws64[w];
    isum += ws32[w];
  }
  const double total = (double)(int)isum;
  const double limit = (double)thr;
  // only the thread whose 64 bins cross the limit has to walk them (the shares grow monotonically for total > 0; for a
  // wrapped, non-positive total every thread walks, as the reference's loop would)
  const bool may_cross = !(total > 0.0) || ((double)(running + s64) / total) > limit;
  if (may_cross) {
#pragma unroll 1
    for (int k = 0; k < 64; ++k) {
      running += bins[tid * 64 + k];
      if (((double)running / total) > limit) { atomicMin(&best, (uint32_t)(tid * 64 + k)); break; }
    }
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t m = best & 0xffffu;   // (uint16_t)support: 65536 -> 0
    out[4 * blockIdx.x + 0] = best;
    out[4 * blockIdx.x + 1] = m ? bins[m] : 0u;
    out[4 * blockIdx.x + 2] = m ? bins[m - 1] : 0u;
    out[4 * blockIdx.x + 3] = 0u;
  }
}

// u16 -> u8 through a 65536-entry table staged in shared memory (64 KiB)
__global__ void __launch_bounds__(512) lut_apply_kernel(const uint16_t* __restrict__ in, uint8_t* __restrict__ out,
                                                        uint64_t n, const uint8_t* __restrict__ lut) {
  extern __shared__ uint32_t shw[];
  uint8_t* s = reinterpret_cast<uint8_t*>(shw);
  {
    const uint4* l4 = reinterpret_cast<const uint4*>(lut);
    uint4* s4 = reinterpret_cast<uint4*>(shw);
    for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) s4[i] = l4[i];
  }
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool aligned = ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 7) == 0;
  const uint64_t nv = aligned ? n / 8 : 0;
  for (uint64_t i = tid; i < nv; i += stride) {
    const uint4 v = ld_stream(reinterpret_cast<const uint4*>(in) + i);
    uint2 o;
    o.x = s[v.x & 0xffff] | (s[v.x >> 16] << 8) | (s[v.y & 0xffff] << 16) | (s[v.y >> 16] << 24);
    o.y = s[v.z & 0xffff] | (s[v.z >> 16] << 8) | (s[v.w & 0xffff] << 16) | (s[v.w >> 16] << 24);
    st_stream(reinterpret_cast<uint2*>(out) + i, o);
  }
  for (uint64_t i = nv * 8 + tid; i < n; i += stride) out[i] = s[in[i]];
}

// u8 -> u16 through a 256-entry table
__global__ void __launch_bounds__(512) lut_decode_kernel(const uint8_t* __restrict__ in, uint16_t* __restrict__ out,
                                                         uint64_t n, const uint16_t* __restrict__ lut) {
  __shared__ uint16_t s[256];
  if (threadIdx.x < 256) s[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool aligned = ((uintptr_t)in & 7) == 0 && ((uintptr_t)out & 15) == 0;
  const uint64_t nv = aligned ? n / 8 : 0;
  for (uint64_t i = tid; i < nv; i += stride) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(in) + i);
    uint4 o;
    o.x = s[v.x & 0xff] | ((uint32_t)s[(v.x >> 8) & 0xff] << 16);
    o.y = s[(v.x >> 16) & 0xff] | ((uint32_t)s[v.x >> 24] << 16);
    o.z = s[v.y & 0xff] | ((uint32_t)s[(v.y >> 8) & 0xff] << 16);
    o.w = s[(v.y >> 16) & 0xff] | ((uint32_t)s[v.y >> 24] << 16);
    st_stream(reinterpret_cast<uint4*>(out) + i, o);
  }
  for (uint64_t i = nv * 8 + tid; i < n; i += stride) out[i] = s[in[i]];
}

}  // namespace

int k_histogram_u16(const uint16_t* in, uint64_t n, uint32_t* hist, cudaStream_t st) {
  if (n == 0) return 0;
  if (n < (1u << 20)) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    histogram_u16_small_kernel<<<blocks, 256, 0, st>>>(in, n, hist);
    SQYB_COUNT_LAUNCH(1);
    return (int)cudaGetLastError();
  }
  const size_t smem = kSmemBins * sizeof(uint32_t);
  SQYB_CUDA_OK(cudaFuncSetAttribute(histogram_u16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  histogram_u16_kernel<<<kNumSMs, kHistThreads, smem, st>>>(in, n, hist);
  SQYB_COUNT_LAUNCH(1);
  return (int)cudaGetLastError();
}

int k_support_index(const uint32_t* hist_dev, int nhist, float threshold, uint32_t* out_dev, cudaStream_t st) {
  if (nhist <= 0) return 0;
  support_index_kernel<<<nhist, 1024, 0, st>>>(hist_dev, threshold, out_dev);
  SQYB_COUNT_LAUNCH(1);
  return (int)cudaGetLastError();
}

int k_lut_apply(const uint16_t* in, uint8_t* out, uint64_t n, const uint8_t* lut_dev, cudaStream_t st) {
  if (n == 0) return 0;
  SQYB_CUDA_OK(cudaFuncSetAttribute(lut_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  uint64_t blocks = (n / 8 + 511) / 512;
  if (blocks > (uint64_t)kNumSMs * 3) blocks = kNumSMs * 3;  // 3 CTAs (64 KiB each) per SM, grid-stride
  if (blocks < 1) blocks = 1;
  lut_apply_kernel<<<(int)blocks, 512, 65536, st>>>(in, out, n, lut_dev);
  SQYB_COUNT_LAUNCH(1);
  return (int)cudaGetLastError();
}

int k_lut_decode(const uint8_t* in, uint16_t* out, uint64_t n, const uint16_t* lut_dev, cudaStream_t st) {
  if (n == 0) return 0;
  uint64_t blocks = (n / 8 + 511) / 512;
  if (blocks > (uint64_t)kNumSMs * 16) blocks = kNumSMs * 16;
  if (blocks < 1) blocks = 1;
  lut_decode_kernel<<<(int)blocks, 512, 0, st>>>(in, out, n, lut_dev);
  SQYB_COUNT_LAUNCH(1);
  return (int)cudaGetLastError();
}

}  // namespace sqyb
