// Register-level bit-matrix transpose shared by bitshuffle.cu and its CPU simulation (tests/helpers/bitshuffle_sim.cpp:
// the thread program of the fast kernels is replayed on the host, there is no GPU where the code is written).
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace sqyb {

// 16x16 bit-matrix transpose of the low halves and of the high halves of w[0..15] at once: afterwards bit c of
// half(w[i]) is what bit i of half(w[c]) was. An involution.
template <int J>
__host__ __device__ __forceinline__ void delta_swap_stage(uint32_t* w) {
  constexpr uint32_t m = J == 8 ? 0x00FF00FFu : (J == 4 ? 0x0F0F0F0Fu : (J == 2 ? 0x33333333u : 0x55555555u));
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if ((k & J) == 0) {
      const uint32_t t = ((w[k] >> J) ^ w[k + J]) & m;
      w[k + J] ^= t;
      w[k] ^= t << J;
    }
  }
}
__host__ __device__ __forceinline__ void transpose16x16_pairs(uint32_t* w) {
  delta_swap_stage<8>(w);
  delta_swap_stage<4>(w);
  delta_swap_stage<2>(w);
  delta_swap_stage<1>(w);
}

// 8x8 bit-matrix transpose of the 8 bytes of x (byte j = row j, least significant bit = column 0): afterwards bit j of byte r
// is what bit r of byte j was (the TRANS_BIT_8X8 step of the bitshuffle library on little-endian words). An involution.
__host__ __device__ __forceinline__ uint64_t transpose8x8(uint64_t x) {
  uint64_t t;
  t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;
  x ^= t ^ (t << 7);
  t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull;
  x ^= t ^ (t << 14);
  t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull;
  x ^= t ^ (t << 28);
  return x;
}

}  // namespace sqyb
