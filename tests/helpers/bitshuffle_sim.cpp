// CPU replay of the thread program of bitshuffle16_encode_fast / bitshuffle16_decode_fast (sqeazy_b200/csrc/device/bitshuffle.cu):
// same register transpose (bit_transpose16.h), same PRMT selectors, same addressing. Test infrastructure: it lets the
// index arithmetic be checked against the oracle where no GPU exists.  g++ -O2 -shared -fPIC -I sqeazy_b200/csrc/device
#include <cstdint>
#include <cstring>

#include "bit_transpose16.h"

static uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {   // PRMT, default mode
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}

extern "C" void sim_bitshuffle16_encode_fast(const uint16_t* in, uint8_t* out, uint64_t n32, uint32_t bs) {
  const uint32_t row_bytes = bs / 8, per_block = bs / 32;
  for (uint64_t t = 0; t < n32; ++t) {
    uint32_t L[16], w[16];
    std::memcpy(L, in + t * 32, 64);
    for (int k = 0; k < 8; ++k) {
      w[2 * k] = byte_perm(L[k], L[8 + k], 0x5410);
      w[2 * k + 1] = byte_perm(L[k], L[8 + k], 0x7632);
    }
    sqyb::transpose16x16_pairs(w);
    const uint64_t blk = t / per_block;
    const uint32_t c = (uint32_t)(t - blk * per_block);
    uint8_t* o = out + blk * 2ull * bs + 4u * c;
    for (int r = 0; r < 16; ++r) std::memcpy(o + (uint64_t)r * row_bytes, &w[r], 4);
  }
}

extern "C" void sim_bitshuffle16_decode_fast(const uint8_t* in, uint16_t* out, uint64_t n32, uint32_t bs) {
  const uint32_t row_bytes = bs / 8, per_block = bs / 32;
  for (uint64_t t = 0; t < n32; ++t) {
    const uint64_t blk = t / per_block;
    const uint32_t c = (uint32_t)(t - blk * per_block);
    const uint8_t* s = in + blk * 2ull * bs + 4u * c;
    uint32_t w[16], L[16];
    for (int r = 0; r < 16; ++r) std::memcpy(&w[r], s + (uint64_t)r * row_bytes, 4);
    sqyb::transpose16x16_pairs(w);
    for (int k = 0; k < 8; ++k) {
      L[k] = byte_perm(w[2 * k], w[2 * k + 1], 0x5410);
      L[8 + k] = byte_perm(w[2 * k], w[2 * k + 1], 0x7632);
    }
    std::memcpy(out + t * 32, L, 64);
  }
}

// ---- uint8 kernels (bitshuffle8_encode_fast / bitshuffle8_decode_fast) ----
extern "C" void sim_bitshuffle8_encode_fast(const uint8_t* in, uint8_t* out, uint64_t n32, uint32_t bs) {
  const uint32_t row_bytes = bs / 8, per_block = bs / 32;
  for (uint64_t t = 0; t < n32; ++t) {
    uint32_t L[8], lo[4], hi[4];
    std::memcpy(L, in + t * 32, 32);
    for (int g = 0; g < 4; ++g) {
      const uint64_t x = sqyb::transpose8x8((uint64_t)L[2 * g] | ((uint64_t)L[2 * g + 1] << 32));
      lo[g] = (uint32_t)x;
      hi[g] = (uint32_t)(x >> 32);
    }
    const uint64_t blk = t / per_block;
    const uint32_t c = (uint32_t)(t - blk * per_block);
    uint8_t* o = out + blk * (uint64_t)bs + 4u * c;
    for (int r = 0; r < 8; ++r) {
      const uint32_t* q = r < 4 ? lo : hi;
      const int b = r & 3;
      const uint32_t s01 = byte_perm(q[0], q[1], 0x0040 + b * 0x0011);
      const uint32_t s23 = byte_perm(q[2], q[3], 0x0040 + b * 0x0011);
      const uint32_t w = byte_perm(s01, s23, 0x5410);
      std::memcpy(o + (uint64_t)r * row_bytes, &w, 4);
    }
  }
}

extern "C" void sim_bitshuffle8_decode_fast(const uint8_t* in, uint8_t* out, uint64_t n32, uint32_t bs) {
  const uint32_t row_bytes = bs / 8, per_block = bs / 32;
  for (uint64_t t = 0; t < n32; ++t) {
    const uint64_t blk = t / per_block;
    const uint32_t c = (uint32_t)(t - blk * per_block);
    const uint8_t* s = in + blk * (uint64_t)bs + 4u * c;
    uint32_t w[8], L[8];
    for (int r = 0; r < 8; ++r) std::memcpy(&w[r], s + (uint64_t)r * row_bytes, 4);
    for (int g = 0; g < 4; ++g) {
      const uint32_t sel = 0x0040 + g * 0x0011;
      const uint32_t lo = byte_perm(byte_perm(w[0], w[1], sel), byte_perm(w[2], w[3], sel), 0x5410);
      const uint32_t hi = byte_perm(byte_perm(w[4], w[5], sel), byte_perm(w[6], w[7], sel), 0x5410);
      const uint64_t x = sqyb::transpose8x8((uint64_t)lo | ((uint64_t)hi << 32));
      L[2 * g] = (uint32_t)x;
      L[2 * g + 1] = (uint32_t)(x >> 32);
    }
    std::memcpy(out + t * 32, L, 32);
  }
}
