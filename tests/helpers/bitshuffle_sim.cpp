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
