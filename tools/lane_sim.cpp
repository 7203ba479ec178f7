// CPU harness for the lane-serial LZ4 block decoder (sqeazy_b200/csrc/device/lz4_lane.inl): the state machine is
// compiled for the host and checked against blocks compressed by liblz4 (development tool, not shipped).
// build: g++ -O2 -std=c++17 -o /tmp/lane_sim tools/lane_sim.cpp -ldl ; usage: lane_sim file [block_bytes]
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

struct uint4 { uint32_t x, y, z, w; };
#define SQYB_LANE_FN static inline
constexpr uint32_t kLaneStride = 1;
static inline uint32_t lane_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh); }
static inline uint32_t lane_funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi << 32) | lo) << sh) >> 32); }
static inline int lane_ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline uint4 lane_load_chunk(const uint8_t* a, const uint8_t* sbeg, const uint8_t* send) {
  uint8_t b[16];
  for (int k = 0; k < 16; ++k) b[k] = (a + k >= sbeg && a + k < send) ? a[k] : 0;
  uint4 v;
  memcpy(&v, b, 16);
  return v;
}
static const uint8_t* g_dlo; static const uint8_t* g_dhi;
static inline uint32_t lane_load_out32(const uint8_t* p) {
  if (p < g_dlo || p + 4 > g_dhi || ((uintptr_t)p & 3)) { fprintf(stderr, "OOB/unaligned far read\n"); abort(); }
  uint32_t v; memcpy(&v, p, 4); return v;
}
static inline void lane_store_out16(uint8_t* p, uint4 v) {
  if (p < g_dlo || p + 16 > g_dhi || ((uintptr_t)p & 15)) { fprintf(stderr, "OOB/unaligned store\n"); abort(); }
  memcpy(p, &v, 16);
}
#include "../sqeazy_b200/csrc/device/lz4_lane.inl"

typedef int (*compress_fn)(const char*, char*, int, int, int);

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  void* h = dlopen("/usr/lib/x86_64-linux-gnu/liblz4.so.1", RTLD_NOW);
  if (!h) { fprintf(stderr, "no liblz4\n"); return 2; }
  compress_fn fast = (compress_fn)dlsym(h, "LZ4_compress_fast");
  compress_fn hc = (compress_fn)dlsym(h, "LZ4_compress_HC");
  FILE* f = fopen(argv[1], "rb");
  fseek(f, 0, SEEK_END); long total = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf(total);
  if (fread(buf.data(), 1, total, f) != (size_t)total) return 3;
  const int block = argc > 2 ? atoi(argv[2]) : 16384;
  const long limit = argc > 3 ? atol(argv[3]) : total;
  std::vector<uint8_t> comp(block + block / 255 + 64 + 32), cbuf(comp.size() + 64);
  uint8_t* out = (uint8_t*)aligned_alloc(64, block + 64);
  long nblk = 0, steps = 0, bad = 0, seqbytes = 0;
  for (long o = 0; o < total && o < limit; o += block, ++nblk) {
    const int n = o + block <= total ? block : (int)(total - o);
    const int mode = (int)(nblk % 3);
    const int cs = mode == 2 ? hc((const char*)buf.data() + o, (char*)comp.data(), n, (int)comp.size(), 9)
                             : fast((const char*)buf.data() + o, (char*)comp.data(), n, (int)comp.size(), mode == 0 ? 1 : 8);
    if (cs <= 0) return 4;
    const int shift = (int)(nblk % 16);
    // stream = [shift junk bytes][block]; sbeg/send bound the whole stream
    uint8_t* sb = (uint8_t*)(((uintptr_t)cbuf.data() + 15) & ~(uintptr_t)15);
    memset(sb, 0xEE, shift);
    memcpy(sb + shift, comp.data(), cs);
    const bool fuzz = argc > 4;
    if (fuzz) for (int q = 0; q < 1 + (int)(nblk % 4); ++q) sb[shift + (rand() % cs)] ^= (uint8_t)(1 + rand() % 255);
    memset(out, 0xAA, block + 64);
    g_dlo = out; g_dhi = out + ((n + 15) & ~15);
    uint32_t sm[32];
    Lane L;
    lane_begin(L, sm, sb + shift, cs, out, n, sb, sb + shift + cs);
    uint32_t rc = 0;
    long st = 0;
    while (L.mode != kLaneIdle) { rc = lane_step(L, sm, sb, sb + shift + cs); ++st; if (rc) break; if (st > 40L * block) { rc = 99; break; } }
    steps += st;
    seqbytes += n;
    if (fuzz ? (rc == 99 || out[(n + 15) & ~15] != 0xAA) : (rc || memcmp(out, buf.data() + o, n) != 0 || out[n] != 0xAA)) {
      int fd = 0; while (fd < n && out[fd] == buf[o + fd]) fd++;
      if (bad < 5) fprintf(stderr, "block %ld (n=%d cs=%d shift=%d mode=%d): rc=%u first diff %d guard %02x\n", nblk, n, cs, shift, mode, rc, fd, out[n]);
      bad++;
    }
  }
  printf("%s block=%d: %ld blocks, %ld bad, %.1f steps/KiB\n", argv[1], block, nblk, bad, steps / (seqbytes / 1024.0));
  return bad ? 10 : 0;
}
