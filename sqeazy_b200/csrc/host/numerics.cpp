#include "numerics.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#if defined(__x86_64__) || defined(__i386__)
#include <cpuid.h>
#endif

namespace sqyb {

// ------------------------------------------------------------------------------------------
// quantiser LUTs. Strict float32 in source order; weights are 1.0f ("none").
// ------------------------------------------------------------------------------------------
void quantiser_luts_from_histogram(const uint32_t* hist, uint8_t* enc, uint16_t* dec) {
  std::memset(enc, 0, 65536);
  std::memset(dec, 0, 256 * sizeof(uint16_t));
  std::vector<float> imp(65536);
  double acc = 0.;
  uint32_t levels = 0;
  for (uint32_t v = 0; v < 65536; ++v) {
    imp[v] = float(hist[v]) * 1.0f;  // computeImportance, quantiser_utils.hpp:155-168
    acc += imp[v];                   // std::accumulate(..., 0.) is a double sum
    levels += imp[v] != 0.f;
  }
  const float impSum = (float)acc;
  if (!(impSum != 0)) return;

  if (levels <= 256) {
    // linear_mapping_quantisation, quantiser_utils.hpp:286-306
    uint32_t c = 0;
    for (uint32_t v = 0; v < 65536 && c < 256; ++v) {
      enc[v] = (uint8_t)c;
      dec[c] = (uint16_t)v;
      if (imp[v] != 0.f) c++;
    }
    if (c < 256 && c > 0 && dec[c] == 65535) {
      for (uint32_t k = c; k < 256; ++k) dec[k] = dec[c - 1];
    }
    return;
  }

  // adaptive_lloyd_com, quantiser_utils.hpp:227-284
  size_t levels_available = 256;
  volatile float bucket = impSum / levels_available;  // volatile: forbid fused/extended evaluation
  float integral = imp[0];
  float q = imp[0];
  uint32_t c = 0;
  float wm = 0.f * imp[0];
  float idx = 0.f;
  for (uint32_t v = 1; v < 65536; ++v) {
    if (q >= bucket && c < 255) {
      dec[c] = (uint16_t)idx;
      c++;
      levels_available--;
      q = imp[v];
      wm = float(v) * imp[v];
      if (integral < impSum) bucket = (impSum - integral) / levels_available;
      if (q != 0.f) idx = roundf(wm / q);
    } else {
      q += imp[v];
      volatile float prod = float(v) * imp[v];
      wm += prod;
      if (q != 0.f) idx = roundf(wm / q);
    }
    enc[v] = (uint8_t)c;
    integral += imp[v];
  }
  dec[c] = (uint16_t)idx;
}

// ------------------------------------------------------------------------------------------
// calc_support: scan range is all 65536 bins because add_from_image never refreshes the
// populated-bin bounds (hist_impl.hpp:129-135,200-205,371-374).
// ------------------------------------------------------------------------------------------
float histogram_support(const uint32_t* bins, float threshold) {
  float result = 0;
  if (threshold > 1.) return result;
  if (threshold < 0.) return result;
  int isum = 0;  // std::accumulate(begin, end, 0): int accumulator, wraps
  for (uint32_t i = 0; i < 65536; ++i) isum = (int)((unsigned)isum + bins[i]);
  const double total = isum;
  double running = 0;
  uint32_t support = 65536;
  for (uint32_t i = 0; i < 65536; ++i) {
    running += bins[i];
    if ((running / total) > threshold) { support = i; break; }
  }
  const uint16_t m = (uint16_t)support;  // 65536 -> 0
  return support_from_index(support, m ? bins[m] : 0u, m ? bins[m - 1] : 0u);
}

// hist_impl.hpp:63-84: interpolation between the two bins around the support index (products and sum in uint32)
float support_from_index(uint32_t support, uint32_t bin_m, uint32_t bin_m1) {
  const uint16_t m = (uint16_t)support;  // 65536 -> 0
  if (m == 0) return 0.f;
  const uint32_t num = bin_m * (uint32_t)m + bin_m1 * (uint32_t)(m - 1);  // u32 wrap-around
  return float(num) / float(bin_m1 + bin_m);
}

// ------------------------------------------------------------------------------------------
// compass::runtime::size::cache::level(2), including its exclusive-end bit ranges
// (bit_view::range(b,e) masks e-b bits, compass.hpp:409-420).
// ------------------------------------------------------------------------------------------
static inline uint32_t bits(uint32_t v, unsigned b, unsigned e) { return (v >> b) & ~(~0u << (e - b)); }

size_t host_l2_cache_bytes() {
  if (const char* env = std::getenv("SQY_L2_BYTES")) {
    char* endp = nullptr;
    unsigned long long v = std::strtoull(env, &endp, 10);
    if (endp && endp != env) return (size_t)v;
  }
#if defined(__x86_64__) || defined(__i386__)
  unsigned a = 0, b = 0, c = 0, d = 0;
  __cpuid_count(0, 0, a, b, c, d);
  char vendor[13];
  std::memcpy(vendor + 0, &b, 4);
  std::memcpy(vendor + 4, &d, 4);
  std::memcpy(vendor + 8, &c, 4);
  vendor[12] = 0;
  const std::string brand(vendor);
  std::vector<uint32_t> sizes;
  if (brand.find("AMD") != std::string::npos) {
    __cpuid_count(0x80000005u, 1, a, b, c, d);
    if (bits(c, 0, 7)) {
      sizes.push_back(bits(c, 24, 31) * 1024);
      __cpuid_count(0x80000006u, 1, a, b, c, d);
      sizes.push_back((bits(c, 16, 31) & 0xffff) * 1024);
      sizes.push_back(bits(d, 19, 31) * 512 * 1024);
    }
  }
  if (brand.find("Intel") != std::string::npos) {
    for (uint32_t l = 0; l < 8; ++l) {
      __cpuid_count(4, l, a, b, c, d);
      if (!((a >> 1) & 1)) continue;
      if (bits(a, 5, 8) != l) continue;
      const uint32_t ways = 1 + bits(b, 22, 31);
      const uint32_t partitions = 1 + bits(b, 12, 21);
      const uint32_t line = 1 + bits(b, 0, 11);
      const uint32_t sets = 1 + c;
      sizes.push_back(ways * partitions * line * sets);
    }
  }
  return sizes.size() >= 2 ? sizes[1] : 0;
#else
  return 0;
#endif
}

uint64_t rmest_frame_portion(uint64_t frame_elems, size_t l2_bytes) {
  // index_type frame_portion = frame_size > L2 ? L2*.75 : frame_size  (double product truncated)
  return frame_elems > l2_bytes ? (uint64_t)(double(l2_bytes) * .75) : frame_elems;
}

}  // namespace sqyb
