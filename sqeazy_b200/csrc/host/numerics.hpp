// Host-side sequential numerics of the path. These are 65536-step float/double recurrences whose
// bit-exactness against the reference matters more than speed (micro-seconds on the host); the
// GPU supplies the histograms they consume.
//   quantiser LUTs      : encoders/quantiser_utils.hpp:227-306 (adaptive_lloyd_com, linear_mapping_quantisation), :386-418 (setup_com)
//   99 % support        : hist_impl.hpp:63-84 (support_index), :359-381 (calc_support)
//   host L2 cache bytes : compass.hpp:948-1067,1312-1326 (feeds background_scheme_utils.hpp:44-45)
#pragma once
#include <cstddef>
#include <cstdint>

namespace sqyb {

// hist: 65536 x u32 (wraps mod 2^32 like the reference's bins). enc: 65536 x u8 codes, dec: 256 x u16.
void quantiser_luts_from_histogram(const uint32_t* hist, uint8_t* enc, uint16_t* dec);

// calc_support(0.99f) of a 65536-bin u32 histogram exactly as the reference computes it.
float support_from_index(uint32_t support, uint32_t bin_m, uint32_t bin_m1);
float histogram_support(const uint32_t* bins, float threshold);

// L2 cache size in bytes as compass::runtime::size::cache::level(2) reports it on this host;
// SQY_L2_BYTES in the environment overrides (needed to reproduce blobs made on another host).
size_t host_l2_cache_bytes();

// number of leading elements of a z-frame the reference histograms (background_scheme_utils.hpp:44-45)
uint64_t rmest_frame_portion(uint64_t frame_elems, size_t l2_bytes);

}  // namespace sqyb
