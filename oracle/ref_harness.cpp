// oracle/_ref harness: compiles the REFERENCE's own stage headers, where they
// lie under /root/reference/src/cpp/src, into libsqyref.so with a tiny C ABI so
// that tests (ctypes) and bench.py's cpu_baseline / --impl reference leg can run
// the reference's real code.  TEST INFRASTRUCTURE ONLY: nothing under
// sqeazy_b200/ may link or load this.  No reference source is copied; this file
// only instantiates and calls the reference's templates:
//   bitswap_scheme<uint16_t,W>            encoders/bitswap_scheme_impl.hpp:27-197
//   remove_background_scheme<uint16_t>    encoders/remove_background_scheme_impl.hpp:25-131
//   remove_estimated_background_scheme    encoders/remove_estimated_background_scheme_impl.hpp:20-149
//   extract_darkest_face_supports         encoders/background_scheme_utils.hpp:35-105
//   quantiser<uint16_t,char>              encoders/quantiser_utils.hpp:44-545
//   lz4_scheme<uint16_t>/<char>           encoders/lz4.hpp:34-347, lz4_utils.hpp:99-274
//   compass::runtime::size::cache::level  compass.hpp:1312-1326
//   diff_scheme<uint16_t>/<uint8_t>       encoders/diff_scheme_impl.hpp:15-213 (last_plane_neighborhood<3>, "diff3x3x1")
// Boost is absent in this image: oracle/refshim/ provides ~40 lines of stand-ins
// and a stub of string_parsers.hpp. The full
// dynamic_pipeline/header (boost::property_tree) cannot be compiled here; the
// stage sequence of dynamic_pipeline::detail_encode (dynamic_pipeline.hpp:619-690)
// is re-driven below in ref_pipeline_encode_stages().
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "string_parsers_stub.hpp"  // must precede any reference header

#include "encoders/bitswap_scheme_impl.hpp"
#include "encoders/remove_background_scheme_impl.hpp"
#include "encoders/remove_estimated_background_scheme_impl.hpp"
#include "encoders/quantiser_utils.hpp"
#include "encoders/lz4.hpp"
#include "encoders/diff_scheme_impl.hpp"

namespace sqy = sqeazy;

extern "C" {

int ref_abi_version() { return 1; }

long ref_l2_cache_bytes() { return (long)compass::runtime::size::cache::level(2); }
int ref_has_sse4() { return compass::runtime::has(compass::feature::sse4()) ? 1 : 0; }

// ---- bitswap -------------------------------------------------------------
int ref_bitswap_encode(int w, const uint16_t* in, uint16_t* out, long n, int nthreads) {
  uint16_t* end = nullptr;
  if (w == 1) { sqy::bitswap_scheme<uint16_t, 1> s; s.set_n_threads(nthreads); end = s.encode(in, out, (std::size_t)n); }
  else if (w == 2) { sqy::bitswap_scheme<uint16_t, 2> s; s.set_n_threads(nthreads); end = s.encode(in, out, (std::size_t)n); }
  else if (w == 4) { sqy::bitswap_scheme<uint16_t, 4> s; s.set_n_threads(nthreads); end = s.encode(in, out, (std::size_t)n); }
  else if (w == 8) { sqy::bitswap_scheme<uint16_t, 8> s; s.set_n_threads(nthreads); end = s.encode(in, out, (std::size_t)n); }
  else return 2;
  return end == out + n ? 0 : 1;
}
// scalar path forced (bitplane_reorder_scalar.hpp:27-74), single thread to avoid the RMW race
int ref_bitswap_encode_scalar(int w, const uint16_t* in, uint16_t* out, long n) {
  long m;
  switch (w) {
    case 1: m = n - (n % 16); std::copy(in + m, in + n, out + m); return sqy::detail::scalar_bitplane_reorder_encode<1>(in, out, (std::size_t)m, 1);
    case 2: m = n - (n % 8);  std::copy(in + m, in + n, out + m); return sqy::detail::scalar_bitplane_reorder_encode<2>(in, out, (std::size_t)m, 1);
    case 4: m = n - (n % 4);  std::copy(in + m, in + n, out + m); return sqy::detail::scalar_bitplane_reorder_encode<4>(in, out, (std::size_t)m, 1);
    case 8: m = n - (n % 2);  std::copy(in + m, in + n, out + m); return sqy::detail::scalar_bitplane_reorder_encode<8>(in, out, (std::size_t)m, 1);
  }
  return 2;
}
int ref_bitswap_decode(int w, const uint16_t* in, uint16_t* out, long n) {
  if (w == 1) { sqy::bitswap_scheme<uint16_t, 1> s; return s.decode(in, out, (std::size_t)n); }
  if (w == 2) { sqy::bitswap_scheme<uint16_t, 2> s; return s.decode(in, out, (std::size_t)n); }
  if (w == 4) { sqy::bitswap_scheme<uint16_t, 4> s; return s.decode(in, out, (std::size_t)n); }
  if (w == 8) { sqy::bitswap_scheme<uint16_t, 8> s; return s.decode(in, out, (std::size_t)n); }
  return 2;
}

// ---- uint8 volumes: the same templates instantiated for raw_type = uint8_t (dypeline<uint8_t>, src/sqeazy.cpp:72-106).
// One thread: the scalar encode read-modify-writes shared output words (bitplane_reorder_scalar.hpp:45-69).
int ref_bitswap8_encode(int w, const uint8_t* in, uint8_t* out, long n) {
  uint8_t* end = nullptr;
  if (w == 1) { sqy::bitswap_scheme<uint8_t, 1> s; s.set_n_threads(1); end = s.encode(in, out, (std::size_t)n); }
  else if (w == 2) { sqy::bitswap_scheme<uint8_t, 2> s; s.set_n_threads(1); end = s.encode(in, out, (std::size_t)n); }
  else if (w == 4) { sqy::bitswap_scheme<uint8_t, 4> s; s.set_n_threads(1); end = s.encode(in, out, (std::size_t)n); }
  else return 2;
  return end == out + n ? 0 : 1;
}
int ref_bitswap8_decode(int w, const uint8_t* in, uint8_t* out, long n) {
  if (w == 1) { sqy::bitswap_scheme<uint8_t, 1> s; return s.decode(in, out, (std::size_t)n); }
  if (w == 2) { sqy::bitswap_scheme<uint8_t, 2> s; return s.decode(in, out, (std::size_t)n); }
  if (w == 4) { sqy::bitswap_scheme<uint8_t, 4> s; return s.decode(in, out, (std::size_t)n); }
  return 2;
}
int ref_remove_background8(int threshold, const uint8_t* in, uint8_t* out, long n) {
  sqy::remove_background_scheme<uint8_t> s("threshold=" + std::to_string(threshold));
  s.set_n_threads(1);
  return s.encode(in, out, (std::size_t)n) == out + n ? 0 : 1;
}

// ---- background removal --------------------------------------------------
int ref_remove_background(int threshold, const uint16_t* in, uint16_t* out, long n, int nthreads) {
  sqy::remove_background_scheme<uint16_t> s("threshold=" + std::to_string(threshold));
  s.set_n_threads(nthreads);
  return s.encode(in, out, (std::size_t)n) == out + n ? 0 : 1;
}
void ref_darkest_face_supports(const uint16_t* in, long z, long y, long x, float* out4) {
  std::vector<std::size_t> dims = {(std::size_t)z, (std::size_t)y, (std::size_t)x};
  std::vector<float> v = sqy::extract_darkest_face_supports(in, dims, 0.99f, 1);
  for (int i = 0; i < 4; ++i) out4[i] = v[i];
}
int ref_rmestbkrd_encode(const uint16_t* in, uint16_t* out, long z, long y, long x, int nthreads) {
  sqy::remove_estimated_background_scheme<uint16_t> s;
  s.set_n_threads(nthreads);
  std::vector<std::size_t> dims = {(std::size_t)z, (std::size_t)y, (std::size_t)x};
  return s.encode(in, out, dims) == out + z * y * x ? 0 : 1;
}

// ---- diff: difference to the mean of the 3x3 neighbours in the previous z plane --------------
int ref_diff_encode(const uint16_t* in, uint16_t* out, long z, long y, long x, int nthreads) {
  sqy::diff_scheme<uint16_t> s;
  s.set_n_threads(nthreads);
  std::vector<std::size_t> dims = {(std::size_t)z, (std::size_t)y, (std::size_t)x};
  return s.encode(in, out, dims) == out + z * y * x ? 0 : 1;
}
int ref_diff_decode(const uint16_t* in, uint16_t* out, long z, long y, long x) {
  sqy::diff_scheme<uint16_t> s;
  s.set_n_threads(1);                       // the reference's decode loop races across planes with more threads
  std::vector<std::size_t> dims = {(std::size_t)z, (std::size_t)y, (std::size_t)x};
  return s.decode(in, out, dims);
}
int ref_diff8_encode(const uint8_t* in, uint8_t* out, long z, long y, long x) {
  sqy::diff_scheme<uint8_t> s;
  s.set_n_threads(1);
  std::vector<std::size_t> dims = {(std::size_t)z, (std::size_t)y, (std::size_t)x};
  return s.encode(in, out, dims) == out + z * y * x ? 0 : 1;
}
int ref_diff8_decode(const uint8_t* in, uint8_t* out, long z, long y, long x) {
  sqy::diff_scheme<uint8_t> s;
  s.set_n_threads(1);
  std::vector<std::size_t> dims = {(std::size_t)z, (std::size_t)y, (std::size_t)x};
  return s.decode(in, out, dims);
}
int ref_diff_name(char* out, int cap) {
  const std::string n = sqy::diff_scheme<uint16_t>().name();
  if ((int)n.size() + 1 > cap) return 1;
  std::memcpy(out, n.c_str(), n.size() + 1);
  return 0;
}

// ---- quantiser -----------------------------------------------------------
// hist: 65536 x u32, enc: 65536 x i8, dec: 256 x u16 (quantiser<uint16_t,char> as in sqeazy_pipelines.hpp:47-56)
void ref_quantiser_setup(const uint16_t* in, long n, uint32_t* hist, char* enc, uint16_t* dec) {
  sqy::quantiser<uint16_t, char> q;
  q.setup_com(in, in + n);
  std::copy(q.histo_.begin(), q.histo_.end(), hist);
  std::copy(q.lut_encode_.begin(), q.lut_encode_.end(), enc);
  std::copy(q.lut_decode_.begin(), q.lut_decode_.end(), dec);
}
// LUT from a given histogram (lets tests feed >2^31-voxel histograms accumulated slab-wise)
void ref_quantiser_luts_from_hist(const uint32_t* hist, char* enc, uint16_t* dec) {
  sqy::quantiser<uint16_t, char> q;
  std::copy(hist, hist + 65536, q.histo_.begin());
  q.computeWeights();
  q.computeImportance();
  float importanceSum = std::accumulate(q.importance_.begin(), q.importance_.end(), 0.);
  if (importanceSum != 0) {
    uint32_t n_levels = std::count_if(q.importance_.begin(), q.importance_.end(), [](float v) { return v != 0.f; });
    if (n_levels <= 256) q.linear_mapping_quantisation();
    else q.adaptive_lloyd_com(importanceSum);
  }
  std::copy(q.lut_encode_.begin(), q.lut_encode_.end(), enc);
  std::copy(q.lut_decode_.begin(), q.lut_decode_.end(), dec);
}
void ref_quantiser_encode(const uint16_t* in, long n, char* out, uint16_t* dec, int nthreads) {
  sqy::quantiser<uint16_t, char> q;
  q.set_n_threads(1);  // F12: the scheme's shrinker always runs single-threaded
  q.encode(in, (std::size_t)n, out);
  (void)nthreads;
  std::copy(q.lut_decode_.begin(), q.lut_decode_.end(), dec);
}
// reference decode semantics restricted to codes < 128 being well-defined (F11)
void ref_quantiser_decode(const char* in, long n, const uint16_t* dec, uint16_t* out) {
  sqy::quantiser<uint16_t, char> q;
  std::copy(dec, dec + 256, q.lut_decode_.begin());
  for (long i = 0; i < n; ++i) out[i] = q.lut_decode_[(unsigned char)in[i]];
}
// the decode-LUT string as the scheme writes it into the header (quantiser_utils.hpp:490-497)
long ref_quantiser_lut_string(const uint16_t* dec, char* out, long cap) {
  sqy::quantiser<uint16_t, char> q;
  std::copy(dec, dec + 256, q.lut_decode_.begin());
  std::string s = q.lut_to_string(q.lut_decode_);
  if ((long)s.size() + 1 > cap) return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (long)s.size();
}

// ---- lz4 -----------------------------------------------------------------
long ref_lz4_max_encoded_size(const char* config, long nbytes, int nthreads) {
  sqy::lz4_scheme<uint16_t> s(config ? config : "");
  s.set_n_threads(nthreads);
  return (long)s.max_encoded_size(nbytes);
}
long ref_lz4_config(const char* config, char* out, long cap) {
  sqy::lz4_scheme<uint16_t> s(config ? config : "");
  std::string c = s.config();
  if ((long)c.size() + 1 > cap) return -1;
  std::memcpy(out, c.c_str(), c.size() + 1);
  return (long)c.size();
}
// encode n uint16 elements; returns compressed bytes or -1
long ref_lz4_encode_u16(const char* config, const uint16_t* in, long n, char* out, int nthreads) {
  sqy::lz4_scheme<uint16_t> s(config ? config : "");
  s.set_n_threads(nthreads);
  std::vector<std::size_t> shape = {(std::size_t)n};
  char* end = s.encode(in, out, shape);
  return end ? (long)(end - out) : -1;
}
// tail-filter flavour lz4_scheme<char> (after the quantiser): n bytes
long ref_lz4_encode_bytes(const char* config, const char* in, long n, char* out, int nthreads) {
  sqy::lz4_scheme<char> s(config ? config : "");
  s.set_n_threads(nthreads);
  std::vector<std::size_t> shape = {(std::size_t)n};
  char* end = s.encode(in, out, shape);
  return end ? (long)(end - out) : -1;
}
// the reference's multi-frame decode loop (lz4.hpp:257-339); returns its rc (0 ok)
int ref_lz4_decode_u16(const char* in, long nbytes_in, uint16_t* out, long n_out) {
  sqy::lz4_scheme<uint16_t> s;
  return s.decode(in, out, (std::size_t)nbytes_in, (std::size_t)n_out);
}
int ref_lz4_decode_bytes(const char* in, long nbytes_in, char* out, long n_out) {
  sqy::lz4_scheme<char> s;
  return s.decode(in, out, (std::size_t)nbytes_in, (std::size_t)n_out);
}

// ---- stage chain in detail_encode order, timed like verbs/bench.hpp:181-194 ----
// pipeline_id: 0 bitswap1->lz4, 1 rmestbkrd->bitswap1->lz4, 2 quantiser->lz4,
//              3 remove_background(threshold=T)->bitswapW->lz4 (w, T passed), 4 lz4 only
// Returns payload bytes (or -1). seconds_out[0] = wall seconds of the stage chain.
long ref_pipeline_encode_stages(int pipeline_id, const uint16_t* in, long z, long y, long x, char* out,
                                uint16_t* scratch_a, uint16_t* scratch_b, int nthreads, int w, int threshold,
                                double* seconds_out) {
  const long n = z * y * x;
  std::vector<std::size_t> dims = {(std::size_t)z, (std::size_t)y, (std::size_t)x};
  long bytes = -1;
  auto t0 = std::chrono::high_resolution_clock::now();
  if (pipeline_id == 0) {
    sqy::bitswap_scheme<uint16_t, 1> b; b.set_n_threads(nthreads);
    if (!b.encode(in, scratch_a, (std::size_t)n)) return -1;
    bytes = ref_lz4_encode_u16("", scratch_a, n, out, nthreads);
  } else if (pipeline_id == 1) {
    sqy::remove_estimated_background_scheme<uint16_t> r; r.set_n_threads(nthreads);
    if (!r.encode(in, scratch_a, dims)) return -1;
    sqy::bitswap_scheme<uint16_t, 1> b; b.set_n_threads(nthreads);
    if (!b.encode(scratch_a, scratch_b, (std::size_t)n)) return -1;
    bytes = ref_lz4_encode_u16("", scratch_b, n, out, nthreads);
  } else if (pipeline_id == 2) {
    sqy::quantiser<uint16_t, char> q; q.set_n_threads(1);
    q.setup_com(in, in + n);
    char* codes = reinterpret_cast<char*>(scratch_a);
    sqy::applyLUT<uint16_t, char> lut(q.lut_encode_);
#pragma omp parallel for num_threads(nthreads)
    for (long i = 0; i < n; ++i) codes[i] = lut(in[i]);
    bytes = ref_lz4_encode_bytes("", codes, n, out, nthreads);
  } else if (pipeline_id == 3) {
    sqy::remove_background_scheme<uint16_t> r("threshold=" + std::to_string(threshold)); r.set_n_threads(nthreads);
    if (!r.encode(in, scratch_a, (std::size_t)n)) return -1;
    if (ref_bitswap_encode(w, scratch_a, scratch_b, n, nthreads)) return -1;
    bytes = ref_lz4_encode_u16("", scratch_b, n, out, nthreads);
  } else if (pipeline_id == 4) {
    bytes = ref_lz4_encode_u16("", in, n, out, nthreads);
  }
  auto t1 = std::chrono::high_resolution_clock::now();
  if (seconds_out) seconds_out[0] = std::chrono::duration<double>(t1 - t0).count();
  return bytes;
}

// decode chain for X->bitswapW->lz4 payloads (detail_decode order, dynamic_pipeline.hpp:772-846)
int ref_pipeline_decode_stages(int w, const char* payload, long nbytes, uint16_t* out, uint16_t* scratch, long n,
                               double* seconds_out) {
  auto t0 = std::chrono::high_resolution_clock::now();
  int rc = ref_lz4_decode_u16(payload, nbytes, w ? scratch : out, n);
  if (!rc && w) rc = ref_bitswap_decode(w, scratch, out, n) ? 100 : 0;
  auto t1 = std::chrono::high_resolution_clock::now();
  if (seconds_out) seconds_out[0] = std::chrono::duration<double>(t1 - t0).count();
  return rc;
}

}  // extern "C"
