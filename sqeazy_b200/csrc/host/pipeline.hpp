// Pipeline planning: string -> typed stage plan, canonical names, size bounds.
// Mirrors the reference's format/control layer for the hot-path stages only:
//   dynamic_pipeline.hpp:137-170 (from_string), :177-226 (can_be_built_from), :476-503 (name),
//   :866-890 (max_encoded_size); sqeazy_pipelines.hpp:31-77 (registry);
//   stage configs: bitswap_scheme_impl.hpp:40-90, remove_background_scheme_impl.hpp:32-71,
//   quantiser_scheme_impl.hpp:83-137, lz4.hpp:58-188, diff_scheme_impl.hpp:35-59 ("diff3x3x1", no config).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "text.hpp"

namespace sqyb {

enum class StageKind { Bitswap, Bitshuffle, Diff, RemoveBackground, RmEstBkrd, Quantiser, Lz4, PassThrough };

struct Stage {
  StageKind kind;
  int w = 1;                                     // bitswap: bits per plane
  uint32_t block_size = 0;                       // bitshuffle: elements per block, 0 = the library's default (bitshuffle_scheme_impl.hpp:45-58)
  int threshold = 0;                             // remove_background
  std::map<std::string, std::string> kv;         // quantiser config map (key order = std::map order, like the reference)
  bool has_decode_lut = false;
  uint16_t decode_lut[256];
  // lz4 (lz4.hpp:58-101)
  int accel = 1;
  uint32_t blocksize_kb = 256, framestep_kb = 256, n_chunks_of_input = 0;

  std::string name() const;
  std::string config() const;
};

struct Pipeline {
  int elem = 2;                   // bytes per voxel: 2 (dypeline<uint16_t>) or 1 (dypeline<uint8_t>, the *_UI8 entry points)
  const char* type_name() const { return elem == 1 ? "uint8" : "uint16"; }
  std::vector<Stage> head;        // raw_type -> raw_type filters
  bool has_sink = false;
  Stage sink;                     // lz4 | quantiser | pass_through
  bool has_tail = false;
  Stage tail;                     // lz4 on the sink's bytes
  std::string canonical() const;  // name written into the header
  bool empty() const { return head.empty() && !has_sink; }
};

// reference semantics of dypeline<uint16_t>::can_be_built_from restricted to the accelerated stages
bool pipeline_possible_u16(const std::string& s);
// builds the plan; false if the string is not valid / not supported
bool build_pipeline_u16(const std::string& s, Pipeline& out);

// the same for dypeline<uint8_t> (src/sqeazy.cpp:72-106,144-163,209-231,243-251,309-335), restricted to the stages
// with uint8 kernels: bitswap1|2|4, remove_background (alias rmbkrd) -> lz4 | pass_through [-> lz4]
bool pipeline_possible_u8(const std::string& s);
bool build_pipeline_u8(const std::string& s, Pipeline& out);

// bound of the encoded payload for raw_bytes of input, and of the whole blob (2*header + max(stage bounds))
uint64_t max_encoded_size(const Pipeline& p, uint64_t raw_bytes);   // raw_bytes = voxels * p.elem

// bytes reserved for the (right-aligned) header in front of the payload for this pipeline/shape
size_t header_reserve_bytes(const Pipeline& p, const std::vector<uint64_t>& shape);

uint32_t lz4_closest_blocksize_kb(uint32_t kb);   // lz4_utils.hpp:60-93

}  // namespace sqyb
