// Pipeline-string grammar, base64 and the JSON blob header of the sqeazy format.
// Host-side only (no CUDA). Behavioural contract = the reference's
//   string_parsers.hpp:285-471   (pipeline_parser::to_pairs / ::minors, verbatim regions)
//   dynamic_pipeline.hpp:177-226 (can_be_built_from: registry + re-assembled length check)
//   sqeazy_header.hpp:147-193    (header::pack, boost::property_tree pretty JSON)
//   sqeazy_header.hpp:295-344    (header::unpack), :506-538 (validity / end search)
//   base64.hpp:135-202
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace sqyb {

static const char kVerbatimOpen[] = "<verbatim>";
static const char kVerbatimClose[] = "</verbatim>";
static const char kHeaderDelim[] = "|01307#!";

// ---- base64 (standard alphabet, '=' padding) ----
std::string base64_encode(const void* data, size_t nbytes);
// decodes until the first non-alphabet char; returns bytes written (<= cap)
size_t base64_decode(const char* s, size_t n, void* out, size_t cap);

// ---- pipeline string ----
struct StageSpec {
  std::string name;                          // e.g. "lz4"
  std::string args;                          // text between the outer parentheses ("" if none)
  std::map<std::string, std::string> kv;     // args split on ',' and first '=' (verbatim-aware)
};
// split on sep outside <verbatim>..</verbatim>
std::vector<std::string> split_outside_verbatim(const std::string& s, const std::string& sep);
std::vector<std::pair<std::string, std::string>> to_pairs(const std::string& pipeline);
std::map<std::string, std::string> minors(const std::string& args);
std::vector<StageSpec> parse_pipeline(const std::string& pipeline);

// ---- header ----
struct Header {
  std::string pipeline;
  std::string raw_type;                      // "uint16"
  std::vector<uint64_t> shape;
  uint64_t compressed_bytes = 0;
  size_t size = 0;                           // bytes up to and including the delimiter (+ leading pad)
  bool valid = false;
};
std::string json_escape(const std::string& s);
std::string pack_header(const std::string& raw_type, unsigned sizeof_raw, const std::vector<uint64_t>& shape,
                        const std::string& pipeline, uint64_t payload_bytes);
Header unpack_header(const char* buf, size_t nbytes);
unsigned sizeof_typename(const std::string& t);   // header_utils.hpp:40-61 ; 0 if unknown

}  // namespace sqyb
