#include "pipeline.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "../device/lz4_format.h"

namespace sqyb {

// ------------------------------------------------------------------------------------------------
// stage names / configs as the reference serialises them
// ------------------------------------------------------------------------------------------------
std::string Stage::name() const {
  switch (kind) {
    case StageKind::Bitswap: return "bitswap" + std::to_string(w);
    case StageKind::Bitshuffle: return "bitshuffle";
    case StageKind::Diff: return "diff3x3x1";   // diff_scheme_impl.hpp:35-45 with last_plane_neighborhood<3>
    case StageKind::RemoveBackground: return "remove_background";
    case StageKind::RmEstBkrd: return "rmestbkrd";
    case StageKind::Quantiser: return "quantiser";
    case StageKind::Lz4: return "lz4";
    case StageKind::PassThrough: return "pass_through";
  }
  return "";
}

std::string Stage::config() const {
  std::ostringstream msg;
  switch (kind) {
    case StageKind::Bitswap:  // bitswap_scheme_impl.hpp:83-90
      msg << "num_bits_per_plane=" << w;
      break;
    case StageKind::Bitshuffle:  // bitshuffle_scheme_impl.hpp:78-84
      msg << "block_size=" << block_size;
      break;
    case StageKind::RemoveBackground:  // remove_background_scheme_impl.hpp:62-67
      msg << "threshold=" << threshold;
      break;
    case StageKind::RmEstBkrd:
    case StageKind::PassThrough:
    case StageKind::Diff:  // diff_scheme_impl.hpp:55-59: empty
      break;
    case StageKind::Quantiser: {  // quantiser_scheme_impl.hpp:104-120 (map order)
      size_t count = 0;
      for (auto& kvp : kv) {
        msg << kvp.first << "=" << kvp.second;
        if ((count++) < kv.size() - 1) msg << ",";
      }
      break;
    }
    case StageKind::Lz4:  // lz4.hpp:132-141
      msg << "accel=" << accel << ",blocksize_kb=" << blocksize_kb << ",framestep_kb=" << framestep_kb
          << ",n_chunks_of_input=" << n_chunks_of_input;
      break;
  }
  return msg.str();
}

static std::string stage_repr(const Stage& s) {
  const std::string cfg = s.config();
  return cfg.empty() ? s.name() : s.name() + "(" + cfg + ")";
}

std::string Pipeline::canonical() const {
  // dynamic_pipeline.hpp:476-503 + dynamic_stage_chain.hpp:157-183
  std::string v;
  for (size_t i = 0; i < head.size(); ++i) {
    v += stage_repr(head[i]);
    if (i + 1 < head.size()) v += "->";
  }
  if (has_sink) {
    if (!head.empty()) v += "->";
    v += stage_repr(sink);
    if (has_tail) v += "->" + stage_repr(tail);
  }
  return v;
}

// ------------------------------------------------------------------------------------------------
// registry (sqeazy_pipelines.hpp:31-77, hot-path stages only) + aliases
// ------------------------------------------------------------------------------------------------
static bool is_head_name(const std::string& n) {
  return n == "bitswap1" || n == "bitswap2" || n == "bitswap4" || n == "bitswap8" || n == "bitshuffle" || n == "diff3x3x1" || n == "remove_background" ||
         n == "rmbkrd" || n == "rmestbkrd";
}
static bool is_sink_name(const std::string& n) { return n == "lz4" || n == "quantiser" || n == "pass_through"; }
static bool is_tail_name(const std::string& n) { return n == "lz4"; }

uint32_t lz4_closest_blocksize_kb(uint32_t kb) {
  static const uint32_t sizes[4] = {64, 256, 1024, 4096};
  const uint32_t* itr = std::lower_bound(sizes, sizes + 4, kb);
  if (itr == sizes + 4) return sizes[3];
  if (itr == sizes) return sizes[0];
  const size_t upper = size_t(itr - sizes), lower = upper - 1;
  const uint32_t middle = sizes[lower] + (sizes[upper] - sizes[lower]) / 2;
  return kb >= middle ? sizes[upper] : sizes[lower];
}

static bool parse_float_like(const std::string& s, float& out) {
  char* e = nullptr;
  out = std::strtof(s.c_str(), &e);
  return e && e != s.c_str();
}

static bool make_stage(const std::string& name, const std::string& args, Stage& st, int elem) {
  const std::map<std::string, std::string> kv = minors(args);
  if (name == "bitshuffle") {
    st.kind = StageKind::Bitshuffle;
    st.block_size = 0;
    auto f = kv.find("block_size");
    if (f != kv.end()) {
      char* e = nullptr;
      const long v = std::strtol(f->second.c_str(), &e, 10);
      if (!e || e == f->second.c_str() || v < 0 || v > (1l << 30)) return false;   // std::stoi would throw in the reference
      st.block_size = (uint32_t)v;
    }
    return true;                                     // a block size that is no multiple of 8 fails at encode time, like bshuf's -81
  }
  if (name == "diff3x3x1") { st.kind = StageKind::Diff; return true; }   // a shape it cannot take fails at encode time
  if (name.rfind("bitswap", 0) == 0) {
    st.kind = StageKind::Bitswap;
    st.w = std::atoi(name.c_str() + 7);
    // bitswap1(num_bits_per_plane=4) only warns in the reference and stays at the static width
    return st.w == 1 || st.w == 2 || st.w == 4 || (st.w == 8 && elem == 2);
  }
  if (name == "remove_background" || name == "rmbkrd") {
    st.kind = StageKind::RemoveBackground;
    auto f = kv.find("threshold");
    st.threshold = 0;
    if (f != kv.end()) {
      char* e = nullptr;
      const long v = std::strtol(f->second.c_str(), &e, 10);
      if (!e || e == f->second.c_str()) return false;  // std::stoi would throw in the reference
      st.threshold = elem == 1 ? (int)(uint8_t)v : (int)(uint16_t)v;   // stored as raw_type
    }
    return true;
  }
  if (name == "rmestbkrd") { st.kind = StageKind::RmEstBkrd; return elem == 2; }
  if (name == "pass_through") { st.kind = StageKind::PassThrough; return true; }
  if (name == "quantiser") {
    if (elem != 2) return false;
    st.kind = StageKind::Quantiser;
    st.kv = kv;
    auto w = kv.find("weighting_function");
    if (w != kv.end() && w->second.find("none") == std::string::npos) return false;  // only `none` is in scope
    if (kv.count("decode_lut_path")) return false;                                    // file side channel unsupported
    auto l = kv.find("decode_lut_string");
    if (l != kv.end()) {
      // quantiser_utils.hpp:519-530 -> parsing::verbatim_to_range
      const std::string& v = l->second;
      const size_t no = sizeof(kVerbatimOpen) - 1;
      size_t endp = v.rfind(kVerbatimClose);
      if (endp == std::string::npos) endp = v.size();
      std::memset(st.decode_lut, 0, sizeof(st.decode_lut));
      if (v.size() >= no && endp >= no) {
        base64_decode(v.data() + no, endp - no, st.decode_lut, sizeof(st.decode_lut));
        st.has_decode_lut = true;
      }
    }
    return true;
  }
  if (name == "lz4") {
    st.kind = StageKind::Lz4;
    float f;
    auto it = kv.find("accel");
    if (it != kv.end() && parse_float_like(it->second, f)) st.accel = (int)f;
    it = kv.find("blocksize_kb");
    if (it != kv.end() && parse_float_like(it->second, f)) st.blocksize_kb = (uint32_t)f;
    it = kv.find("framestep_kb");
    if (it != kv.end() && parse_float_like(it->second, f)) st.framestep_kb = (uint32_t)f;
    it = kv.find("n_chunks_of_input");
    if (it != kv.end() && parse_float_like(it->second, f)) st.n_chunks_of_input = (uint32_t)f;
    if (st.blocksize_kb == 0) st.blocksize_kb = 256;
    if (st.framestep_kb < st.blocksize_kb) st.framestep_kb = st.blocksize_kb;
    else st.framestep_kb = (uint32_t)(std::round(st.framestep_kb / float(st.blocksize_kb)) * st.blocksize_kb);
    if (st.n_chunks_of_input != 0) st.framestep_kb = 0;
    return true;
  }
  return false;
}

static bool plan(const std::string& s, Pipeline* out, int elem = 2) {
  if (s.empty()) return false;
  const auto pairs = to_pairs(s);
  if (pairs.empty()) return false;
  // re-assembled length check (dynamic_pipeline.hpp:214-224): catches stray characters
  size_t rebuilt = 2 * (pairs.size() - 1);
  for (auto& p : pairs) rebuilt += p.first.size() + (p.second.empty() ? 0 : 2 + p.second.size());
  if (rebuilt != s.size()) return false;
  Pipeline pl;
  pl.elem = elem;
  for (auto& p : pairs) {
    Stage st;
    if (!pl.has_sink && is_head_name(p.first)) {
      if (!make_stage(p.first, p.second, st, elem)) return false;
      pl.head.push_back(st);
    } else if (!pl.has_sink && is_sink_name(p.first)) {
      if (!make_stage(p.first, p.second, st, elem)) return false;
      pl.sink = st;
      pl.has_sink = true;
    } else if (pl.has_sink && !pl.has_tail && is_tail_name(p.first) && pl.sink.kind != StageKind::Lz4) {
      if (!make_stage(p.first, p.second, st, elem)) return false;
      pl.tail = st;
      pl.has_tail = true;
    } else {
      return false;
    }
  }
  if (out) *out = pl;
  return true;
}

bool pipeline_possible_u16(const std::string& s) { return plan(s, nullptr); }
bool build_pipeline_u16(const std::string& s, Pipeline& out) { return plan(s, &out); }
bool pipeline_possible_u8(const std::string& s) { return plan(s, nullptr, 1); }
bool build_pipeline_u8(const std::string& s, Pipeline& out) { return plan(s, &out, 1); }

// ------------------------------------------------------------------------------------------------
// bounds
// ------------------------------------------------------------------------------------------------
// LZ4F_compressBound(srcSize, prefs{autoFlush=0, no checksums}) of liblz4 >= 1.8 (published formula)
static uint64_t lz4f_compress_bound(uint64_t src, uint64_t block) {
  const uint64_t max_src = src + (block - 1);
  const uint64_t full = max_src / block;
  const uint64_t partial = max_src & (block - 1);
  const uint64_t last = (src == 0) ? partial : 0;
  const uint64_t nblocks = full + (last > 0);
  return 4 * nblocks + block * full + last + 4;
}

static uint64_t lz4_stage_bound(const Stage& st, uint64_t bytes) {
  // lz4.hpp:146-188 with n_threads == 1
  uint64_t chunk = st.framestep_kb ? (uint64_t)st.framestep_kb << 10
                                   : (st.n_chunks_of_input ? (uint64_t)std::ceil(bytes / st.n_chunks_of_input) : bytes);
  if (chunk >= bytes || st.n_chunks_of_input >= bytes) chunk = bytes;
  const uint64_t block = (uint64_t)lz4_closest_blocksize_kb(st.blocksize_kb) << 10;
  uint64_t ref;
  if (chunk >= bytes) ref = 19 + lz4f_compress_bound(chunk, block);
  else {
    const uint64_t nchunks = (bytes + chunk - 1) / chunk;
    ref = nchunks * (lz4f_compress_bound(chunk, block) + 19);
  }
  return ref;
}

uint64_t max_encoded_size(const Pipeline& p, uint64_t raw_bytes) {
  // header of the bound query: rank-1 shape {raw_bytes}, payload = 2*raw_bytes (sqeazy_header.hpp:216-241)
  const std::string hdr = pack_header(p.type_name(), p.elem, {raw_bytes}, p.canonical(), raw_bytes * 2);
  uint64_t stage_max = 0;
  if (!p.head.empty()) stage_max = std::max<uint64_t>(stage_max, raw_bytes);
  uint64_t ours = raw_bytes;  // what this implementation can actually emit
  if (p.has_sink) {
    switch (p.sink.kind) {
      case StageKind::Lz4:
        stage_max = std::max(stage_max, lz4_stage_bound(p.sink, raw_bytes));
        ours = lz4_payload_bound(raw_bytes);
        break;
      case StageKind::Quantiser:
        stage_max = std::max(stage_max, raw_bytes * 2 + 512);  // quantiser_scheme_impl.hpp:131-137
        ours = raw_bytes / 2;
        break;
      default:
        stage_max = std::max(stage_max, raw_bytes);
        break;
    }
    if (p.has_tail) {
      stage_max = std::max(stage_max, lz4_stage_bound(p.tail, raw_bytes));
      ours = lz4_payload_bound(p.sink.kind == StageKind::Quantiser ? raw_bytes / 2 : raw_bytes);
    }
  }
  // the right-aligned header slot of this implementation is header_reserve_bytes() <= 2*hdr + 2048
  return 2 * hdr.size() + 2048 + std::max(stage_max, ours);
}

size_t header_reserve_bytes(const Pipeline& p, const std::vector<uint64_t>& shape) {
  std::string name = p.canonical();
  size_t extra = 0;
  const bool quant = p.has_sink && p.sink.kind == StageKind::Quantiser;
  if (quant && !p.sink.kv.count("decode_lut_string")) {
    // decode_lut_string=<verbatim> 684 base64 chars </verbatim>, every '/' doubled by JSON escaping
    extra = 32 + 2 * 684 + 32;
  }
  const std::string h = pack_header(p.type_name(), p.elem, shape, name, UINT64_MAX);
  size_t n = h.size() + extra + name.size() / 8 + 64;
  return (n + 255) & ~size_t(255);
}

}  // namespace sqyb
