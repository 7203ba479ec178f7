// Stand-in for the reference's string_parsers.hpp (which needs boost::string_ref
// and boost::algorithm). Defines the reference's include guard so the real file
// is skipped, and supplies the three entry points the stage headers call:
//   pipeline_parser::minors            (string_parsers.hpp:434-467)
//   parsing::range_to_verbatim         (string_parsers.hpp:508-535)
//   parsing::verbatim_to_range         (string_parsers.hpp:551-583)
// Behaviour (verbatim-aware ',' split, first '=' splits key/value, std base64
// with '=' padding) restated from those lines; pinned against the reference's
// own fixtures in tests/test_parser.py. Test infrastructure only.
#ifndef _STRING_PARSERS_H_
#define _STRING_PARSERS_H_
#include <cstdint>
#include <cstring>
#include <iterator>
#include <map>
#include <string>
#include <vector>
#include "sqeazy_common.hpp"

namespace sqeazy {
typedef std::vector<std::string> vec_of_strings_t;
typedef std::vector<std::pair<std::string, std::string> > vec_of_pairs_t;
typedef std::map<std::string, std::string> parsed_map_t;

namespace stub_detail {
inline std::vector<std::string> verbatim_aware_split(const std::string& s, const std::string& sep) {
  std::vector<std::string> out;
  if (s.empty() || sep.empty()) return out;
  const std::string& open = ignore_this_delimiters.first;
  const std::string& close = ignore_this_delimiters.second;
  std::size_t start = 0, i = 0;
  bool inside = false;
  while (i < s.size()) {
    if (!inside && s.compare(i, open.size(), open) == 0) { inside = true; i += open.size(); continue; }
    if (inside && s.compare(i, close.size(), close) == 0) { inside = false; i += close.size(); continue; }
    if (!inside && s.compare(i, sep.size(), sep) == 0) {
      out.push_back(s.substr(start, i - start));
      i += sep.size();
      start = i;
      continue;
    }
    ++i;
  }
  out.push_back(s.substr(start));
  return out;
}
static const char b64chars[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
inline std::string b64_encode(const unsigned char* p, std::size_t n) {
  std::string o;
  o.reserve(((n + 2) / 3) * 4);
  std::size_t i = 0;
  for (; i + 2 < n; i += 3) {
    std::uint32_t v = (p[i] << 16) | (p[i + 1] << 8) | p[i + 2];
    o += b64chars[(v >> 18) & 63]; o += b64chars[(v >> 12) & 63]; o += b64chars[(v >> 6) & 63]; o += b64chars[v & 63];
  }
  if (n - i == 1) {
    std::uint32_t v = p[i] << 16;
    o += b64chars[(v >> 18) & 63]; o += b64chars[(v >> 12) & 63]; o += "==";
  } else if (n - i == 2) {
    std::uint32_t v = (p[i] << 16) | (p[i + 1] << 8);
    o += b64chars[(v >> 18) & 63]; o += b64chars[(v >> 12) & 63]; o += b64chars[(v >> 6) & 63]; o += '=';
  }
  return o;
}
inline std::vector<unsigned char> b64_decode(const char* p, std::size_t n) {
  std::vector<unsigned char> o;
  std::uint32_t acc = 0; int bits = 0;
  for (std::size_t i = 0; i < n; ++i) {
    const char* f = (p[i] == '=') ? nullptr : std::strchr(b64chars, p[i]);
    if (!f || !p[i]) break;
    acc = (acc << 6) | std::uint32_t(f - b64chars); bits += 6;
    if (bits >= 8) { bits -= 8; o.push_back((unsigned char)((acc >> bits) & 0xff)); }
  }
  return o;
}
}  // namespace stub_detail

struct pipeline_parser {
  template <typename iter_t>
  parsed_map_t minors(iter_t _begin, iter_t _end) {
    parsed_map_t value;
    std::string msg(_begin, _end);
    if (msg.empty()) return value;
    for (const std::string& kv : stub_detail::verbatim_aware_split(msg, ",")) {
      std::size_t dist = kv.find("=");
      if (dist == std::string::npos) dist = kv.size();
      std::string key = kv.substr(0, dist);
      std::string val = (dist + 1 < kv.size()) ? kv.substr(dist + 1) : kv;
      value[key] = val;
    }
    return value;
  }
};

namespace parsing {
template <typename iter_t>
static std::string range_to_verbatim(iter_t _begin, iter_t _end) {
  typedef typename std::iterator_traits<iter_t>::value_type value_t;
  const std::size_t len = (std::size_t)std::distance(_begin, _end);
  std::string value;
  if (!len) return value;
  value = ignore_this_delimiters.first;
  value += stub_detail::b64_encode(reinterpret_cast<const unsigned char*>(&*_begin), len * sizeof(value_t));
  value += ignore_this_delimiters.second;
  return value;
}
template <typename string_t, typename iter_t>
static iter_t verbatim_to_range(string_t _verbatim, iter_t _begin, iter_t _end) {
  typedef typename std::iterator_traits<iter_t>::value_type value_t;
  const std::size_t len = (std::size_t)std::distance(_begin, _end);
  iter_t value = _begin;
  if (!len) return value;
  const std::size_t skip = ignore_this_delimiters.first.size();
  std::size_t end_pos = _verbatim.rfind(ignore_this_delimiters.second);
  if (end_pos == std::string::npos) end_pos = _verbatim.size();
  if (end_pos < skip) return value;
  std::vector<unsigned char> bytes = stub_detail::b64_decode(_verbatim.data() + skip, end_pos - skip);
  if (bytes.size() > len * sizeof(value_t)) return value;
  std::memcpy(reinterpret_cast<char*>(&*_begin), bytes.data(), bytes.size());
  value += bytes.size() / sizeof(value_t);
  return value;
}
}  // namespace parsing
}  // namespace sqeazy
#endif
