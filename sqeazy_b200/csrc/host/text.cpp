#include "text.hpp"

#include <cstdio>
#include <cstring>
#include <sstream>

namespace sqyb {

// --------------------------------------------------------------------------------------------
// base64 — reference: base64.hpp:135-162 (encode), :177-202 (decode)
// --------------------------------------------------------------------------------------------
static const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

std::string base64_encode(const void* data, size_t n) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  std::string o;
  o.reserve(((n + 2) / 3) * 4);
  size_t i = 0;
  for (; i + 2 < n; i += 3) {
    uint32_t v = (uint32_t(p[i]) << 16) | (uint32_t(p[i + 1]) << 8) | p[i + 2];
    o.push_back(kB64[(v >> 18) & 63]);
    o.push_back(kB64[(v >> 12) & 63]);
    o.push_back(kB64[(v >> 6) & 63]);
    o.push_back(kB64[v & 63]);
  }
  if (n - i == 1) {
    uint32_t v = uint32_t(p[i]) << 16;
    o.push_back(kB64[(v >> 18) & 63]);
    o.push_back(kB64[(v >> 12) & 63]);
    o += "==";
  } else if (n - i == 2) {
    uint32_t v = (uint32_t(p[i]) << 16) | (uint32_t(p[i + 1]) << 8);
    o.push_back(kB64[(v >> 18) & 63]);
    o.push_back(kB64[(v >> 12) & 63]);
    o.push_back(kB64[(v >> 6) & 63]);
    o.push_back('=');
  }
  return o;
}

size_t base64_decode(const char* s, size_t n, void* out, size_t cap) {
  int8_t rev[256];
  std::memset(rev, -1, sizeof(rev));
  for (int i = 0; i < 64; ++i) rev[(unsigned char)kB64[i]] = (int8_t)i;
  unsigned char* o = static_cast<unsigned char*>(out);
  size_t w = 0;
  uint32_t acc = 0;
  int bits = 0;
  for (size_t i = 0; i < n; ++i) {
    int8_t v = rev[(unsigned char)s[i]];
    if (v < 0) break;  // '=' or anything foreign ends the payload
    acc = (acc << 6) | uint32_t(v);
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      if (w >= cap) return w;
      o[w++] = (unsigned char)((acc >> bits) & 0xff);
    }
  }
  return w;
}

// --------------------------------------------------------------------------------------------
// pipeline string — reference: string_parsers.hpp:124-147 (informed_split), :355-395, :434-467
// --------------------------------------------------------------------------------------------
std::vector<std::string> split_outside_verbatim(const std::string& s, const std::string& sep) {
  std::vector<std::string> out;
  if (s.empty() || sep.empty()) return out;
  const size_t no = sizeof(kVerbatimOpen) - 1, nc = sizeof(kVerbatimClose) - 1;
  size_t start = 0, i = 0;
  bool inside = false;
  while (i < s.size()) {
    if (!inside && s.compare(i, no, kVerbatimOpen) == 0) { inside = true; i += no; continue; }
    if (inside && s.compare(i, nc, kVerbatimClose) == 0) { inside = false; i += nc; continue; }
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

std::vector<std::pair<std::string, std::string>> to_pairs(const std::string& pipeline) {
  std::vector<std::pair<std::string, std::string>> value;
  for (const std::string& major : split_outside_verbatim(pipeline, "->")) {
    size_t dist = major.find('(');
    if (dist == std::string::npos) dist = major.size();
    std::string key = major.substr(0, dist);
    std::string in_brackets;
    if (key.size() < major.size() && major.size() >= dist + 2)
      in_brackets = major.substr(dist + 1, major.size() - 1 - (dist + 1));  // drops the last char (")")
    value.emplace_back(key, in_brackets);
  }
  return value;
}

std::map<std::string, std::string> minors(const std::string& args) {
  std::map<std::string, std::string> value;
  for (const std::string& item : split_outside_verbatim(args, ",")) {
    size_t dist = item.find('=');
    if (dist == std::string::npos) dist = item.size();
    std::string key = item.substr(0, dist);
    std::string val = (dist + 1 < item.size()) ? item.substr(dist + 1) : item;
    value[key] = val;
  }
  return value;
}

std::vector<StageSpec> parse_pipeline(const std::string& pipeline) {
  std::vector<StageSpec> out;
  for (auto& p : to_pairs(pipeline)) {
    StageSpec s;
    s.name = p.first;
    s.args = p.second;
    s.kv = minors(p.second);
    out.push_back(std::move(s));
  }
  return out;
}

// --------------------------------------------------------------------------------------------
// header — reference: sqeazy_header.hpp:147-193 (pack), :295-344 (unpack), :506-538
// JSON text layout = boost::property_tree::write_json(pretty): 4-space indent, all values quoted,
// '/' escaped as "\/" (Boost.PropertyTree json create_escapes).
// --------------------------------------------------------------------------------------------
std::string json_escape(const std::string& s) {
  std::string o;
  o.reserve(s.size() + 16);
  for (unsigned char c : s) {
    if (c == 0x20 || c == 0x21 || (c >= 0x23 && c <= 0x2E) || (c >= 0x30 && c <= 0x5B) || c >= 0x5D) {
      o.push_back((char)c);
    } else if (c == '\b') o += "\\b";
    else if (c == '\f') o += "\\f";
    else if (c == '\n') o += "\\n";
    else if (c == '\r') o += "\\r";
    else if (c == '\t') o += "\\t";
    else if (c == '/') o += "\\/";
    else if (c == '"') o += "\\\"";
    else if (c == '\\') o += "\\\\";
    else {
      char buf[8];
      std::snprintf(buf, sizeof(buf), "\\u%04X", (unsigned)c);
      o += buf;
    }
  }
  return o;
}

#ifndef SQYB_VERSION_STRING
#define SQYB_VERSION_STRING "0.7.2"
#endif
#ifndef SQYB_VERSION_HEADREF
#define SQYB_VERSION_HEADREF "b200"
#endif

std::string pack_header(const std::string& raw_type, unsigned sizeof_raw, const std::vector<uint64_t>& shape,
                        const std::string& pipeline, uint64_t payload_bytes) {
  std::ostringstream js;
  js << "{\n";
  js << "    \"pipename\": \"" << json_escape(pipeline) << "\",\n";
  js << "    \"raw\": {\n";
  js << "        \"type\": \"" << raw_type << "\",\n";
  js << "        \"rank\": \"" << shape.size() << "\"";
  if (!shape.empty()) {
    js << ",\n        \"shape\": {\n";
    for (size_t i = 0; i < shape.size(); ++i)
      js << "            \"dim\": \"" << shape[i] << "\"" << (i + 1 < shape.size() ? ",\n" : "\n");
    js << "        }\n";
  } else {
    js << "\n";
  }
  js << "    },\n";
  js << "    \"encoded\": {\n";
  js << "        \"bytes\": \"" << payload_bytes << "\"\n";
  js << "    },\n";
  js << "    \"sqy\": {\n";
  js << "        \"version\": \"" << SQYB_VERSION_STRING << "\",\n";
  js << "        \"headref\": \"" << SQYB_VERSION_HEADREF << "\"\n";
  js << "    }\n";
  js << "}\n";
  js << kHeaderDelim;
  std::string s = js.str();
  if (sizeof_raw > 1 && s.size() % sizeof_raw != 0) s = std::string(sizeof_raw - (s.size() % sizeof_raw), ' ') + s;
  return s;
}

unsigned sizeof_typename(const std::string& t) {
  if (t == "uint8" || t == "int8") return 1;
  if (t == "uint16" || t == "int16") return 2;
  if (t == "uint32" || t == "int32") return 4;
  if (t == "uint64" || t == "int64") return 8;
  return 0;
}

namespace {
// Minimal JSON reader for the header: nested objects, string / bare scalar values, repeated keys.
struct JsonCursor {
  const char* p;
  const char* e;
  bool ok = true;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p; }
  bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  std::string str() {
    std::string o;
    ws();
    if (p >= e || *p != '"') { ok = false; return o; }
    ++p;
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) {
        ++p;
        switch (*p) {
          case 'b': o.push_back('\b'); break;
          case 'f': o.push_back('\f'); break;
          case 'n': o.push_back('\n'); break;
          case 'r': o.push_back('\r'); break;
          case 't': o.push_back('\t'); break;
          case 'u': {
            if (p + 4 < e) {
              unsigned v = 0;
              for (int i = 1; i <= 4; ++i) {
                char c = p[i];
                v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : 0);
              }
              o.push_back((char)(v & 0xff));
              p += 4;
            }
            break;
          }
          default: o.push_back(*p); break;  // '"', '\\', '/'
        }
        ++p;
      } else {
        o.push_back(*p++);
      }
    }
    if (p >= e) { ok = false; return o; }
    ++p;
    return o;
  }
  // flattens into (dotted path, value) pairs, preserving order
  void object(const std::string& prefix, std::vector<std::pair<std::string, std::string>>& out, int depth) {
    if (!eat('{') || depth > 8) { ok = false; return; }
    if (eat('}')) return;
    while (ok) {
      std::string key = str();
      if (!ok || !eat(':')) { ok = false; return; }
      std::string path = prefix.empty() ? key : prefix + "." + key;
      ws();
      if (p < e && *p == '{') object(path, out, depth + 1);
      else if (p < e && *p == '"') out.emplace_back(path, str());
      else {  // bare scalar
        const char* b = p;
        while (p < e && *p != ',' && *p != '}' && *p != '\n') ++p;
        out.emplace_back(path, std::string(b, p));
      }
      if (eat(',')) continue;
      if (eat('}')) return;
      ok = false;
    }
  }
};
}  // namespace

Header unpack_header(const char* buf, size_t nbytes) {
  Header h;
  if (!buf || !nbytes) return h;
  const size_t nd = sizeof(kHeaderDelim) - 1;
  // first occurrence of the delimiter (sqeazy_header.hpp:521-535)
  const char* end = nullptr;
  for (size_t i = 0; i + nd <= nbytes; ++i) {
    if (buf[i] == kHeaderDelim[0] && std::memcmp(buf + i, kHeaderDelim, nd) == 0) { end = buf + i; break; }
  }
  if (!end) return h;
  size_t open = 0, close = 0, colons = 0;
  for (const char* p = buf; p < end + nd; ++p) { open += *p == '{'; close += *p == '}'; colons += *p == ':'; }
  if (!close || open != close || colons < 2) return h;  // valid_header :506-517
  JsonCursor c{buf, end};
  std::vector<std::pair<std::string, std::string>> kv;
  c.object("", kv, 0);
  if (!c.ok) return h;
  h.raw_type = "";
  for (auto& e : kv) {
    if (e.first == "pipename") h.pipeline = e.second;
    else if (e.first == "raw.type") h.raw_type = e.second;
    else if (e.first == "raw.shape.dim") h.shape.push_back(std::strtoull(e.second.c_str(), nullptr, 10));
    else if (e.first == "encoded.bytes") h.compressed_bytes = std::strtoull(e.second.c_str(), nullptr, 10);
  }
  h.size = size_t(end - buf) + nd;
  h.valid = true;
  return h;
}

}  // namespace sqyb
