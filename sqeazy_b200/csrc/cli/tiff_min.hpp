// Minimal TIFF 6.0 reader/writer for the stacks the sqy CLI handles: uncompressed, one sample per pixel, 8 or 16
// bits, any number of pages, strips of any height, little or big endian (classic TIFF, 32-bit offsets).
// The reference reads and writes the same subset through libtiff (tiff_utils.hpp:270-315 write_tiff_from_array,
// :562-582 load_to_buffer: scanline I/O of PLANARCONFIG_CONTIG / PHOTOMETRIC_MINISBLACK / COMPRESSION_NONE pages);
// libtiff is not in this image, and the format subset is small enough to state directly.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

namespace sqycli {

struct TiffStack {
  std::vector<uint64_t> shape;   // {z, y, x} (C order, like image_stack / tiff_facet)
  int bits = 16;                 // 8 or 16
  std::vector<char> data;        // native little-endian samples
  uint64_t voxels() const { return shape.size() == 3 ? shape[0] * shape[1] * shape[2] : 0; }
};

namespace detail {
struct Reader {
  FILE* f = nullptr;
  bool be = false;
  uint16_t u16(const unsigned char* p) const { return be ? (uint16_t)((p[0] << 8) | p[1]) : (uint16_t)(p[0] | (p[1] << 8)); }
  uint32_t u32(const unsigned char* p) const {
    return be ? ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]
              : ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0];
  }
  bool at(uint64_t off, void* dst, size_t n) { return fseek(f, (long)off, SEEK_SET) == 0 && fread(dst, 1, n, f) == n; }
};
}  // namespace detail

// returns an empty string on success, an error text otherwise
inline std::string tiff_read(const std::string& path, TiffStack& out) {
  detail::Reader r;
  r.f = fopen(path.c_str(), "rb");
  if (!r.f) return "unable to open " + path;
  struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{r.f};
  unsigned char hdr[8];
  if (!r.at(0, hdr, 8)) return "not a TIFF file (too short)";
  if (hdr[0] == 'I' && hdr[1] == 'I') r.be = false;
  else if (hdr[0] == 'M' && hdr[1] == 'M') r.be = true;
  else return "not a TIFF file (byte order mark)";
  if (r.u16(hdr + 2) != 42) return r.u16(hdr + 2) == 43 ? "BigTIFF is not supported" : "not a TIFF file (magic)";
  uint64_t ifd = r.u32(hdr + 4);
  out.shape.clear();
  out.data.clear();
  uint64_t pages = 0, W = 0, H = 0;
  // a classic TIFF is at most 4 GiB: its size bounds the page data, and every IFD lies strictly inside it; an IFD chain
  // that returns to an offset already walked (a cycle) is refused instead of being followed for ever
  if (fseek(r.f, 0, SEEK_END) != 0) return "unable to size " + path;
  const uint64_t file_bytes = (uint64_t)ftell(r.f);
  std::vector<uint64_t> seen;
  while (ifd != 0) {
    if (ifd + 6 > file_bytes) return "IFD offset outside the file";
    if (std::find(seen.begin(), seen.end(), ifd) != seen.end()) return "IFD chain loops";
    seen.push_back(ifd);
    if (seen.size() > (1u << 20)) return "too many TIFF pages";
    unsigned char nb[2];
    if (!r.at(ifd, nb, 2)) return "truncated IFD";
    const unsigned n = r.u16(nb);
    std::vector<unsigned char> ent(12 * n + 4);
    if (!r.at(ifd + 2, ent.data(), ent.size())) return "truncated IFD";
    uint64_t w = 0, h = 0, bits = 1, spp = 1, comp = 1, rps = 0xFFFFFFFFull;
    uint64_t so_count = 0, so_off = 0, so_type = 4, sb_count = 0, sb_off = 0, sb_type = 4;
    for (unsigned i = 0; i < n; ++i) {
      const unsigned char* e = ent.data() + 12 * i;
      const unsigned tag = r.u16(e), type = r.u16(e + 2);
      const uint64_t count = r.u32(e + 4);
      const uint64_t v = (type == 3 && count == 1) ? r.u16(e + 8) : r.u32(e + 8);
      switch (tag) {
        case 256: w = v; break;
        case 257: h = v; break;
        case 258: bits = v; break;
        case 259: comp = v; break;
        case 277: spp = v; break;
        case 278: rps = v; break;
        case 273: so_count = count; so_type = type; so_off = (count == 1) ? v : r.u32(e + 8); if (count == 2 && type == 3) so_off = ifd + 2 + 12 * i + 8; break;
        case 279: sb_count = count; sb_type = type; sb_off = (count == 1) ? v : r.u32(e + 8); if (count == 2 && type == 3) sb_off = ifd + 2 + 12 * i + 8; break;
        default: break;
      }
    }
    if (comp != 1) return "compressed TIFF pages are not supported";
    if (spp != 1) return "only single-channel TIFF stacks are supported";
    if (bits != 8 && bits != 16) return "only 8 and 16 bit TIFF stacks are supported";
    if (w == 0 || h == 0 || so_count == 0 || so_count != sb_count) return "TIFF page without usable strips";
    if (pages == 0) { W = w; H = h; out.bits = (int)bits; }
    else if (w != W || h != H || (int)bits != out.bits) return "pages of different shape or depth";
    // strip tables
    std::vector<uint64_t> offs(so_count), lens(so_count);
    auto table = [&](uint64_t count, uint64_t type, uint64_t off, std::vector<uint64_t>& dst) -> bool {
      if (count == 1) { dst[0] = off; return true; }
      const size_t es = type == 3 ? 2 : 4;
      std::vector<unsigned char> raw(count * es);
      if (!r.at(off, raw.data(), raw.size())) return false;
      for (uint64_t k = 0; k < count; ++k) dst[k] = type == 3 ? r.u16(raw.data() + 2 * k) : r.u32(raw.data() + 4 * k);
      return true;
    };
    if (!table(so_count, so_type, so_off, offs) || !table(sb_count, sb_type, sb_off, lens)) return "truncated strip table";
    if (W > file_bytes || H > file_bytes || W * H > file_bytes) return "TIFF page larger than the file";
    const uint64_t page_bytes = W * H * (bits / 8);
    if (page_bytes > file_bytes) return "TIFF page larger than the file";
    const size_t base = out.data.size();
    out.data.resize(base + page_bytes);
    uint64_t got = 0;
    for (uint64_t k = 0; k < so_count && got < page_bytes; ++k) {
      const uint64_t take = lens[k] < page_bytes - got ? lens[k] : page_bytes - got;
      if (!r.at(offs[k], out.data.data() + base + got, take)) return "truncated strip";
      got += take;
    }
    if (got != page_bytes) return "TIFF page shorter than its dimensions";
    (void)rps;
    if (bits == 16 && r.be) {
      char* p = out.data.data() + base;
      for (uint64_t k = 0; k + 1 < page_bytes; k += 2) { const char t = p[k]; p[k] = p[k + 1]; p[k + 1] = t; }
    }
    ++pages;
    ifd = r.u32(ent.data() + 12 * n);
  }
  if (pages == 0) return "TIFF file without pages";
  out.shape = {pages, H, W};
  return "";
}

// classic little-endian TIFF, one strip per page, tags as write_tiff_from_array sets them (tiff_utils.hpp:286-297)
inline std::string tiff_write(const std::string& path, const std::vector<uint64_t>& shape_in, int bits, const char* data) {
  std::vector<uint64_t> shape = shape_in;
  while (shape.size() < 3) shape.insert(shape.begin(), 1);
  if (shape.size() != 3) return "only stacks of rank <= 3 can be written as TIFF";
  const uint64_t Z = shape[0], H = shape[1], W = shape[2];
  const uint64_t page_bytes = W * H * (uint64_t)(bits / 8);
  const unsigned ntags = 13;
  const uint64_t ifd_bytes = 2 + 12ull * ntags + 4;
  const uint64_t total = 8 + Z * (page_bytes + ifd_bytes + (page_bytes & 1));
  if (total >= 0xFFFFFFF0ull) return "stack too large for classic TIFF (4 GiB)";
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return "unable to open " + path;
  auto p16 = [](unsigned char* p, uint32_t v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); };
  auto p32 = [](unsigned char* p, uint32_t v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); p[2] = (unsigned char)(v >> 16); p[3] = (unsigned char)(v >> 24); };
  unsigned char hdr[8] = {'I', 'I', 42, 0, 0, 0, 0, 0};
  uint64_t pos = 8;
  p32(hdr + 4, (uint32_t)(Z ? pos + page_bytes + (page_bytes & 1) : 0));
  bool ok = fwrite(hdr, 1, 8, f) == 8;
  for (uint64_t z = 0; z < Z && ok; ++z) {
    const uint64_t data_off = pos;
    ok = fwrite(data + z * page_bytes, 1, page_bytes, f) == page_bytes;
    pos += page_bytes;
    if (page_bytes & 1) { const char pad = 0; ok = ok && fwrite(&pad, 1, 1, f) == 1; pos += 1; }
    std::vector<unsigned char> ifd(ifd_bytes, 0);
    p16(ifd.data(), ntags);
    unsigned k = 0;
    auto tag = [&](unsigned id, unsigned type, uint32_t count, uint32_t v) {
      unsigned char* e = ifd.data() + 2 + 12 * k++;
      p16(e, id); p16(e + 2, type); p32(e + 4, count);
      if (type == 3 && count == 1) p16(e + 8, v);
      else if (type == 3 && count == 2) { p16(e + 8, v & 0xffffu); p16(e + 10, v >> 16); }
      else p32(e + 8, v);
    };
    tag(254, 4, 1, 2);                         // SubfileType: FILETYPE_PAGE
    tag(256, 4, 1, (uint32_t)W);               // ImageWidth
    tag(257, 4, 1, (uint32_t)H);               // ImageLength
    tag(258, 3, 1, (uint32_t)bits);            // BitsPerSample
    tag(259, 3, 1, 1);                         // Compression: none
    tag(262, 3, 1, 1);                         // Photometric: MINISBLACK
    tag(273, 4, 1, (uint32_t)data_off);        // StripOffsets
    tag(277, 3, 1, 1);                         // SamplesPerPixel
    tag(278, 4, 1, (uint32_t)H);               // RowsPerStrip
    tag(279, 4, 1, (uint32_t)page_bytes);      // StripByteCounts
    tag(284, 3, 1, 1);                         // PlanarConfig: CONTIG
    tag(297, 3, 2, (uint32_t)(z & 0xffffu) | ((uint32_t)(Z & 0xffffu) << 16));   // PageNumber
    tag(339, 3, 1, 1);                         // SampleFormat: unsigned integer
    const uint64_t next = (z + 1 < Z) ? pos + ifd_bytes + page_bytes + (page_bytes & 1) : 0;
    p32(ifd.data() + 2 + 12 * ntags, (uint32_t)next);
    ok = ok && fwrite(ifd.data(), 1, ifd.size(), f) == ifd.size();
    pos += ifd_bytes;
  }
  fclose(f);
  return ok ? "" : "short write to " + path;
}

}  // namespace sqycli
