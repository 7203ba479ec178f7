/* sqeazy hot-path ORACLE — plain C restatement of the reference's algorithms.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under sqeazy_b200/ may include, link or load this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it, and only
 * as the checker. Parity status: PINNED — every function below is checked (tests/test_oracle.py)
 * against the reference's own known-answer vectors (tests/test_bitswap_scheme_impl.cpp:286-329,
 * test_quantiser_impl.cpp:862-1020, test_histogram_fill.cpp:31-53, test_lz4_utils_impl.cpp:147-180)
 * and, in this container, against the reference's real stage code compiled into oracle/_ref
 * (oracle/ref_harness.cpp). The LZ4 arithmetic lives in an un-vendored third-party dependency
 * (github.com/lz4/lz4, version unpinned by the reference; liblz4 1.9.4 is the executable stand-in in
 * this image): the block/frame decoder below restates the published LZ4 Block / Frame format and is
 * pinned by decoding liblz4-produced frames; compressed-byte parity is undefined by the reference.
 *
 * Each function cites the reference file:line it follows (paths relative to src/cpp/src/).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * bitswapN — encoders/bitplane_reorder_scalar.hpp:27-74 (encode), :81-116 (decode),
 *            encoders/bitswap_scheme_impl.hpp:97-103 (verbatim tail)
 * ------------------------------------------------------------------------------------------ */
int orc_bitswap_encode(int w, const uint16_t* in, uint16_t* out, uint64_t n) {
  if (w != 1 && w != 2 && w != 4 && w != 8) return 1;
  const uint64_t P = 16 / (uint64_t)w;
  const uint64_t np = n - (n % P);
  const uint64_t S = np / P;
  const uint32_t mask = (1u << w) - 1u;
  for (uint64_t i = 0; i < np; ++i) out[i] = 0;
  for (uint64_t i = 0; i < np; ++i) {
    const uint32_t v = in[i];
    const uint64_t g = i / P, j = i % P;
    for (uint64_t p = 0; p < P; ++p) {
      const uint32_t field = (v >> (p * w)) & mask;
      out[(P - 1 - p) * S + g] |= (uint16_t)(field << ((16 - w) - j * w));
    }
  }
  for (uint64_t i = np; i < n; ++i) out[i] = in[i];
  return 0;
}

int orc_bitswap_decode(int w, const uint16_t* in, uint16_t* out, uint64_t n) {
  if (w != 1 && w != 2 && w != 4 && w != 8) return 1;
  const uint64_t P = 16 / (uint64_t)w;
  const uint64_t np = n - (n % P);
  const uint64_t S = np / P;
  const uint32_t mask = (1u << w) - 1u;
  for (uint64_t i = 0; i < np; ++i) {
    const uint64_t g = i / P, j = i % P;
    uint32_t v = 0;
    for (uint64_t p = 0; p < P; ++p) {
      const uint32_t word = in[(P - 1 - p) * S + g];
      v |= ((word >> ((16 - w) - j * w)) & mask) << (p * w);
    }
    out[i] = (uint16_t)v;
  }
  for (uint64_t i = np; i < n; ++i) out[i] = in[i];
  return 0;
}

/* uint8 volumes (dypeline<uint8_t>, src/sqeazy.cpp:72-106): the same scalar template with raw_type = uint8_t
 * (type_width 8; bitswap_scheme_impl.hpp:106-121 excludes the SSE path for sizeof(raw_type) == 1) */
int orc_bitswap8_encode(int w, const uint8_t* in, uint8_t* out, uint64_t n) {
  if (w != 1 && w != 2 && w != 4) return 1;
  const uint64_t P = 8 / (uint64_t)w;
  const uint64_t np = n - (n % P);
  const uint64_t S = np / P;
  const uint32_t mask = (1u << w) - 1u;
  for (uint64_t i = 0; i < np; ++i) out[i] = 0;
  for (uint64_t i = 0; i < np; ++i) {
    const uint32_t v = in[i];
    const uint64_t g = i / P, j = i % P;
    for (uint64_t p = 0; p < P; ++p) {
      const uint32_t field = (v >> (p * w)) & mask;
      out[(P - 1 - p) * S + g] |= (uint8_t)(field << ((8 - w) - j * w));
    }
  }
  for (uint64_t i = np; i < n; ++i) out[i] = in[i];
  return 0;
}

int orc_bitswap8_decode(int w, const uint8_t* in, uint8_t* out, uint64_t n) {
  if (w != 1 && w != 2 && w != 4) return 1;
  const uint64_t P = 8 / (uint64_t)w;
  const uint64_t np = n - (n % P);
  const uint64_t S = np / P;
  const uint32_t mask = (1u << w) - 1u;
  for (uint64_t i = 0; i < np; ++i) {
    const uint64_t g = i / P, j = i % P;
    uint32_t v = 0;
    for (uint64_t p = 0; p < P; ++p) {
      const uint32_t word = in[(P - 1 - p) * S + g];
      v |= ((word >> ((8 - w) - j * w)) & mask) << (p * w);
    }
    out[i] = (uint8_t)v;
  }
  for (uint64_t i = np; i < n; ++i) out[i] = in[i];
  return 0;
}

void orc_remove_background8(const uint8_t* in, uint8_t* out, uint64_t n, uint8_t threshold) {
  for (uint64_t i = 0; i < n; ++i) out[i] = in[i] > threshold ? (uint8_t)(in[i] - threshold) : 0;
}

/* ------------------------------------------------------------------------------------------
 * remove_background — encoders/remove_background_scheme_impl.hpp:73-95
 * ------------------------------------------------------------------------------------------ */
void orc_remove_background(const uint16_t* in, uint16_t* out, uint64_t n, uint16_t threshold) {
  for (uint64_t i = 0; i < n; ++i) out[i] = in[i] > threshold ? (uint16_t)(in[i] - threshold) : 0;
}

/* ------------------------------------------------------------------------------------------
 * histogram — encoders/histogram_utils.hpp:41-55 ; bins are uint32 and wrap
 * ------------------------------------------------------------------------------------------ */
void orc_histogram(const uint16_t* in, uint64_t n, uint32_t* bins /* 65536, accumulated */) {
  for (uint64_t i = 0; i < n; ++i) bins[in[i]]++;
}

/* calc_support — hist_impl.hpp:63-84 (support_index), :359-381 (calc_support).
 * add_from_image never refreshes small/large populated bin, so all 65536 bins are scanned. */
float orc_support(const uint32_t* bins, float threshold) {
  float result = 0;
  if (threshold > 1. || threshold < 0.) return result;
  int isum = 0;
  for (uint32_t i = 0; i < 65536; ++i) isum = (int)((unsigned)isum + bins[i]); /* std::accumulate(.., 0) */
  const double total = isum;
  double running = 0;
  uint32_t support = 65536;
  for (uint32_t i = 0; i < 65536; ++i) {
    running += bins[i];
    if ((running / total) > threshold) { support = i; break; }
  }
  const uint16_t m = (uint16_t)support;
  if (m > 0) {
    const uint32_t num = bins[m] * (uint32_t)m + bins[m - 1] * (uint32_t)(m - 1);
    result = (float)num / (float)(bins[m - 1] + bins[m]);
  }
  return result;
}

/* extract_darkest_face_supports — encoders/background_scheme_utils.hpp:35-105
 * l2_bytes = compass::runtime::size::cache::level(2) of the host that made the blob. */
void orc_darkest_face_supports(const uint16_t* in, uint64_t Z, uint64_t Y, uint64_t X, uint64_t l2_bytes, float* out4) {
  const uint64_t frame = Y * X;
  const uint64_t portion = frame > l2_bytes ? (uint64_t)((double)l2_bytes * .75) : frame;
  uint32_t* bins = (uint32_t*)calloc(65536, sizeof(uint32_t));
  const uint64_t zi[2] = {0, Z - 1};
  for (int i = 0; i < 2; ++i) {
    memset(bins, 0, 65536 * sizeof(uint32_t));
    if (zi[i] < Z) orc_histogram(in + zi[i] * frame, portion, bins);
    out4[i] = orc_support(bins, 0.99f);
  }
  const uint64_t yi[2] = {0, Y - 1};
  const uint64_t zo[3] = {1, Z / 2, Z - 2};
  for (int i = 0; i < 2; ++i) {
    memset(bins, 0, 65536 * sizeof(uint32_t));
    for (int k = 0; k < 3; ++k)
      if (zo[k] < Z) orc_histogram(in + zo[k] * frame + yi[i] * X, X, bins);
    out4[2 + i] = orc_support(bins, 0.99f);
  }
  free(bins);
}

/* rmestbkrd — encoders/remove_estimated_background_scheme_impl.hpp:71-112 */
int orc_rmestbkrd(const uint16_t* in, uint16_t* out, uint64_t Z, uint64_t Y, uint64_t X, uint64_t l2_bytes, int* threshold) {
  float s[4];
  orc_darkest_face_supports(in, Z, Y, X, l2_bytes, s);
  float mn = s[0];
  for (int i = 1; i < 4; ++i) if (s[i] < mn) mn = s[i];
  const uint16_t t = (uint16_t)mn;
  if (threshold) *threshold = t;
  orc_remove_background(in, out, Z * Y * X, t);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * quantiser LUTs — encoders/quantiser_utils.hpp:386-418 (setup_com), :155-168 (importance),
 *                  :286-306 (linear_mapping_quantisation), :227-284 (adaptive_lloyd_com)
 * enc: 65536 x u8 (the reference's char codes, compared as unsigned), dec: 256 x u16
 * ------------------------------------------------------------------------------------------ */
void orc_quantiser_luts(const uint32_t* hist, uint8_t* enc, uint16_t* dec) {
  memset(enc, 0, 65536);
  memset(dec, 0, 512);
  float* imp = (float*)malloc(65536 * sizeof(float));
  double acc = 0.;
  uint32_t levels = 0;
  for (uint32_t v = 0; v < 65536; ++v) {
    imp[v] = (float)hist[v] * 1.0f;
    acc += imp[v];
    if (imp[v] != 0.f) levels++;
  }
  const float sum = (float)acc;
  if (!(sum != 0)) { free(imp); return; }
  if (levels <= 256) {
    uint32_t c = 0;
    for (uint32_t v = 0; v < 65536 && c < 256; ++v) {
      enc[v] = (uint8_t)c;
      dec[c] = (uint16_t)v;
      if (imp[v] != 0.f) c++;
    }
    if (c < 256 && c > 0 && dec[c] == 65535)
      for (uint32_t k = c; k < 256; ++k) dec[k] = dec[c - 1];
    free(imp);
    return;
  }
  size_t levels_available = 256;
  volatile float bucket = sum / levels_available;
  float integral = imp[0], q = imp[0];
  uint32_t c = 0;
  float wm = 0 * imp[0];
  float idx = 0;
  for (uint32_t v = 1; v < 65536; ++v) {
    if (q >= bucket && c < 255) {
      dec[c] = (uint16_t)idx;
      c++;
      levels_available--;
      q = imp[v];
      wm = (float)v * imp[v];
      if (integral < sum) bucket = (sum - integral) / levels_available;
      if (q != 0.) idx = roundf(wm / q);
    } else {
      q += imp[v];
      volatile float prod = (float)v * imp[v];
      wm += prod;
      if (q != 0.) idx = roundf(wm / q);
    }
    enc[v] = (uint8_t)c;
    integral += imp[v];
  }
  dec[c] = (uint16_t)idx;
  free(imp);
}

void orc_lut_apply(const uint16_t* in, uint8_t* out, uint64_t n, const uint8_t* enc) {
  for (uint64_t i = 0; i < n; ++i) out[i] = enc[in[i]];
}
/* unsigned index: the intended semantics (the reference's signed char index is UB for codes >= 128) */
void orc_lut_decode(const uint8_t* in, uint16_t* out, uint64_t n, const uint16_t* dec) {
  for (uint64_t i = 0; i < n; ++i) out[i] = dec[in[i]];
}

/* ------------------------------------------------------------------------------------------
 * lz4 helper tables — encoders/lz4_utils.hpp:60-93 (closest_blocksize), encoders/lz4.hpp:91-101
 * ------------------------------------------------------------------------------------------ */
uint32_t orc_lz4_closest_blocksize_kb(uint32_t kb) {
  static const uint32_t sizes[4] = {64, 256, 1024, 4096};
  int up = 0;
  while (up < 4 && sizes[up] < kb) up++;
  if (up == 4) return sizes[3];
  if (up == 0) return sizes[0];
  const uint32_t middle = sizes[up - 1] + (sizes[up] - sizes[up - 1]) / 2;
  return kb >= middle ? sizes[up] : sizes[up - 1];
}

/* ------------------------------------------------------------------------------------------
 * LZ4 block decoder (published LZ4 Block Format). `prefix` = bytes available in front of dst
 * (linked blocks). Returns decoded size or -1.
 * ------------------------------------------------------------------------------------------ */
static int64_t lz4_block_decode(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, uint64_t prefix) {
  uint64_t ip = 0, op = 0;
  while (ip < n) {
    const uint32_t token = src[ip++];
    uint64_t lit = token >> 4;
    if (lit == 15) {
      uint32_t b;
      do {
        if (ip >= n) return -1;
        b = src[ip++];
        lit += b;
      } while (b == 255);
    }
    if (ip + lit > n || op + lit > cap) return -1;
    memcpy(dst + op, src + ip, lit);
    ip += lit;
    op += lit;
    if (ip >= n) break;
    if (ip + 2 > n) return -1;
    const uint32_t offset = src[ip] | (src[ip + 1] << 8);
    ip += 2;
    uint64_t ml = token & 15;
    if (ml == 15) {
      uint32_t b;
      do {
        if (ip >= n) return -1;
        b = src[ip++];
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (offset == 0 || offset > op + prefix || op + ml > cap) return -1;
    for (uint64_t k = 0; k < ml; ++k) dst[op + k] = dst[op + k - offset];
    op += ml;
  }
  return (int64_t)op;
}

static uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

/* Multi-frame LZ4F decode, the loop of encoders/lz4.hpp:257-339 (frames concatenated, skippable frames
 * skipped, linked or independent blocks, stored blocks, optional checksums ignored).
 * Returns 0 and *decoded on success. */
int orc_lz4_frames_decode(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, uint64_t* decoded) {
  uint64_t pos = 0, op = 0;
  while (pos + 4 <= n) {
    const uint32_t magic = rd32(src + pos);
    if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
      if (pos + 8 > n) return 1;
      pos += 8ull + rd32(src + pos + 4);
      continue;
    }
    if (magic != 0x184D2204u || pos + 7 > n) return 1;
    const uint32_t flg = src[pos + 4], bd = src[pos + 5];
    const int indep = (flg >> 5) & 1, bchk = (flg >> 4) & 1, csz = (flg >> 3) & 1, cchk = (flg >> 2) & 1, dict = flg & 1;
    if ((flg >> 6) != 1) return 1;
    const uint32_t bsid = (bd >> 4) & 7;
    if (bsid < 4) return 1;
    const uint64_t maxblock = 1ull << (8 + 2 * bsid);
    pos += 7 + (csz ? 8 : 0) + (dict ? 4 : 0);
    const uint64_t frame_start = op;
    for (;;) {
      if (pos + 4 > n) return 1;
      const uint32_t word = rd32(src + pos);
      pos += 4;
      if (word == 0) break;
      const uint64_t sz = word & 0x7FFFFFFFu;
      if (sz > maxblock || pos + sz > n) return 1;
      if (word & 0x80000000u) {
        if (op + sz > cap) return 1;
        memcpy(dst + op, src + pos, sz);
        op += sz;
      } else {
        const uint64_t room = cap - op < maxblock ? cap - op : maxblock;
        const int64_t got = lz4_block_decode(src + pos, sz, dst + op, room, indep ? 0 : op - frame_start);
        if (got < 0) return 1;
        op += (uint64_t)got;
      }
      pos += sz + (bchk ? 4 : 0);
    }
    if (cchk) pos += 4;
  }
  if (decoded) *decoded = op;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * LZ4 block compressor (greedy, 4096-entry hash of 4-byte sequences; published format rules:
 * last 5 bytes literal, last match starts >= 12 bytes before the end). CPU baseline "port" and
 * a source of valid frames for decoder tests; not byte-identical to liblz4 (undefined by the reference).
 * ------------------------------------------------------------------------------------------ */
static uint64_t put_len(uint8_t* p, uint64_t v) {
  uint64_t k = 0;
  v -= 15;
  while (v >= 255) { p[k++] = 255; v -= 255; }
  p[k++] = (uint8_t)v;
  return k;
}

static int64_t lz4_block_compress(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap) {
  uint32_t table[4096];
  memset(table, 0xff, sizeof(table));
  uint64_t ip = 0, anchor = 0, op = 0;
  if (n >= 13) {
    const uint64_t mflimit = n - 12, matchlimit = n - 5;
    while (ip <= mflimit) {
      uint32_t v;
      memcpy(&v, src + ip, 4);
      const uint32_t h = (v * 2654435761u) >> 20;
      const uint32_t c = table[h];
      table[h] = (uint32_t)ip;
      uint32_t cv = 0;
      if (c != 0xffffffffu) memcpy(&cv, src + c, 4);
      if (c == 0xffffffffu || cv != v || ip - c > 65535) { ip++; continue; }
      uint64_t ml = 4;
      while (ip + ml < matchlimit && src[ip + ml] == src[c + ml]) ml++;
      const uint64_t lit = ip - anchor;
      if (op + 1 + lit / 255 + 1 + lit + 2 + ml / 255 + 1 > cap) return -1;
      uint8_t* token = dst + op++;
      *token = (uint8_t)((lit < 15 ? lit : 15) << 4);
      if (lit >= 15) op += put_len(dst + op, lit);
      memcpy(dst + op, src + anchor, lit);
      op += lit;
      dst[op++] = (uint8_t)((ip - c) & 0xff);
      dst[op++] = (uint8_t)((ip - c) >> 8);
      *token |= (uint8_t)(ml - 4 < 15 ? ml - 4 : 15);
      if (ml - 4 >= 15) op += put_len(dst + op, ml - 4);
      ip += ml;
      anchor = ip;
    }
  }
  const uint64_t lit = n - anchor;
  if (op + 1 + lit / 255 + 1 + lit > cap) return -1;
  uint8_t* token = dst + op++;
  *token = (uint8_t)((lit < 15 ? lit : 15) << 4);
  if (lit >= 15) op += put_len(dst + op, lit);
  memcpy(dst + op, src + anchor, lit);
  op += lit;
  return (int64_t)op;
}

/* One frame per `chunk` bytes (the layout of the reference's parallel mode, encoders/lz4_utils.hpp:193-274),
 * independent 256 KiB-max blocks. Returns payload bytes or -1. */
int64_t orc_lz4_frames_encode(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, uint64_t chunk) {
  static const uint8_t hdr[7] = {0x04, 0x22, 0x4D, 0x18, 0x60, 0x50, 0xFB};
  const uint64_t block = 262144;
  uint64_t op = 0;
  if (chunk == 0 || chunk > n) chunk = n ? n : 1;
  uint8_t* tmp = (uint8_t*)malloc(block + block / 255 + 64);
  for (uint64_t c0 = 0; c0 < n || (n == 0 && c0 == 0); c0 += chunk) {
    const uint64_t c1 = c0 + chunk < n ? c0 + chunk : n;
    if (op + 7 > cap) { free(tmp); return -1; }
    memcpy(dst + op, hdr, 7);
    op += 7;
    for (uint64_t b0 = c0; b0 < c1; b0 += block) {
      const uint64_t bn = b0 + block < c1 ? block : c1 - b0;
      int64_t cs = lz4_block_compress(src + b0, bn, tmp, block + block / 255 + 64);
      uint32_t word;
      const uint8_t* data;
      if (cs < 0 || (uint64_t)cs >= bn) { word = (uint32_t)bn | 0x80000000u; data = src + b0; cs = (int64_t)bn; }
      else { word = (uint32_t)cs; data = tmp; }
      if (op + 4 + (uint64_t)cs > cap) { free(tmp); return -1; }
      dst[op] = (uint8_t)word; dst[op + 1] = (uint8_t)(word >> 8); dst[op + 2] = (uint8_t)(word >> 16); dst[op + 3] = (uint8_t)(word >> 24);
      memcpy(dst + op + 4, data, (size_t)cs);
      op += 4 + (uint64_t)cs;
    }
    if (op + 4 > cap) { free(tmp); return -1; }
    memset(dst + op, 0, 4);
    op += 4;
    if (n == 0) break;
  }
  free(tmp);
  return (int64_t)op;
}

/* ------------------------------------------------------------------------------------------
 * base64 — base64.hpp:135-162 (encode), :177-202 (decode)
 * ------------------------------------------------------------------------------------------ */
static const char b64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
uint64_t orc_base64_encode(const uint8_t* p, uint64_t n, char* out) {
  uint64_t o = 0, i = 0;
  for (; i + 2 < n; i += 3) {
    const uint32_t v = (p[i] << 16) | (p[i + 1] << 8) | p[i + 2];
    out[o++] = b64[(v >> 18) & 63]; out[o++] = b64[(v >> 12) & 63]; out[o++] = b64[(v >> 6) & 63]; out[o++] = b64[v & 63];
  }
  if (n - i == 1) {
    const uint32_t v = p[i] << 16;
    out[o++] = b64[(v >> 18) & 63]; out[o++] = b64[(v >> 12) & 63]; out[o++] = '='; out[o++] = '=';
  } else if (n - i == 2) {
    const uint32_t v = (p[i] << 16) | (p[i + 1] << 8);
    out[o++] = b64[(v >> 18) & 63]; out[o++] = b64[(v >> 12) & 63]; out[o++] = b64[(v >> 6) & 63]; out[o++] = '=';
  }
  return o;
}

/* ------------------------------------------------------------------------------------------
 * bitshuffle — encoders/bitshuffle_scheme_impl.hpp:91-160 calls bshuf_bitshuffle_nthreads / bshuf_bitunshuffle of the
 * THIRD-PARTY bitshuffle library (github.com/kiyo-masui/bitshuffle; the reference downloads its sources at cmake time,
 * none of them is under /root/reference and no version is pinned). PARITY UNPINNED for this stage: the reference's own
 * tests hold round trips only (tests/test_bitshuffle_scheme_impl.cpp:21-190), and the library cannot be compiled here.
 * What follows restates the library's published scalar algorithm (src/bitshuffle_core.c) step by step:
 *   bshuf_blocked_wrap_fun   : whole blocks of `block_size` elements (0 -> bshuf_default_block_size: 8192 bytes / element
 *                              size, rounded down to a multiple of BSHUF_BLOCKED_MULT = 8, at least 128), one block of the
 *                              rest rounded down to a multiple of 8, the last size % 8 elements copied
 *   bshuf_trans_bit_elem     : per block  A bshuf_trans_byte_elem_scal (byte j of every element -> byte row j)
 *                                         B bshuf_trans_bit_byte_scal  (8 bytes -> one byte in each of 8 bit rows, TRANS_BIT_8X8,
 *                                                                       byte 8i+k of the input at bit k)
 *                                         C bshuf_trans_bitrow_eight   (8 x elem_size matrix of rows -> elem_size x 8)
 * tests/test_oracle.py checks it against an independent numpy formulation (np.unpackbits, bitorder="little").
 * ------------------------------------------------------------------------------------------ */
static uint32_t orc_bshuf_default_block_size(uint32_t elem_size) {
  uint32_t b = 8192u / elem_size;
  b = (b / 8u) * 8u;
  return b < 128u ? 128u : b;
}

static void orc_bshuf_trans_bit_elem(const uint8_t* in, uint8_t* out, uint8_t* tmp, size_t size, size_t elem_size) {
  const size_t nbyte = size * elem_size, nrow = nbyte / 8;
  /* A: out[j * size + i] = in[i * elem_size + j] */
  for (size_t i = 0; i < size; ++i)
    for (size_t j = 0; j < elem_size; ++j) out[j * size + i] = in[i * elem_size + j];
  /* B: tmp[k * nrow + i] bit b = bit k of out[8 i + b] */
  for (size_t i = 0; i < nrow; ++i)
    for (size_t k = 0; k < 8; ++k) {
      uint8_t v = 0;
      for (size_t b = 0; b < 8; ++b) v |= (uint8_t)(((out[8 * i + b] >> k) & 1u) << b);
      tmp[k * nrow + i] = v;
    }
  /* C: bshuf_trans_elem(tmp, out, 8, elem_size, size / 8): element (k, j) of an 8 x elem_size matrix of (size/8)-byte
   * items goes to (j, k) */
  const size_t item = size / 8;
  for (size_t k = 0; k < 8; ++k)
    for (size_t j = 0; j < elem_size; ++j) memcpy(out + (j * 8 + k) * item, tmp + (k * elem_size + j) * item, item);
}

static void orc_bshuf_untrans_bit_elem(const uint8_t* in, uint8_t* out, uint8_t* tmp, size_t size, size_t elem_size) {
  const size_t nbyte = size * elem_size, nrow = nbyte / 8, item = size / 8;
  for (size_t k = 0; k < 8; ++k)
    for (size_t j = 0; j < elem_size; ++j) memcpy(tmp + (k * elem_size + j) * item, in + (j * 8 + k) * item, item);
  uint8_t* mid = (uint8_t*)malloc(nbyte ? nbyte : 1);
  memset(mid, 0, nbyte);
  for (size_t i = 0; i < nrow; ++i)
    for (size_t k = 0; k < 8; ++k) {
      const uint8_t v = tmp[k * nrow + i];
      for (size_t b = 0; b < 8; ++b) mid[8 * i + b] |= (uint8_t)(((v >> b) & 1u) << k);
    }
  for (size_t i = 0; i < size; ++i)
    for (size_t j = 0; j < elem_size; ++j) out[i * elem_size + j] = mid[j * size + i];
  free(mid);
}

/* returns 0, or -81 (the library's error for a block size that is no multiple of 8) */
int orc_bitshuffle(int decode, const uint8_t* in, uint8_t* out, uint64_t size, uint32_t elem_size, uint32_t block_size) {
  if (block_size == 0) block_size = orc_bshuf_default_block_size(elem_size);
  if (block_size % 8u) return -81;
  uint8_t* tmp = (uint8_t*)malloc((size_t)block_size * elem_size + 8);
  uint64_t pos = 0;
  for (uint64_t b = 0; b < size / block_size; ++b, pos += block_size) {
    if (decode) orc_bshuf_untrans_bit_elem(in + pos * elem_size, out + pos * elem_size, tmp, block_size, elem_size);
    else orc_bshuf_trans_bit_elem(in + pos * elem_size, out + pos * elem_size, tmp, block_size, elem_size);
  }
  uint64_t last = size % block_size;
  last -= last % 8u;
  if (last) {
    if (decode) orc_bshuf_untrans_bit_elem(in + pos * elem_size, out + pos * elem_size, tmp, last, elem_size);
    else orc_bshuf_trans_bit_elem(in + pos * elem_size, out + pos * elem_size, tmp, last, elem_size);
    pos += last;
  }
  memcpy(out + pos * elem_size, in + pos * elem_size, (size_t)((size - pos) * elem_size));
  free(tmp);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * diff3x3x1 — encoders/diff_scheme_impl.hpp:78-199 with last_plane_neighborhood<3>
 * (neighborhood_utils.hpp:72-92), halo::compute_offsets_in_x (:186-226), naive_sum (diff_scheme_utils.hpp:75-103).
 * Pinned against oracle/_ref (the compiled reference) in tests/test_oracle.py.
 *
 *   out[i] = in[i] - (sum of the 9 voxels around i in the previous z plane, ACCUMULATED IN THE VOXEL TYPE, i.e. mod 2^16
 *            or 2^8) / 9          for i in the covered set, out[i] = in[i] elsewhere; decode adds the same term, computed
 *            from already decoded voxels, in index order.
 * naive_sum only uses linear offsets: the 9 voxels are i - Y*X + dy*X + dx, dy, dx in {-1,0,1}.
 * Covered set, as the reference's loops produce it (halo is built with world = {Z,Y,X} but asked per axis in x,y,z
 * order, so the x range comes from the Z extent and the z range from the X extent):
 *   rows (z, y) with 1 <= z < min(X, Z), 1 <= y < Y-1; of each row the indices z*Y*X + y*X + 1 + [0, Z-2).
 * For Z > X the runs spill into the following rows. Shapes where a run would leave its plane, where fewer than two
 * rows qualify (the reference then sweeps to the end of the buffer and reads in front of it) or where an extent does
 * not fit the coordinates naive_sum keeps in the SIGNED type of the voxel width (int16 for uint16 stacks, int8 for uint8
 * stacks: from 129 rows on the reference reads from wrapped coordinates, in front of the buffer) are refused (return 1).
 * elem = 2 (uint16) or 1 (uint8).
 * ------------------------------------------------------------------------------------------ */
int orc_diff_supported(uint64_t Z, uint64_t Y, uint64_t X, int elem) {
  if (Z < 3 || Y < 3 || X < 2) return 0;
  const uint64_t lim = elem == 1 ? 128 : 32767;                /* coord_t = int8 / int16 (diff_scheme_utils.hpp:81-89) */
  if (Z > lim || Y > lim || X > lim) return 0;
  if ((X - 1) * (Y - 2) <= 1) return 0;                        /* num_offsets_required <= 1: the degenerate sweep */
  const uint64_t zend = X < Z ? X : Z;
  if ((zend - 1) * (Y - 2) <= 1) return 0;                     /* a single offset pushed: same sweep */
  if ((Y - 2) * X + 1 + (Z - 2) > Y * X) return 0;             /* the last run would leave the plane */
  return 1;
}

int orc_diff(int decode, const void* in_v, void* out_v, uint64_t Z, uint64_t Y, uint64_t X, int elem) {
  if ((elem != 1 && elem != 2) || !orc_diff_supported(Z, Y, X, elem)) return 1;
  const uint64_t frame = Y * X, n = Z * frame, zend = X < Z ? X : Z;
  memcpy(out_v, in_v, n * (uint64_t)elem);
  const uint8_t* in8 = (const uint8_t*)in_v;   uint8_t* out8 = (uint8_t*)out_v;
  const uint16_t* in16 = (const uint16_t*)in_v; uint16_t* out16 = (uint16_t*)out_v;
  for (uint64_t z = 1; z < zend; ++z)
    for (uint64_t y = 1; y + 1 < Y; ++y)
      for (uint64_t k = 0; k + 2 < Z; ++k) {
        const uint64_t i = z * frame + y * X + 1 + k;
        if (elem == 2) {
          const uint16_t* nb = decode ? out16 : in16;          /* decode reads what it has already written (plane z-1) */
          uint16_t sum = 0;
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) sum = (uint16_t)(sum + nb[i - frame + (int64_t)dy * (int64_t)X + dx]);
          const uint32_t q = (uint32_t)sum / 9u;
          out16[i] = decode ? (uint16_t)(in16[i] + q) : (uint16_t)(in16[i] - q);
        } else {
          const uint8_t* nb = decode ? out8 : in8;
          uint8_t sum = 0;
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) sum = (uint8_t)(sum + nb[i - frame + (int64_t)dy * (int64_t)X + dx]);
          const uint32_t q = (uint32_t)sum / 9u;
          out8[i] = decode ? (uint8_t)(in8[i] + q) : (uint8_t)(in8[i] - q);
        }
      }
  return 0;
}
