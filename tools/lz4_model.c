/* CPU model of the GPU LZ4 encoder's match finding + greedy parse, for tuning the compression ratio
 * without a GPU (development tool, not shipped, not an oracle). Computes encoded sizes only.
 * build: gcc -O2 -o /tmp/lz4_model tools/lz4_model.c
 * usage: lz4_model file block_bytes round cut hashlog mode [win minlen early_bytes early_div]
 *   mode 72 (= 64 | 8) is the shipped policy: short offsets 1,2,4,3 first, hash rounds only for positions without one, minimum match 5
 *   mode bit 4096 + win/minlen: (rejected) no lookup when a short-offset match of >= minlen bytes starts within `win` positions
 *   early_bytes/early_div: the early-store policy — a block with fewer than early_bytes/early_div candidates (>= 5 bytes) in
 *   its first early_bytes bytes is stored; MODEL_DUMP=1 prints per block "BLK <candidates in the prefix> <short-offset candidates
 *   in the whole block> <encoded size>", which is where kEarlyMin = 128 per 4 KiB (lz4_encode.cu) comes from
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int ext_bytes(int v) { return v < 15 ? 0 : 1 + (v - 15) / 255; }
static uint32_t ld4(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* candidates: cand[i] = position or -1. round-based table (positions of earlier rounds) + short offsets */
static int g_cut = 1 << 30, g_win = 3, g_minlen = 8, g_early = 0, g_div = 16;
static int g_last_cnt = 0, g_last_short = 0;
static long g_early_stored = 0, g_early_lost = 0;
static long g_lookups = 0, g_lookup_segs = 0;
static int seg_has_short(const uint8_t* d, int n, int seg) {
  static const int ds[4] = {1, 2, 4, 3};
  for (int i = seg * 32; i < seg * 32 + 32 && i + 12 <= n; ++i)
    for (int q = 0; q < 4; ++q) {
      int dd = ds[q], l = 0;
      if (i < dd) continue;
      while (l < 5 && i + l < n - 5 && d[i + l] == d[i + l - dd]) l++;
      if (l >= 5) return 1;
    }
  return 0;
}
static void find_candidates(const uint8_t* d, int n, int round, int hashlog, int mode, int* cand) {
  int tsize = 1 << hashlog;
  int* tab = malloc(sizeof(int) * tsize);
  static int noshort[1 << 20];
  for (int i = 0; i < tsize; ++i) tab[i] = -1;
  /* 4096: g_win > 0: a position does not look up (nor insert) when a short-offset candidate of >= g_minlen bytes starts
   * within the next g_win positions: the hash match would save at most g_win literals and cost a 3-byte sequence */
  static unsigned char slen[(1 << 20) + 16];
  if (mode & 4096) {
    static const int ds[4] = {1, 2, 4, 3};
    for (int i = 0; i < n + 8; ++i) slen[i] = 0;
    for (int i = 0; i + 12 <= n; ++i)
      for (int q = 0; q < 4; ++q) {
        int dd = ds[q], l = 0;
        if (i < dd) continue;
        int cut_hi = ((i / g_cut) + 1) * g_cut, lim = n - 5 < cut_hi ? n - 5 : cut_hi;
        while (l < 64 && i + l < lim && d[i + l] == d[i + l - dd]) l++;
        if (l >= 5) { slen[i] = l; break; }
      }
  }
  for (int r0 = 0; r0 < n; r0 += round) {
    int r1 = r0 + round < n ? r0 + round : n;
    for (int i = r0; i < r1; ++i) {
      cand[i] = -1;
      if (i + 12 > n) continue;
      uint32_t v = ld4(d + i);
      uint32_t h = (v * 2654435761u) >> (32 - hashlog);
      int c = tab[h];
      int found = -1;
      if (mode & 1) { /* prefer short offsets first */
        if (i >= 1 && ld4(d + i - 1) == v) found = i - 1;
        else if (i >= 2 && ld4(d + i - 2) == v) found = i - 2;
        else if (i >= 4 && ld4(d + i - 4) == v) found = i - 4;
        else if (i >= 3 && ld4(d + i - 3) == v) found = i - 3;
        else if (c >= 0 && ld4(d + c) == v && !((mode & 32) && (i & 1))) found = c;
      } else {
        if (c >= 0 && ld4(d + c) == v) found = c;
        else if (i >= 1 && ld4(d + i - 1) == v) found = i - 1;
        else if (i >= 2 && ld4(d + i - 2) == v) found = i - 2;
        else if (i >= 4 && ld4(d + i - 4) == v) found = i - 4;
        else if (i >= 3 && ld4(d + i - 3) == v) found = i - 3;
      }
      if ((mode & 2) && found >= 0 && c >= 0 && c != found && ld4(d + c) == v) {
        /* both available: keep the one with the longer 8-byte agreement */
        int a = 4, b = 4;
        while (a < 12 && i + a < n - 5 && d[i + a] == d[found + a]) a++;
        while (b < 12 && i + b < n - 5 && d[i + b] == d[c + b]) b++;
        if (b > a) found = c;
      }
      if (mode & 64) {
        static const int ds[4] = {1, 2, 4, 3};
        found = -1;
        for (int q = 0; q < 4 && found < 0; ++q) {
          int dd = ds[q];
          if (i < dd) continue;
          int l = 0;
          while (l < 5 && i + l < n - 5 && d[i + l] == d[i + l - dd]) l++;
          if (l >= 5) found = i - dd;
        }
        int sup = 0;
        if (mode & 4096) for (int k = 1; k <= g_win; ++k) if (slen[i + k] >= g_minlen) sup = 1;
        noshort[i] = found < 0 && !sup && !((mode & 512) && i > 0 && d[i] == d[i - 1]) && !((mode & 1024) && (i & 3) && !seg_has_short(d, n, i / 32)) && !((mode & 2048) && ((i / 32) & 3) && !seg_has_short(d, n, i / 32)); /* 2048: run-free segments look up only in every 4th segment */ /* 512: look up / insert only where a run breaks */
        if (found < 0 && noshort[i] && c >= 0 && ld4(d + c) == v && !(mode & 128) && !((mode & 256) && c / g_cut != i / g_cut)) found = c; /* 256: same sub-block only */ /* 128: short offsets only */
      }
      cand[i] = found;
      if (mode & 64) { g_lookups += noshort[i]; }
    }
    if (mode & 64) for (int sg = r0 / 32; sg < (r1 + 31) / 32; ++sg) { int any = 0; for (int i = sg * 32; i < sg * 32 + 32 && i < r1; ++i) any |= noshort[i]; g_lookup_segs += any; }
    for (int i = r0; i < r1; ++i) {
      if (i + 4 > n) continue;
      if ((mode & 64) && !noshort[i]) continue;
      uint32_t v = ld4(d + i);
      uint32_t h = (v * 2654435761u) >> (32 - hashlog);
      if (!((mode & 32) && (i & 1))) tab[h] = i; /* last writer of the round wins (GPU: racy, any) */
    }
  }
  free(tab);
}

static long encode_block_size(const uint8_t* d, int n, int round, int cut, int hashlog, int mode, long* nseq_out) {
  int* cand = malloc(sizeof(int) * (n + 16));
  g_cut = cut;
  find_candidates(d, n, round, hashlog, mode, cand);
  long out = 0;
  int anchor = 0, pos = 0;
  int early = 0;
  if (g_early && n > g_early) { /* early-store policy: few candidates in the first g_early bytes => the block is stored unparsed */
    int cnt = 0;
    for (int i = 0; i < g_early; ++i) { /* the GPU's candidates agree in at least 5 bytes */
      int c = cand[i];
      if (c >= 0 && i + 5 <= n - 5 && d[i + 4] == d[c + 4] && (i + 5 <= ((i / cut) + 1) * cut)) cnt++;
    }
    g_last_short = 0;
    for (int i = 0; i < n; ++i) {
      int c = cand[i];
      if (c >= 0 && i - c <= 4 && i + 5 <= n - 5 && d[i + 4] == d[c + 4] && (i + 5 <= ((i / cut) + 1) * cut)) g_last_short++;
    }
    g_last_cnt = cnt;
    early = cnt < g_early / g_div;
  }
  long nseq = 0;
  while (pos < n) {
    int c = cand[pos];
    if (c < 0) { pos++; continue; }
    int limit = n - 5;
    int cut_hi = ((pos / cut) + 1) * cut;
    if (cut_hi < limit) limit = cut_hi;
    int maxlen = limit - pos;
    if (maxlen < 4) { pos++; continue; }
    int len = 4;
    while (len < maxlen && d[pos + len] == d[c + len]) len++;
    if ((mode & 4) && len < 5 && pos > anchor) { pos++; continue; } /* skip len-4 matches that need a new token */
    if ((mode & 8) && len < 5) { pos++; continue; }                 /* minimum match length 5 */
    if ((mode & 16) && len < 6) { pos++; continue; }                /* minimum match length 6 */
    int lit = pos - anchor;
    out += 1 + ext_bytes(lit) + lit + 2 + ext_bytes(len - 4);
    nseq++;
    pos += len;
    anchor = pos;
  }
  int lit = n - anchor;
  out += 1 + ext_bytes(lit) + lit;
  free(cand);
  if (getenv("MODEL_DUMP")) fprintf(stdout, "BLK %d %d %ld\n", g_last_cnt, g_last_short, out >= n ? (long)n : out);
  if (early) { g_early_stored++; if (out < n) g_early_lost += n - out; return n; }
  if (nseq_out) *nseq_out += nseq;
  return out >= n ? n : out;
}

int main(int argc, char** argv) {
  if (argc < 7) return 1;
  FILE* f = fopen(argv[1], "rb");
  fseek(f, 0, SEEK_END);
  long total = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t* buf = malloc(total + 64);
  if (fread(buf, 1, total, f) != (size_t)total) return 2;
  memset(buf + total, 0, 64);
  int block = atoi(argv[2]), round = atoi(argv[3]), cut = atoi(argv[4]), hashlog = atoi(argv[5]), mode = atoi(argv[6]);
  if (argc > 7) g_win = atoi(argv[7]);
  if (argc > 8) g_minlen = atoi(argv[8]);
  if (argc > 9) g_early = atoi(argv[9]);
  if (argc > 10) g_div = atoi(argv[10]);
  long out = 0, nseq = 0;
  for (long o = 0; o < total; o += block) {
    int n = o + block <= total ? block : (int)(total - o);
    int same = 1;
    for (int i = 1; i < n; ++i) if (buf[o + i] != buf[o]) { same = 0; break; }
    if (same && n >= 16) { out += 4 + 4 + ext_bytes(n - 10) + 6; continue; }
    out += 4 + encode_block_size(buf + o, n, round, cut, hashlog, mode, &nseq);
  }
  printf("block=%d round=%d cut=%d hashlog=%d mode=%d: %ld -> %ld ratio %.3f seqs %ld (%.1f B/seq)\n", block, round, cut, hashlog, mode,
         total, out, (double)total / out, nseq, nseq ? (double)total / nseq : 0.0);
  if (g_early) fprintf(stderr, "early-stored blocks %ld, bytes lost %ld\n", g_early_stored, g_early_lost);
  fprintf(stderr, "lookups %ld (%.1f%% of bytes), segments with lookups %ld (%.1f%% of segments)\n", g_lookups, 100.0 * g_lookups / total, g_lookup_segs, 100.0 * g_lookup_segs / (total / 32.0));
  return 0;
}
