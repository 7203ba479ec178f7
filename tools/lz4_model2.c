/* CPU model #2 of the GPU LZ4 encoder's candidate policy (development tool, sizes only): candidates come from a list of
 * fixed offsets tried in priority order (the short offsets 1,2,4,3 and, for bit planes / code rows, the row pitch of the
 * stack in the stream), optionally followed by the round-based hash table for positions that found none.
 * build: gcc -O2 -o /tmp/lz4_model2 tools/lz4_model2.c
 * usage: lz4_model2 file block cut hash(0|1) minmatch off1,off2,...
 * environment: ROUND=n (waves of n positions, last occurrence in earlier waves: the round-1 policy), LISTMAX=n (cap of the
 *   lookup list), HLOG=n, FIRST=1 (one block-wide table of first occurrences), FIRST=2 RH=bits RLOG=log2(region bytes)
 *   PROBES=n [PROBES_N=n NTHR=n] (per-region first-occurrence tables probed nearest first; measured on B200 with
 *   RH=8 RLOG=10 PROBES=16 for blocks that list <= 4096 positions, RH=10 RLOG=12 PROBES=4 for the others: 4.9x fewer far
 *   matches than FIRST=1, decode of cfg2 6.27 -> 5.92 ms, but the probe loops cost the encoder 7.44 -> 8.07 ms: rejected,
 *   FIRST=1 is the shipped policy),
 *   NOINTERIOR=1 NIQ=n KEEPTAIL=k (positions inside a run of the first n offsets do not look up unless fewer than k bytes
 *   of the run lie ahead: NIQ=1 KEEPTAIL=3 is shipped — the position with exactly four run bytes ahead — 12 % fewer lookups
 *   on background-removed planes at the same size; KEEPTAIL=2 / 1 / 0: -26 / -39 / -51 % lookups for +0.5 / 1.3 / 2.2 % size),
 *   FARMIN=n NEAR=n (minimum length of matches further than NEAR bytes away); prints the number of far matches
 *   (distance > 1472 = what the decoder's output ring does not hold) in front of the result line */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static int ext_bytes(int v) { return v < 15 ? 0 : 1 + (v - 15) / 255; }
static uint32_t ld4(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
int main(int argc, char** argv) {
  if (argc < 7) return 1;
  FILE* f = fopen(argv[1], "rb");
  fseek(f, 0, SEEK_END); long total = ftell(f); fseek(f, 0, SEEK_SET);
  uint8_t* buf = malloc(total + 64);
  if (fread(buf, 1, total, f) != (size_t)total) return 2;
  int block = atoi(argv[2]), cut = atoi(argv[3]), use_hash = atoi(argv[4]), minmatch = atoi(argv[5]);
  int offs[16], noff = 0;
  for (char* t = strtok(argv[6], ","); t && noff < 16; t = strtok(NULL, ",")) offs[noff++] = atoi(t);
  int hashlog = getenv("HLOG") ? atoi(getenv("HLOG")) : 12; int listmax = getenv("LISTMAX") ? atoi(getenv("LISTMAX")) : 1<<30; int round = getenv("ROUND") ? atoi(getenv("ROUND")) : 512;
  long nskip = 0; long nfar = 0, early_lost = 0, nearly = 0; long out = 0, nseq = 0, nconst = 0, nstored = 0, lookups = 0, by_off[17] = {0}, bytes_off[17] = {0};
  int* cand = malloc(sizeof(int) * (block + 16));
  int* tab = malloc(sizeof(int) << hashlog);
  for (long o = 0; o < total; o += block) {
    int n = o + block <= total ? block : (int)(total - o);
    const uint8_t* d = buf + o;
    int same = 1;
    for (int i = 1; i < n; ++i) if (d[i] != d[0]) { same = 0; break; }
    if (same && n >= 16) { out += 4 + 4 + ext_bytes(n - 10) + 6; nconst++; continue; }
    for (int i = 0; i < (1 << hashlog); ++i) tab[i] = -1;
    int listed = 0; int sub = 0;
    if (getenv("SUBSAMPLE")) { int wants = 0; for (int i = 0; i + 12 <= n; ++i) { int has = 0; int cut_hi = ((i / cut) + 1) * cut, lim = n - 5 < cut_hi ? n - 5 : cut_hi; for (int q = 0; q < noff && !has; ++q) { int dd = offs[q], l = 0; if (i < dd) continue; while (l < minmatch && i + l < lim && d[i + l] == d[i + l - dd]) l++; if (l >= minmatch) has = 1; } wants += !has; } if (wants > listmax) sub = 1; }
    int first_mode = getenv("FIRST") ? atoi(getenv("FIRST")) : 0;
    int rh = getenv("RH") ? atoi(getenv("RH")) : 8, probes = getenv("PROBES") ? atoi(getenv("PROBES")) : 2, rlog = getenv("RLOG") ? atoi(getenv("RLOG")) : 10;
    static int rtab[64][1 << 12];
    if (first_mode == 2 && use_hash) {
      for (int r = 0; r < 64; ++r) for (int k = 0; k < (1 << rh); ++k) rtab[r][k] = -1;
      for (int i = 0; i + 12 <= n; ++i) {
        int has = 0; int cut_hi = ((i / cut) + 1) * cut, lim = n - 5 < cut_hi ? n - 5 : cut_hi;
        for (int q = 0; q < noff && !has; ++q) { int dd = offs[q], l = 0; if (i < dd) continue; while (l < minmatch && i + l < lim && d[i + l] == d[i + l - dd]) l++; if (l >= minmatch) has = 1; }
        if (has) continue;
        uint32_t v = ld4(d + i), h = (v * 2654435761u) >> (32 - rh);
        if (rtab[i >> rlog][h] < 0) rtab[i >> rlog][h] = i;
      }
    }
    if (first_mode == 2) {
      first_mode = 3;
      int wants = 0;
      for (int i = 0; i + 12 <= n; ++i) { int has = 0; int cut_hi = ((i / cut) + 1) * cut, lim = n - 5 < cut_hi ? n - 5 : cut_hi;
        for (int q = 0; q < noff && !has; ++q) { int dd = offs[q], l = 0; if (i < dd) continue; while (l < minmatch && i + l < lim && d[i + l] == d[i + l - dd]) l++; if (l >= minmatch) has = 1; }
        wants += !has; }
      probes = getenv("PROBES") ? atoi(getenv("PROBES")) : 2;
      if (getenv("PROBES_N") && wants > (getenv("NTHR") ? atoi(getenv("NTHR")) : 4096)) probes = atoi(getenv("PROBES_N"));
    }
    if (first_mode == 1 && use_hash) {
      for (int i = 0; i + 12 <= n; ++i) {
        int has = 0; int cut_hi = ((i / cut) + 1) * cut, lim = n - 5 < cut_hi ? n - 5 : cut_hi;
        for (int q = 0; q < noff && !has; ++q) { int dd = offs[q], l = 0; if (i < dd) continue; while (l < minmatch && i + l < lim && d[i + l] == d[i + l - dd]) l++; if (l >= minmatch) has = 1; }
        if (has) continue;
        uint32_t v = ld4(d + i), h = (v * 2654435761u) >> (32 - hashlog);
        if (tab[h] < 0) tab[h] = i;
      }
    }
    for (int r0 = 0; r0 < n; r0 += round) {
      int r1 = r0 + round < n ? r0 + round : n;
      for (int i = r0; i < r1; ++i) {
        cand[i] = -1;
        if (i + 12 > n) continue;
        int cut_hi = ((i / cut) + 1) * cut, lim = n - 5 < cut_hi ? n - 5 : cut_hi;
        for (int q = 0; q < noff && cand[i] < 0; ++q) {
          int dd = offs[q], l = 0;
          if (i < dd) continue;
          while (l < minmatch && i + l < lim && d[i + l] == d[i + l - dd]) l++;
          if (l >= minmatch) cand[i] = i - dd;
        }
        if (cand[i] < 0 && use_hash && getenv("NOINTERIOR")) {
          /* interior of a fixed-offset run of >= 5: some q with d[i]==d[i-dd] and the run containing i has length >= 5 */
          int interior = 0;
          for (int q = 0; q < (getenv("NIQ") ? atoi(getenv("NIQ")) : noff) && !interior; ++q) {
            int dd = offs[q]; if (i < dd || d[i] != d[i - dd]) continue;
            int a = i; int sub_lo = (i / cut) * cut; while (a - 1 >= dd && a - 1 >= sub_lo && d[a - 1] == d[a - 1 - dd]) a--;
            int b2 = i; int cut_hi2 = ((i / cut) + 1) * cut, lim2 = n - 5 < cut_hi2 ? n - 5 : cut_hi2; while (b2 + 1 < lim2 && d[b2 + 1] == d[b2 + 1 - dd]) b2++;
            if (b2 - a + 1 >= 5 && a < i) { int keep = getenv("KEEPTAIL") ? atoi(getenv("KEEPTAIL")) : 0; /* b2 = last position of the run */ if (b2 - i >= keep) interior = 1; }
          }
          if (interior) { nskip++; continue; }
        }
        if (cand[i] < 0 && use_hash && (sub ? !(i & 1) : listed++ < listmax)) {
          uint32_t v = ld4(d + i), h = (v * 2654435761u) >> (32 - hashlog);
          int c = tab[h];
          if (first_mode == 3) {
            uint32_t h2 = (v * 2654435761u) >> (32 - rh);
            c = -1;
            for (int pr = 0; pr < probes && (i >> rlog) - pr >= 0; ++pr) {
              int cc = rtab[(i >> rlog) - pr][h2];
              if (cc >= 0 && cc < i && ld4(d + cc) == v) { c = cc; break; }
            }
          }
          lookups++;
          if (c >= 0 && c < i && ld4(d + c) == v) cand[i] = c;
        }
      }
      if (use_hash && !first_mode) for (int i = r0; i < r1; ++i) {
        if (i + 4 > n) continue;
        int has_fixed = 0;
        for (int q = 0; q < noff; ++q) if (cand[i] == i - offs[q]) has_fixed = 1;
        if (has_fixed) continue;
        if (sub && (i & 1)) continue;
        uint32_t v = ld4(d + i), h = (v * 2654435761u) >> (32 - hashlog);
        tab[h] = i;
      }
    }
    long bo = 0; int anchor = 0, pos = 0;
    int early_p = getenv("EARLY") ? atoi(getenv("EARLY")) : 0, early_min = getenv("EARLYMIN") ? atoi(getenv("EARLYMIN")) : 128, early_store = 0;
    if (early_p && n > early_p) {
      int cnt = 0;
      for (int i = 0; i + 12 <= n; ++i) {
        int c = cand[i]; if (c < 0) continue;
        int isfixed = 0; for (int q = 0; q < noff; ++q) if (i - c == offs[q]) isfixed = 1;
        if (isfixed) { cnt++; continue; }
        if (i >= early_p) continue;
        int cut_hi = ((i / cut) + 1) * cut, lim = n - 5 < cut_hi ? n - 5 : cut_hi, l = 0;
        while (l < 5 && i + l < lim && d[i + l] == d[c + l]) l++;
        if (l >= 5) cnt++;
      }
      early_store = cnt < early_min;
    }
    while (pos < n) {
      int c = cand[pos];
      if (c < 0) { pos++; continue; }
      int limit = n - 5, cut_hi = ((pos / cut) + 1) * cut;
      if (cut_hi < limit) limit = cut_hi;
      int maxlen = limit - pos, len = 0;
      while (len < maxlen && d[pos + len] == d[c + len]) len++;
      if (len < minmatch) { pos++; continue; }
      { static int farmin = -1, nearlim = 1472; if (farmin < 0) { farmin = getenv("FARMIN") ? atoi(getenv("FARMIN")) : 0; if (getenv("NEAR")) nearlim = atoi(getenv("NEAR")); }
        if (pos - c > nearlim) { nfar++; if (len < farmin) { pos++; nfar--; continue; } } }
      int lit = pos - anchor;
      bo += 1 + ext_bytes(lit) + lit + 2 + ext_bytes(len - 4);
      nseq++;
      int q; for (q = 0; q < noff; ++q) if (pos - c == offs[q]) break;
      by_off[q]++; bytes_off[q] += len;
      pos += len; anchor = pos;
    }
    int lit = n - anchor;
    bo += 1 + ext_bytes(lit) + lit;
    if (early_store) { if (bo < n) early_lost += n - bo; nearly++; bo = n; }
    if (bo >= n) { bo = n; nstored++; }
    out += 4 + bo;
  }
  printf("interior skipped %.1f%%  ", 100.0 * nskip / total);
  printf("far matches %ld  early-stored %ld (lost %ld B)  ", nfar, nearly, early_lost);
  printf("%s block=%d cut=%d hash=%d min=%d: %ld -> %ld ratio %.3f seqs %ld const %ld stored %ld lookups %.1f%%\n", argv[1], block, cut, use_hash, minmatch,
         total, out, (double)total / out, nseq, nconst, nstored, 100.0 * lookups / total);
  for (int q = 0; q <= noff; ++q) printf("   off %d: %ld matches, %ld bytes\n", q < noff ? offs[q] : -1, by_off[q], bytes_off[q]);
  return 0;
}
