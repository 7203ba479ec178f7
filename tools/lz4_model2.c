/* CPU model #2 of the GPU LZ4 encoder's candidate policy (development tool, sizes only): candidates come from a list of
 * fixed offsets tried in priority order (the short offsets 1,2,4,3 and, for bit planes / code rows, the row pitch of the
 * stack in the stream), optionally followed by the round-based hash table for positions that found none.
 * build: gcc -O2 -o /tmp/lz4_model2 tools/lz4_model2.c
 * usage: lz4_model2 file block cut hash(0|1) minmatch off1,off2,...  */
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
  const int hashlog = 12; int round = getenv("ROUND") ? atoi(getenv("ROUND")) : 512;
  long out = 0, nseq = 0, nconst = 0, nstored = 0, lookups = 0, by_off[17] = {0}, bytes_off[17] = {0};
  int* cand = malloc(sizeof(int) * (block + 16));
  int* tab = malloc(sizeof(int) << hashlog);
  for (long o = 0; o < total; o += block) {
    int n = o + block <= total ? block : (int)(total - o);
    const uint8_t* d = buf + o;
    int same = 1;
    for (int i = 1; i < n; ++i) if (d[i] != d[0]) { same = 0; break; }
    if (same && n >= 16) { out += 4 + 4 + ext_bytes(n - 10) + 6; nconst++; continue; }
    for (int i = 0; i < (1 << hashlog); ++i) tab[i] = -1;
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
        if (cand[i] < 0 && use_hash) {
          uint32_t v = ld4(d + i), h = (v * 2654435761u) >> (32 - hashlog);
          int c = tab[h];
          lookups++;
          if (c >= 0 && ld4(d + c) == v) cand[i] = c;
        }
      }
      if (use_hash) for (int i = r0; i < r1; ++i) {
        if (i + 4 > n) continue;
        int has_fixed = 0;
        for (int q = 0; q < noff; ++q) if (cand[i] == i - offs[q]) has_fixed = 1;
        if (has_fixed) continue;
        uint32_t v = ld4(d + i), h = (v * 2654435761u) >> (32 - hashlog);
        tab[h] = i;
      }
    }
    long bo = 0; int anchor = 0, pos = 0;
    while (pos < n) {
      int c = cand[pos];
      if (c < 0) { pos++; continue; }
      int limit = n - 5, cut_hi = ((pos / cut) + 1) * cut;
      if (cut_hi < limit) limit = cut_hi;
      int maxlen = limit - pos, len = 0;
      while (len < maxlen && d[pos + len] == d[c + len]) len++;
      if (len < minmatch) { pos++; continue; }
      int lit = pos - anchor;
      bo += 1 + ext_bytes(lit) + lit + 2 + ext_bytes(len - 4);
      nseq++;
      int q; for (q = 0; q < noff; ++q) if (pos - c == offs[q]) break;
      by_off[q]++; bytes_off[q] += len;
      pos += len; anchor = pos;
    }
    int lit = n - anchor;
    bo += 1 + ext_bytes(lit) + lit;
    if (bo >= n) { bo = n; nstored++; }
    out += 4 + bo;
  }
  printf("%s block=%d cut=%d hash=%d min=%d: %ld -> %ld ratio %.3f seqs %ld const %ld stored %ld lookups %.1f%%\n", argv[1], block, cut, use_hash, minmatch,
         total, out, (double)total / out, nseq, nconst, nstored, 100.0 * lookups / total);
  for (int q = 0; q <= noff; ++q) printf("   off %d: %ld matches, %ld bytes\n", q < noff ? offs[q] : -1, by_off[q], bytes_off[q]);
  return 0;
}
