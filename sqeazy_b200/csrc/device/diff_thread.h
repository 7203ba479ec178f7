// Thread program and launch schedule of the diff3x3x1 kernels (diff.cu), compilable for the device and for the host:
// tests/helpers/diff_sim.cpp replays them thread by thread against the oracle where no GPU exists (any thread order,
// poisoned output).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SQYB_HD __device__ __forceinline__
#else
#define SQYB_HD inline
#endif
// `in`, `out` and `nb` of one launch never touch the same voxel through two names: encode reads `in` (= `nb`) and writes
// `out`; decode reads the stored plane through `in`, the decoded plane before it through `nb` and writes the plane itself
// through `out` (`nb` and `out` are the same buffer, one plane apart). Saying so lets the loads of a strip leave together
// instead of waiting behind the stores of the row before.
#define SQYB_RESTRICT __restrict__

namespace sqyb {

// rows of a strip: a thread walks 8 voxels x R rows and keeps the row sums it can reuse (R + 2 row loads for R output rows).
// Encode is one big launch: tall strips. A decode launch is one plane: shorter strips keep enough threads in it.
constexpr int kDiffRowsEncode = 8, kDiffRowsDecode = 4;
constexpr int diff_rows(bool decode) { return decode ? kDiffRowsDecode : kDiffRowsEncode; }

template <typename T> struct alignas(sizeof(T) * 8) Pack8 { T v[8]; };

struct DiffGeom {
  uint32_t Z, Y, X, zend;   // zend = min(X, Z): the planes the reference codes are 1 .. zend-1
  uint64_t frame;           // Y * X
  uint32_t ppr, strips;     // 8-voxel packs per row, strips of diff_rows() rows per plane
};

inline DiffGeom diff_geom(uint64_t Z, uint64_t Y, uint64_t X, bool decode) {
  DiffGeom g;
  g.Z = (uint32_t)Z; g.Y = (uint32_t)Y; g.X = (uint32_t)X; g.zend = (uint32_t)(X < Z ? X : Z);
  g.frame = Y * X;
  g.ppr = (uint32_t)((X + 7) / 8);
  g.strips = (uint32_t)((Y + diff_rows(decode) - 1) / diff_rows(decode));
  return g;
}

// shapes on which the reference's loops stay inside their plane (see diff.cu); elem = bytes per voxel
inline bool diff_shape_ok(uint64_t Z, uint64_t Y, uint64_t X, int elem) {
  if (Z < 3 || Y < 3 || X < 2) return false;
  // naive_sum keeps z, y, x in the signed type of the voxel width (diff_scheme_utils.hpp:81-89): int16 for uint16 stacks,
  // int8 for uint8 stacks - beyond that the reference reads from wrapped coordinates (in front of the buffer)
  const uint64_t lim = elem == 1 ? 128 : 32767;
  if (Z > lim || Y > lim || X > lim) return false;
  if ((X - 1) * (Y - 2) <= 1) return false;                    // compute_offsets_in_x: a single offset = sweep to the end
  const uint64_t zend = X < Z ? X : Z;
  if ((zend - 1) * (Y - 2) <= 1) return false;
  if ((Y - 2) * X + 1 + (Z - 2) > Y * X) return false;         // the last run of a plane would leave it
  return true;
}

// the launches of one encode (all planes at once) or decode (plane 0 as stored, one launch per coded plane in z order,
// the planes the reference skips): f(first plane, planes), stops at the first non-zero return
template <typename F>
inline int diff_for_each_launch(bool decode, const DiffGeom& g, F f) {
  if (!decode) return f(0u, g.Z);
  int rc = f(0u, 1u);
  for (uint32_t z = 1; z < g.zend && !rc; ++z) rc = f(z, 1u);
  if (!rc && g.zend < g.Z) rc = f(g.zend, g.Z - g.zend);
  return rc;
}
inline uint64_t diff_threads_per_plane(const DiffGeom& g) { return (uint64_t)g.strips * g.ppr; }

// is voxel (y, x) of a coded plane one the reference visits: in the run of its own row, or in the end of the run of the
// row above that spilled over (Z > X + 1)
SQYB_HD bool diff_covered(const DiffGeom& g, uint32_t y, uint32_t x) {
  const bool own = y >= 1 && y + 1 < g.Y && x >= 1 && x + 1 < g.Z;
  const bool spilled = y >= 2 && x + g.X + 1 < g.Z;
  return own || spilled;
}

// one voxel the slow way: linear offsets exactly as naive_sum forms them. In decode a neighbour inside the plane being
// written (row 0, reached by the spilled end of the last run) is never coded and is taken from `in`.
template <typename T, bool DECODE>
SQYB_HD T diff_voxel(const T* SQYB_RESTRICT in, const T* SQYB_RESTRICT nb, const DiffGeom& g, uint64_t plane, bool z_coded, uint32_t y, uint32_t x) {
  const uint64_t i = plane + (uint64_t)y * g.X + x;
  uint32_t q = 0;
  if (z_coded && diff_covered(g, y, x)) {
    T sum = 0;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const uint64_t k = (uint64_t)((int64_t)(i - g.frame) + (int64_t)dy * (int64_t)g.X + dx);
        sum = (T)(sum + ((DECODE && k >= plane) ? in[k] : nb[k]));
      }
    q = (uint32_t)sum / 9u;
  }
  return DECODE ? (T)(in[i] + q) : (T)(in[i] - q);
}

// ---- packed lanes: 8 voxels = 4 words of 2 x uint16 or 2 words of 4 x uint8; every lane wraps on its own
template <typename T> struct alignas(sizeof(T) * 8) DiffWords { uint32_t w[2 * sizeof(T)]; };

template <typename T> struct DiffLanes;
template <> struct DiffLanes<uint16_t> {
  static constexpr int kBits = 16, kPerWord = 2, kWords = 4;
  static constexpr uint32_t kLaneMask = 0xffffu;
  static SQYB_HD uint32_t add(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __vadd2(a, b);                                    // VIADD.16x2
#else
    return ((a + b) & 0xffffu) | ((((a >> 16) + (b >> 16)) & 0xffffu) << 16);
#endif
  }
  static SQYB_HD uint32_t sub(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __vsub2(a, b);
#else
    return ((a - b) & 0xffffu) | ((((a >> 16) - (b >> 16)) & 0xffffu) << 16);
#endif
  }
  static SQYB_HD uint32_t div9(uint32_t a) { return ((a & 0xffffu) / 9u) | (((a >> 16) / 9u) << 16); }
};
template <> struct DiffLanes<uint8_t> {
  static constexpr int kBits = 8, kPerWord = 4, kWords = 2;
  static constexpr uint32_t kLaneMask = 0xffu;
  static SQYB_HD uint32_t add(uint32_t a, uint32_t b) {      // carries stopped at the lane tops
    return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u);
  }
  static SQYB_HD uint32_t sub(uint32_t a, uint32_t b) {      // borrows stopped at the lane tops
    return ((a | 0x80808080u) - (b & 0x7f7f7f7fu)) ^ ((a ^ ~b) & 0x80808080u);
  }
  static SQYB_HD uint32_t div9(uint32_t a) {                 // x * 57 >> 9 == x / 9 for x < 256
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) r |= ((((a >> (8 * k)) & 0xffu) * 57u) >> 9) << (8 * k);
    return r;
  }
};

// 3-tap sums (per lane, in the voxel type) of 8 voxels of the previous plane starting at `p` (16-byte aligned). The voxel
// in front of p[0] and the one behind p[7] come by linear offsets like naive_sum's: at a row end they are the neighbouring
// row's. `p == nb` only for z = 1, first row, x0 = 0: the voxel in front does not exist and feeds no coded voxel (p[0] is
// read instead). `behind` is where the voxel after p[7] is read from: p + 8, except for the end of the last row in decode -
// that voxel is row 0 of the plane being written, never coded, so `in` has it.
template <typename T>
SQYB_HD void diff_row_sums(const T* SQYB_RESTRICT p, const T* SQYB_RESTRICT nb, const T* SQYB_RESTRICT behind, uint32_t* h) {
  using L = DiffLanes<T>;
  constexpr int W = L::kWords;
  const DiffWords<T> c = *reinterpret_cast<const DiffWords<T>*>(p);
  const uint32_t left = (uint32_t)p[p != nb ? -1 : 0];
  const uint32_t right = (uint32_t)*behind;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const uint32_t before = w > 0 ? c.w[w - 1] : left << (32 - L::kBits);     // its top lane = the voxel in front of this word
    const uint32_t after = w + 1 < W ? c.w[w + 1] : right;                      // its bottom lane = the voxel behind it
    const uint32_t l = (c.w[w] << L::kBits) | (before >> (32 - L::kBits));
    const uint32_t rr = (c.w[w] >> L::kBits) | (after << (32 - L::kBits));
    h[w] = L::add(c.w[w], L::add(l, rr));
  }
}

// Thread t of plane z: 8 voxels x diff_rows(DECODE) rows. `nb` holds the neighbours: `in` itself for encode, `out` for decode
// (one plane per launch: every neighbour lies in the plane before).
template <typename T, bool DECODE>
SQYB_HD void diff_thread(const T* SQYB_RESTRICT in, T* SQYB_RESTRICT out, const T* SQYB_RESTRICT nb, const DiffGeom& g, uint32_t z, uint32_t t) {
  constexpr int kDiffRows = DECODE ? kDiffRowsDecode : kDiffRowsEncode;
  const uint32_t s = t / g.ppr, p = t - s * g.ppr;
  if (s >= g.strips) return;
  const uint32_t x0 = 8 * p, y0 = s * kDiffRows;
  const uint32_t rows = g.Y - y0 < (uint32_t)kDiffRows ? g.Y - y0 : (uint32_t)kDiffRows;
  const uint32_t cnt = g.X - x0 < 8u ? g.X - x0 : 8u;
  const uint64_t plane = (uint64_t)z * g.frame;
  const bool z_coded = z >= 1 && z < g.zend;
  constexpr uintptr_t kMask = sizeof(T) * 8 - 1;
  const bool aligned = (g.X & 7u) == 0 && (((uintptr_t)in | (uintptr_t)out | (uintptr_t)nb) & kMask) == 0;
  const bool has_spill = g.Z > g.X + 1;

  if (!aligned) {   // any row length, any buffer alignment: voxel by voxel
    for (uint32_t k = 0; k < rows; ++k)
      for (uint32_t j = 0; j < cnt; ++j)
        out[plane + (uint64_t)(y0 + k) * g.X + x0 + j] = diff_voxel<T, DECODE>(in, nb, g, plane, z_coded, y0 + k, x0 + j);
    return;
  }
  if (!z_coded) {
#pragma unroll
    for (int k = 0; k < kDiffRows; ++k)
      if ((uint32_t)k < rows) {
        const uint64_t i = plane + (uint64_t)(y0 + k) * g.X + x0;
        *reinterpret_cast<Pack8<T>*>(out + i) = *reinterpret_cast<const Pack8<T>*>(in + i);
      }
    return;
  }

  // The strip in packed lanes (two uint16 or four uint8 voxels per 32-bit word, voxel j of a pack in lane j): the kernel
  // is bound by its instruction count, not by HBM, so every add, the /9 and the select work on whole words, addresses
  // advance by a row, and the lanes a row covers are two masks ANDed with a per-row all-or-nothing word.
  using L = DiffLanes<T>;
  constexpr int W = L::kWords;
  uint32_t own_mask[W], spill_mask[W];                         // lanes of this pack a coded row covers (diff_covered)
  if (x0 >= 1 && x0 + 8 < g.Z) {                                // inside the run of the row: all of them
#pragma unroll
    for (int w = 0; w < W; ++w) own_mask[w] = 0xffffffffu;
  } else {
#pragma unroll
    for (int w = 0; w < W; ++w) own_mask[w] = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t x = x0 + (uint32_t)j;
      if (x >= 1 && x + 1 < g.Z) own_mask[j / L::kPerWord] |= L::kLaneMask << (L::kBits * (j % L::kPerWord));
    }
  }
#pragma unroll
  for (int w = 0; w < W; ++w) spill_mask[w] = 0;
  if (has_spill) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (x0 + (uint32_t)j + g.X + 1 < g.Z) spill_mask[j / L::kPerWord] |= L::kLaneMask << (L::kBits * (j % L::kPerWord));
  }
  const bool last_row_spills = has_spill && y0 + rows == g.Y;   // that row looks into this plane's row 0: voxel by voxel below

  if (DECODE) {
    // A decode launch is one plane = one wave of CTAs: nothing hides a thread's memory round trips, so all loads of the strip
    // leave together (R + 2 packs of the plane before with the voxels at their ends, R packs of the stored plane) and the
    // arithmetic starts when they are back - one round trip per strip instead of one per row.
    const uint64_t first = plane + (uint64_t)y0 * g.X + x0;
    const T* pi = in + first;
    T* po = out + first;
    const T* pn = nb + (first - g.frame) - g.X;                 // row y0 - 1 of the plane before
    const bool row_end = x0 + 8 == g.X;
    DiffWords<T> c[kDiffRows + 2], v[kDiffRows];
    uint32_t lf[kDiffRows + 2], rt[kDiffRows + 2];
#pragma unroll
    for (int r = 0; r < kDiffRows + 2; ++r) {
      const uint32_t yy = y0 + (uint32_t)r;                     // this is row yy - 1
      if (yy >= 1 && yy <= g.Y) {
        const T* p = pn + (uint64_t)r * g.X;
        c[r] = *reinterpret_cast<const DiffWords<T>*>(p);
        lf[r] = (uint32_t)p[p != nb ? -1 : 0];
        rt[r] = (uint32_t)((row_end && yy == g.Y) ? in[plane] : p[8]);   // behind the last row: row 0 of this plane, from `in`
      } else {
#pragma unroll
        for (int w = 0; w < W; ++w) c[r].w[w] = 0;
        lf[r] = rt[r] = 0;
      }
    }
#pragma unroll
    for (int k = 0; k < kDiffRows; ++k) {
      if ((uint32_t)k < rows) v[k] = *reinterpret_cast<const DiffWords<T>*>(pi + (uint64_t)k * g.X);
      else {
#pragma unroll
        for (int w = 0; w < W; ++w) v[k].w[w] = 0;
      }
    }
#ifdef __CUDA_ARCH__
    asm volatile("" ::: "memory");                              // the loads above stay above the stores below
#endif
    uint32_t h[kDiffRows + 2][W];
#pragma unroll
    for (int r = 0; r < kDiffRows + 2; ++r) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const uint32_t before = w > 0 ? c[r].w[w - 1] : lf[r] << (32 - L::kBits);
        const uint32_t after = w + 1 < W ? c[r].w[w + 1] : rt[r];
        const uint32_t l = (c[r].w[w] << L::kBits) | (before >> (32 - L::kBits));
        const uint32_t rr = (c[r].w[w] >> L::kBits) | (after << (32 - L::kBits));
        h[r][w] = L::add(c[r].w[w], L::add(l, rr));
      }
    }
#pragma unroll
    for (int k = 0; k < kDiffRows; ++k) {
      const uint32_t y = y0 + (uint32_t)k;
      if ((uint32_t)k >= rows || (last_row_spills && y + 1 == g.Y)) continue;
      const uint32_t own_row = (y >= 1 && y + 1 < g.Y) ? 0xffffffffu : 0u;
      const uint32_t spill_row = y >= 2 ? 0xffffffffu : 0u;
      DiffWords<T> o;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const uint32_t mask = (own_mask[w] & own_row) | (spill_mask[w] & spill_row);
        const uint32_t q = L::div9(L::add(h[k][w], L::add(h[k + 1][w], h[k + 2][w]))) & mask;
        o.w[w] = L::add(v[k].w[w], q);
      }
      *reinterpret_cast<DiffWords<T>*>(po + (uint64_t)k * g.X) = o;
    }
  } else {
    const uint64_t first = plane + (uint64_t)y0 * g.X + x0;       // the strip's first pack; rows are g.X apart
  const T* pi = in + first;
  T* po = out + first;
  const T* pn = nb + (first - g.frame);                         // the same pack one plane back
  const bool row_end = x0 + 8 == g.X;
  uint32_t hm[W], hc[W], hp[W];
#pragma unroll
  for (int w = 0; w < W; ++w) hm[w] = hc[w] = hp[w] = 0;
  if (y0 >= 1) diff_row_sums<T>(pn - g.X, nb, pn - g.X + 8, hm);
  diff_row_sums<T>(pn, nb, (DECODE && row_end && y0 + 1 == g.Y) ? in + plane : pn + 8, hc);
#pragma unroll
  for (int k = 0; k < kDiffRows; ++k) {
    if ((uint32_t)k >= rows) break;
    const uint32_t y = y0 + (uint32_t)k;
    if (y + 1 < g.Y) diff_row_sums<T>(pn + g.X, nb, (DECODE && row_end && y + 2 == g.Y) ? in + plane : pn + g.X + 8, hp);
    if (!(last_row_spills && y + 1 == g.Y)) {
      const DiffWords<T> v = *reinterpret_cast<const DiffWords<T>*>(pi);
      const uint32_t own_row = (y >= 1 && y + 1 < g.Y) ? 0xffffffffu : 0u;      // rows 0 and Y-1 (no spill): nothing coded
      const uint32_t spill_row = y >= 2 ? 0xffffffffu : 0u;
      DiffWords<T> o;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const uint32_t mask = (own_mask[w] & own_row) | (spill_mask[w] & spill_row);
        const uint32_t q = L::div9(L::add(hm[w], L::add(hc[w], hp[w]))) & mask;
        o.w[w] = DECODE ? L::add(v.w[w], q) : L::sub(v.w[w], q);
      }
      *reinterpret_cast<DiffWords<T>*>(po) = o;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) { hm[w] = hc[w]; hc[w] = hp[w]; }
    pi += g.X; po += g.X; pn += g.X;
  }
  }
  if (last_row_spills) {
#pragma unroll 1
    for (uint32_t j = 0; j < 8; ++j)
      out[plane + (uint64_t)(g.Y - 1) * g.X + x0 + j] = diff_voxel<T, DECODE>(in, nb, g, plane, true, g.Y - 1, x0 + j);
  }
}

}  // namespace sqyb
