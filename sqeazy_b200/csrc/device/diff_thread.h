// Thread program of the diff3x3x1 kernels (diff.cu), compilable for the device and for the host: tests/helpers/diff_sim.cpp
// replays it thread by thread against the oracle where no GPU exists (any thread order, poisoned output).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SQYB_HD __device__ __forceinline__
#else
#define SQYB_HD inline
#endif

namespace sqyb {

template <typename T> struct alignas(sizeof(T) * 8) Pack8 { T v[8]; };

struct DiffGeom {
  uint64_t Z, Y, X, frame, zend;
};

inline DiffGeom diff_geom(uint64_t Z, uint64_t Y, uint64_t X) { return DiffGeom{Z, Y, X, Y * X, X < Z ? X : Z}; }

// shapes on which the reference's loops stay inside their plane (see diff.cu)
inline bool diff_shape_ok(uint64_t Z, uint64_t Y, uint64_t X) {
  if (Z < 3 || Y < 3 || X < 2) return false;
  if (Z > 32767 || Y > 32767 || X > 32767) return false;       // int16 coordinates in naive_sum (diff_scheme_utils.hpp:81-89)
  if ((X - 1) * (Y - 2) <= 1) return false;                    // compute_offsets_in_x: a single offset = sweep to the end
  const uint64_t zend = X < Z ? X : Z;
  if ((zend - 1) * (Y - 2) <= 1) return false;
  if ((Y - 2) * X + 1 + (Z - 2) > Y * X) return false;         // the last run of a plane would leave it
  return true;
}

// the launches of one encode (the whole volume at once) or decode (plane 0 as stored, one launch per coded plane in z
// order, the planes the reference skips): f(begin, end) over element ranges, stops at the first non-zero return
template <typename F>
inline int diff_for_each_launch(bool decode, const DiffGeom& g, F f) {
  if (!decode) return f(0, g.Z * g.frame);
  int rc = f(0, g.frame);
  for (uint64_t z = 1; z < g.zend && !rc; ++z) rc = f(z * g.frame, (z + 1) * g.frame);
  if (!rc && g.zend < g.Z) rc = f(g.zend * g.frame, g.Z * g.frame);
  return rc;
}

SQYB_HD bool diff_covered(const DiffGeom& g, uint64_t z, uint64_t y, uint64_t x) {
  if (z < 1 || z >= g.zend) return false;
  const bool own = y >= 1 && y + 1 < g.Y && x >= 1 && x + 1 < g.Z;          // the run of row y
  const bool spilled = y >= 2 && y <= g.Y - 1 && x + g.X + 1 < g.Z;          // the end of the run of row y-1
  return own || spilled;
}

// Elements [begin, end) of the volume, 8 per thread. `nb` holds the neighbours: `in` itself for encode, `out` for decode
// (decode: the launch covers one plane, every neighbour lies in the plane before it or, for the spilled end of the last
// run, in row 0 of the same plane, which is never coded and therefore read from `in`).
template <typename T, bool DECODE>
SQYB_HD void diff_thread(const T* in, T* out, const T* nb, uint64_t begin, uint64_t end, const DiffGeom& g, uint64_t tid) {
  const uint64_t i0 = begin + 8 * tid;
  if (i0 >= end) return;
  const int cnt = end - i0 < 8 ? (int)(end - i0) : 8;
  const uint64_t z = i0 / g.frame, r = i0 - z * g.frame, y = r / g.X, x = r - y * g.X;
  constexpr uintptr_t kMask = sizeof(T) * 8 - 1;

  const bool one_row = cnt == 8 && x + 8 <= g.X;
  const bool io_aligned = ((((uintptr_t)(in + i0)) | ((uintptr_t)(out + i0))) & kMask) == 0;
  if (one_row && io_aligned && (z < 1 || z >= g.zend || y < 1)) {   // nothing coded here: copy
    *reinterpret_cast<Pack8<T>*>(out + i0) = *reinterpret_cast<const Pack8<T>*>(in + i0);
    return;
  }
  if (one_row && io_aligned && y + 1 < g.Y) {
    const Pack8<T> v = *reinterpret_cast<const Pack8<T>*>(in + i0);
    T acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0;
    const uint64_t base = i0 - g.frame;   // z >= 1 here
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const uint64_t row = dy < 0 ? base - g.X : (dy > 0 ? base + g.X : base);   // y >= 1: row >= 0
      const T* p = nb + row;
      T e[10];
      if ((((uintptr_t)p) & kMask) == 0) {
        const Pack8<T> c = *reinterpret_cast<const Pack8<T>*>(p);
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j + 1] = c.v[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j + 1] = p[j];
      }
      e[0] = row > 0 ? p[-1] : (T)0;      // row == 0 only for z = 1, y = 1, x = 0, which is not coded
      // x + 8 == X in row Y-2: the element after the row below is row 0 of the plane being decoded (never coded: from `in`)
      e[9] = (DECODE && row + 8 >= z * g.frame) ? in[row + 8] : p[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = (T)(acc[j] + (T)(e[j] + (T)(e[j + 1] + e[j + 2])));
    }
    Pack8<T> o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t q = diff_covered(g, z, y, x + j) ? (uint32_t)acc[j] / 9u : 0u;
      o.v[j] = DECODE ? (T)(v.v[j] + q) : (T)(v.v[j] - q);
    }
    *reinterpret_cast<Pack8<T>*>(out + i0) = o;
    return;
  }

  // general route: any alignment, packs that cross a row, the last row of a plane
  uint64_t zz = z, yy = y, xx = x;
  for (int j = 0; j < cnt; ++j) {
    const uint64_t i = i0 + j;
    uint32_t q = 0;
    if (diff_covered(g, zz, yy, xx)) {
      T sum = 0;
      const uint64_t plane_begin = zz * g.frame;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const uint64_t k = (uint64_t)((int64_t)(i - g.frame) + (int64_t)dy * (int64_t)g.X + dx);
          sum = (T)(sum + ((DECODE && k >= plane_begin) ? in[k] : nb[k]));
        }
      q = (uint32_t)sum / 9u;
    }
    out[i] = DECODE ? (T)(in[i] + q) : (T)(in[i] - q);
    if (++xx == g.X) {
      xx = 0;
      if (++yy == g.Y) { yy = 0; ++zz; }
    }
  }
}

}  // namespace sqyb
