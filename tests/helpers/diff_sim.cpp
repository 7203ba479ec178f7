// CPU replay of the diff3x3x1 kernels (sqeazy_b200/csrc/device/diff.cu): the same thread program (diff_thread.h) and the
// same launch schedule, threads of a launch run in the order asked for (0 forward, 1 backward, 2 odd threads first) over an
// output that is poisoned wherever the launch has not written yet - a thread that read what another thread of the same
// launch writes would show. Test infrastructure.  g++ -O2 -shared -fPIC -I sqeazy_b200/csrc/device
#include <cstdint>
#include <cstring>

#include "diff_thread.h"

template <typename T>
static int run(int decode, const T* in, T* out, uint64_t Z, uint64_t Y, uint64_t X, int order) {
  using namespace sqyb;
  if (!diff_shape_ok(Z, Y, X, (int)sizeof(T))) return 1;
  const DiffGeom g = diff_geom(Z, Y, X, decode != 0);
  std::memset(out, 0xA5, Z * Y * X * sizeof(T));
  return diff_for_each_launch(decode != 0, g, [&](uint32_t z0, uint32_t planes) {
    const uint32_t threads = (uint32_t)((diff_threads_per_plane(g) + 255) / 256 * 256);     // whole CTAs, like the grid
    auto one = [&](uint32_t z, uint32_t t) {
      if (decode) diff_thread<T, true>(in, out, out, g, z, t);
      else diff_thread<T, false>(in, out, in, g, z, t);
    };
    for (uint32_t zi = 0; zi < planes; ++zi) {
      const uint32_t z = order == 1 ? z0 + planes - 1 - zi : z0 + zi;       // planes of one launch in any order too
      if (order == 0) for (uint32_t t = 0; t < threads; ++t) one(z, t);
      else if (order == 1) for (uint32_t t = threads; t-- > 0;) one(z, t);
      else { for (uint32_t t = 1; t < threads; t += 2) one(z, t); for (uint32_t t = 0; t < threads; t += 2) one(z, t); }
    }
    return 0;
  });
}

extern "C" int sim_diff(int decode, const void* in, void* out, uint64_t Z, uint64_t Y, uint64_t X, int elem, int order) {
  return elem == 1 ? run<uint8_t>(decode, static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), Z, Y, X, order)
                   : run<uint16_t>(decode, static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), Z, Y, X, order);
}
