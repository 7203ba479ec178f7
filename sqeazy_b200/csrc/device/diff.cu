// diff3x3x1 head filter: every voxel minus the mean of the 3x3 voxels around it in the previous z plane.
// Reference: encoders/diff_scheme_impl.hpp:78-199 with last_plane_neighborhood<3> (neighborhood_utils.hpp:72-92), the
// row list of halo::compute_offsets_in_x (neighborhood_utils.hpp:186-226) and naive_sum (diff_scheme_utils.hpp:75-103).
//
// What the reference computes (restated in oracle/sqy_oracle.c: orc_diff, pinned against the compiled reference):
//   q(i)   = ((sum of v[i - Y*X + dy*X + dx], dy, dx in {-1,0,1}) mod 2^bits) / 9        -- the sum wraps in the voxel type
//   encode : out[i] = in[i] - q(i) on `in`            for covered i, out[i] = in[i] elsewhere
//   decode : out[i] = in[i] + q(i) on `out`, in index order (plane z needs the decoded plane z-1)
//   covered: rows (z, y), 1 <= z < min(X, Z), 1 <= y < Y-1, of each the Z-2 indices from x = 1 on (the reference takes
//            the x range from the Z extent and the z range from the X extent); for Z > X the runs spill into the next row.
// Shapes the reference itself mishandles (runs leaving the plane, the one-row sweep, extents beyond the int16 / int8
// coordinates it keeps for uint16 / uint8 stacks) are refused.
//
// Encode is one launch over the volume: 4 B/voxel of HBM traffic, the previous plane is re-read from L2. A thread walks a
// strip of 8 voxels x R rows (8 in encode, 4 in decode) and keeps the 3-tap row sums of the previous plane it can reuse
// (R + 2 row loads of 16 bytes + 2 halo voxels for R output rows). Decode is a recurrence along z only: one launch per coded
// plane (all voxels of a plane are independent), planes 0 and >= min(X, Z) are copies.
#include "common.cuh"
#include "diff_thread.h"
#include "kernels.h"

namespace sqyb {

namespace {

constexpr int kDiffThreads = 128;   // 12 CTAs per SM (<= 40 registers): a 2048 x 2048 plane = 1024 CTAs is one wave of the 148 SMs

template <typename T, bool DECODE>
__global__ void __launch_bounds__(kDiffThreads, DECODE ? 8 : 12) diff_kernel(const T* __restrict__ in, T* __restrict__ out, const T* __restrict__ nb, DiffGeom g,
                                                            uint32_t z0) {
  if (DECODE) {
    // decode launches follow each other plane by plane with programmatic stream serialization: the CTAs of the next plane
    // are placed while this one drains and wait here until the plane before them is complete and visible
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  diff_thread<T, DECODE>(in, out, nb, g, z0 + blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

template <typename T, bool DECODE>
int launch_planes(const T* in, T* out, const T* nb, const DiffGeom& g, uint32_t z0, uint32_t planes, cudaStream_t st) {
  if (planes == 0) return 0;
  const uint64_t blocks = (diff_threads_per_plane(g) + kDiffThreads - 1) / kDiffThreads;     // <= 2^20
  for (uint32_t done = 0; done < planes; done += 65535) {              // gridDim.y limit (not reached: Z <= 32767)
    const uint32_t now = planes - done < 65535 ? planes - done : 65535;
    if (DECODE) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)blocks, now);
      cfg.blockDim = dim3(kDiffThreads);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      const cudaError_t e = cudaLaunchKernelEx(&cfg, diff_kernel<T, DECODE>, in, out, nb, g, z0 + done);
      if (e != cudaSuccess) return (int)e;
    } else {
      diff_kernel<T, DECODE><<<dim3((unsigned)blocks, now), kDiffThreads, 0, st>>>(in, out, nb, g, z0 + done);
    }
    SQYB_COUNT_LAUNCH(1);
  }
  return (int)cudaGetLastError();
}

template <typename T>
int diff_run(bool decode, const T* in, T* out, uint64_t Z, uint64_t Y, uint64_t X, cudaStream_t st) {
  if (!diff_shape_ok(Z, Y, X, (int)sizeof(T)) || in == out) return 1;
  const DiffGeom g = diff_geom(Z, Y, X, decode);
  return diff_for_each_launch(decode, g, [&](uint32_t z0, uint32_t planes) {
    return decode ? launch_planes<T, true>(in, out, out, g, z0, planes, st) : launch_planes<T, false>(in, out, in, g, z0, planes, st);
  });
}

}  // namespace

bool diff_shape_supported(uint64_t Z, uint64_t Y, uint64_t X, int elem) { return diff_shape_ok(Z, Y, X, elem); }

int k_diff_encode(int elem, const void* in, void* out, uint64_t Z, uint64_t Y, uint64_t X, cudaStream_t st) {
  return elem == 1 ? diff_run<uint8_t>(false, static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), Z, Y, X, st)
                   : diff_run<uint16_t>(false, static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), Z, Y, X, st);
}
int k_diff_decode(int elem, const void* in, void* out, uint64_t Z, uint64_t Y, uint64_t X, cudaStream_t st) {
  return elem == 1 ? diff_run<uint8_t>(true, static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), Z, Y, X, st)
                   : diff_run<uint16_t>(true, static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), Z, Y, X, st);
}

}  // namespace sqyb
