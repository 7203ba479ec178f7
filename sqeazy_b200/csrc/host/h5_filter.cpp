// HDF5 filter plugin entry points (include/sqeazy_h5_filter.h; reference: inc/sqeazy_h5_filter.hpp:28-226). Plain host
// code over the library's own C ABI: the chunk is encoded / decoded by SQY_PipelineEncode_* / SQY_Decode_* (GPU stages,
// caller-owned host buffers). Nothing of libhdf5 is called, so nothing of it is linked.
#include <climits>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/sqeazy.h"
#include "../../../include/sqeazy_h5_filter.h"
#include "pipeline.hpp"
#include "text.hpp"

using namespace sqyb;

namespace {

uint64_t product(const std::vector<uint64_t>& shape) {
  uint64_t n = 1;
  for (uint64_t d : shape) n *= d;
  return shape.empty() ? 0 : n;
}

// sqeazy_h5_filter.hpp:43-104: the blob's own header says what to build
size_t filter_read(size_t* buf_size, void** buf) {
  const char* in = static_cast<const char*>(*buf);
  const Header hdr = unpack_header(in, *buf_size);
  if (!hdr.valid) return 0;                                                    // ret = 100 in the reference
  const unsigned elem = sizeof_typename(hdr.raw_type);
  if (elem != 1 && elem != 2) return 0;
  if (hdr.size + hdr.compressed_bytes > *buf_size) return 0;                   // (the reference would read past the chunk)
  const uint64_t raw_bytes = product(hdr.shape) * elem;
  if (raw_bytes == 0) return 0;
  char* out = static_cast<char*>(std::malloc(raw_bytes));
  if (!out) return 0;
  const long blob = (long)(hdr.size + hdr.compressed_bytes);
  const int rc = elem == 2 ? SQY_Decode_UI16(in, blob, out, 1) : SQY_Decode_UI8(in, blob, out, 1);
  if (rc != 0) {
    std::free(out);
    return 0;
  }
  std::free(*buf);
  *buf = out;
  *buf_size = raw_bytes;
  return raw_bytes;
}

// sqeazy_h5_filter.hpp:106-186: pipeline, voxel type and shape come from the header text in cd_values
size_t filter_write(size_t cd_nelmts, const unsigned cd_values[], size_t nbytes, size_t* buf_size, void** buf) {
  const char* in = static_cast<const char*>(*buf);
  const size_t cd_bytes = cd_nelmts * sizeof(unsigned);
  if (!cd_values || cd_bytes == 0) return 0;
  const Header want = unpack_header(reinterpret_cast<const char*>(cd_values), cd_bytes);

  // a chunk that already carries a sqeazy header is stored as it is (:118-132). The reference looks for that header in the
  // first 2 * cd_bytes of the chunk; blobs of this library keep their header right-aligned in a slot of 256-byte multiples
  // with the canonical stage configs spelled out, so the window is 2 KiB wider.
  const size_t window = 2 * cd_bytes + 2048;
  const size_t look = window > nbytes ? nbytes : window;
  const Header have = unpack_header(in, look);
  if (have.valid) {
    const uint64_t blob = have.size + have.compressed_bytes;
    if (blob > nbytes) return 0;
    char* out = static_cast<char*>(std::malloc(blob));
    if (!out) return 0;
    std::memcpy(out, in, blob);
    std::free(*buf);
    *buf = out;
    *buf_size = blob;
    return blob;
  }

  if (!want.valid || !pipeline_possible_u16(want.pipeline)) return 0;          // :138-143 asks the uint16 registry for both types
  const unsigned elem = sizeof_typename(want.raw_type);
  if (elem != 1 && elem != 2) return 0;
  if (elem == 1 && !pipeline_possible_u8(want.pipeline)) return 0;
  if (product(want.shape) * elem != nbytes) return 0;                          // the chunk is not the stack the header describes
  long cap = (long)nbytes;
  const std::string& p = want.pipeline;
  if ((elem == 2 ? SQY_Pipeline_Max_Compressed_Length_UI16(p.c_str(), (long)p.size(), &cap)
                 : SQY_Pipeline_Max_Compressed_Length_UI8(p.c_str(), (long)p.size(), &cap)) != 0)
    return 0;
  char* out = static_cast<char*>(std::malloc((size_t)cap));
  if (!out) return 0;
  std::vector<long> shape(want.shape.begin(), want.shape.end());
  long len = 0;
  const int rc = elem == 2 ? SQY_PipelineEncode_UI16(p.c_str(), in, shape.data(), (unsigned)shape.size(), out, &len, 1)
                           : SQY_PipelineEncode_UI8(p.c_str(), in, shape.data(), (unsigned)shape.size(), out, &len, 1);
  if (rc != 0 || len <= 0) {
    std::free(out);
    return 0;
  }
  if (char* fit = static_cast<char*>(std::realloc(out, (size_t)len))) out = fit;
  std::free(*buf);
  *buf = out;
  *buf_size = (size_t)len;
  return (size_t)len;
}

const sqy_h5z_class2 kSqyFilter[1] = {{
    SQY_H5Z_CLASS_T_VERS, SQY_H5Z_FILTER_ID, 1, 1,
    "HDF5 sqy filter; see https://github.org/sqeazy/sqeazy",   // sqeazy_h5_filter.hpp:218
    nullptr, nullptr, &H5Z_filter_sqy,
}};

}  // namespace

extern "C" {

size_t H5Z_filter_sqy(unsigned flags, size_t cd_nelmts, const unsigned cd_values[], size_t nbytes, size_t* buf_size, void** buf) {
  if (!buf || !*buf || !buf_size) return 0;
  try {
    return (flags & SQY_H5Z_FLAG_REVERSE) ? filter_read(buf_size, buf) : filter_write(cd_nelmts, cd_values, nbytes, buf_size, buf);
  } catch (...) {
    return 0;
  }
}

int H5PLget_plugin_type(void) { return SQY_H5PL_TYPE_FILTER; }
const void* H5PLget_plugin_info(void) { return kSqyFilter; }

}  // extern "C"
