// LZ4 frame / block format constants shared by the GPU encoder, the GPU decoder and the host code.
// The reference's `lz4` stage emits LZ4 *Frame* streams through liblz4 (encoders/lz4.hpp:58-114,
// 214-242; encoders/lz4_utils.hpp:99-173); liblz4 itself is an un-vendored dependency, so the
// format below follows the published LZ4 Frame Format / Block Format descriptions (lz4 >= 1.8).
//
// Payload layout written by this library for `...->lz4`:
//   [skippable frame  : u32 magic 0x184D2A5B | u32 size | SqybIndex header | u32 block_word[nblocks]]
//   [LZ4 frame header : 04 22 4D 18 | FLG 0x60 (v1, independent blocks) | BD 0x40 (64 KiB max) | HC 0x82]
//   [block]*          : u32 size (bit31 = stored) | data        (blocks hold kLz4BlockBytes of input)
//   [EndMark          : u32 0]
// The skippable frame is legal LZ4F (decoders skip it; verified against the reference's own decode
// loop, encoders/lz4.hpp:257-339 + liblz4 1.9.4) and lets the GPU decoder find every block without
// walking the stream sequentially.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SQYB_HD __host__ __device__
#else
#define SQYB_HD
#endif

namespace sqyb {

constexpr uint32_t kLz4FrameMagic = 0x184D2204u;
constexpr uint32_t kLz4SkippableMagicBase = 0x184D2A50u;   // ..5F
constexpr uint32_t kSqybSkippableMagic = 0x184D2A5Bu;
constexpr uint32_t kSqybIndexMagic = 0x42595153u;           // "SQYB"
constexpr uint32_t kLz4StoredFlag = 0x80000000u;

constexpr int kLz4BlockBytes = 16384;   // input bytes per LZ4 block produced by the GPU encoder
constexpr int kLz4MinMatch = 4;
constexpr int kLz4LastLiterals = 5;     // the last 5 bytes of a block are always literals
constexpr int kLz4MFLimit = 12;         // the last match starts at least 12 bytes before the block end

struct SqybIndexHeader {                // 32 bytes, little endian
  uint32_t magic;                       // kSqybIndexMagic
  uint32_t version;                     // 1
  uint32_t block_bytes;                 // input bytes per block (last block may be shorter)
  uint32_t nblocks;
  uint64_t raw_bytes;                   // total decoded bytes of the frame that follows
  uint64_t frame_bytes;                 // bytes of the LZ4 frame that follows (header .. EndMark), filled by the encoder
};

constexpr uint64_t kSkippableHeaderBytes = 8;
constexpr uint64_t kLz4FrameHeaderBytes = 7;
constexpr uint64_t kLz4EndMarkBytes = 4;

SQYB_HD static inline uint64_t lz4_nblocks(uint64_t raw_bytes) { return (raw_bytes + kLz4BlockBytes - 1) / kLz4BlockBytes; }
// bytes in front of the first LZ4 block header
SQYB_HD static inline uint64_t lz4_prefix_bytes(uint64_t nblocks) {
  return kSkippableHeaderBytes + sizeof(SqybIndexHeader) + 4 * nblocks + kLz4FrameHeaderBytes;
}
// worst case payload: every block stored
SQYB_HD static inline uint64_t lz4_payload_bound(uint64_t raw_bytes) {
  const uint64_t nb = lz4_nblocks(raw_bytes);
  return lz4_prefix_bytes(nb) + raw_bytes + 4 * nb + kLz4EndMarkBytes;
}

}  // namespace sqyb
