// Internal launch interface of the sm_100a kernels (device pointers, explicit stream).
// Every function enqueues work on `st` and returns 0 or a cudaError_t / negative library code.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>

namespace sqyb {

// cumulative count of kernels launched by this library (reported as gpu_launches by bench.py)
extern std::atomic<long> g_kernel_launches;
#define SQYB_COUNT_LAUNCH(n) ::sqyb::g_kernel_launches.fetch_add((n), std::memory_order_relaxed)

// bitswap.cu
int k_bitswap_encode(int w, const uint16_t* in, uint16_t* out, uint64_t n, int threshold, cudaStream_t st);
int k_bitswap_decode(int w, const uint16_t* in, uint16_t* out, uint64_t n, cudaStream_t st);
int k_remove_background(const uint16_t* in, uint16_t* out, uint64_t n, int threshold, cudaStream_t st);
// voxels [first, first+count) of an n-voxel volume (in/out = whole-volume base pointers; multiples of 128 voxels): lets the
// host entry points transpose a z-slab as soon as it has arrived / ship a slab as soon as it is transposed back
int k_bitswap_encode_range(int w, const uint16_t* in, uint16_t* out, uint64_t n, uint64_t first, uint64_t count, int threshold,
                           cudaStream_t st);
int k_bitswap_decode_range(int w, const uint16_t* in, uint16_t* out, uint64_t n, uint64_t first, uint64_t count, cudaStream_t st);

// bitshuffle.cu: the bitshuffle library's blocked bit transpose for uint16 (encoders/bitshuffle_scheme_impl.hpp:91-160).
// block_size in elements, 0 = the library's default (4096); -81 (the library's own code) when it is not a multiple of 8
int k_bitshuffle16_encode(const uint16_t* in, uint16_t* out, uint64_t n, uint32_t block_size, cudaStream_t st);
int k_bitshuffle16_decode(const uint16_t* in, uint16_t* out, uint64_t n, uint32_t block_size, cudaStream_t st);
int k_bitshuffle8_encode(const uint8_t* in, uint8_t* out, uint64_t n, uint32_t block_size, cudaStream_t st);   // uint8: default 8192
int k_bitshuffle8_decode(const uint8_t* in, uint8_t* out, uint64_t n, uint32_t block_size, cudaStream_t st);
uint32_t bitshuffle_block_elems(uint32_t block_size, int elem_size);

// diff.cu: diff3x3x1 (encoders/diff_scheme_impl.hpp:78-199), elem = bytes per voxel (1 or 2); in != out; 1 = shape refused
bool diff_shape_supported(uint64_t Z, uint64_t Y, uint64_t X, int elem);
int k_diff_encode(int elem, const void* in, void* out, uint64_t Z, uint64_t Y, uint64_t X, cudaStream_t st);
int k_diff_decode(int elem, const void* in, void* out, uint64_t Z, uint64_t Y, uint64_t X, cudaStream_t st);

// bitswap8.cu (uint8 volumes)
int k_bitswap8_encode(int w, const uint8_t* in, uint8_t* out, uint64_t n, int threshold, cudaStream_t st);
int k_bitswap8_decode(int w, const uint8_t* in, uint8_t* out, uint64_t n, cudaStream_t st);
int k_remove_background8(const uint8_t* in, uint8_t* out, uint64_t n, int threshold, cudaStream_t st);

// quantise.cu
int k_histogram_u16(const uint16_t* in, uint64_t n, uint32_t* hist /* 65536 x u32, accumulated into */, cudaStream_t st);
// per histogram h: out[4h] = first bin whose cumulative share exceeds `threshold` (65536: none), out[4h+1] = bins[m],
// out[4h+2] = bins[m-1] (hist_impl.hpp:63-84,359-381)
int k_support_index(const uint32_t* hist_dev, int nhist, float threshold, uint32_t* out_dev, cudaStream_t st);
int k_lut_apply(const uint16_t* in, uint8_t* out, uint64_t n, const uint8_t* lut_dev /* 65536 */, cudaStream_t st);
int k_lut_decode(const uint8_t* in, uint16_t* out, uint64_t n, const uint16_t* lut_dev /* 256 */, cudaStream_t st);

// lz4_encode.cu
size_t k_lz4_encode_workspace_bytes(uint64_t raw_bytes);
// workspace[0..8) receives the payload size (u64) when the stream has drained; [16,28) block-kind counters
// pitch_bytes: distance in the stream between vertically adjacent voxels (a row of a bit plane: X * w / 8 bytes; a row of
// 8-bit codes: X bytes), tried as a fixed match offset like offset 1; 0 (or not a multiple of 32): none. The
// compressed bytes are a pure function of (input bytes, pitch_bytes).
// kLz4HintNoNoise OR-ed into pitch_bytes: the caller expects no incompressible blocks (bit planes behind a background
// removal). It changes only the order of work inside a block — without it twelve of a block's sixteen warps wait until four
// sampled ones have settled whether the block is noise and is stored as it is, which is what makes noise cheap (8-bit
// quantiser codes, the low planes of an unfiltered stack) and costs a compressible block a few per cent — never a byte.
constexpr uint32_t kLz4HintNoNoise = 0x80000000u;
int k_lz4_encode(const uint8_t* src, uint64_t raw_bytes, uint8_t* dst, void* workspace, uint32_t pitch_bytes, cudaStream_t st);
// the same in three steps, for input that becomes available piece by piece: begin (frame prefix), any number of block
// sets — `nsets` sets of `count` consecutive 16 KiB blocks, set j starting at block first + j * set_stride (the pieces of
// the 16/w bit planes that one z-slab contributes) —, end (offsets + compaction). Every block exactly once.
int k_lz4_encode_begin(uint64_t raw_bytes, uint8_t* dst, void* workspace, cudaStream_t st);
int k_lz4_encode_blocks(const uint8_t* src, uint64_t raw_bytes, uint8_t* dst, void* workspace, uint32_t first, uint32_t count,
                        uint32_t nsets, uint32_t set_stride, uint32_t pitch_bytes, cudaStream_t st);
int k_lz4_encode_end(const uint8_t* src, uint64_t raw_bytes, uint8_t* dst, void* workspace, cudaStream_t st);

// lz4_decode.cu
size_t k_lz4_decode_workspace_bytes(uint64_t dst_bytes);
// allow_deferred: streams with many block-LINKED blocks (the reference's serial mode) only get their block table here; their
// blocks are decoded by k_lz4_decode_linked (deferred cross-block references) once the caller has seen deferred_blocks > 0
// in the status and provided the origin buffer.
int k_lz4_decode(const uint8_t* src, uint64_t src_bytes, uint8_t* dst, uint64_t dst_bytes, void* workspace, int measure_all,
                 int allow_deferred, cudaStream_t st);
// k_lz4_decode in two steps, for callers that want the blocks of one z-slab at a time (the host entry points ship a slab while
// the next one is decoded): the block table, then any number of block-set launches over it (count == 0: the whole stream).
int k_lz4_decode_tables(const uint8_t* src, uint64_t src_bytes, uint64_t dst_bytes, void* workspace, int measure_all, int allow_deferred,
                        cudaStream_t st);
int k_lz4_decode_run(const uint8_t* src, uint64_t src_bytes, uint8_t* dst, uint64_t dst_bytes, void* workspace, uint32_t first,
                     uint32_t count, uint32_t nsets, uint32_t set_stride, uint32_t slot, cudaStream_t st);
int k_lz4_decode_peek(void* workspace, uint32_t* error, uint32_t* own_frame, uint32_t* nblocks, uint32_t* block_bytes, cudaStream_t st);
size_t k_lz4_decode_linked_workspace_bytes(uint64_t dst_bytes);
int k_lz4_decode_linked(const uint8_t* src, uint64_t src_bytes, uint8_t* dst, uint64_t dst_bytes, void* workspace, void* origins,
                        cudaStream_t st);
// decoded-size limit of the lane-serial block decoder (0 = warp-per-block decoder only); returns the previous value
// linked blocks from which a stream takes the deferred-reference path (default 8, 0 = never); returns the previous value
long k_lz4_set_defer_min(long nblocks);
int k_lz4_decode_status(void* workspace, uint32_t* error, uint64_t* total_decoded, uint32_t* deferred_blocks, cudaStream_t st);

}  // namespace sqyb
