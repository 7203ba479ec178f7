/* Hand-written declarations of the LZ4 frame API subset used by the reference
 * (encoders/lz4.hpp, encoders/lz4_utils.hpp), matching the liblz4 1.9.4 ABI of
 * /usr/lib/x86_64-linux-gnu/liblz4.so.1 (lz4 dev headers are not installed).
 * Test infrastructure only (oracle/_ref and oracle C port). */
#ifndef SQYB_LZ4FRAME_DECL_H
#define SQYB_LZ4FRAME_DECL_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LZ4F_VERSION 100
#define LZ4F_HEADER_SIZE_MIN 7
#define LZ4F_HEADER_SIZE_MAX 19

typedef size_t LZ4F_errorCode_t;

typedef enum { LZ4F_default = 0, LZ4F_max64KB = 4, LZ4F_max256KB = 5, LZ4F_max1MB = 6, LZ4F_max4MB = 7 } LZ4F_blockSizeID_t;
typedef enum { LZ4F_blockLinked = 0, LZ4F_blockIndependent } LZ4F_blockMode_t;
typedef enum { LZ4F_noContentChecksum = 0, LZ4F_contentChecksumEnabled } LZ4F_contentChecksum_t;
typedef enum { LZ4F_noBlockChecksum = 0, LZ4F_blockChecksumEnabled } LZ4F_blockChecksum_t;
typedef enum { LZ4F_frame = 0, LZ4F_skippableFrame } LZ4F_frameType_t;

typedef struct {
  LZ4F_blockSizeID_t blockSizeID;
  LZ4F_blockMode_t blockMode;
  LZ4F_contentChecksum_t contentChecksumFlag;
  LZ4F_frameType_t frameType;
  unsigned long long contentSize;
  unsigned dictID;
  LZ4F_blockChecksum_t blockChecksumFlag;
} LZ4F_frameInfo_t;

typedef struct {
  LZ4F_frameInfo_t frameInfo;
  int compressionLevel;
  unsigned autoFlush;
  unsigned favorDecSpeed;
  unsigned reserved[3];
} LZ4F_preferences_t;

typedef struct LZ4F_cctx_s LZ4F_cctx;
typedef LZ4F_cctx* LZ4F_compressionContext_t;
typedef struct LZ4F_dctx_s LZ4F_dctx;
typedef LZ4F_dctx* LZ4F_decompressionContext_t;

typedef struct { unsigned stableSrc; unsigned reserved[3]; } LZ4F_compressOptions_t;
typedef struct { unsigned stableDst; unsigned skipChecksums; unsigned reserved1; unsigned reserved0; } LZ4F_decompressOptions_t;

unsigned LZ4F_isError(LZ4F_errorCode_t code);
const char* LZ4F_getErrorName(LZ4F_errorCode_t code);
size_t LZ4F_compressFrameBound(size_t srcSize, const LZ4F_preferences_t* prefsPtr);
size_t LZ4F_compressFrame(void* dst, size_t dstCap, const void* src, size_t srcSize, const LZ4F_preferences_t* prefsPtr);
LZ4F_errorCode_t LZ4F_createCompressionContext(LZ4F_cctx** cctxPtr, unsigned version);
LZ4F_errorCode_t LZ4F_freeCompressionContext(LZ4F_cctx* cctx);
size_t LZ4F_compressBegin(LZ4F_cctx* cctx, void* dstBuffer, size_t dstCapacity, const LZ4F_preferences_t* prefsPtr);
size_t LZ4F_compressBound(size_t srcSize, const LZ4F_preferences_t* prefsPtr);
size_t LZ4F_compressUpdate(LZ4F_cctx* cctx, void* dstBuffer, size_t dstCapacity, const void* srcBuffer, size_t srcSize, const LZ4F_compressOptions_t* cOptPtr);
size_t LZ4F_flush(LZ4F_cctx* cctx, void* dstBuffer, size_t dstCapacity, const LZ4F_compressOptions_t* cOptPtr);
size_t LZ4F_compressEnd(LZ4F_cctx* cctx, void* dstBuffer, size_t dstCapacity, const LZ4F_compressOptions_t* cOptPtr);
LZ4F_errorCode_t LZ4F_createDecompressionContext(LZ4F_dctx** dctxPtr, unsigned version);
LZ4F_errorCode_t LZ4F_freeDecompressionContext(LZ4F_dctx* dctx);
size_t LZ4F_headerSize(const void* src, size_t srcSize);
size_t LZ4F_getFrameInfo(LZ4F_dctx* dctx, LZ4F_frameInfo_t* frameInfoPtr, const void* srcBuffer, size_t* srcSizePtr);
size_t LZ4F_decompress(LZ4F_dctx* dctx, void* dstBuffer, size_t* dstSizePtr, const void* srcBuffer, size_t* srcSizePtr, const LZ4F_decompressOptions_t* dOptPtr);
void LZ4F_resetDecompressionContext(LZ4F_dctx* dctx);

/* block API (lz4.h), used only by tests/bench helpers */
int LZ4_compress_default(const char* src, char* dst, int srcSize, int dstCapacity);
int LZ4_compress_fast(const char* src, char* dst, int srcSize, int dstCapacity, int acceleration);
int LZ4_decompress_safe(const char* src, char* dst, int compressedSize, int dstCapacity);
int LZ4_compressBound(int inputSize);
int LZ4_versionNumber(void);

#ifdef __cplusplus
}
#endif
#endif
