/* sqeazy C API — drop-in boundary of the B200-native implementation.
 *
 * Symbol-for-symbol replacement of the reference's src/cpp/inc/sqeazy.h for the uint16 volume
 * pipeline path (library name libsqeazy.so, as bound by the Java BridJ wrapper
 * src/java/src/sqeazy/bindings/SqeazyLibrary.java:15-19 and by tests/test_pipeline_interface.cpp).
 * All buffers are caller-owned HOST memory; every call is stateless and re-entrant.
 * Return convention of the reference: int 0 = success, non-zero = failure; bool for *_Possible*.
 * The heavy lifting behind these entry points runs as hand-written sm_100a CUDA kernels; there is
 * no CPU fallback: without a CUDA device the compute entry points return 1.
 */
#ifndef SQEAZY_B200_SQEAZY_H
#define SQEAZY_B200_SQEAZY_H

#ifdef __cplusplus
extern "C" {
#else
#include <stdbool.h>
#endif

/* reference: src/cpp/inc/sqeazy.h:26-46, src/cpp/src/sqeazy.cpp:16-22
 * in: *length = bytes available at src; out: *length = header bytes (incl. delimiter and padding) */
int SQY_Header_Size(const char* src, long* length);

/* reference: sqeazy.cpp:24-33 — in: *num = bytes at src; out: rank of the stored volume */
int SQY_Decompressed_NDims(const char* src, long* num);

/* reference: sqeazy.cpp:35-46 — in: shape[0] = bytes at src; out: shape[0..rank) */
int SQY_Decompressed_Shape(const char* src, long* shape);

/* reference: sqeazy.cpp:48-58 — in: *Sizeof = bytes at src; out: bytes per voxel */
int SQY_Decompressed_Sizeof(const char* src, long* Sizeof);

/* reference: sqeazy.cpp:270-279 — in: *length = bytes at src; out: raw bytes of the decoded volume */
int SQY_Decompressed_Length(const char* src, long* length);

/* reference: sqeazy.cpp:61-68 — version[0..3) = major, minor, patch */
int SQY_Version_Triple(int* version);

/* reference: sqeazy.h:143-166, sqeazy.cpp:108-142
 * pipeline: NUL-terminated, e.g. "rmestbkrd->bitswap1->lz4"; src: shape-product uint16 voxels (C order);
 * dst: at least SQY_Pipeline_Max_Compressed_Length_UI16 bytes; *dstlength (out only) = blob bytes.
 * nthreads: host threads, <= 0 or more than the machine has = all cores (sqeazy_algorithms.hpp:14-22). Here they stage
 * pageable src/dst buffers for the PCIe hop (page-locked buffers need none); the kernels have no thread knob. */
int SQY_PipelineEncode_UI16(const char* pipeline, const char* src, long* shape, unsigned shape_size, char* dst,
                            long* dstlength, int nthreads);

/* reference: sqeazy.cpp:72-106 (dypeline<uint8_t>) — uint8 voxels; stages with uint8 kernels: bitswap1|2|4, bitshuffle, diff3x3x1,
 * remove_background(threshold=N) -> lz4 | pass_through [-> lz4]; any other stage name returns 1 */
int SQY_PipelineEncode_UI8(const char* pipeline, const char* src, long* shape, unsigned shape_size, char* dst,
                           long* dstlength, int nthreads);

/* reference: sqeazy.cpp:166-183 — in: *length = raw bytes; out: upper bound of the blob size */
int SQY_Pipeline_Max_Compressed_Length_UI16(const char* pipeline, long pipeline_length, long* length);
int SQY_Pipeline_Max_Compressed_Length_UI8(const char* pipeline, long pipeline_length, long* length);

/* reference: sqeazy.cpp:185-207 — in: *length = strlen(pipeline), raw bytes from shape; out: bound */
int SQY_Pipeline_Max_Compressed_Length_3D_UI16(const char* pipeline, long* shape, unsigned shape_size, long* length);
int SQY_Pipeline_Max_Compressed_Length_3D_UI8(const char* pipeline, long* shape, unsigned shape_size, long* length);

/* reference: sqeazy.cpp:233-268 */
bool SQY_Pipeline_Possible_UI16(const char* pipeline);
bool SQY_Pipeline_Possible_UI8(const char* pipeline);
bool SQY_Pipeline_Possible(const char* pipeline, int sizeofpixel);

/* reference: sqeazy.h:274-277, sqeazy.cpp:281-307 — src: blob of srclength bytes; dst: SQY_Decompressed_Length bytes.
 * SQY_PipelineDecode_UI16 is an alias (the name used by BASELINE.json; the reference has no such symbol). */
int SQY_Decode_UI16(const char* src, long srclength, char* dst, int nthreads);
int SQY_PipelineDecode_UI16(const char* src, long srclength, char* dst, int nthreads);
int SQY_Decode_UI8(const char* src, long srclength, char* dst, int nthreads);

/* reference: sqeazy.h:329-452, sqeazy.cpp:341-554 — HDF5 entry points are outside this path (no libhdf5
 * in the image): exported for link compatibility, each returns 1 */
int SQY_h5_query_sizeof(const char* fname, const char* dname, unsigned* sizeof_out);
int SQY_h5_query_dtype(const char* fname, const char* dname, unsigned* dtype);
int SQY_h5_query_ndims(const char* fname, const char* dname, unsigned* ndims);
int SQY_h5_query_shape(const char* fname, const char* dname, unsigned* shape);
int SQY_h5_read_UI16(const char* fname, const char* dname, unsigned short* data);
int SQY_h5_write_UI16(const char* fname, const char* dname, const unsigned short* data, unsigned shape_size,
                      const unsigned* shape, const char* filter);
int SQY_h5_write(const char* fname, const char* dname, const char* data, unsigned long data_size);
int SQY_h5_link(const char* pSrcFileName, const char* pSrcLinkPath, const char* pSrcLinkName, const char* pTargetFile,
                const char* pTargetDatasetPath, const char* pTargetDatasetName);

#ifdef __cplusplus
}
#endif
#endif
