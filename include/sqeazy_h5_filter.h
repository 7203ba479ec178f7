/* HDF5 filter plugin entry points of libsqeazy.so.
 *
 * reference: src/cpp/inc/sqeazy_h5_filter.hpp:28-224 (H5Z_filter_sqy, H5Z_FILTER_SQY = 01307 - an OCTAL literal in the
 * reference, i.e. filter id 711 -, H5Z_SQY[1], H5PLget_plugin_type, H5PLget_plugin_info) and inc/H5PLextern.h.
 *
 * HDF5 finds a filter plugin by dlopen()ing the libraries on HDF5_PLUGIN_PATH and asking them H5PLget_plugin_type() /
 * H5PLget_plugin_info(); the plugin itself calls nothing of libhdf5. This image has no libhdf5 and no hdf5.h, so the few
 * ABI items the plugin needs are declared here with the values and layout of H5Zpublic.h / H5PLpublic.h (HDF5 1.8.11+,
 * unchanged through 1.14): a program linked against libhdf5 can load libsqeazy.so as the sqy filter. The SQY_h5_* file
 * functions of sqeazy.h (sqeazy.cpp:341-554) do need libhdf5 and stay stubs that return 1.
 *
 * The filter works on caller-owned HOST buffers like the rest of the C API; the stages run on the GPU.
 */
#ifndef SQEAZY_B200_H5_FILTER_H
#define SQEAZY_B200_H5_FILTER_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQY_H5Z_FILTER_ID 01307          /* sqeazy_h5_filter.hpp:211 - octal: 711 */
#define SQY_H5Z_FLAG_REVERSE 0x0100u     /* H5Zpublic.h: H5Z_FLAG_REVERSE, set when the chunk is read */
#define SQY_H5Z_CLASS_T_VERS 1           /* H5Zpublic.h: H5Z_CLASS_T_VERS */
#define SQY_H5PL_TYPE_FILTER 0           /* H5PLpublic.h: H5PL_TYPE_FILTER */

/* layout of H5Z_class2_t (H5Zpublic.h) */
typedef struct sqy_h5z_class2 {
  int version;                 /* H5Z_CLASS_T_VERS */
  int id;                      /* H5Z_filter_t */
  unsigned encoder_present;
  unsigned decoder_present;
  const char* name;
  void* can_apply;             /* H5Z_can_apply_func_t, NULL like the reference */
  void* set_local;             /* H5Z_set_local_func_t, NULL like the reference */
  size_t (*filter)(unsigned flags, size_t cd_nelmts, const unsigned cd_values[], size_t nbytes, size_t* buf_size, void** buf);
} sqy_h5z_class2;

/* reference: sqeazy_h5_filter.hpp:28-206.
 * write (flags without H5Z_FLAG_REVERSE): cd_values hold the sqeazy header text (pipeline, voxel type, shape) as
 *   hdf5_utils.hpp:728-737 packs it; *buf holds nbytes of raw voxels -> replaced by the blob [header][payload]. A chunk that
 *   already starts with a sqeazy header is passed through (its header size + encoded bytes).
 * read (H5Z_FLAG_REVERSE): *buf holds a blob -> replaced by the raw voxels (size from the blob's header).
 * Returns the number of valid bytes in the new *buf, or 0 on failure (HDF5's convention; *buf is then left alone).
 * The new buffer comes from malloc() and the old one is released with free(), the allocator HDF5 uses for filter
 * buffers (the reference uses new[] / delete[] and notes the mismatch). */
size_t H5Z_filter_sqy(unsigned flags, size_t cd_nelmts, const unsigned cd_values[], size_t nbytes, size_t* buf_size, void** buf);

/* reference: sqeazy_h5_filter.hpp:225-226 */
int H5PLget_plugin_type(void);           /* H5PL_type_t: H5PL_TYPE_FILTER */
const void* H5PLget_plugin_info(void);   /* -> sqy_h5z_class2 (H5Z_class2_t) */

#ifdef __cplusplus
}
#endif
#endif
