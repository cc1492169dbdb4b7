/*
 * rsk_io.h -- C ABI of librsk's SkyH5 (HDF5) loader: the step *before* the hot path.
 *
 * The reference reads and writes its catalogue format through h5py:
 *     SkyModel.read_skyh5    src/pyradiosky/skymodel.py:3491-3766   (helper _get_value_hdf5_group :270-302)
 *     SkyModel.write_skyh5   src/pyradiosky/skymodel.py:4871-5029   (helper _add_value_hdf5_group :229-267)
 * and then hands the host arrays to numpy.  Here the (4, Nfreqs, Ncomponents) Stokes cube and the
 * per-component columns go from the file into HBM without a host-side copy of the whole array: chunks
 * are gathered (and inflated, un-shuffled, converted to double) by a few host threads into a ring of
 * pinned staging slabs, each slab leaving with one asynchronous copy while the next is being filled.
 *
 * libhdf5 / h5py are not used (they are absent from the image); this is a self-contained reader/writer
 * for the part of the HDF5 file format SkyH5 files use:
 *   - superblock version 0/1, 8-byte offsets and lengths, little-endian data
 *   - version 1 object headers (with continuation blocks), symbol-table groups (B-tree v1 + local heap)
 *   - datatypes: fixed-point 1/2/4/8 bytes, IEEE float 4/8 bytes, fixed-length strings, variable-length
 *     strings (global heap), enums over integers (h5py booleans)
 *   - layouts: compact, contiguous, chunked (B-tree v1 index) with the shuffle (2) and deflate (1) filters
 *     (what write_skyh5(data_compression="gzip") produces); other filters give RSK_H5_ERR_UNSUPPORTED
 *   - attributes (message versions 1-3) holding scalars, small arrays or strings
 * Files written by libhdf5 >= 1.10 with libver='latest' (superblock v2/v3, version 2 object headers,
 * fractal-heap groups) are rejected with RSK_H5_ERR_UNSUPPORTED, not mis-read.
 *
 * Conventions: extern "C", plain pointers and sizes; every pointer is a HOST pointer unless marked
 * DEVICE; strings are NUL-terminated UTF-8; 0 = success, negative = RSK_H5_ERR_*, positive =
 * cudaError_t.  A handle may be used from one thread at a time.
 */
#ifndef RSK_IO_H
#define RSK_IO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSK_H5_OK 0
#define RSK_H5_ERR_NULL (-101)        /* required pointer is NULL */
#define RSK_H5_ERR_OPEN (-102)        /* cannot open / map / create the file */
#define RSK_H5_ERR_FORMAT (-103)      /* not HDF5, or structurally corrupt (address out of range, bad signature) */
#define RSK_H5_ERR_UNSUPPORTED (-104) /* valid HDF5 outside the supported subset (see above) */
#define RSK_H5_ERR_NOTFOUND (-105)    /* no such object / attribute */
#define RSK_H5_ERR_TYPE (-106)        /* object is not of the requested kind, or the type cannot be converted */
#define RSK_H5_ERR_SIZE (-107)        /* destination too small / count mismatch */
#define RSK_H5_ERR_FILTER (-108)      /* a filter failed (corrupt deflate stream, checksum...) */
#define RSK_H5_ERR_STATE (-109)       /* writer: object exists already, parent group missing, closed handle */

/* object kinds */
#define RSK_H5_ABSENT 0
#define RSK_H5_GROUP 1
#define RSK_H5_DATASET 2

/* element classes reported by rsk_h5_info / accepted by rsk_h5w_dataset */
#define RSK_H5_INT 0
#define RSK_H5_UINT 1
#define RSK_H5_FLOAT 2
#define RSK_H5_STRING 3  /* fixed length, elem_size bytes, NUL/space padded */
#define RSK_H5_VLEN_STRING 4

/* destination element types of rsk_h5_read_as / rsk_h5_read_device */
#define RSK_H5_AS_F64 0
#define RSK_H5_AS_I64 1

typedef struct RskH5File RskH5File;     /* reader handle */
typedef struct RskH5Writer RskH5Writer; /* writer handle */

typedef struct RskH5Info {
  int32_t kind;       /* RSK_H5_GROUP / RSK_H5_DATASET */
  int32_t type_class; /* RSK_H5_INT ... */
  int32_t elem_size;  /* bytes per element in the file (16 for variable-length strings: the heap reference) */
  int32_t rank;       /* 0 = scalar */
  int64_t dims[8];
  int64_t count;      /* number of elements */
  int32_t layout;     /* 0 compact, 1 contiguous, 2 chunked */
  int32_t nfilters;
  int64_t chunk[8];   /* chunk shape when layout == 2 */
  int64_t nchunks;    /* allocated chunks */
} RskH5Info;

/* ---- reader (replaces h5py.File(filename, "r") at skymodel.py:3525, :3531) ------------------------- */
int rsk_h5_open(const char* path, RskH5File** out);
void rsk_h5_close(RskH5File* f);
/* RSK_H5_ABSENT / RSK_H5_GROUP / RSK_H5_DATASET, or a negative error.  ("/Header" in fileobj, :3526) */
int rsk_h5_kind(RskH5File* f, const char* path);
/* Member names of a group, sorted, '\n'-separated, NUL-terminated.  Returns the number of bytes needed
 * (including the NUL); nothing is written beyond cap.  (for key in header["skycoord"], :3571) */
int64_t rsk_h5_list(RskH5File* f, const char* group, char* buf, int64_t cap);
int rsk_h5_info(RskH5File* f, const char* path, RskH5Info* info);
/* Raw read: the dataset's elements in file byte order (little-endian), C order, into dst.  Strings:
 * fixed-length ones as elem_size-byte records; use rsk_h5_read_strings for either string kind. */
int rsk_h5_read(RskH5File* f, const char* path, void* dst, int64_t dst_bytes, int nthreads);
/* Read with conversion to double (RSK_H5_AS_F64) or int64 (RSK_H5_AS_I64) from any numeric type. */
int rsk_h5_read_as(RskH5File* f, const char* path, int as_type, void* dst, int64_t count, int nthreads);
/* Strings of a (scalar or 1-D) string dataset, each NUL-terminated, concatenated.  Returns bytes needed. */
int64_t rsk_h5_read_strings(RskH5File* f, const char* path, char* buf, int64_t cap);
/* String value of an attribute (dset.attrs["unit"], "object_type", "angtype", :275-283).  Returns the
 * bytes needed (including the NUL), RSK_H5_ERR_NOTFOUND if the attribute does not exist. */
int64_t rsk_h5_attr_string(RskH5File* f, const char* path, const char* attr, char* buf, int64_t cap);
/* Numeric attribute converted to double; returns the element count or a negative error. */
int64_t rsk_h5_attr_f64(RskH5File* f, const char* path, const char* attr, double* dst, int64_t cap);
/* File -> HBM.  dst is a DEVICE buffer of ``count`` doubles (or int64) in the dataset's C order.  Chunks
 * are assembled by ``nthreads`` host threads into pinned slabs of about ``slab_bytes`` (0 = 32 MiB) that
 * are copied asynchronously on ``stream`` while the next slab is filled.  Returns after the last copy has
 * been enqueued *and* the staging memory has been released (i.e. it synchronises the copies, not the
 * stream's earlier work). */
int rsk_h5_read_device(RskH5File* f, const char* path, int as_type, void* dst_device, int64_t count, int nthreads,
                       int64_t slab_bytes, void* stream);
/* Human-readable text for the last error on this handle (never NULL). */
const char* rsk_h5_last_error(RskH5File* f);

/* ---- writer (replaces h5py.File(filename, "w") at skymodel.py:4927) -------------------------------- */
/* Everything is kept in memory and laid out when the file is closed. */
int rsk_h5w_create(const char* path, RskH5Writer** out);
int rsk_h5w_group(RskH5Writer* w, const char* path);
/* data: HOST buffer of count elements (fixed-length strings: count * elem_size bytes).  rank 0 = scalar.
 * chunk == NULL: contiguous.  gzip_level 1-9 with a chunk shape adds shuffle + deflate. */
int rsk_h5w_dataset(RskH5Writer* w, const char* path, int type_class, int elem_size, int rank, const int64_t* dims,
                    const void* data, const int64_t* chunk, int gzip_level);
/* The same from a DEVICE array (HBM -> file): up to three dimensions; the array is fetched in slabs (one chunk-row
 * along the first axis times a run of chunk-rows along the second) through two pinned staging buffers, each pitched
 * device->host copy on ``stream`` overlapping the chunking / deflating / writing of the previous slab.  Returns when
 * the data has been written. */
int rsk_h5w_dataset_device(RskH5Writer* w, const char* path, int type_class, int elem_size, int rank,
                           const int64_t* dims, const void* device_data, const int64_t* chunk, int gzip_level,
                           void* stream);
int rsk_h5w_attr_string(RskH5Writer* w, const char* path, const char* name, const char* value);
/* Writes the file; the handle is freed whatever the result. */
int rsk_h5w_close(RskH5Writer* w);

const char* rsk_h5_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif /* RSK_IO_H */
