/*
 * hs_gpu.h -- C ABI of the B200-native covering-index engine (libhs_gpu.so).
 *
 * The reference (microsoft/hyperspace, Scala on the JVM) has no FFI boundary: its whole data path is three
 * Spark calls.  This header is the seam a maintainer binds with JNI / ctypes; each entry
 * point names the reference interface whose body it replaces.  Paths below are relative to
 * src/main/scala/com/microsoft/hyperspace/ in the reference.
 *
 * Conventions: plain pointers and sizes only; strings are UTF-8, caller-owned, copied before return; every
 * function returns 0 on success or a negative HS_E* code and writes a message into (err, errlen) when non-NULL.
 * The library owns all device memory and every handle it returns until the matching *_free call.  There is no
 * CPU fallback: without a CUDA device hs_init fails with HS_ENODEVICE and nothing else can be called.
 * One hs_ctx drives one GPU: a main stream for the kernels plus one copy stream per direction for the staged / asynchronous
 * entry points; use one ctx per host thread / per rank.
 *
 * Environment switches (diagnostics and A/B measurements only; results are identical either way):
 *   HS_EXCHANGE=nccl   multi-GPU: NCCL all-to-all instead of the fused partition + NVLink peer stores
 *   HS_NO_CARRY=1      decode dictionary-encoded included columns to values instead of carrying 16-bit codes
 *                      (on several GPUs all ranks must agree)
 *   HS_FULL_SORT=1     radix-sort every varying key byte instead of the high bytes + tie fix-up
 *   HS_PART_REHASH=1   partition kernel hashes the keys again instead of reading the stored bucket ids
 *   HS_NO_ZEROCOPY=1   decode aligned PLAIN pages into column arrays instead of reading them in place
 *   HS_PART_BULK=0|1   partition kernel: per-thread stores (0) or cp.async.bulk stores of whole runs (1); default: bulk only for
 *                      runs that leave over NVLink
 *   HS_PEER_TILE=small multi-GPU: the 4096-row partition tile of the single-GPU path instead of the 8192-row one
 *   HS_DEBUG_LOCAL_PEERS=1  multi-GPU: every rank keeps its rows (peer stores go to local memory; wrong results, isolates
 *                      the NVLink share of the exchange time) -- the one switch that changes results
 *   HS_IO_THREADS=n    host threads that read source files / write bucket files (default 16, at most half the cores)
 *   HS_TIMELINE=1      print the event timeline (H2D / build / D2H begin and end) of every staged or asynchronous call
 */
#ifndef HS_GPU_H
#define HS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_ABI_VERSION 1

/* error codes */
#define HS_OK 0
#define HS_EINVAL (-1)       /* bad argument / unknown column / unsupported configuration */
#define HS_ENODEVICE (-2)    /* no CUDA device or CUDA runtime failure at init */
#define HS_ECUDA (-3)        /* CUDA error during execution */
#define HS_EFORMAT (-4)      /* malformed or unsupported Parquet input */
#define HS_EIO (-5)          /* file system error */
#define HS_EUNSUPPORTED (-6) /* valid input the GPU path does not handle yet (no CPU fallback exists) */
#define HS_ENOMEM (-7)
#define HS_ECOMM (-8)        /* NCCL / multi-GPU exchange failure */

/* physical column types (Parquet physical type x Spark SQL type actually supported on the GPU path) */
#define HS_TYPE_INT32 0  /* Spark int / date            */
#define HS_TYPE_INT64 1  /* Spark long / timestamp(us)  */
#define HS_TYPE_FLOAT 2
#define HS_TYPE_DOUBLE 3
#define HS_TYPE_BOOL 4
#define HS_TYPE_STRING 5 /* BYTE_ARRAY (Spark string / binary): write path (keys and included columns, one GPU); the index
                            scans and joins do not read it yet -> HS_EUNSUPPORTED */

typedef struct hs_ctx hs_ctx;
typedef struct hs_index_result hs_index_result;
typedef struct hs_batch hs_batch;

/* ------------------------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------------------------- */

/* ABI version of the loaded library (== HS_ABI_VERSION it was built with). */
int hs_abi_version(void);
/* Static description: compile arch, CUDA version, build flags. */
const char* hs_build_info(void);

/* Create a context on CUDA device `device_id`.  `cuda_stream` may be NULL (the library creates its own
 * non-blocking stream) or a cudaStream_t owned by the caller (e.g. torch's current stream) on which every kernel
 * and copy of this ctx is then issued.  Replaces: the Spark executor pool a Hyperspace action runs on
 * (actions/Action.scala:84-105 runs op() on the driver thread; Spark schedules the tasks). */
int hs_init(int device_id, void* cuda_stream, hs_ctx** out, char* err, size_t errlen);
void hs_shutdown(hs_ctx* ctx);
/* Return cached device/pinned buffers to the driver. */
void hs_trim(hs_ctx* ctx);
/* Per-kernel timing: when enabled, the hot kernels are bracketed with CUDA events on the ctx stream; hs_profile_report
 * writes a JSON object {"kernel": {"launches": n, "ms": total}} covering the calls since the last report and resets
 * it.  Used by bench.py for the roofline of the dominant kernel. */
void hs_profile_enable(hs_ctx* ctx, int on);
int hs_profile_report(hs_ctx* ctx, char* out_json, size_t outlen);
/* Page-locked host memory for file images handed to hs_create_index (a JNI direct ByteBuffer can wrap it); pageable
 * memory works too but copies at a fraction of the PCIe rate. */
void* hs_host_alloc(hs_ctx* ctx, size_t bytes);
void hs_host_free(hs_ctx* ctx, void* p);

/* Multi-GPU (one process per GPU).  Rank 0 calls hs_comm_unique_id and ships the 128 bytes to the other ranks
 * out of band; every rank then calls hs_comm_init.  Replaces: Spark's shuffle service behind
 * `indexData.repartition(numBuckets, indexedColumns)` (index/covering/CoveringIndex.scala:60). */
int hs_comm_unique_id(void* out_id128, char* err, size_t errlen);
int hs_comm_init(hs_ctx* ctx, int rank, int world_size, const void* id128, char* err, size_t errlen);

/* ------------------------------------------------------------------------------------------------------------
 * Write side: createIndex / refreshIndex(full|incremental) / optimizeIndex data path
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct {
  const char* path;  /* read from the file system when data == NULL */
  const void* data;  /* whole Parquet file image; host memory, or device memory when on_device != 0
                        (device images must be 16-byte aligned) */
  uint64_t size;     /* image size in bytes (ignored when data == NULL) */
  int64_t file_id;   /* lineage id from FileIdTracker (index/IndexLogEntry.scala:627-703); -1 when lineage is off */
  int32_t on_device;
  int32_t reserved;
} hs_source_file;

#define HS_SAVE_OVERWRITE 0 /* create / refresh full / optimize  (covering/CoveringIndexTrait.scala:45-47) */
#define HS_SAVE_APPEND 1    /* incremental refresh into an existing version dir (CoveringIndexTrait.scala:87-93) */

#define HS_CODEC_UNCOMPRESSED 0
#define HS_CODEC_SNAPPY 1

#define HS_OUT_FILES 0  /* write <out_dir>/part-<bbbbb>-<uuid>_<bbbbb>.c000.parquet (the reference's effect) */
#define HS_OUT_HOST 1   /* keep the bucket file images in pinned host memory owned by the result handle */
#define HS_OUT_DEVICE 2 /* keep the bucket file images in device memory owned by the result handle */

typedef struct {
  const hs_source_file* files; /* this rank's share of the source files (all of them on one GPU) */
  int32_t n_files;
  const char* const* indexed_columns; /* resolved names, in IndexConfig order (CoveringIndex.scala:34) */
  int32_t n_indexed;
  const char* const* included_columns; /* CoveringIndex.scala:35 */
  int32_t n_included;
  int32_t num_buckets; /* spark.hyperspace.index.numBuckets, frozen in the log entry (CoveringIndex.scala:37) */
  int32_t save_mode;   /* HS_SAVE_* */
  int32_t output;      /* HS_OUT_* */
  int32_t lineage;     /* != 0: append `_data_file_id` (int64) from hs_source_file.file_id (CoveringIndex.scala:152-186) */
  const char* out_dir; /* ctx.indexDataPath = <index>/v__=<N> (actions/CreateActionBase.scala:32-37); HS_OUT_FILES only */
  const char* job_uuid; /* the <uuid> of the part file names; NULL -> generated */
  int64_t rows_per_page;      /* 0 -> 131072 */
  int64_t rows_per_row_group; /* 0 -> 4194304 */
  /* rows to drop: lineage ids whose rows must not reach the output (refreshIncremental's deleted files,
   * CoveringIndexTrait.scala:78-94).  Requires a `_data_file_id` column in the source files. */
  const int64_t* deleted_file_ids;
  int32_t n_deleted_file_ids;
  int32_t disable_dictionary; /* 0 (default): dictionary-encode columns whose distinct values fit a dictionary page, as
                                 parquet-mr does; != 0: PLAIN only */
  int32_t compression;        /* HS_CODEC_*: codec of the index pages.  HS_CODEC_SNAPPY is what Spark writes by default
                                 (files are then named ...c000.snappy.parquet, T/index/VacuumOutdatedActionTest.scala:67) */
  int32_t reserved;
} hs_index_spec;

typedef struct {
  int64_t rows_in;       /* rows decoded on this rank */
  int64_t rows_out;      /* rows written by this rank (after the exchange) */
  int64_t bytes_in;      /* encoded source bytes consumed */
  int64_t bytes_out;     /* encoded index bytes produced */
  int64_t bytes_exchanged; /* bytes this rank sent through the all-to-all */
  int32_t files_out;
  int32_t gpu_launches;  /* kernels launched by this call */
  float ms_total;        /* CUDA-event time of the whole call on the ctx stream */
  float ms_h2d, ms_plan, ms_decode, ms_hash, ms_partition, ms_exchange, ms_sort, ms_gather, ms_encode, ms_d2h, ms_write;
} hs_stats;

/* Body of CoveringIndex.write(ctx, indexData, mode) (index/covering/CoveringIndex.scala:56-71): scan the source
 * Parquet, project indexed ++ included columns, bucket = pmod(murmur3(keys, 42), num_buckets), sort each bucket
 * ascending nulls-first on the indexed columns, and emit one Parquet file per non-empty bucket
 * (index/DataFrameWriterExtensions.scala:50-68).  All-or-nothing: on failure nothing is left in out_dir. */
int hs_create_index(hs_ctx* ctx, const hs_index_spec* spec, hs_index_result** out, hs_stats* stats, char* err,
                    size_t errlen);

/* ---- software pipelining across calls -------------------------------------------------------------------------------
 * One createIndex cannot overlap its own input and output copies: every index file depends on every source file (the hash
 * repartition), so the device->host drain of a call starts after its last source byte has arrived.  What CAN overlap are
 * the copies of NEIGHBOURING calls, in both directions at once (PCIe is full duplex) -- the overlap Spark gets from running
 * the scan tasks of one job beside the write tasks of another (index/DataFrameWriterExtensions.scala:76-80 hands the write
 * to Spark's task scheduler).  Three steps, each on its own CUDA stream of the ctx:
 *
 *   hs_stage_sources       host (or file system) Parquet images -> device, asynchronously on the H2D copy stream; footers
 *                          are parsed from host memory on the way.  The handle's descriptors (on_device = 1) are valid
 *                          inputs of hs_create_index[_async], hs_filter_scan and hs_bucket_join, which wait for the copy.
 *                          The caller's host images must stay valid until hs_staged_wait / hs_staged_free returns or a
 *                          call that consumed the descriptors has returned.
 *   hs_create_index_async  everything hs_create_index does up to the encoded index files in device memory (returns when
 *                          the kernels have run), then starts the device->host copy on the D2H copy stream.
 *   hs_pending_wait        waits for that copy (and writes the files for HS_OUT_FILES); yields the result + stats.
 *                          Consumes the handle, also on failure.
 *
 *   staged[i+1] = hs_stage_sources(...);  pending[i] = hs_create_index_async(staged[i]...);  hs_pending_wait(pending[i-1])
 *
 * keeps the H2D engine, the SMs and the D2H engine busy at the same time.  hs_create_index == async + wait. */
typedef struct hs_staged hs_staged;
typedef struct hs_pending hs_pending;
int hs_stage_sources(hs_ctx* ctx, const hs_source_file* files, int32_t n_files, hs_staged** out, char* err, size_t errlen);
int32_t hs_staged_num_files(const hs_staged* s);
int hs_staged_file(const hs_staged* s, int32_t i, hs_source_file* out); /* out->path points into the handle */
int hs_staged_wait(hs_staged* s, float* ms_copy); /* blocks until the copies have completed; ms_copy (optional): their
                                                      duration on the H2D stream */
void hs_staged_free(hs_staged* s); /* waits for the copies and for the ctx stream, then releases the device images */
int hs_create_index_async(hs_ctx* ctx, const hs_index_spec* spec, hs_pending** out, char* err, size_t errlen);
int hs_pending_wait(hs_pending* p, hs_index_result** out, hs_stats* stats, char* err, size_t errlen);
void hs_pending_cancel(hs_pending* p); /* waits for the copy and drops the result */

int32_t hs_result_num_files(const hs_index_result* r);
/* File i of the result: bucket id, file name (no directory), image pointer (host or device, NULL for
 * HS_OUT_FILES), image size, row count. */
int hs_result_file(const hs_index_result* r, int32_t i, int32_t* bucket, const char** name, const void** data,
                   uint64_t* size, int64_t* rows);
void hs_result_free(hs_index_result* r);

/* ------------------------------------------------------------------------------------------------------------
 * Read side: FilterIndexRule scan, JoinIndexRule bucket-aligned sort-merge join
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct {
  const hs_source_file* files; /* index (or source) Parquet files */
  int32_t n_files;
  int32_t sorted_on_key;       /* != 0: every file is sorted ascending on key_column (index files) -> binary search;
                                  0: full predicate scan (appended source files under Hybrid Scan) */
  const char* key_column;      /* predicate column (first indexed column, covering/FilterIndexRule.scala:33-103) */
  const char* const* projected_columns;
  int32_t n_projected;
  int32_t has_lo, has_hi;      /* inclusive bounds lo <= key <= hi on an integer key */
  int64_t lo, hi;
  const int64_t* deleted_file_ids; /* Hybrid Scan: NOT (_data_file_id IN ids) (covering/CoveringIndexRuleUtils.scala:244-253) */
  int32_t n_deleted_file_ids;
  int32_t output;              /* HS_OUT_HOST (or 0): result columns in pinned host memory; HS_OUT_DEVICE: left in device
                                  memory for the next GPU operator (hs_batch_column then yields device pointers) */
  /* string / binary key column (HS_TYPE_STRING): the inclusive bounds as bytes, compared in UTF8String byte order (unsigned
   * bytes, the shorter value first on a common prefix); has_lo / has_hi say which are set, lo / hi are ignored.  Equality
   * -- the predicate of the reference's own filter-rule tests, `c3 == "facebook"` (T/index/E2EHyperspaceRulesTest.scala) --
   * is lo == hi. */
  const void* lo_bytes;
  const void* hi_bytes;
  uint32_t lo_len, hi_len;
} hs_scan_spec;

/* Executes the scan FilterIndexRule.applyIndex substitutes for the source scan
 * (index/covering/FilterIndexRule.scala:135-149; CoveringIndexRuleUtils.scala:98-130). */
int hs_filter_scan(hs_ctx* ctx, const hs_scan_spec* spec, hs_batch** out, hs_stats* stats, char* err, size_t errlen);

typedef struct {
  const hs_source_file* left_files;  /* index files of the left side, any order; bucket id parsed from the name */
  int32_t n_left;
  const hs_source_file* right_files;
  int32_t n_right;
  const int32_t* left_buckets;       /* bucket id per left file  (BucketingUtils.getBucketId on the file name) */
  const int32_t* right_buckets;
  int32_t num_buckets;
  int32_t output;                    /* HS_OUT_HOST (or 0) / HS_OUT_DEVICE, as in hs_scan_spec */
  const char* left_key;              /* single join key on each side: int32 / int64 / string, the same type on both */
  const char* right_key;
  const char* const* left_columns;   /* projected from the left side */
  int32_t n_left_columns;
  const char* const* right_columns;
  int32_t n_right_columns;
} hs_join_spec;

/* Inner equi-join bucket b of the left index with bucket b of the right index, no exchange -- what Spark plans
 * after JoinIndexRule.applyIndex (index/covering/JoinIndexRule.scala:653-687; T/index/E2EHyperspaceRulesTest.scala:487-512).
 * Buckets holding several files (after an incremental refresh) are merged first, as Spark's SortExec would. */
int hs_bucket_join(hs_ctx* ctx, const hs_join_spec* spec, hs_batch** out, hs_stats* stats, char* err, size_t errlen);

int64_t hs_batch_num_rows(const hs_batch* b);
int32_t hs_batch_on_device(const hs_batch* b); /* != 0: the column pointers are device pointers (output = HS_OUT_DEVICE) */
int32_t hs_batch_num_columns(const hs_batch* b);
/* Column i: name, HS_TYPE_*, pointer to num_rows values, pointer to one validity byte per row (NULL when the column
 * has no nulls); host pointers unless hs_batch_on_device. */
int hs_batch_column(const hs_batch* b, int32_t i, const char** name, int32_t* type, const void** data,
                    const uint8_t** valid);
/* A HS_TYPE_STRING column: `data` of hs_batch_column points to the values' bytes back to back, and value r occupies
 * bytes [offsets[r], offsets[r + 1]) (num_rows + 1 offsets; a null has length 0).  HS_EINVAL for other columns. */
int hs_batch_string_offsets(const hs_batch* b, int32_t i, const uint64_t** offsets, uint64_t* total_bytes);
void hs_batch_free(hs_batch* b);

/* ------------------------------------------------------------------------------------------------------------
 * Verification of a written index (any size; bench.py checks the 1 B-row benchmark output with it on every run)
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct {
  int64_t rows;               /* rows found in the files */
  int64_t bucket_mismatches;  /* rows whose pmod(murmur3(indexed columns, 42), num_buckets) != bucket id of their file */
  int64_t order_violations;   /* adjacent rows of one file whose indexed columns are not ascending, nulls first */
  uint64_t row_checksum;      /* sum over rows, mod 2^64, of a 64-bit mix of all the row's values (indexed ++ included order):
                                 independent of row order; changes when a value moves to another row */
  uint64_t column_checksum[16]; /* the same per column */
  int32_t n_columns;
  int32_t reserved;
} hs_verify_report;

/* The three properties the reference's write-path test pins (T/index/DataFrameWriterExtensionsTest.scala:93-158: bucket id
 * of every row == HashPartitioning's, every file sorted on the indexed columns, row multiset preserved), evaluated on the
 * GPU over whole index files.  buckets[i] is the bucket id of files[i] (BucketingUtils.getBucketId of its name). */
int hs_verify_index(hs_ctx* ctx, const hs_source_file* files, const int32_t* buckets, int32_t n_files,
                    const char* const* indexed_columns, int32_t n_indexed, const char* const* included_columns,
                    int32_t n_included, int32_t num_buckets, hs_verify_report* out, char* err, size_t errlen);
/* rows / row_checksum / column_checksum of rows [first_row, first_row + nrows) of the synthetic table T as generated
 * (never encoded): what hs_verify_index must report for an index built over those rows. */
int hs_synth_checksum(hs_ctx* ctx, int64_t first_row, int64_t nrows, int32_t ncols, hs_verify_report* out, char* err,
                      size_t errlen);

/* ------------------------------------------------------------------------------------------------------------
 * Kernel-level entry points (host arrays in, host arrays out).  They exist so the parity tests can pin each
 * kernel against the oracle in isolation; the JVM binding does not need them.
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct {
  int32_t type;         /* HS_TYPE_* */
  int32_t reserved;
  const void* data;     /* n values */
  const uint8_t* valid; /* one byte per row or NULL */
} hs_host_column;

/* K2: bucket id per row (int32) and the num_buckets-bin histogram (int64) of Spark's HashPartitioning. */
int hs_k_bucket_ids(hs_ctx* ctx, const hs_host_column* keys, int32_t nkeys, int64_t nrows, int32_t num_buckets,
                    int32_t* out_bucket, int64_t* out_hist, char* err, size_t errlen);
/* K3+K4: permutation ordering rows by (bucket, keys ascending nulls-first); bucket_offsets has num_buckets+1 entries. */
int hs_k_sort_perm(hs_ctx* ctx, const hs_host_column* keys, int32_t nkeys, int64_t nrows, int32_t num_buckets,
                   int64_t* out_perm, int64_t* out_bucket_offsets, char* err, size_t errlen);

/* Synthetic table T of the benchmark: rows [first_row, first_row+nrows) of
 * (k:int64, v1:int64, v2:float64, v3:int32, v4:float32)[:ncols], generated and Parquet-encoded on the GPU into
 * n_files file images of row_groups_per_file row groups each (HS_OUT_HOST or HS_OUT_DEVICE).  dictionary != 0: low-cardinality
 * columns (v1, v3, v4) are PLAIN_DICTIONARY-encoded like a parquet-mr / pyarrow writer would; 0: PLAIN only. */
int hs_synth_table(hs_ctx* ctx, int64_t first_row, int64_t nrows, int32_t ncols, int32_t n_files,
                   int32_t row_groups_per_file, int32_t dictionary, int32_t output, hs_index_result** out, char* err,
                   size_t errlen);
/* the same with a page codec (HS_CODEC_*): the SNAPPY variant of T that SURVEY.md 8d asks for */
int hs_synth_table_ex(hs_ctx* ctx, int64_t first_row, int64_t nrows, int32_t ncols, int32_t n_files,
                      int32_t row_groups_per_file, int32_t dictionary, int32_t compression, int32_t output,
                      hs_index_result** out, char* err, size_t errlen);
/* Snappy-compresses n bytes of host memory on the GPU (the page compressor, fragment by fragment) into out (capacity cap);
 * kernel-level entry point for the parity tests: any Snappy decoder must give the input back. */
int hs_k_snappy_compress(hs_ctx* ctx, const void* in, uint64_t n, void* out, uint64_t cap, uint64_t* out_len, char* err,
                         size_t errlen);
/* The page decompressor on one raw Snappy stream of n bytes whose uncompressed length is out_len (Parquet's page header
 * carries it); *sequential = 1 when the stream's 64 KB blocks were not independent and one warp decoded it front to back.
 * HS_EFORMAT for a damaged stream.  Kernel-level entry point for the parity tests. */
int hs_k_snappy_decompress(hs_ctx* ctx, const void* in, uint64_t n, void* out, uint64_t out_len, int32_t* sequential, char* err,
                           size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* HS_GPU_H */
