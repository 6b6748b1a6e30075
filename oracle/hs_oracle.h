/*
 * hs_oracle.h -- CPU restatement of the covering-index hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for the CUDA path in hyperspace_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may link or call it;
 * the product library (libhs_gpu.so) never does.
 *
 * What it restates (the arithmetic lives in Apache Spark 3.1.1, a `provided` dependency of the
 * reference that is absent from /root/reference -- build.sbt:58-65, project/Dependencies.scala:23-25):
 *   - CoveringIndex.write's repartition(numBuckets, indexedColumns)
 *       (src/main/scala/com/microsoft/hyperspace/index/covering/CoveringIndex.scala:56-71)
 *       == Spark HashPartitioning.partitionIdExpression == pmod(Murmur3Hash(cols, seed 42), n)
 *   - Bucketizer.saveWithBuckets' BucketSpec(n, cols, cols)
 *       (src/main/scala/com/microsoft/hyperspace/index/DataFrameWriterExtensions.scala:50-68)
 *       == FileFormatWriter sorts each task's rows by (bucketId, cols) ascending, nulls first
 *   - JoinIndexRule's bucket-aligned sort-merge join and FilterIndexRule's predicate scan
 *       (index/covering/JoinIndexRule.scala:653-687, FilterIndexRule.scala:135-149) on decoded columns.
 *
 * Parity pin: golden vector src/test/scala/com/microsoft/hyperspace/index/BucketUnionTest.scala:101-123
 * (int keys {2,3} into 10 partitions -> per-partition sums Seq(0,6,0,0,4,0,0,0,0,0)) and Spark's documented
 * hash(1L) = -1712319331; both checked in tests/test_oracle.py.  Page-level Parquet encoding choices,
 * compression and tie order among equal keys are unpinned in the reference.
 */
#ifndef HS_ORACLE_H
#define HS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* physical column types understood by the oracle (same codes as include/hs_gpu.h HS_TYPE_*) */
enum {
  HSO_INT32 = 0,
  HSO_INT64 = 1,
  HSO_FLOAT = 2,
  HSO_DOUBLE = 3,
  HSO_BOOL = 4,   /* one byte per value */
  HSO_STRING = 5  /* data = int64 offsets[n+1], aux = utf8 bytes */
};

typedef struct {
  int32_t type;
  const void* data;       /* n fixed-width values, or n+1 int64 offsets for HSO_STRING */
  const uint8_t* aux;     /* string bytes (HSO_STRING) else NULL */
  const uint8_t* valid;   /* one byte per row, 0 = null; NULL = no nulls */
} hso_column;

/* Spark's org.apache.spark.unsafe.hash.Murmur3_x86_32 */
int32_t hso_hash_int(int32_t v, int32_t seed);
int32_t hso_hash_long(int64_t v, int32_t seed);
int32_t hso_hash_bytes(const uint8_t* p, int32_t len, int32_t seed);

/* Spark Murmur3Hash(children, 42) folded over the key columns of row i (null leaves the hash unchanged) */
int32_t hso_row_hash(const hso_column* keys, int32_t nkeys, int64_t row);

/* bucket id = pmod(row_hash, num_buckets) for every row; nthreads<=1 -> scalar loop */
void hso_bucket_ids(const hso_column* keys, int32_t nkeys, int64_t nrows, int32_t num_buckets,
                    int32_t* out_bucket, int32_t nthreads);

/* Permutation that orders rows by (bucket, key columns ascending nulls-first, original row index).
 * bucket_offsets has num_buckets+1 entries (exclusive prefix of bucket sizes). */
void hso_sort_perm(const hso_column* keys, int32_t nkeys, int64_t nrows, int32_t num_buckets,
                   const int32_t* bucket, int64_t* out_perm, int64_t* bucket_offsets, int32_t nthreads);

/* lower/upper bound of [lo, hi] (inclusive) in a sorted int64 key array: rows [*first, *last) qualify */
void hso_range_select_i64(const int64_t* keys, int64_t n, int64_t lo, int64_t hi, int64_t* first,
                          int64_t* last);

/* inner equi-join of two ascending int64 key arrays; writes up to cap (li, ri) pairs ordered by
 * (left row, right row); returns the total number of pairs (may exceed cap) */
int64_t hso_merge_join_i64(const int64_t* lk, int64_t nl, const int64_t* rk, int64_t nr, int64_t* out_li,
                           int64_t* out_ri, int64_t cap);

/* splitmix64 generator used by the synthetic tables of SURVEY.md section 8d: value i of stream `seed` */
uint64_t hso_splitmix64(uint64_t seed, uint64_t i);

/* Global row numbers of the rows of the synthetic table T (k_i = splitmix64(42, i), i in [first_row, first_row + nrows))
 * whose bucket pmod(hashLong(k_i, 42), num_buckets) == bucket, in source order; writes at most cap of them, returns how
 * many there are. */
int64_t hso_synth_bucket_rows(uint64_t first_row, int64_t nrows, int32_t num_buckets, int32_t bucket, int64_t* out_rows,
                              int64_t cap, int32_t nthreads);

#ifdef __cplusplus
}
#endif
#endif
