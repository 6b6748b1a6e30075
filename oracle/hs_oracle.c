/*
 * hs_oracle.c -- CPU restatement of the covering-index hot path.  TEST INFRASTRUCTURE ONLY
 * (see hs_oracle.h for the scope statement and the reference citations).
 *
 * Spark-side algorithms restated here (Spark 3.1.1, not vendored in /root/reference):
 *   org.apache.spark.unsafe.hash.Murmur3_x86_32      hashInt / hashLong / hashUnsafeBytes (per-byte tail)
 *   org.apache.spark.sql.catalyst.expressions.Murmur3Hash / HashExpression   seed 42, fold left, null skips
 *   org.apache.spark.sql.catalyst.expressions.Pmod                              ((h % n) + n) % n
 *   org.apache.spark.sql.execution.SortExec via FileFormatWriter               ascending, nulls first
 * Reference call sites: CoveringIndex.scala:60 (repartition), DataFrameWriterExtensions.scala:64
 * (BucketSpec(n, cols, cols)), JoinIndexRule.scala:653-687, FilterIndexRule.scala:135-149.
 */
#define _GNU_SOURCE
#include "hs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* tiny pthread work-sharing helper (libgomp is not in this image): items [0,n) are handed out
 * dynamically in chunks to nthreads workers */
typedef void (*hso_body)(int64_t lo, int64_t hi, void* arg);
typedef struct {
  hso_body body;
  void* arg;
  int64_t n, chunk;
  int64_t next; /* guarded by mu */
  pthread_mutex_t mu;
} hso_pool;

static void* hso_worker(void* vp) {
  hso_pool* p = (hso_pool*)vp;
  for (;;) {
    pthread_mutex_lock(&p->mu);
    int64_t lo = p->next;
    p->next += p->chunk;
    pthread_mutex_unlock(&p->mu);
    if (lo >= p->n) break;
    int64_t hi = lo + p->chunk < p->n ? lo + p->chunk : p->n;
    p->body(lo, hi, p->arg);
  }
  return NULL;
}

static void hso_parallel_for(int64_t n, int64_t chunk, int32_t nthreads, hso_body body, void* arg) {
  if (nthreads <= 1 || n <= chunk) {
    body(0, n, arg);
    return;
  }
  hso_pool p = {body, arg, n, chunk, 0, PTHREAD_MUTEX_INITIALIZER};
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 0; t < nthreads - 1; t++)
    if (pthread_create(&th[started], NULL, hso_worker, &p) == 0) started++;
  hso_worker(&p);
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static inline uint32_t mix_k1(uint32_t k1) {
  k1 *= 0xcc9e2d51u;
  k1 = rotl32(k1, 15);
  k1 *= 0x1b873593u;
  return k1;
}

static inline uint32_t mix_h1(uint32_t h1, uint32_t k1) {
  h1 ^= k1;
  h1 = rotl32(h1, 13);
  h1 = h1 * 5u + 0xe6546b64u;
  return h1;
}

static inline uint32_t fmix(uint32_t h1, uint32_t len) {
  h1 ^= len;
  h1 ^= h1 >> 16;
  h1 *= 0x85ebca6bu;
  h1 ^= h1 >> 13;
  h1 *= 0xc2b2ae35u;
  h1 ^= h1 >> 16;
  return h1;
}

int32_t hso_hash_int(int32_t v, int32_t seed) {
  return (int32_t)fmix(mix_h1((uint32_t)seed, mix_k1((uint32_t)v)), 4);
}

int32_t hso_hash_long(int64_t v, int32_t seed) {
  uint32_t lo = (uint32_t)((uint64_t)v & 0xffffffffu);
  uint32_t hi = (uint32_t)((uint64_t)v >> 32);
  uint32_t h1 = mix_h1((uint32_t)seed, mix_k1(lo));
  h1 = mix_h1(h1, mix_k1(hi));
  return (int32_t)fmix(h1, 8);
}

/* Murmur3_x86_32.hashUnsafeBytes: 4-byte little-endian words, then every tail byte mixed on its own
 * as a sign-extended int (Spark's legacy, non-standard tail). */
int32_t hso_hash_bytes(const uint8_t* p, int32_t len, int32_t seed) {
  uint32_t h1 = (uint32_t)seed;
  int32_t aligned = len - (len % 4);
  for (int32_t i = 0; i < aligned; i += 4) {
    uint32_t w = (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) |
                 ((uint32_t)p[i + 3] << 24);
    h1 = mix_h1(h1, mix_k1(w));
  }
  for (int32_t i = aligned; i < len; i++) {
    int32_t b = (int8_t)p[i];
    h1 = mix_h1(h1, mix_k1((uint32_t)b));
  }
  return (int32_t)fmix(h1, (uint32_t)len);
}

static inline int32_t hash_value(const hso_column* c, int64_t row, int32_t seed) {
  if (c->valid && !c->valid[row]) return seed; /* null: hash unchanged */
  switch (c->type) {
    case HSO_INT32:
      return hso_hash_int(((const int32_t*)c->data)[row], seed);
    case HSO_INT64:
      return hso_hash_long(((const int64_t*)c->data)[row], seed);
    case HSO_FLOAT: {
      float f = ((const float*)c->data)[row];
      uint32_t bits;
      if (f == 0.0f) f = 0.0f;              /* -0.0f -> +0.0f */
      if (f != f) bits = 0x7fc00000u;       /* Float.floatToIntBits canonical NaN */
      else memcpy(&bits, &f, 4);
      return hso_hash_int((int32_t)bits, seed);
    }
    case HSO_DOUBLE: {
      double d = ((const double*)c->data)[row];
      uint64_t bits;
      if (d == 0.0) d = 0.0;
      if (d != d) bits = 0x7ff8000000000000ull;
      else memcpy(&bits, &d, 8);
      return hso_hash_long((int64_t)bits, seed);
    }
    case HSO_BOOL:
      return hso_hash_int(((const uint8_t*)c->data)[row] ? 1 : 0, seed);
    case HSO_STRING: {
      const int64_t* off = (const int64_t*)c->data;
      return hso_hash_bytes(c->aux + off[row], (int32_t)(off[row + 1] - off[row]), seed);
    }
  }
  return seed;
}

int32_t hso_row_hash(const hso_column* keys, int32_t nkeys, int64_t row) {
  int32_t h = 42;
  for (int32_t k = 0; k < nkeys; k++) h = hash_value(&keys[k], row, h);
  return h;
}

static inline int32_t pmod(int32_t h, int32_t n) {
  int32_t r = h % n;
  return r < 0 ? r + n : r;
}

typedef struct {
  const hso_column* keys;
  int32_t nkeys, num_buckets;
  int32_t* out;
} bucket_args;

static void bucket_body(int64_t lo, int64_t hi, void* va) {
  bucket_args* a = (bucket_args*)va;
  for (int64_t i = lo; i < hi; i++) a->out[i] = pmod(hso_row_hash(a->keys, a->nkeys, i), a->num_buckets);
}

void hso_bucket_ids(const hso_column* keys, int32_t nkeys, int64_t nrows, int32_t num_buckets,
                    int32_t* out_bucket, int32_t nthreads) {
  bucket_args a = {keys, nkeys, num_buckets, out_bucket};
  hso_parallel_for(nrows, 1 << 20, nthreads, bucket_body, &a);
}

/* ---- ordering ----------------------------------------------------------------------------- */

static inline int cmp_value(const hso_column* c, int64_t a, int64_t b) {
  int va = c->valid ? c->valid[a] != 0 : 1, vb = c->valid ? c->valid[b] != 0 : 1;
  if (!va || !vb) return va - vb; /* nulls first */
  switch (c->type) {
    case HSO_INT32: {
      int32_t x = ((const int32_t*)c->data)[a], y = ((const int32_t*)c->data)[b];
      return (x > y) - (x < y);
    }
    case HSO_INT64: {
      int64_t x = ((const int64_t*)c->data)[a], y = ((const int64_t*)c->data)[b];
      return (x > y) - (x < y);
    }
    case HSO_FLOAT: { /* SQLOrderingUtil.compareFloats: -0.0 == 0.0, NaN greatest */
      float x = ((const float*)c->data)[a], y = ((const float*)c->data)[b];
      if (x == y) return 0;
      int nx = x != x, ny = y != y;
      if (nx || ny) return nx - ny;
      return (x > y) - (x < y);
    }
    case HSO_DOUBLE: {
      double x = ((const double*)c->data)[a], y = ((const double*)c->data)[b];
      if (x == y) return 0;
      int nx = x != x, ny = y != y;
      if (nx || ny) return nx - ny;
      return (x > y) - (x < y);
    }
    case HSO_BOOL: {
      int x = ((const uint8_t*)c->data)[a] != 0, y = ((const uint8_t*)c->data)[b] != 0;
      return x - y;
    }
    case HSO_STRING: { /* UTF8String.compareTo: unsigned byte-wise, shorter prefix first */
      const int64_t* off = (const int64_t*)c->data;
      int64_t la = off[a + 1] - off[a], lb = off[b + 1] - off[b];
      int r = memcmp(c->aux + off[a], c->aux + off[b], (size_t)(la < lb ? la : lb));
      if (r) return r < 0 ? -1 : 1;
      return (la > lb) - (la < lb);
    }
  }
  return 0;
}

typedef struct {
  const hso_column* keys;
  int32_t nkeys;
} cmp_ctx;

static int cmp_rows(const void* pa, const void* pb, void* vctx) {
  const cmp_ctx* ctx = (const cmp_ctx*)vctx;
  int64_t a = *(const int64_t*)pa, b = *(const int64_t*)pb;
  for (int32_t k = 0; k < ctx->nkeys; k++) {
    int r = cmp_value(&ctx->keys[k], a, b);
    if (r) return r;
  }
  return (a > b) - (a < b); /* deterministic tie order: original row index */
}

/* fast path: one non-null int64 key -> stable LSD radix sort of (key^sign, row) pairs */
static void radix_sort_i64(const int64_t* keycol, int64_t* perm, int64_t n) {
  if (n < 2) return;
  uint64_t* ka = (uint64_t*)malloc((size_t)n * 8);
  uint64_t* kb = (uint64_t*)malloc((size_t)n * 8);
  int64_t* pb = (int64_t*)malloc((size_t)n * 8);
  int64_t* pa = perm;
  for (int64_t i = 0; i < n; i++) ka[i] = (uint64_t)keycol[perm[i]] ^ 0x8000000000000000ull;
  for (int pass = 0; pass < 8; pass++) {
    size_t cnt[256] = {0};
    int sh = pass * 8;
    for (int64_t i = 0; i < n; i++) cnt[(ka[i] >> sh) & 255]++;
    int skip = 0;
    for (int d = 0; d < 256; d++)
      if (cnt[d] == (size_t)n) skip = 1;
    if (skip) continue;
    size_t sum = 0;
    for (int d = 0; d < 256; d++) {
      size_t c = cnt[d];
      cnt[d] = sum;
      sum += c;
    }
    for (int64_t i = 0; i < n; i++) {
      size_t d = cnt[(ka[i] >> sh) & 255]++;
      kb[d] = ka[i];
      pb[d] = pa[i];
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    int64_t* tp = pa; pa = pb; pb = tp;
  }
  if (pa != perm) {
    memcpy(perm, pa, (size_t)n * 8);
    free(pa);
  } else {
    free(pb);
  }
  free(ka);
  free(kb);
}

typedef struct {
  const hso_column* keys;
  int32_t nkeys;
  int fast;
  int64_t* perm;
  const int64_t* offsets;
} sort_args;

static void sort_body(int64_t blo, int64_t bhi, void* va) {
  sort_args* a = (sort_args*)va;
  cmp_ctx ctx = {a->keys, a->nkeys};
  for (int64_t b = blo; b < bhi; b++) {
    int64_t lo = a->offsets[b], n = a->offsets[b + 1] - lo;
    if (a->fast) radix_sort_i64((const int64_t*)a->keys[0].data, a->perm + lo, n);
    else qsort_r(a->perm + lo, (size_t)n, sizeof(int64_t), cmp_rows, &ctx);
  }
}

void hso_sort_perm(const hso_column* keys, int32_t nkeys, int64_t nrows, int32_t num_buckets,
                   const int32_t* bucket, int64_t* out_perm, int64_t* bucket_offsets, int32_t nthreads) {
  /* stable counting sort by bucket id */
  memset(bucket_offsets, 0, sizeof(int64_t) * (size_t)(num_buckets + 1));
  for (int64_t i = 0; i < nrows; i++) bucket_offsets[bucket[i] + 1]++;
  for (int32_t b = 0; b < num_buckets; b++) bucket_offsets[b + 1] += bucket_offsets[b];
  int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)num_buckets);
  memcpy(cursor, bucket_offsets, sizeof(int64_t) * (size_t)num_buckets);
  for (int64_t i = 0; i < nrows; i++) out_perm[cursor[bucket[i]]++] = i;
  free(cursor);

  int fast = nkeys == 1 && keys[0].type == HSO_INT64 && keys[0].valid == NULL;
  sort_args a = {keys, nkeys, fast, out_perm, bucket_offsets};
  hso_parallel_for(num_buckets, 1, nthreads, sort_body, &a);
}

/* ---- read side ----------------------------------------------------------------------------- */

static int64_t lower_bound_i64(const int64_t* a, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = lo + (hi - lo) / 2;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static int64_t upper_bound_i64(const int64_t* a, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = lo + (hi - lo) / 2;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

void hso_range_select_i64(const int64_t* keys, int64_t n, int64_t lo, int64_t hi, int64_t* first,
                          int64_t* last) {
  *first = lower_bound_i64(keys, n, lo);
  *last = upper_bound_i64(keys, n, hi);
  if (*last < *first) *last = *first;
}

int64_t hso_merge_join_i64(const int64_t* lk, int64_t nl, const int64_t* rk, int64_t nr, int64_t* out_li,
                           int64_t* out_ri, int64_t cap) {
  int64_t total = 0, j = 0;
  for (int64_t i = 0; i < nl; i++) {
    while (j < nr && rk[j] < lk[i]) j++;
    for (int64_t jj = j; jj < nr && rk[jj] == lk[i]; jj++) {
      if (total < cap) {
        out_li[total] = i;
        out_ri[total] = jj;
      }
      total++;
    }
  }
  return total;
}

uint64_t hso_splitmix64(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* Rows of the synthetic table T (k_i = splitmix64(42, i)) that fall into `bucket`, in source order.  Two passes over
 * [first_row, first_row + nrows) in 1 M-row chunks: count, then fill -- what a scan of the whole table followed by
 * `WHERE pmod(hash(k), n) = bucket` yields, without materialising the table (bench.py checks whole buckets of the
 * 1 B-row build against it). */
typedef struct {
  uint64_t first_row;
  int64_t nrows;
  int32_t nb, bucket;
  int64_t* counts;   /* per chunk */
  int64_t* offsets;  /* per chunk, exclusive prefix (fill pass) */
  int64_t* out;
  int64_t cap;
} synth_bucket_arg;

#define SYNTH_CHUNK (1 << 20)

static inline int32_t synth_row_bucket(uint64_t i, int32_t nb) {
  int32_t h = hso_hash_long((int64_t)hso_splitmix64(42, i), 42);
  int32_t r = h % nb;
  return r < 0 ? r + nb : r;
}

static void synth_bucket_count(int64_t lo, int64_t hi, void* vp) {
  synth_bucket_arg* a = (synth_bucket_arg*)vp;
  for (int64_t c = lo; c < hi; c++) {
    int64_t r0 = c * SYNTH_CHUNK, r1 = r0 + SYNTH_CHUNK < a->nrows ? r0 + SYNTH_CHUNK : a->nrows, n = 0;
    for (int64_t r = r0; r < r1; r++) n += synth_row_bucket(a->first_row + (uint64_t)r, a->nb) == a->bucket;
    a->counts[c] = n;
  }
}

static void synth_bucket_fill(int64_t lo, int64_t hi, void* vp) {
  synth_bucket_arg* a = (synth_bucket_arg*)vp;
  for (int64_t c = lo; c < hi; c++) {
    int64_t r0 = c * SYNTH_CHUNK, r1 = r0 + SYNTH_CHUNK < a->nrows ? r0 + SYNTH_CHUNK : a->nrows, o = a->offsets[c];
    for (int64_t r = r0; r < r1; r++)
      if (synth_row_bucket(a->first_row + (uint64_t)r, a->nb) == a->bucket) {
        if (o < a->cap) a->out[o] = (int64_t)(a->first_row + (uint64_t)r);
        o++;
      }
  }
}

int64_t hso_synth_bucket_rows(uint64_t first_row, int64_t nrows, int32_t num_buckets, int32_t bucket, int64_t* out_rows,
                              int64_t cap, int32_t nthreads) {
  int64_t nchunks = (nrows + SYNTH_CHUNK - 1) / SYNTH_CHUNK;
  if (nchunks <= 0) return 0;
  synth_bucket_arg a = {first_row, nrows, num_buckets, bucket, NULL, NULL, out_rows, cap};
  a.counts = (int64_t*)calloc((size_t)nchunks, sizeof(int64_t));
  a.offsets = (int64_t*)calloc((size_t)nchunks + 1, sizeof(int64_t));
  hso_parallel_for(nchunks, 1, nthreads, synth_bucket_count, &a);
  for (int64_t c = 0; c < nchunks; c++) a.offsets[c + 1] = a.offsets[c] + a.counts[c];
  int64_t total = a.offsets[nchunks];
  if (out_rows) hso_parallel_for(nchunks, 1, nthreads, synth_bucket_fill, &a);
  free(a.counts);
  free(a.offsets);
  return total;
}
