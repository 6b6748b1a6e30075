// read_side.cu -- K7 sorted range select and K8 per-bucket merge join over decoded index buckets.
//
// K7 replaces the FileSourceScanExec + pushed Filter Spark runs after FilterIndexRule.applyIndex swaps the source
// relation for the index files (index/covering/FilterIndexRule.scala:135-149).  Index files are sorted on the key, so a
// range predicate is two binary searches per file instead of a predicate over every row.
// K8 replaces the bucketed scan + SortMergeJoinExec (no ShuffleExchangeExec) Spark plans after
// JoinIndexRule.applyIndex (index/covering/JoinIndexRule.scala:653-687): bucket b of the left index joins bucket b of
// the right index; every left row binary-searches its match range in the right bucket, a scan turns match counts into
// output offsets, and a second kernel emits the (left row, right row) pairs in (left, right) order.
#include "device_utils.cuh"
#include "kernels.h"

namespace hs {

namespace {

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t* a, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int64_t upper_bound_i64(const int64_t* a, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] <= v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void k_range_bounds(const int64_t* __restrict__ keys, const uint64_t* __restrict__ seg_offsets, int nseg,
                               int has_lo, int64_t lo, int has_hi, int64_t hi, int64_t* __restrict__ bounds) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int64_t b = (int64_t)seg_offsets[s], n = (int64_t)seg_offsets[s + 1] - b;
  int64_t first = has_lo ? lower_bound_i64(keys + b, n, lo) : 0;
  int64_t last = has_hi ? upper_bound_i64(keys + b, n, hi) : n;
  if (last < first) last = first;
  bounds[2 * s] = first;
  bounds[2 * s + 1] = last;
}

// one thread per left row; the segment of a row is found by binary search over the (few hundred) segment offsets
template <bool STR>
__global__ void k_join_count(const int64_t* __restrict__ lkeys, const uint64_t* __restrict__ lseg,
                             const int64_t* __restrict__ rkeys, const uint64_t* __restrict__ rseg, int nseg, int64_t nl,
                             uint32_t* __restrict__ counts, uint32_t* __restrict__ first_match) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += stride) {
    int lo = 0, hi = nseg;  // last segment with lseg[s] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (lseg[mid] <= (uint64_t)i) lo = mid;
      else hi = mid;
    }
    const int64_t rb = (int64_t)rseg[lo], rn = (int64_t)rseg[lo + 1] - rb;
    int64_t f, l;
    if (STR) {  // keys are string references: the same two searches in byte order
      const uint64_t k = (uint64_t)lkeys[i];
      const uint64_t* a = (const uint64_t*)rkeys + rb;
      int64_t x = 0, y = rn;
      while (x < y) {
        const int64_t mid = x + ((y - x) >> 1);
        if (string_compare(a[mid], k) < 0) x = mid + 1;
        else y = mid;
      }
      f = x;
      y = rn;
      while (x < y) {
        const int64_t mid = x + ((y - x) >> 1);
        if (string_compare(a[mid], k) <= 0) x = mid + 1;
        else y = mid;
      }
      l = x;
    } else {
      const int64_t k = lkeys[i];
      f = lower_bound_i64(rkeys + rb, rn, k);
      l = upper_bound_i64(rkeys + rb, rn, k);
    }
    counts[i] = (uint32_t)(l - f);
    first_match[i] = (uint32_t)(rb + f);
  }
}

__global__ void k_join_emit(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ first_match,
                            const uint64_t* __restrict__ out_offsets, int64_t nl, uint32_t* __restrict__ out_li,
                            uint32_t* __restrict__ out_ri) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += stride) {
    const uint32_t c = counts[i];
    const uint64_t o = out_offsets[i];
    const uint32_t f = first_match[i];
    for (uint32_t j = 0; j < c; j++) {
      out_li[o + j] = (uint32_t)i;
      out_ri[o + j] = f + j;
    }
  }
}

// ---- exclusive scan uint32 -> uint64 (three kernels; block of 256 threads x 8 items) --------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) k_scan_block_sums(const uint32_t* __restrict__ in, int64_t n,
                                                                   unsigned long long* __restrict__ block_sums) {
  __shared__ unsigned long long s[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  unsigned long long v = 0;
  for (int k = 0; k < kScanItems; k++) {
    const int64_t i = base + k * kScanThreads + threadIdx.x;
    if (i < n) v += in[i];
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kScanThreads / 32; w++) t += s[w];
    block_sums[blockIdx.x] = t;
  }
}

// one CTA of 1024 threads: each thread owns a contiguous slice of the block sums
__global__ void __launch_bounds__(1024) k_scan_block_offsets(unsigned long long* __restrict__ block_sums, int64_t nblocks,
                                                              unsigned long long* __restrict__ total_out) {
  __shared__ unsigned long long s[1025];
  const int64_t per = (nblocks + 1023) / 1024;
  const int64_t b0 = (int64_t)threadIdx.x * per, b1 = min(b0 + per, nblocks);
  unsigned long long local = 0;
  for (int64_t b = b0; b < b1; b++) local += block_sums[b];
  s[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; i++) {
      const unsigned long long v = s[i];
      s[i] = run;
      run += v;
    }
    s[1024] = run;
  }
  __syncthreads();
  unsigned long long run = s[threadIdx.x];
  for (int64_t b = b0; b < b1; b++) {
    const unsigned long long v = block_sums[b];
    block_sums[b] = run;
    run += v;
  }
  if (threadIdx.x == 0) *total_out = s[1024];
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint32_t* __restrict__ in, int64_t n,
                                                              const unsigned long long* __restrict__ block_offsets,
                                                              uint64_t* __restrict__ out) {
  __shared__ uint32_t warp_sums[40];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t x[kScanItems];
  uint32_t local = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    x[k] = (base + k < n) ? in[base + k] : 0;
    local += x[k];
  }
  uint32_t pre = block_exclusive_scan(local, warp_sums, nullptr);
  unsigned long long run = block_offsets[blockIdx.x] + pre;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    if (base + k < n) out[base + k] = run;
    run += x[k];
  }
}

__global__ void k_filter_mask(const int64_t* __restrict__ keys, const uint8_t* __restrict__ valid, int64_t n, int has_lo,
                              int64_t lo, int has_hi, int64_t hi, uint32_t* __restrict__ mask) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t k = keys[i];
    bool ok = !valid || valid[i];  // a null key never satisfies a comparison
    if (has_lo) ok = ok && k >= lo;
    if (has_hi) ok = ok && k <= hi;
    mask[i] = ok ? 1u : 0u;
  }
}

__global__ void k_not_in_mask(const int64_t* __restrict__ ids, int64_t n, const int64_t* __restrict__ deleted,
                              int ndeleted, uint32_t* __restrict__ mask) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t id = ids[i];
    bool hit = false;
    for (int d = 0; d < ndeleted; d++) hit |= deleted[d] == id;
    if (hit) mask[i] = 0;
  }
}

__global__ void k_compact(const uint32_t* __restrict__ mask, const uint64_t* __restrict__ offsets, int64_t n,
                          uint32_t* __restrict__ out_idx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (mask[i]) out_idx[offsets[i]] = (uint32_t)i;
}

inline int grid_for(hs_ctx* ctx, int64_t n, int threads, int per_sm) {
  int64_t want = ceil_div(n, threads);
  int64_t cap = (int64_t)ctx->sm_count * per_sm;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

// ---- string keys: the same scans over references, compared in UTF8String byte order ------------------------------------
__global__ void k_range_bounds_str(const uint64_t* __restrict__ refs, const uint64_t* __restrict__ seg_offsets, int nseg,
                                   int has_lo, uint64_t lo_ref, int has_hi, uint64_t hi_ref, int64_t* __restrict__ bounds) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int64_t b = (int64_t)seg_offsets[s], n = (int64_t)seg_offsets[s + 1] - b;
  const uint64_t* a = refs + b;
  int64_t first = 0, last = n;
  if (has_lo) {  // first row with value >= lo
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (string_compare(a[mid], lo_ref) < 0) lo = mid + 1;
      else hi = mid;
    }
    first = lo;
  }
  if (has_hi) {  // first row with value > hi
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (string_compare(a[mid], hi_ref) <= 0) lo = mid + 1;
      else hi = mid;
    }
    last = lo;
  }
  if (last < first) last = first;
  bounds[2 * s] = first;
  bounds[2 * s + 1] = last;
}

__global__ void k_filter_mask_str(const uint64_t* __restrict__ refs, const uint8_t* __restrict__ valid, int64_t n, int has_lo,
                                  uint64_t lo_ref, int has_hi, uint64_t hi_ref, uint32_t* __restrict__ mask) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    bool ok = !valid || valid[i];  // a null key never satisfies a comparison
    if (ok) {
      const uint64_t r = refs[i];
      if (has_lo) ok = string_compare(r, lo_ref) >= 0;
      if (ok && has_hi) ok = string_compare(r, hi_ref) <= 0;
    }
    mask[i] = ok ? 1u : 0u;
  }
}

__global__ void k_string_lengths(const uint64_t* __restrict__ refs, const uint8_t* __restrict__ valid,
                                 const uint32_t* __restrict__ idx, int64_t n, uint32_t* __restrict__ lens) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t r = idx[i];
    lens[i] = (!valid || valid[r]) ? ref_len(refs[r]) : 0u;
  }
}

// one warp per output value: short values are a single coalesced pass, long ones loop
__global__ void k_copy_strings(const uint64_t* __restrict__ refs, const uint8_t* __restrict__ valid,
                               const uint32_t* __restrict__ idx, int64_t n, const uint64_t* __restrict__ offsets,
                               uint8_t* __restrict__ out) {
  const unsigned lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const uint32_t r = idx[i];
    if (valid && !valid[r]) continue;
    const uint64_t ref = refs[r];
    const uint8_t* src = ref_ptr(ref);
    uint8_t* dst = out + offsets[i];
    for (uint32_t j = lane, len = ref_len(ref); j < len; j += 32) dst[j] = src[j];
  }
}

}  // namespace

void launch_range_bounds_strings(hs_ctx* ctx, const uint64_t* refs, const uint64_t* seg_offsets, int nseg, int has_lo,
                                 uint64_t lo_ref, int has_hi, uint64_t hi_ref, int64_t* bounds) {
  if (nseg == 0) return;
  k_range_bounds_str<<<(nseg + 127) / 128, 128, 0, ctx->stream>>>(refs, seg_offsets, nseg, has_lo, lo_ref, has_hi, hi_ref, bounds);
  HS_LAUNCH_CHECK(ctx);
}

void launch_filter_mask_strings(hs_ctx* ctx, const uint64_t* refs, const uint8_t* valid, int64_t n, int has_lo,
                                uint64_t lo_ref, int has_hi, uint64_t hi_ref, uint32_t* mask) {
  if (n == 0) return;
  k_filter_mask_str<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(refs, valid, n, has_lo, lo_ref, has_hi, hi_ref, mask);
  HS_LAUNCH_CHECK(ctx);
}

void launch_string_lengths(hs_ctx* ctx, const uint64_t* refs, const uint8_t* valid, const uint32_t* idx, int64_t n,
                           uint32_t* lens) {
  if (n == 0) return;
  k_string_lengths<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(refs, valid, idx, n, lens);
  HS_LAUNCH_CHECK(ctx);
}

void launch_copy_strings(hs_ctx* ctx, const uint64_t* refs, const uint8_t* valid, const uint32_t* idx, int64_t n,
                         const uint64_t* offsets, uint8_t* out) {
  if (n == 0) return;
  k_copy_strings<<<grid_for(ctx, n * 32, 256, 16), 256, 0, ctx->stream>>>(refs, valid, idx, n, offsets, out);
  HS_LAUNCH_CHECK(ctx);
}

void launch_range_bounds(hs_ctx* ctx, const int64_t* keys, const uint64_t* seg_offsets, int nseg, int has_lo,
                         int64_t lo, int has_hi, int64_t hi, int64_t* bounds) {
  if (nseg == 0) return;
  k_range_bounds<<<(nseg + 127) / 128, 128, 0, ctx->stream>>>(keys, seg_offsets, nseg, has_lo, lo, has_hi, hi, bounds);
  HS_LAUNCH_CHECK(ctx);
}

void launch_join_count(hs_ctx* ctx, const int64_t* lkeys, const uint64_t* lseg, const int64_t* rkeys,
                       const uint64_t* rseg, int nseg, int64_t nl, uint32_t* counts, uint32_t* first_match, bool string_keys) {
  KernelScope _ks(ctx, "k_join_count");
  if (nl == 0) return;
  if (string_keys)
    k_join_count<true><<<grid_for(ctx, nl, 256, 16), 256, 0, ctx->stream>>>(lkeys, lseg, rkeys, rseg, nseg, nl, counts, first_match);
  else
    k_join_count<false><<<grid_for(ctx, nl, 256, 16), 256, 0, ctx->stream>>>(lkeys, lseg, rkeys, rseg, nseg, nl, counts, first_match);
  HS_LAUNCH_CHECK(ctx);
}

void launch_join_emit(hs_ctx* ctx, const uint32_t* counts, const uint32_t* first_match, const uint64_t* out_offsets,
                      int64_t nl, uint32_t* out_li, uint32_t* out_ri) {
  KernelScope _ks(ctx, "k_join_emit");
  if (nl == 0) return;
  k_join_emit<<<grid_for(ctx, nl, 256, 16), 256, 0, ctx->stream>>>(counts, first_match, out_offsets, nl, out_li, out_ri);
  HS_LAUNCH_CHECK(ctx);
}

void exclusive_scan_u32_u64(hs_ctx* ctx, const uint32_t* in, int64_t n, uint64_t* out) {
  // out has n+1 entries; out[n] = total
  const int64_t nblocks = std::max<int64_t>(1, ceil_div(n, kScanTile));
  Buf<unsigned long long> block_sums(ctx, (size_t)nblocks);
  k_scan_block_sums<<<(unsigned)nblocks, kScanThreads, 0, ctx->stream>>>(in, n, block_sums.get());
  HS_LAUNCH_CHECK(ctx);
  k_scan_block_offsets<<<1, 1024, 0, ctx->stream>>>(block_sums.get(), nblocks, (unsigned long long*)(out + n));
  HS_LAUNCH_CHECK(ctx);
  if (n > 0) {
    k_scan_apply<<<(unsigned)nblocks, kScanThreads, 0, ctx->stream>>>(in, n, block_sums.get(), out);
    HS_LAUNCH_CHECK(ctx);
  }
}

void launch_filter_mask(hs_ctx* ctx, const int64_t* keys, const uint8_t* valid, int64_t n, int has_lo, int64_t lo,
                        int has_hi, int64_t hi, uint32_t* mask) {
  if (n == 0) return;
  k_filter_mask<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(keys, valid, n, has_lo, lo, has_hi, hi, mask);
  HS_LAUNCH_CHECK(ctx);
}

void launch_compact_indices(hs_ctx* ctx, const uint32_t* mask, const uint64_t* offsets, int64_t n, uint32_t* out_idx) {
  if (n == 0) return;
  k_compact<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(mask, offsets, n, out_idx);
  HS_LAUNCH_CHECK(ctx);
}

void launch_not_in_mask(hs_ctx* ctx, const int64_t* file_ids, int64_t n, const int64_t* deleted, int ndeleted,
                        uint32_t* mask) {
  if (n == 0 || ndeleted == 0) return;
  k_not_in_mask<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(file_ids, n, deleted, ndeleted, mask);
  HS_LAUNCH_CHECK(ctx);
}

}  // namespace hs
