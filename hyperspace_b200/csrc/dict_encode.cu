// dict_encode.cu -- dictionary encoding on the GPU (PLAIN_DICTIONARY pages with bit-packed indices).
//
// parquet-mr, which writes the reference's index files (index/DataFrameWriterExtensions.scala:58-67 -> ParquetOutputWriter),
// dictionary-encodes every column chunk whose distinct values fit its dictionary page and falls back to PLAIN otherwise.
// Here: one open-addressing hash set per candidate column (64-bit CAS on the raw value bits, no locks) collects the
// distinct values of the whole column and gives up as soon as more than kMaxDictEntries are seen; the host sorts the
// (small) dictionary, maps every table slot to its dictionary index, and k_dict_encode looks each gathered value up and
// bit-packs the indices of a 4096-row tile in shared memory.  For the benchmark table this shrinks v1/v3/v4 from
// 8/4/4 bytes per row to 10/7/12 bits, i.e. 32 -> 19.7 bytes per row over PCIe and on disk.
#include "device_utils.cuh"
#include "kernels.h"

namespace hs {

namespace {

constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t dict_hash(uint64_t v) {
  v ^= v >> 33;
  v *= 0xff51afd7ed558ccdull;
  v ^= v >> 33;
  v *= 0xc4ceb9fe1a85ec53ull;
  v ^= v >> 33;
  return (uint32_t)v;
}

__device__ __forceinline__ uint64_t load_raw_value(const void* src, int width, int64_t i) {
  return width == 8 ? ((const uint64_t*)src)[i] : (uint64_t)((const uint32_t*)src)[i];
}

// state[0] = number of distinct values inserted, state[1] = overflow flag, state[2] = the value kEmpty itself occurs
__global__ void __launch_bounds__(kThreads) k_dict_build(const void* __restrict__ src, int width, int64_t begin, int64_t end,
                                                          unsigned long long* __restrict__ keys, uint32_t mask,
                                                          uint32_t max_distinct, uint32_t* __restrict__ state) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
    if (*(volatile uint32_t*)&state[1]) return;
    const uint64_t v = load_raw_value(src, width, i);
    if (v == kEmpty) {
      state[2] = 1;
      continue;
    }
    uint32_t h = dict_hash(v) & mask;
    // Successful inserts are reserved through state[0] BEFORE the CAS, so at most max_distinct slots of the
    // 4 x max_distinct table are ever occupied and probing always terminates.
    for (uint32_t probes = 0; probes <= mask; probes++) {
      const unsigned long long cur = keys[h];
      if (cur == v) break;
      if (cur == kEmpty) {
        if (atomicAdd(&state[0], 1u) >= max_distinct) {
          state[1] = 1;
          return;
        }
        const unsigned long long old = atomicCAS(&keys[h], kEmpty, (unsigned long long)v);
        if (old == kEmpty) break;          // inserted
        atomicSub(&state[0], 1u);          // lost the race for this slot: give the reservation back
        if (old == v) break;
      }
      h = (h + 1) & mask;
    }
  }
}

template <int W>
__global__ void __launch_bounds__(kThreads) k_dict_encode(const SortTile* __restrict__ tiles,
                                                           const uint64_t* __restrict__ seg_start,
                                                           const uint32_t* __restrict__ perm, const void* __restrict__ src,
                                                           const unsigned long long* __restrict__ keys,
                                                           const uint32_t* __restrict__ slot_index, uint32_t mask,
                                                           uint32_t empty_index, uint32_t bw,
                                                           const uint64_t* __restrict__ page_value_offset,
                                                           const uint32_t* __restrict__ bucket_page_begin,
                                                           int64_t rows_per_page, uint8_t* __restrict__ arena) {
  __shared__ uint32_t s_bits[kSortTile * 16 / 32 + 1];  // up to 16 bits per index
  const SortTile t = tiles[blockIdx.x];
  const uint32_t words = (kSortTile * bw + 31) / 32 + 1;
  for (uint32_t w = threadIdx.x; w < words; w += kThreads) s_bits[w] = 0;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < t.count; i += kThreads) {
    const uint32_t row = perm[t.start + i];
    const uint64_t v = W == 8 ? ((const uint64_t*)src)[row] : (uint64_t)((const uint32_t*)src)[row];
    uint32_t idx = empty_index;
    if (v != kEmpty) {
      uint32_t h = dict_hash(v) & mask;
      uint32_t probes = 0;
      while (keys[h] != v && probes++ <= mask) h = (h + 1) & mask;  // present by construction; bounded regardless
      idx = slot_index[h];
    }
    const uint32_t bit = i * bw;
    const uint32_t sh = bit & 31;
    atomicOr(&s_bits[bit >> 5], idx << sh);
    if (sh + bw > 32) atomicOr(&s_bits[(bit >> 5) + 1], idx >> (32 - sh));
  }
  __syncthreads();
  const uint64_t lr0 = t.start - seg_start[t.seg];
  const uint64_t page = lr0 / (uint64_t)rows_per_page;
  const uint64_t in_page = lr0 - page * (uint64_t)rows_per_page;
  uint8_t* const out = arena + page_value_offset[bucket_page_begin[t.seg] + page] + in_page * bw / 8;
  const uint32_t nbytes = ((t.count + 7) / 8) * bw;  // whole groups of 8 values; the padding indices are zero
  const uint8_t* sb = reinterpret_cast<const uint8_t*>(s_bits);
  for (uint32_t b = threadIdx.x; b < nbytes; b += kThreads) out[b] = sb[b];
}

inline int grid_for(hs_ctx* ctx, int64_t n, int threads, int per_sm) {
  int64_t want = ceil_div(n, threads);
  int64_t cap = (int64_t)ctx->sm_count * per_sm;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

}  // namespace

void launch_dict_build(hs_ctx* ctx, const void* src, int width, int64_t begin, int64_t end, unsigned long long* keys,
                       uint32_t capacity, uint32_t max_distinct, uint32_t* state) {
  KernelScope _ks(ctx, "k_dict_build");
  if (end <= begin) return;
  k_dict_build<<<grid_for(ctx, end - begin, kThreads, 16), kThreads, 0, ctx->stream>>>(src, width, begin, end, keys,
                                                                                       capacity - 1, max_distinct, state);
  HS_LAUNCH_CHECK(ctx);
}

void launch_dict_encode(hs_ctx* ctx, const SortTile* tiles, int64_t ntiles, const uint64_t* seg_start, const uint32_t* perm,
                        const void* src, int width, const unsigned long long* keys, const uint32_t* slot_index,
                        uint32_t capacity, uint32_t empty_index, uint32_t bw, const uint64_t* page_value_offset,
                        const uint32_t* bucket_page_begin, int64_t rows_per_page, uint8_t* arena) {
  KernelScope _ks(ctx, "k_dict_encode");
  if (ntiles == 0) return;
  if (width == 8)
    k_dict_encode<8><<<(unsigned)ntiles, kThreads, 0, ctx->stream>>>(tiles, seg_start, perm, src, keys, slot_index, capacity - 1,
                                                                      empty_index, bw, page_value_offset, bucket_page_begin,
                                                                      rows_per_page, arena);
  else
    k_dict_encode<4><<<(unsigned)ntiles, kThreads, 0, ctx->stream>>>(tiles, seg_start, perm, src, keys, slot_index, capacity - 1,
                                                                      empty_index, bw, page_value_offset, bucket_page_begin,
                                                                      rows_per_page, arena);
  HS_LAUNCH_CHECK(ctx);
}

}  // namespace hs
