// radix_sort.cu -- K4: segmented stable LSD radix sort of (encoded key u64, row index u32) pairs.
//
// Replaces the per-bucket SortExec Spark's FileFormatWriter adds for BucketSpec(n, cols, cols)
// (index/DataFrameWriterExtensions.scala:64; SURVEY.md section 3.1 HOT LOOP 2: UnsafeExternalRowSorter).  Each bucket is a
// segment; tiles of 4096 pairs never straddle segments, so one launch per pass sorts all buckets at once.
//
// Per 8-bit digit pass:  k_sort_hist (tile digit histograms) -> k_seg_* (per-segment column scan giving every
// (tile, digit) its global destination) -> k_sort_scatter (stable in-tile ranking with __match_any_sync, exchange
// through shared memory so each digit's items leave the tile as one contiguous run).  Passes whose digit is constant
// over the whole input (OR == AND on those bits) are skipped.
#include "device_utils.cuh"
#include "kernels.h"

namespace hs {

namespace {

constexpr int kThreads = 512;                  // 512 x 8 items: <= 64 registers -> 2 CTAs (32 warps) per SM
constexpr int kWarps = kThreads / 32;
constexpr int kItems = kSortTile / kThreads;   // 8
constexpr int kWarpRows = kSortTile / kWarps;  // 256 consecutive pairs per warp
constexpr int kChunkTiles = 128;
constexpr int kHistThreads = 256;

struct SortChunk {
  uint32_t seg;
  uint32_t tile_begin, tile_end;
};

struct DigitShift {
  int shift;
  __device__ __forceinline__ uint32_t operator()(uint64_t key, uint32_t) const { return (uint32_t)(key >> shift) & 255u; }
};
struct DigitTable {
  const uint8_t* table;
  __device__ __forceinline__ uint32_t operator()(uint64_t, uint32_t val) const { return table[val]; }
};

// Where a pass reads its (key, row) pairs from: the ping-pong arrays, or -- for the first pass over a freshly partitioned
// key column -- the raw column itself, encoded on the fly with the row's own position as its value (so the encoded-key
// and iota arrays are never materialised).
struct SrcPairs {
  const uint64_t* __restrict__ keys;
  const uint32_t* __restrict__ vals;
  __device__ __forceinline__ void load(uint64_t p, uint64_t& k, uint32_t& v) const {
    k = keys[p];
    v = vals[p];
  }
};
template <int TYPE>  // HS_TYPE_*: a compile-time type keeps the loads of a thread's items back to back
struct SrcRaw {
  const void* __restrict__ raw;
  __device__ __forceinline__ void load(uint64_t p, uint64_t& k, uint32_t& v) const {
    const uint64_t r = (TYPE == HS_TYPE_INT64 || TYPE == HS_TYPE_DOUBLE) ? ((const uint64_t*)raw)[p]
                                                                          : (uint64_t)((const uint32_t*)raw)[p];
    k = sort_encode(TYPE, r);
    v = (uint32_t)p;
  }
};

template <typename Src, typename Digit>
__global__ void __launch_bounds__(kHistThreads) k_sort_hist(const SortTile* __restrict__ tiles, Src src, Digit digit,
                                                         uint32_t* __restrict__ tile_hist) {
  __shared__ uint32_t s_hist[256];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const SortTile t = tiles[blockIdx.x];
  for (uint32_t i = threadIdx.x; i < t.count; i += kHistThreads) {
    uint64_t k;
    uint32_t v;
    src.load(t.start + i, k, v);
    atomicAdd(&s_hist[digit(k, v)], 1u);
  }
  __syncthreads();
  tile_hist[(size_t)blockIdx.x * 256 + threadIdx.x] = s_hist[threadIdx.x];
}

// A: per chunk, per digit sums
__global__ void __launch_bounds__(256) k_seg_chunk_sums(const SortChunk* __restrict__ chunks,
                                                         const uint32_t* __restrict__ tile_hist,
                                                         uint32_t* __restrict__ chunk_sums) {
  const SortChunk c = chunks[blockIdx.x];
  uint32_t s = 0;
  for (uint32_t t = c.tile_begin; t < c.tile_end; t++) s += tile_hist[(size_t)t * 256 + threadIdx.x];
  chunk_sums[(size_t)blockIdx.x * 256 + threadIdx.x] = s;
}

// B: one CTA per segment: prefix over the segment's chunks, then digit bases
__global__ void __launch_bounds__(256) k_seg_scan(const uint32_t* __restrict__ seg_chunk_begin,
                                                   const uint64_t* __restrict__ seg_start,
                                                   uint32_t* __restrict__ chunk_sums) {
  __shared__ uint32_t warp_sums[40];
  const uint32_t seg = blockIdx.x;
  const uint32_t c0 = seg_chunk_begin[seg], c1 = seg_chunk_begin[seg + 1];
  uint32_t tot = 0;
  for (uint32_t c = c0; c < c1; c++) {
    const uint32_t v = chunk_sums[(size_t)c * 256 + threadIdx.x];
    chunk_sums[(size_t)c * 256 + threadIdx.x] = tot;
    tot += v;
  }
  const uint32_t dbase = block_exclusive_scan(tot, warp_sums, nullptr) + (uint32_t)seg_start[seg];
  for (uint32_t c = c0; c < c1; c++) chunk_sums[(size_t)c * 256 + threadIdx.x] += dbase;
}

// C: per chunk, turn tile histograms into destinations.  1024 threads = 4 groups x 256 digits; each group owns a
// quarter of the chunk's tiles.  Input and output are distinct arrays so the loads can be issued ahead of the stores.
__global__ void __launch_bounds__(1024) k_seg_apply(const SortChunk* __restrict__ chunks,
                                                     const uint32_t* __restrict__ tile_hist,
                                                     uint32_t* __restrict__ tile_dst,
                                                     const uint32_t* __restrict__ chunk_sums) {
  __shared__ uint32_t s_group[4][256];
  const SortChunk c = chunks[blockIdx.x];
  const uint32_t d = threadIdx.x & 255, g = threadIdx.x >> 8;
  const uint32_t n = c.tile_end - c.tile_begin;
  const uint32_t per = (n + 3) / 4;
  const uint32_t t0 = c.tile_begin + min(g * per, n), t1 = c.tile_begin + min((g + 1) * per, n);
  uint32_t s = 0;
#pragma unroll 8
  for (uint32_t t = t0; t < t1; t++) s += tile_hist[(size_t)t * 256 + d];
  s_group[g][d] = s;
  __syncthreads();
  uint32_t run = chunk_sums[(size_t)blockIdx.x * 256 + d];
  for (uint32_t gg = 0; gg < g; gg++) run += s_group[gg][d];
#pragma unroll 8
  for (uint32_t t = t0; t < t1; t++) {
    const uint32_t v = tile_hist[(size_t)t * 256 + d];
    tile_dst[(size_t)t * 256 + d] = run;
    run += v;
  }
}

struct ScatterShared {
  uint64_t keys[kSortTile];
  uint32_t vals[kSortTile];
  uint16_t cnt[kWarps][256];   // ranking: per-warp digit counts; afterwards: first in-tile position of (warp, digit)
  uint32_t out_adj[256];       // global destination of a digit's run minus its first in-tile position
  uint32_t warp_sums[40];
};

// All warp-collectives below run with the full mask and outside any branch: slots past the end of a partial tile carry
// digit 255 and, being the last slots of the tile, rank behind every real item of that digit, so they never disturb a
// real item's position and are simply not written out.
template <typename Src, typename Digit>
__global__ void __launch_bounds__(kThreads, 2) k_sort_scatter(const SortTile* __restrict__ tiles, Src src, Digit digit,
                                                               const uint32_t* __restrict__ tile_dst,
                                                               uint64_t* __restrict__ out_keys,
                                                               uint32_t* __restrict__ out_vals) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  ScatterShared& sm = *reinterpret_cast<ScatterShared*>(smem_raw);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1;
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(&sm.cnt[0][0]);
#pragma unroll
    for (int i = 0; i < kWarps * 256 / 2 / kThreads; i++) z[threadIdx.x + i * kThreads] = 0;
  }
  const SortTile t = tiles[blockIdx.x];
  uint32_t dst0 = 0;
  if (threadIdx.x < 256) dst0 = tile_dst[(size_t)blockIdx.x * 256 + threadIdx.x];
  const uint32_t first = warp * kWarpRows + lane;
  const uint64_t p0 = t.start + first;
  uint64_t k[kItems];
  uint32_t v[kItems];
  uint32_t bin[kItems];
#pragma unroll
  for (int j = 0; j < kItems; j++) {
    const bool a = first + j * 32 < t.count;
    k[j] = ~0ull;
    v[j] = 0;
    if (a) src.load(p0 + j * 32, k[j], v[j]);
    bin[j] = a ? digit(k[j], v[j]) : 255u;
  }
  __syncthreads();
  // stable rank of every item among the warp's earlier items with the same digit
  uint16_t* cnt = sm.cnt[warp];
  uint32_t rank[kItems];
#pragma unroll
  for (int j = 0; j < kItems; j++) {
    const unsigned peers = match_any_full<8>(bin[j]);
    const uint32_t before = __popc(peers & lt);
    const uint32_t pre = cnt[bin[j]];   // every peer reads the same counter (broadcast)
    __syncwarp();
    if (before == 0) cnt[bin[j]] = (uint16_t)(pre + __popc(peers));
    __syncwarp();
    rank[j] = pre + before;
  }
  __syncthreads();
  // per digit: exclusive prefix over warps and digits -> first in-tile position of every (warp, digit)
  uint32_t total = 0;
  if (threadIdx.x < 256) {
#pragma unroll
    for (int w = 0; w < kWarps; w++) total += sm.cnt[w][threadIdx.x];
  }
  const uint32_t start = block_exclusive_scan(total, sm.warp_sums, nullptr);
  if (threadIdx.x < 256) {
    uint32_t run = start;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      const uint16_t c = sm.cnt[w][threadIdx.x];
      sm.cnt[w][threadIdx.x] = (uint16_t)run;
      run += c;
    }
    sm.out_adj[threadIdx.x] = dst0 - start;
  }
  __syncthreads();
  // exchange: digit-sorted order inside the tile
#pragma unroll
  for (int j = 0; j < kItems; j++) {
    const uint32_t pos = cnt[bin[j]] + rank[j];
    sm.keys[pos] = k[j];
    sm.vals[pos] = v[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kItems; j++) {
    const uint32_t i = threadIdx.x + j * kThreads;
    if (i < t.count) {
      const uint64_t key = sm.keys[i];
      const uint32_t val = sm.vals[i];
      const uint32_t dst = sm.out_adj[digit(key, val)] + i;
      out_keys[dst] = key;
      out_vals[dst] = val;
    }
  }
}

// ---- tie-run fix-up ---------------------------------------------------------------------------------------------
// After the stable LSD passes over the HIGH bytes of the key, rows are ordered by (key & high_mask) and rows that share
// those bits still sit in their original relative order.  For high-entropy keys such runs are rare and tiny, so instead
// of four more full passes over the data the head of every run insertion-sorts it (stably) on (key & low_mask).  A run
// longer than max_run raises `flag`; the caller then falls back to full LSD passes, which is still correct because equal
// full keys are in original order both inside untouched runs and inside insertion-sorted ones.
//
// Two phases per tile so that the divergent part runs with full warps: (1) every thread tests its rows for "head of a
// run of two or more" -- a straight-line compare against both neighbours -- and the heads are compacted into a list in
// shared memory; (2) the threads walk that list, one run each.  With 24 sorted bits and 5 M-row buckets about a quarter
// of the rows head such a run; testing and sorting in the same loop left three quarters of every warp idle.
__global__ void __launch_bounds__(256) k_fix_runs(const SortTile* __restrict__ tiles, const uint64_t* __restrict__ seg_start,
                                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                   uint64_t high_mask, uint64_t low_mask, uint32_t max_run,
                                                   uint32_t* __restrict__ flag) {
  // The tile's keys are staged in shared memory with all of a thread's loads in flight at once: s_key[1 + i] = key of row i
  // of the tile, [0] / [count + 1] = the rows just outside it (or a key with a different prefix where the segment ends).
  // Run detection, the walk to the end of a run and the order test then never leave shared memory; global memory is touched
  // again only for the rows an out-of-order run actually moves.  With 5 M-row buckets one row in eight heads a run: doing
  // all of that through dependent global loads was latency-bound (4.9 ms per 1 B rows for 8 B/row of traffic).
  __shared__ uint64_t s_key[kSortTile + 2];
  __shared__ uint16_t s_heads[kSortTile / 2], s_len[kSortTile / 2];  // 8 warps x 256 entries
  const SortTile t = tiles[blockIdx.x];
  const uint64_t segb = seg_start[t.seg], sege = seg_start[t.seg + 1];
  const unsigned lane = threadIdx.x & 31, lt = (1u << lane) - 1;
  constexpr int kPer = kSortTile / 256;
  {
    uint64_t k[kPer];
#pragma unroll
    for (int j = 0; j < kPer; j++) {
      const uint32_t i = j * 256 + threadIdx.x;
      k[j] = i < t.count ? keys[t.start + i] : 0;
    }
#pragma unroll
    for (int j = 0; j < kPer; j++) s_key[1 + j * 256 + threadIdx.x] = k[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint64_t flip = high_mask & (~high_mask + 1);  // lowest bit of the prefix: flipping it makes a different prefix
    s_key[0] = t.start > segb ? keys[t.start - 1] : (s_key[1] ^ flip);
    s_key[t.count + 1] = t.start + t.count < sege ? keys[t.start + t.count] : (s_key[t.count] ^ flip);
  }
  __syncthreads();
  // every warp keeps its own list of heads (a warp's rows give at most 256: a head needs a row behind it), so building the
  // lists takes no atomics and no block-wide counter -- 128 contended shared-memory atomics per tile, each followed by a
  // dependent shuffle, were a third of this kernel's time
  const uint32_t wid = threadIdx.x >> 5;
  uint16_t* my_heads = s_heads + wid * (kSortTile / 16);
  uint16_t* my_len = s_len + wid * (kSortTile / 16);
  uint32_t my_n = 0;
#pragma unroll 4
  for (uint32_t i0 = 0; i0 < kSortTile; i0 += 256) {  // uniform trip count: the ballot below needs whole warps
    const uint32_t i = i0 + threadIdx.x;
    bool head = false;
    if (i < t.count) {
      const uint64_t kh = s_key[1 + i] & high_mask;
      head = (s_key[i] & high_mask) != kh && (s_key[2 + i] & high_mask) == kh;
    }
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if (head) my_heads[my_n + __popc(m & lt)] = (uint16_t)i;
    my_n += __popc(m);
  }
  __syncwarp();
  // first every run is measured (reads only: a walk looks one key past its run, i.e. at the head of the next one), then,
  // behind a barrier, every run is sorted (writes stay inside the run)
  for (uint32_t e = lane; e < my_n; e += 32) {
    const uint32_t i = my_heads[e];
    const uint64_t* sk = s_key + 1 + i;
    const uint64_t kh = sk[0] & high_mask;
    uint32_t len = 1;
    while (i + len < t.count && len <= max_run && (sk[len] & high_mask) == kh) len++;
    const bool crosses = i + len == t.count && (sk[len] & high_mask) == kh;  // sk[len] is the next tile's first key here
    my_len[e] = crosses ? (uint16_t)0xffffu : (uint16_t)len;
  }
  __syncthreads();
  for (uint32_t e = lane; e < my_n; e += 32) {
    const uint32_t i = my_heads[e];
    const uint64_t p = t.start + i;
    uint64_t* sk = s_key + 1 + i;  // the run starts at sk[0]
    const uint64_t kh = sk[0] & high_mask;
    uint32_t len = my_len[e];
    if (len == 0xffffu) {
      len = t.count - i;
      // the run continues into the next tile (at most one per tile): settle it in global memory, as a whole.  The next
      // tile never touches these rows -- none of them heads a run there -- and reads only their (unchanging) prefixes.
      uint64_t q = p + len;
      while (q < sege && q - p <= max_run && (keys[q] & high_mask) == kh) q++;
      const uint32_t glen = (uint32_t)(q - p);
      if (glen > max_run) {
        *flag = 1;
        continue;
      }
      for (uint32_t a = 1; a < glen; a++) {  // stable insertion sort on the low bits
        const uint64_t ka = keys[p + a];
        const uint32_t va = vals[p + a];
        const uint64_t la = ka & low_mask;
        uint32_t b = a;
        while (b > 0 && (keys[p + b - 1] & low_mask) > la) {
          keys[p + b] = keys[p + b - 1];
          vals[p + b] = vals[p + b - 1];
          b--;
        }
        if (b != a) {
          keys[p + b] = ka;
          vals[p + b] = va;
        }
      }
      continue;
    }
    if (len > max_run) {
      *flag = 1;
      continue;
    }
    // the run lies inside the tile: stable insertion sort on the low bits, keys in shared memory; the rows that move are
    // written through to global memory (their row indices are fetched only then)
    for (uint32_t a = 1; a < len; a++) {
      const uint64_t ka = sk[a];
      const uint64_t la = ka & low_mask;
      uint32_t b = a;
      while (b > 0 && (sk[b - 1] & low_mask) > la) b--;
      if (b == a) continue;
      const uint32_t va = vals[p + a];
      for (uint32_t c = a; c > b; c--) {
        const uint64_t kc = sk[c - 1];
        sk[c] = kc;
        keys[p + c] = kc;
        vals[p + c] = vals[p + c - 1];
      }
      sk[b] = ka;
      keys[p + b] = ka;
      vals[p + b] = va;
    }
  }
}

template <typename Src, typename Digit>
void run_pass(hs_ctx* ctx, SortPlan* plan, const SortChunk* chunks, int64_t nchunks, const uint32_t* seg_chunk_begin,
              uint32_t* chunk_sums, Src src, uint64_t* out_keys, uint32_t* out_vals, Digit digit) {
  {
    KernelScope _ks(ctx, "k_sort_hist");
    k_sort_hist<Src, Digit><<<(unsigned)plan->ntiles, kHistThreads, 0, ctx->stream>>>(plan->tiles.get(), src, digit,
                                                                                 plan->tile_hist.get());
    HS_LAUNCH_CHECK(ctx);
  }
  KernelScope* _scan = new KernelScope(ctx, "k_seg_scan");
  k_seg_chunk_sums<<<(unsigned)nchunks, 256, 0, ctx->stream>>>(chunks, plan->tile_hist.get(), chunk_sums);
  HS_LAUNCH_CHECK(ctx);
  k_seg_scan<<<(unsigned)plan->nseg, 256, 0, ctx->stream>>>(seg_chunk_begin, plan->seg_start.get(), chunk_sums);
  HS_LAUNCH_CHECK(ctx);
  k_seg_apply<<<(unsigned)nchunks, 1024, 0, ctx->stream>>>(chunks, plan->tile_hist.get(), plan->tile_dst.get(), chunk_sums);
  HS_LAUNCH_CHECK(ctx);
  delete _scan;
  static DeviceOnce attr_once;  // one per (Src, Digit) instantiation
  bool& attr = attr_once(ctx->device);
  if (!attr) {
    HS_CUDA(cudaFuncSetAttribute(k_sort_scatter<Src, Digit>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(ScatterShared)));
    attr = true;
  }
  KernelScope _ks(ctx, "k_sort_scatter");
  k_sort_scatter<Src, Digit><<<(unsigned)plan->ntiles, kThreads, sizeof(ScatterShared), ctx->stream>>>(
      plan->tiles.get(), src, digit, plan->tile_dst.get(), out_keys, out_vals);
  HS_LAUNCH_CHECK(ctx);
}

struct ChunkPlan {
  Buf<SortChunk> chunks;
  Buf<uint32_t> seg_chunk_begin;
  Buf<uint32_t> chunk_sums;
  int64_t nchunks = 0;
};

// chunk lists are derived from the plan's host-side segment -> tile mapping
ChunkPlan build_chunks(hs_ctx* ctx, const std::vector<uint32_t>& seg_tile_begin) {
  ChunkPlan cp;
  const int nseg = (int)seg_tile_begin.size() - 1;
  std::vector<SortChunk> chunks;
  std::vector<uint32_t> scb(nseg + 1);
  for (int s = 0; s < nseg; s++) {
    scb[s] = (uint32_t)chunks.size();
    for (uint32_t t = seg_tile_begin[s]; t < seg_tile_begin[s + 1]; t += kChunkTiles)
      chunks.push_back(SortChunk{(uint32_t)s, t, std::min<uint32_t>(t + kChunkTiles, seg_tile_begin[s + 1])});
  }
  scb[nseg] = (uint32_t)chunks.size();
  cp.nchunks = (int64_t)chunks.size();
  cp.chunks.alloc(ctx, std::max<size_t>(1, chunks.size()));
  cp.seg_chunk_begin.alloc(ctx, scb.size());
  cp.chunk_sums.alloc(ctx, std::max<size_t>(1, chunks.size()) * 256);
  if (!chunks.empty())
    copy_h2d(ctx, cp.chunks.get(), chunks.data(), chunks.size() * sizeof(SortChunk));
  copy_h2d(ctx, cp.seg_chunk_begin.get(), scb.data(), scb.size() * 4);  // (snapshots: the vectors may go out of scope)
  return cp;
}

}  // namespace

// tiles of one segment (one CTA per segment): the list has a quarter of a million entries at 1 B rows, so it is generated
// where it is used instead of being built on the host and copied over from pageable memory
__global__ void k_build_tiles(const uint64_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_tile_begin,
                              SortTile* __restrict__ tiles) {
  const uint32_t s = blockIdx.x;
  const uint64_t b = seg_start[s], e = seg_start[s + 1];
  const uint32_t t0 = seg_tile_begin[s], nt = seg_tile_begin[s + 1] - t0;
  for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) {
    const uint64_t p = b + (uint64_t)i * kSortTile;
    tiles[t0 + i] = SortTile{s, (uint32_t)min((uint64_t)kSortTile, e - p), p};
  }
}

void build_sort_plan(hs_ctx* ctx, const uint64_t* seg_offsets, int nseg, SortPlan* plan) {
  std::vector<uint32_t> stb(nseg + 1);
  std::vector<uint64_t> sstart(nseg + 1);
  uint64_t ntiles = 0;
  for (int s = 0; s < nseg; s++) {
    stb[s] = (uint32_t)ntiles;
    sstart[s] = seg_offsets[s];
    ntiles += ceil_div(seg_offsets[s + 1] - seg_offsets[s], (uint64_t)kSortTile);
  }
  stb[nseg] = (uint32_t)ntiles;
  sstart[nseg] = seg_offsets[nseg];
  if (seg_offsets[nseg] >= (1ull << 32)) fail(HS_EUNSUPPORTED, "more than 2^32-1 rows per GPU per call");
  plan->n = (int64_t)seg_offsets[nseg];
  plan->ntiles = (int64_t)ntiles;
  plan->nseg = nseg;
  plan->tiles.alloc(ctx, std::max<size_t>(1, ntiles));
  plan->seg_tile_begin.alloc(ctx, stb.size());
  plan->seg_start.alloc(ctx, sstart.size());
  plan->tile_hist.alloc(ctx, std::max<size_t>(1, ntiles) * 256);
  plan->tile_dst.alloc(ctx, std::max<size_t>(1, ntiles) * 256);
  copy_h2d(ctx, plan->seg_tile_begin.get(), stb.data(), stb.size() * 4);
  copy_h2d(ctx, plan->seg_start.get(), sstart.data(), sstart.size() * 8);
  if (nseg > 0 && ntiles > 0) {
    k_build_tiles<<<nseg, 256, 0, ctx->stream>>>(plan->seg_start.get(), plan->seg_tile_begin.get(), plan->tiles.get());
    HS_LAUNCH_CHECK(ctx);
  }
  plan->h_seg_tile_begin = stb;  // (copy_h2d took snapshots of the host vectors)
}

void segmented_sort_pairs(hs_ctx* ctx, SortPlan* plan, uint64_t*& keys, uint64_t*& keys_alt, uint32_t*& vals,
                          uint32_t*& vals_alt, uint64_t bit_mask, const RawKeyColumn* first_pass_source) {
  if (plan->ntiles == 0) return;
  ChunkPlan cp = build_chunks(ctx, plan->h_seg_tile_begin);
  bool first = first_pass_source != nullptr;
  for (int pass = 0; pass < 8; pass++) {
    if (((bit_mask >> (pass * 8)) & 0xff) == 0) continue;  // digit constant over the whole input
    if (first) {
      const void* raw = first_pass_source->data;
      const DigitShift digit{pass * 8};
#define HS_RAW_PASS(T)                                                                                                   \
  run_pass(ctx, plan, cp.chunks.get(), cp.nchunks, cp.seg_chunk_begin.get(), cp.chunk_sums.get(), SrcRaw<T>{raw}, keys_alt, \
           vals_alt, digit)
      switch (first_pass_source->type) {
        case HS_TYPE_INT32: HS_RAW_PASS(HS_TYPE_INT32); break;
        case HS_TYPE_INT64: HS_RAW_PASS(HS_TYPE_INT64); break;
        case HS_TYPE_FLOAT: HS_RAW_PASS(HS_TYPE_FLOAT); break;
        case HS_TYPE_DOUBLE: HS_RAW_PASS(HS_TYPE_DOUBLE); break;
        default: fail(HS_EUNSUPPORTED, "sort: key type %d", first_pass_source->type);
      }
#undef HS_RAW_PASS
      first = false;
    } else {
      run_pass(ctx, plan, cp.chunks.get(), cp.nchunks, cp.seg_chunk_begin.get(), cp.chunk_sums.get(), SrcPairs{keys, vals},
               keys_alt, vals_alt, DigitShift{pass * 8});
    }
    std::swap(keys, keys_alt);
    std::swap(vals, vals_alt);
  }
  if (first) fail(HS_EINVAL, "segmented_sort_pairs: raw first-pass source given but no pass ran");
}

void launch_fix_runs(hs_ctx* ctx, SortPlan* plan, uint64_t* keys, uint32_t* vals, uint64_t high_mask, uint64_t low_mask,
                     uint32_t max_run, uint32_t* d_flag) {
  KernelScope _ks(ctx, "k_fix_runs");
  if (plan->ntiles == 0) return;
  k_fix_runs<<<(unsigned)plan->ntiles, 256, 0, ctx->stream>>>(plan->tiles.get(), plan->seg_start.get(), keys, vals, high_mask,
                                                              low_mask, max_run, d_flag);
  HS_LAUNCH_CHECK(ctx);
}

void segmented_sort_pass_by_table(hs_ctx* ctx, SortPlan* plan, uint64_t*& keys, uint64_t*& keys_alt, uint32_t*& vals,
                                  uint32_t*& vals_alt, const uint8_t* digits) {
  if (plan->ntiles == 0) return;
  ChunkPlan cp = build_chunks(ctx, plan->h_seg_tile_begin);
  run_pass(ctx, plan, cp.chunks.get(), cp.nchunks, cp.seg_chunk_begin.get(), cp.chunk_sums.get(), SrcPairs{keys, vals},
           keys_alt, vals_alt, DigitTable{digits});
  std::swap(keys, keys_alt);
  std::swap(vals, vals_alt);
}

}  // namespace hs
