// kernels.h -- host-callable launchers of the CUDA kernels (one .cu per stage).
#pragma once
#include "hs_common.h"

namespace hs {

// ---- Parquet decode (parquet_decode.cu) ---------------------------------------------------------------------------
struct ChunkDesc {          // one per (file, row group, projected column)
  const uint8_t* data;      // device pointer to the first page header of the chunk
  uint64_t size;            // chunk bytes
  int64_t num_values;       // values (rows) in the chunk
  int64_t row_base;         // global row index of the row group's first row
  int32_t col;              // projected column index
  int32_t phys_type;        // pq::PhysType
  int32_t max_def;          // 0 (required) or 1 (optional)
  int32_t file_index;
  int32_t codec;            // pq::Codec of the chunk (UNCOMPRESSED or SNAPPY)
  int32_t pad;
};

struct PageDesc {           // one per data page
  const uint8_t* data;      // page body
  const uint8_t* dict;      // dictionary page body (PLAIN values) or nullptr
  int64_t first_row;        // global row index of the page's first value
  int32_t dict_count;
  int32_t size;             // body bytes
  int32_t num_values;
  int32_t encoding;         // pq::Encoding of the values
  int32_t page_type;        // DATA_PAGE / DATA_PAGE_V2
  int32_t def_bytes;        // v2: definition_levels_byte_length, v1: -1 (length-prefixed block)
  int32_t rep_bytes;        // v2
  int32_t col, phys_type, max_def;
  int32_t file_index;
  int32_t uncompressed_size;  // body bytes after decompression (== size for stored pages)
  // compressed chunks only: the chunk's dictionary page as stored, and whether this page / the dictionary is compressed
  int32_t dict_size, dict_uncompressed_size;
  int32_t is_compressed;      // this page's values are snappy-compressed
  int32_t codec;
  int32_t chunk;              // index of the column chunk (ChunkDesc) the page belongs to
};

struct ColumnOut {          // decoded column destination
  void* data;               // num_rows values of `width` bytes (row order); carry != 0: one uint16 code per row
  uint8_t* valid;           // one byte per row (pre-set to 1) or nullptr for required columns
  int32_t width;
  int32_t type;             // HS_TYPE_*
  // Late-materialised dictionary column (carry != 0): every page of the column is dictionary-encoded and free of nulls, and
  // the union of the chunk dictionaries is already final.  The decoder then translates each chunk-local index into the
  // code of the value in that global dictionary (look-up table `carry_entries`, see DictMapArgs) and never writes values.
  const void* carry_entries;
  uint32_t carry_mask;
  uint32_t carry_empty_index;  // code of the value 0xFFFF...F, which the look-up table cannot hold
  int32_t carry;
  int32_t skip;                // != 0: zero-copy column, its pages are not decoded at all
};

// Per-column OR over the column's data pages (k_classify_pages), read before decoding to pick late-materialised columns
// PAGECLASS_NOT_IN_PLACE: the page cannot be read where it lies -- it is not a PLAIN, stored (uncompressed), null-free
// page of 4- or 8-byte values whose body is value-aligned, large enough, and at least one partition tile long
enum : uint32_t { PAGECLASS_NOT_DICT = 1u, PAGECLASS_MAYBE_NULLS = 2u, PAGECLASS_NOT_IN_PLACE = 4u };
void launch_classify_pages(hs_ctx* ctx, const PageDesc* pages, int64_t n_pages, uint32_t* col_flags, int zc_tile_rows);
// Zero-copy PLAIN columns: a column whose every page passes the PAGECLASS_NOT_IN_PLACE test is never decoded.  Its values
// are read by the hash and partition kernels straight from the page bodies inside the source file images.  Per partition
// tile (zc_tile_rows rows: 4096, or 8192 when the rows leave over NVLink) they need to know where the tile's values lie:
// pages are at least one tile long, so a tile touches at most two of them.
struct ZcTile {
  const uint8_t* p0;  // page body of the tile's first row, rebased: the value of GLOBAL row r is at p0 + r * width ...
  const uint8_t* p1;  // ... for r < split, and at p1 + r * width from row `split` on (the next page)
  int64_t split;
};
// entries of every page of a column with tile_src[col] != nullptr
void launch_fill_zc_tiles(hs_ctx* ctx, const PageDesc* pages, int64_t n_pages, ZcTile* const* tile_src, int zc_tile_rows,
                          int64_t nrows);

// device error word: 0 = ok, else (code << 24 | detail)
enum DecodeError : uint32_t {
  DERR_NONE = 0, DERR_BAD_HEADER = 1, DERR_UNSUPPORTED_ENCODING = 2, DERR_VALUE_COUNT = 3, DERR_COMPRESSED = 4,
  DERR_OVERRUN = 5, DERR_DICT_INDEX = 6, DERR_UNSUPPORTED_TYPE = 7, DERR_SNAPPY = 8, DERR_STRING_TOO_LONG = 9
};
// BYTE_ARRAY dictionary pages -> tables of string references (device_utils.cuh: string_ref): one job per dictionary page,
// walked by one thread (the entries are length-prefixed, so their positions are only found sequentially)
struct StringDictJob {
  const uint8_t* page;   // PLAIN byte arrays: [u32 length][bytes] ...
  uint64_t* refs;        // count references out
  int32_t size, count;
};
void launch_build_string_dicts(hs_ctx* ctx, const StringDictJob* jobs, int64_t n, uint32_t* d_error);

// ---- snappy (snappy.cu) ---------------------------------------------------------------------------------------------
struct SnappyBlob {
  const uint8_t* src;   // stored bytes (device)
  uint64_t dst_off;     // offset of the decompressed bytes in the scratch buffer
  uint32_t src_len;
  uint32_t dst_len;     // prefix + decompressed length
  uint32_t prefix;      // leading bytes copied verbatim (v2 level bytes)
  uint32_t compressed;  // 0: copy, 1: snappy
  uint32_t first_block; // index of this page's first 64 KB output block in the block table (ascending over the blobs)
  uint32_t pad;
};
__host__ __device__ inline uint32_t snappy_blocks_of(uint32_t dst_len, uint32_t prefix) {
  const uint32_t body = dst_len - prefix;
  return body == 0 ? 1u : (body + 65535u) / 65536u;
}
// block_in: one uint32 per block (+1), sequential: one uint32 per blob -- scratch of the two launches
// any_verbatim: some blob has a prefix or is stored uncompressed
void launch_snappy_decompress(hs_ctx* ctx, const SnappyBlob* blobs, int64_t n, int64_t total_blocks, bool any_verbatim,
                              uint32_t* block_in, uint32_t* sequential, uint8_t* scratch, uint32_t* d_error);

// compression of page bodies: one warp per fragment (<= 65536 bytes) of a page; fragment f of raw bytes [src_off, src_off +
// len) is written to scratch at dst_off (room for 32 + len + len / 6 bytes), its compressed length to out_len[f]
struct SnappyFragment {
  uint64_t src_off, dst_off;
  uint32_t len, pad;
};
constexpr uint32_t kSnappyFragment = 65536;
inline uint64_t snappy_max_compressed(uint64_t len) { return 32 + len + len / 6; }
void launch_snappy_compress(hs_ctx* ctx, const SnappyFragment* frags, int64_t n, const uint8_t* raw, uint8_t* scratch,
                            uint32_t* out_len);

// Walks the page headers of every chunk.  mode 0: page_counts[chunk] = number of data pages.
// mode 1: fills pages[page_offsets[chunk] ...].
void launch_walk_pages(hs_ctx* ctx, const ChunkDesc* chunks, int n_chunks, int32_t* page_counts,
                       const int64_t* page_offsets, PageDesc* pages, uint32_t* d_error, int mode);
// Decodes all pages into the column arrays.  row_window (optional, device, 2 x int64 per file: [lo, hi) global rows)
// restricts decoding to pages that intersect the window of their file.
void launch_decode_pages(hs_ctx* ctx, const PageDesc* pages, int64_t n_pages, const ColumnOut* cols,
                         uint32_t* col_has_nulls, const int64_t* row_window, uint32_t* d_error);

// ---- hash / partition (hash_partition.cu) ---------------------------------------------------------------------------
struct KeyColumn {
  const void* data;
  const uint8_t* valid;  // nullptr -> no nulls
  int32_t type;
  int32_t width;
  // zero-copy column (data == nullptr): where each partition tile's values lie inside the source images.  Only a single
  // int32 / int64 key without nulls may come this way (the specialised hash kernels handle it).
  const ZcTile* tiles = nullptr;
};

constexpr int kPartTile = 4096;   // rows per partition tile (256 threads x 16)
constexpr int kMaxBuckets = 4096;

// bucket id per row + per-tile histograms M[tile][nb] + global histogram + OR/AND of the first key's sort encoding
void launch_bucket_hist(hs_ctx* ctx, const KeyColumn* d_keys, int nkeys, int64_t nrows, int num_buckets,
                        uint16_t* bucket, uint32_t* tile_hist, unsigned long long* global_hist);
// same, but bins are the owner ranks (bucket % world) -- the map side of the multi-GPU exchange
void launch_owner_hist(hs_ctx* ctx, const KeyColumn* d_keys, int nkeys, int64_t nrows, int num_buckets, int world,
                       uint16_t* owner, uint32_t* tile_hist, unsigned long long* global_hist);
// in-place: tile_hist[t][b] <- bucket_base[b] + sum_{t'<t} tile_hist[t'][b]   (bucket_base = exclusive scan of global hist)
// explicit_base (optional, device, nb entries): start of every bucket's destination instead of the local scan
void launch_tile_offsets(hs_ctx* ctx, uint32_t* tile_hist, int64_t ntiles, int num_buckets,
                         const unsigned long long* global_hist, unsigned long long* bucket_offsets /* nb+1 or null */,
                         const unsigned long long* explicit_base = nullptr);
// dest[row] = stable position of the row in bucket-major order
void launch_partition_dest(hs_ctx* ctx, const uint16_t* bucket, int64_t nrows, int num_buckets,
                           const uint32_t* tile_offsets, uint32_t* dest);
// out[dest[i]] = in[i]
void launch_scatter_column(hs_ctx* ctx, const void* in, void* out, const uint32_t* dest, int64_t nrows, int width);
// ---- fused partition (hash + stable rank + shared-memory exchange of every column in one kernel) ------------------
constexpr int kFusedTileLocal = 4096;  // rows per tile when the partition writes local memory (256 threads x 16)
constexpr int kFusedTilePeer = 8192;   // ... and when it writes peer GPUs' memory over NVLink (512 threads x 16)
int fused_tile_rows(bool peer_tiles);
constexpr int kFusedMaxBins = 1024;  // above this the per-warp counters no longer fit next to the exchange buffer
struct PartColumn {
  const void* in;
  void* out;
  int32_t width;  // 8, 4 or 1 (validity bytes travel as width-1 columns)
  int32_t pad;
  const ZcTile* tiles = nullptr;  // zero-copy source (in == nullptr), see KeyColumn::tiles
};
bool fused_partition_supported(int nbins);
// bin_ids (optional): receives every row's bin so that launch_partition_rows (same argument) need not hash again
// key_or_and (optional, {0, ~0} on entry): accumulates OR / AND of the sort-encoded values of the last key column
// tile histograms M[tile][bin] for fused_tile_rows(peer_tiles)-row tiles (+ global histogram); bin = bucket, or bucket % owner_mod
void launch_tile_hist(hs_ctx* ctx, const KeyColumn* d_keys, int nkeys, int64_t nrows, int num_buckets, int owner_mod,
                      uint32_t* tile_hist, unsigned long long* global_hist,
                      unsigned long long* key_or_and = nullptr, int single_key_type = -1, uint16_t* bin_ids = nullptr,
                      bool peer_tiles = false);
// single_key_type: HS_TYPE_INT32 / HS_TYPE_INT64 when there is exactly one key column, of that type and without nulls
// (selects a kernel with the hash inlined for it); -1 otherwise.  See single_key_type_of().
inline int single_key_type_of(const KeyColumn* h_keys, int nkeys) {
  return nkeys == 1 && h_keys[0].valid == nullptr && (h_keys[0].type == 0 || h_keys[0].type == 1) ? h_keys[0].type : -1;
}
// tile_dst = launch_tile_offsets(tile_hist); moves all columns into bin-major order, stable
// d_peer_out (optional): [ncols][out_world] peer-mapped output pointers; bucket b is written to GPU b % out_world
// pack (optional, single GPU only): one more round that reads up to four 16-bit code columns and writes them as ONE
// 8-byte record per row (slot s in bits [16 s, 16 s + 16)) -- the layout k_dict_pack_all gathers from
struct CodePackRound {
  const uint16_t* src[4];
  void* out;   // nrows x 8 bytes
  int32_t n;   // code columns in use (0: no such round)
  int32_t pad;
};
void launch_partition_rows(hs_ctx* ctx, const KeyColumn* d_keys, int nkeys, int64_t nrows, int num_buckets, int owner_mod,
                           const uint32_t* tile_dst, const PartColumn* d_cols, int ncols, void* const* d_peer_out = nullptr,
                           int out_world = 1, int single_key_type = -1, const CodePackRound* pack = nullptr,
                           const uint16_t* bin_ids = nullptr);
// out[i] = sort_encode(in[src ? src[i] : i])  (+ global OR / AND reduction into or_and[0], or_and[1])
void launch_encode_keys(hs_ctx* ctx, const void* in, int type, const uint32_t* src, int64_t nrows, uint64_t* out,
                        unsigned long long* or_and);
void launch_iota_u32(hs_ctx* ctx, uint32_t* out, int64_t n);
// out[i] = piece of the string refs[perm[i]]: its length (piece < 0) or its piece-th 8 bytes, big-endian, zero-padded
// (+ OR / AND reduction); see k_string_piece_keys
void launch_string_piece_keys(hs_ctx* ctx, const uint64_t* refs, const uint32_t* perm, int64_t nrows, int piece, uint64_t* out,
                              unsigned long long* or_and);

// ---- segmented radix sort (radix_sort.cu) ---------------------------------------------------------------------------
constexpr int kSortTile = 4096;  // pairs per tile (256 threads x 16)
struct SortTile {
  uint32_t seg;
  uint32_t count;
  uint64_t start;  // global position of the tile's first pair
};
struct SortPlan {
  int64_t n = 0;
  int64_t ntiles = 0;
  int32_t nseg = 0;
  Buf<SortTile> tiles;          // device
  Buf<uint32_t> seg_tile_begin; // device, nseg+1: first tile of each segment
  Buf<uint64_t> seg_start;      // device, nseg+1: global start of each segment
  Buf<uint32_t> tile_hist;      // device, ntiles x 256
  Buf<uint32_t> tile_dst;       // device, ntiles x 256: destination of every (tile, digit) for the current pass
  std::vector<uint32_t> h_seg_tile_begin;  // host mirror of seg_tile_begin
};
// seg_offsets: host array of nseg+1 global offsets
void build_sort_plan(hs_ctx* ctx, const uint64_t* seg_offsets, int nseg, SortPlan* plan);
// Stable LSD radix sort of (key, val) pairs within each segment on the key bits set in `bit_mask` (bytes whose bits
// are all constant are skipped).  Result is left in (keys, vals); (keys_alt, vals_alt) are scratch of the same size.
// first_pass_source (optional): the pairs have not been materialised yet -- the first pass that runs reads the raw key
// column (position p holds the value of row p) and uses p itself as the row index.  bit_mask must then select at least
// one byte.
struct RawKeyColumn {
  const void* data;
  int type, width;
};
void segmented_sort_pairs(hs_ctx* ctx, SortPlan* plan, uint64_t*& keys, uint64_t*& keys_alt, uint32_t*& vals,
                          uint32_t*& vals_alt, uint64_t bit_mask, const RawKeyColumn* first_pass_source = nullptr);
// After sorting on the bits of high_mask: stable insertion sort of every run of equal (key & high_mask) on (key & low_mask);
// *d_flag is set when a run is longer than max_run (the caller then runs the remaining passes instead).
void launch_fix_runs(hs_ctx* ctx, SortPlan* plan, uint64_t* keys, uint32_t* vals, uint64_t high_mask, uint64_t low_mask,
                     uint32_t max_run, uint32_t* d_flag);
// One extra stable pass on an external 8-bit digit: digit = digits[vals[i]]  (null flags for nullable 64-bit keys)
void segmented_sort_pass_by_table(hs_ctx* ctx, SortPlan* plan, uint64_t*& keys, uint64_t*& keys_alt, uint32_t*& vals,
                                  uint32_t*& vals_alt, const uint8_t* digits);

// ---- gather + Parquet encode (gather_encode.cu) -----------------------------------------------------------------
struct GatherColumn {
  const void* src;        // partitioned column values
  const uint64_t* sorted_keys;  // when non-null: take the value from the sorted encoded keys (decode with key_type)
  int32_t key_type;
  int32_t width;
  const uint64_t* page_value_offset;  // device: arena offset of the first value byte of each (bucket, page)
};
// For every sorted position p (tile by tile): value = src[perm[p]] written PLAIN into its page body in the arena.
// page index of local row lr in bucket b = bucket_page_begin[b] + lr / rows_per_page.
void launch_gather_encode(hs_ctx* ctx, const SortTile* tiles, int64_t ntiles, const uint64_t* seg_start,
                          const uint32_t* perm, const GatherColumn& col, const uint32_t* bucket_page_begin,
                          int64_t rows_per_page, uint8_t* arena);
// nullable columns: per-tile non-null counts, then bit-packed definition levels + dense values per tile
void launch_tile_valid_counts(hs_ctx* ctx, const SortTile* tiles, int64_t ntiles, const uint32_t* perm,
                              const uint8_t* valid, uint32_t* counts);
void launch_gather_encode_nullable(hs_ctx* ctx, const SortTile* tiles, int64_t ntiles, const uint32_t* perm,
                                   const void* src, const uint8_t* valid, int width, const uint64_t* tile_value_offset,
                                   const uint64_t* tile_def_offset, uint8_t* arena);
// string columns (8-byte references, device_utils.cuh): per tile the bytes its non-null values take as PLAIN BYTE_ARRAY
// ([u32 length][bytes] each) and their number; then definition bits + values per tile.  valid == nullptr: no nulls
void launch_tile_string_sizes(hs_ctx* ctx, const SortTile* tiles, int64_t ntiles, const uint32_t* perm, const uint64_t* refs,
                              const uint8_t* valid, uint32_t* bytes, uint32_t* counts);
void launch_gather_encode_strings(hs_ctx* ctx, const SortTile* tiles, int64_t ntiles, const uint32_t* perm, const uint64_t* refs,
                                  const uint8_t* valid, const uint64_t* tile_value_offset, const uint64_t* tile_def_offset,
                                  uint8_t* arena);
// ---- dictionary encoding (dict_encode.cu) -------------------------------------------------------------------------
constexpr uint32_t kMaxDictEntries = 65536;          // bit width <= 16
constexpr uint32_t kDictCapacity = 8 * kMaxDictEntries;  // open-addressing slots (power of two)
// inserts the distinct raw values of src[begin, end) into the hash set `keys` (capacity slots preset to all-ones);
// state[0] = distinct count, state[1] = 1 when more than max_distinct values were seen, state[2] = the all-ones value occurs
void launch_dict_build(hs_ctx* ctx, const void* src, int width, int64_t begin, int64_t end, unsigned long long* keys,
                       uint32_t capacity, uint32_t max_distinct, uint32_t* state);
// same hash set, filled from the dictionary pages of the decoded source chunks of projected column `col`
void launch_dict_build_from_pages(hs_ctx* ctx, const PageDesc* pages, int64_t n_pages, int col, int width,
                                  unsigned long long* keys, uint32_t capacity, uint32_t max_distinct, uint32_t* state);
// compacts the distinct values out of the hash set (counter must be zeroed); then slot -> rank in the sorted dictionary
// (at most max_out values are written; the counter still counts all of them)
void launch_dict_collect(hs_ctx* ctx, const unsigned long long* keys, uint32_t capacity, unsigned long long* out,
                         uint32_t* counter, uint32_t max_out = 0xffffffffu);
// entries: capacity x 16 bytes {key lo, key hi, dictionary index, 0}
// all dictionary columns of a table in one map + one pack launch (up to 8 columns per call)
struct DictMapArgs {
  const void* src[8];
  const void* entries[8];  // 16-byte {key, index} hash-table entries (open addressing, linear probing, dict_hash_u64)
  uint32_t mask[8];        // table capacity - 1 (a power of two sized to the dictionary, so that the table stays in L1)
  uint32_t empty_index[8];
  int32_t width[8];
  int32_t ncols;
};
struct DictPackArgs {
  const uint64_t* page_value_offset[8];
  uint32_t bw[8];
  int32_t ncols;
};
// rec_scratch: nrows records of 4 (ncols <= 4) or 8 uint16 indices
// bit-packs the codes of already mapped records (4 or 8 uint16 slots per row) into the dictionary-encoded data pages
void launch_dict_pack(hs_ctx* ctx, const SortTile* tiles, int64_t ntiles, const uint64_t* seg_start, const uint32_t* perm,
                      const DictPackArgs& pack_args, int slots, const uint16_t* rec, const uint32_t* bucket_page_begin,
                      int64_t rows_per_page, uint8_t* arena);
void launch_dict_encode_all(hs_ctx* ctx, const SortTile* tiles, int64_t ntiles, const uint64_t* seg_start, const uint32_t* perm,
                            const DictMapArgs& map_args, const DictPackArgs& pack_args, int64_t nrows, uint32_t capacity,
                            uint16_t* rec_scratch, const uint32_t* bucket_page_begin, int64_t rows_per_page, uint8_t* arena);
// Plain gather: out[i] = src[perm[i]]
void launch_gather_plain(hs_ctx* ctx, const void* src, const uint32_t* perm, int64_t n, int width, void* out);
struct StatPatch {        // min/max of the sorted key column of one row group
  uint64_t first_pos, last_pos;  // sorted positions of the row group's first and last row
  uint64_t min_off[2], max_off[2];  // arena offsets of the footer placeholders
  int32_t width, pad;
};
void launch_patch_key_stats(hs_ctx* ctx, const StatPatch* patches, int64_t n, const uint64_t* sorted_keys, int key_type,
                            uint8_t* arena);
struct ByteCopy {
  uint64_t dst;  // arena offset
  uint32_t src;  // offset into the skeleton byte stream
  uint32_t len;
};
void launch_scatter_bytes(hs_ctx* ctx, const ByteCopy* copies, int64_t n, const uint8_t* skeleton, uint8_t* arena);
// larger pieces at any alignment (compressed page fragments moving into their place in the file): one CTA per blob
struct BlobCopy {
  uint64_t src, dst;  // byte offsets into src_base / dst_base
  uint32_t len, pad;
};
void launch_copy_blobs(hs_ctx* ctx, const BlobCopy* blobs, int64_t n, const uint8_t* src_base, uint8_t* dst_base);
// Synthetic table generator: rows [first_row, first_row+n) of column `col` (0..4) of table T
void launch_synth_column(hs_ctx* ctx, int col, int64_t first_row, int64_t n, void* out);

// ---- read side (read_side.cu) ---------------------------------------------------------------------------
// per segment s: bounds[2s] = first index with key >= lo, bounds[2s+1] = first index with key > hi (segment-relative)
void launch_range_bounds(hs_ctx* ctx, const int64_t* keys, const uint64_t* seg_offsets, int nseg, int has_lo,
                         int64_t lo, int has_hi, int64_t hi, int64_t* bounds);
// match counts of every left row against the right rows of the same bucket
// string_keys: lkeys / rkeys hold string references (device_utils.cuh) compared in byte order
void launch_join_count(hs_ctx* ctx, const int64_t* lkeys, const uint64_t* lseg, const int64_t* rkeys,
                       const uint64_t* rseg, int nseg, int64_t nl, uint32_t* counts, uint32_t* first_match, bool string_keys = false);
void launch_join_emit(hs_ctx* ctx, const uint32_t* counts, const uint32_t* first_match, const uint64_t* out_offsets,
                      int64_t nl, uint32_t* out_li, uint32_t* out_ri);
// exclusive scan of uint32 counts into uint64 offsets (n+1 entries; last = total)
void exclusive_scan_u32_u64(hs_ctx* ctx, const uint32_t* in, int64_t n, uint64_t* out);
// mask[i] = lo <= keys[i] <= hi (and file id not deleted); compaction index list
void launch_filter_mask(hs_ctx* ctx, const int64_t* keys, const uint8_t* valid, int64_t n, int has_lo, int64_t lo,
                        int has_hi, int64_t hi, uint32_t* mask);
// string / binary keys (values are references into the source images, device_utils.cuh: string_ref): the bounds are
// references to device copies of the bound bytes; order = unsigned byte order, shorter first on a common prefix
void launch_range_bounds_strings(hs_ctx* ctx, const uint64_t* refs, const uint64_t* seg_offsets, int nseg, int has_lo,
                                 uint64_t lo_ref, int has_hi, uint64_t hi_ref, int64_t* bounds);
void launch_filter_mask_strings(hs_ctx* ctx, const uint64_t* refs, const uint8_t* valid, int64_t n, int has_lo,
                                uint64_t lo_ref, int has_hi, uint64_t hi_ref, uint32_t* mask);
// lens[i] = length of refs[idx[i]] (0 for a null);  then, with offsets = exclusive scan of lens: out[offsets[i] ..] = bytes
void launch_string_lengths(hs_ctx* ctx, const uint64_t* refs, const uint8_t* valid, const uint32_t* idx, int64_t n,
                           uint32_t* lens);
void launch_copy_strings(hs_ctx* ctx, const uint64_t* refs, const uint8_t* valid, const uint32_t* idx, int64_t n,
                         const uint64_t* offsets, uint8_t* out);
void launch_compact_indices(hs_ctx* ctx, const uint32_t* mask, const uint64_t* offsets, int64_t n, uint32_t* out_idx);
void launch_not_in_mask(hs_ctx* ctx, const int64_t* file_ids, int64_t n, const int64_t* deleted, int ndeleted,
                        uint32_t* mask /* and-ed in place */);

}  // namespace hs
