// snappy.cu -- Snappy (raw block format) page decompression on the GPU.
//
// Spark writes Parquet with the SNAPPY codec by default, so the source tables (and index files written by the reference)
// that users hand to createIndex / refreshIndex / the index scans are usually snappy-compressed; the reference reads them
// through parquet-mr + snappy-java (index/covering/CoveringIndexTrait.scala:82-84, CoveringIndexRuleUtils.scala:113-123).
// One warp per compressed page: the warp keeps a 1 KB window of the compressed stream in shared memory, lane 0 parses the
// element tags from the window, and all 32 lanes execute each literal / back-reference copy (overlapping copies repeat
// the pattern of period `offset`).  Pages are independent, so thousands of warps run concurrently.
#include "device_utils.cuh"
#include "kernels.h"

namespace hs {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr uint32_t kWindow = 1024;

__global__ void __launch_bounds__(kWarpsPerCta * 32) k_snappy_decompress(const SnappyBlob* __restrict__ blobs, int64_t n,
                                                                          uint8_t* __restrict__ scratch,
                                                                          uint32_t* __restrict__ d_error) {
  __shared__ uint8_t s_win[kWarpsPerCta][kWindow];
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t w = (int64_t)blockIdx.x * kWarpsPerCta + wib;
  if (w >= n) return;
  const SnappyBlob b = blobs[w];
  const uint8_t* src = b.src;
  uint8_t* dst = scratch + b.dst_off;
  uint32_t n_src = b.src_len, dst_len = b.dst_len;
  // verbatim prefix (data page v2: repetition / definition levels are stored uncompressed in front of the values)
  for (uint32_t j = lane; j < b.prefix; j += 32) dst[j] = src[j];
  src += b.prefix;
  dst += b.prefix;
  n_src -= b.prefix;
  dst_len -= b.prefix;
  if (!b.compressed) {  // stored page of a compressed chunk (v2 is_compressed = false)
    for (uint32_t j = lane; j < n_src && j < dst_len; j += 32) dst[j] = src[j];
    return;
  }
  uint8_t* win = s_win[wib];
  uint32_t win_base = 0;
  auto refill = [&](uint32_t base) {
    win_base = base;
    __syncwarp();
    for (uint32_t j = lane; j < kWindow && base + j < n_src; j += 32) win[j] = src[base + j];
    __syncwarp();
  };
  refill(0);
  // preamble: uncompressed length as a varint
  uint32_t pos = 0, ulen = 0;
  {
    int shift = 0;
    while (pos < n_src && pos < 5) {
      const uint8_t c = win[pos++];
      ulen |= (uint32_t)(c & 0x7f) << shift;
      if (!(c & 0x80)) break;
      shift += 7;
    }
  }
  if (ulen != dst_len) {
    if (lane == 0) atomicCAS(d_error, 0u, ((uint32_t)DERR_SNAPPY << 24) | 1u);
    return;
  }
  uint32_t out = 0;
  while (out < dst_len && pos < n_src) {
    if (pos + 8 > win_base + kWindow && win_base + kWindow < n_src) refill(pos);
    // lane 0 parses one element: literal (tag & 3 == 0) or copy with 1 / 2 / 4-byte offset
    uint32_t len = 0, offset = 0, hdr = 0;
    if (lane == 0) {
      const uint8_t* t = win + (pos - win_base);
      const uint32_t tag = t[0];
      const uint32_t kind = tag & 3;
      if (kind == 0) {
        uint32_t l = tag >> 2;
        hdr = 1;
        if (l >= 60) {
          const uint32_t nb = l - 59;  // 1..4 length bytes follow
          l = 0;
          for (uint32_t i = 0; i < nb; i++) l |= (uint32_t)t[1 + i] << (8 * i);
          hdr = 1 + nb;
        }
        len = l + 1;
      } else if (kind == 1) {
        len = ((tag >> 2) & 7) + 4;
        offset = ((tag >> 5) << 8) | t[1];
        hdr = 2;
      } else if (kind == 2) {
        len = (tag >> 2) + 1;
        offset = (uint32_t)t[1] | ((uint32_t)t[2] << 8);
        hdr = 3;
      } else {
        len = (tag >> 2) + 1;
        offset = (uint32_t)t[1] | ((uint32_t)t[2] << 8) | ((uint32_t)t[3] << 16) | ((uint32_t)t[4] << 24);
        hdr = 5;
      }
    }
    len = __shfl_sync(0xffffffffu, len, 0);
    offset = __shfl_sync(0xffffffffu, offset, 0);
    hdr = __shfl_sync(0xffffffffu, hdr, 0);
    if (len > dst_len - out) {
      if (lane == 0) atomicCAS(d_error, 0u, ((uint32_t)DERR_SNAPPY << 24) | 2u);
      return;
    }
    if (offset == 0) {  // literal
      if (pos + hdr + len > n_src) {
        if (lane == 0) atomicCAS(d_error, 0u, ((uint32_t)DERR_SNAPPY << 24) | 3u);
        return;
      }
      const uint8_t* lit = src + pos + hdr;
      uint8_t* o = dst + out;
      // byte copies up to a 4-byte boundary of the destination, then words assembled from the (unaligned) source
      for (uint32_t j = lane; j < len; j += 32) o[j] = lit[j];
      pos += hdr + len;
    } else {
      if (offset > out) {
        if (lane == 0) atomicCAS(d_error, 0u, ((uint32_t)DERR_SNAPPY << 24) | 4u);
        return;
      }
      uint8_t* o = dst + out;
      const uint8_t* from = o - offset;
      if (offset >= len) {
        for (uint32_t j = lane; j < len; j += 32) o[j] = from[j];
      } else {  // overlapping copy: the last `offset` bytes repeat
        for (uint32_t j = lane; j < len; j += 32) o[j] = from[j % offset];
      }
      pos += hdr;
    }
    out += len;
    __syncwarp();  // the bytes just written may be the source of the next back-reference
  }
  if (out != dst_len && lane == 0) atomicCAS(d_error, 0u, ((uint32_t)DERR_SNAPPY << 24) | 5u);
}

}  // namespace

void launch_snappy_decompress(hs_ctx* ctx, const SnappyBlob* blobs, int64_t n, uint8_t* scratch, uint32_t* d_error) {
  KernelScope _ks(ctx, "k_snappy_decompress");
  if (n == 0) return;
  k_snappy_decompress<<<(unsigned)ceil_div(n, kWarpsPerCta), kWarpsPerCta * 32, 0, ctx->stream>>>(blobs, n, scratch, d_error);
  HS_LAUNCH_CHECK(ctx);
}

}  // namespace hs

// =====================================================================================================================
// Snappy compression of index pages (raw block format), one warp per 64 KB fragment of a page.
//
// Spark writes its Parquet -- and therefore the reference's index files -- with the SNAPPY codec by default
// (index/DataFrameWriterExtensions.scala:59-66 goes through DataFrameWriter; file names ...c000.snappy.parquet,
// T/index/VacuumOutdatedActionTest.scala:67).  A page's compressed body is the varint of its uncompressed length followed by
// the element streams of its fragments, which are compressed independently (back-references never leave a fragment, as in
// snappy's own 64 KB blocks) and therefore in parallel.
//
// Per fragment: the 32 lanes probe 32 consecutive positions at a time against a 2048-entry hash table of earlier positions
// (4-byte hashes); the first lane whose candidate matches wins, the literal before it is flushed, the match is extended 32
// bytes per ballot and emitted as copies with 2-byte offsets.  Windows without a match make the stride grow (snappy's own
// heuristic), so incompressible data costs little more than the literal copy.  The table takes atomicMax updates: the
// output is the same on every run.
// =====================================================================================================================
namespace hs {
namespace {

constexpr int kCompWarps = 4;
constexpr uint32_t kCompTable = 2048;

__device__ __forceinline__ uint32_t load32_any(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// literal [from, to) of `in` -> out; returns the new output position (all lanes take part)
__device__ __forceinline__ uint32_t emit_literal(const uint8_t* __restrict__ in, uint32_t from, uint32_t to, uint8_t* __restrict__ out,
                                                 uint32_t op, unsigned lane) {
  const uint32_t len = to - from;
  if (len == 0) return op;
  uint32_t hdr;
  if (len <= 60) {
    if (lane == 0) out[op] = (uint8_t)((len - 1) << 2);
    hdr = 1;
  } else if (len <= 256) {
    if (lane == 0) {
      out[op] = (uint8_t)(60 << 2);
      out[op + 1] = (uint8_t)(len - 1);
    }
    hdr = 2;
  } else {  // len <= 65536 (a fragment)
    if (lane == 0) {
      out[op] = (uint8_t)(61 << 2);
      out[op + 1] = (uint8_t)((len - 1) & 0xff);
      out[op + 2] = (uint8_t)((len - 1) >> 8);
    }
    hdr = 3;
  }
  for (uint32_t j = lane; j < len; j += 32) out[op + hdr + j] = in[from + j];
  return op + hdr + len;
}

__global__ void __launch_bounds__(kCompWarps * 32) k_snappy_compress(const SnappyFragment* __restrict__ frags, int64_t n,
                                                                      const uint8_t* __restrict__ raw, uint8_t* __restrict__ scratch,
                                                                      uint32_t* __restrict__ out_len) {
  __shared__ uint32_t s_table[kCompWarps][kCompTable];
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t f = (int64_t)blockIdx.x * kCompWarps + wib;
  if (f >= n) return;
  const SnappyFragment fr = frags[f];
  const uint8_t* in = raw + fr.src_off;
  uint8_t* out = scratch + fr.dst_off;
  const uint32_t len = fr.len;
  uint32_t* table = s_table[wib];
  for (uint32_t i = lane; i < kCompTable; i += 32) table[i] = 0;
  __syncwarp();
  uint32_t ip = 0, lit = 0, op = 0, misses = 0;
  while (ip + 4 <= len) {
    const uint32_t pos = ip + lane;
    const bool can = pos + 4 <= len;
    uint32_t w = 0, h = 0, cand = 0;
    if (can) {
      w = load32_any(in + pos);
      h = (w * 0x1e35a7bdu) >> 21;  // 11 bits
      cand = table[h];
    }
    __syncwarp();
    const bool match = can && cand < pos && load32_any(in + cand) == w;
    const unsigned m = __ballot_sync(0xffffffffu, match);
    // only positions up to the match enter the table: the scan resumes right behind the match, and an entry that points
    // past the scan position can never be a candidate (it would shadow the useful, earlier one)
    const int fl = m ? __ffs(m) - 1 : 31;
    if (can && (int)lane <= fl) atomicMax(&table[h], pos);
    __syncwarp();
    if (m == 0) {
      misses++;
      ip += 32u << min(misses >> 2, 4u);
      continue;
    }
    misses = 0;
    const uint32_t q = ip + fl;
    const uint32_t c = __shfl_sync(0xffffffffu, cand, fl);
    op = emit_literal(in, lit, q, out, op, lane);
    // extend the match 32 bytes at a time
    uint32_t mlen = 4;
    for (;;) {
      const uint32_t a = q + mlen + lane;
      const bool eq = a < len && in[a] == in[c + mlen + lane];
      const unsigned e = __ballot_sync(0xffffffffu, eq);
      const uint32_t run = e == 0xffffffffu ? 32u : (uint32_t)(__ffs(~e) - 1);
      mlen += run;
      if (run < 32) break;
    }
    const uint32_t offset = q - c;
    if (mlen <= 11 && offset < 2048) {  // short match nearby: the 2-byte copy element (4..11 bytes, 11-bit offset)
      if (lane == 0) {
        out[op] = (uint8_t)(1u | ((mlen - 4) << 2) | ((offset >> 8) << 5));
        out[op + 1] = (uint8_t)(offset & 0xff);
      }
      op += 2;
    } else {
      if (lane == 0) {  // copies with a 2-byte offset carry 1..64 bytes each
        uint32_t left = mlen, o = op;
        while (left > 0) {
          const uint32_t l = left > 64 ? 64 : left;
          out[o] = (uint8_t)(2u | ((l - 1) << 2));
          out[o + 1] = (uint8_t)(offset & 0xff);
          out[o + 2] = (uint8_t)(offset >> 8);
          o += 3;
          left -= l;
        }
      }
      op += 3 * ((mlen + 63) / 64);
    }
    ip = q + mlen;
    lit = ip;
    __syncwarp();
  }
  op = emit_literal(in, lit, len, out, op, lane);
  if (lane == 0) out_len[f] = op;
}

}  // namespace

void launch_snappy_compress(hs_ctx* ctx, const SnappyFragment* frags, int64_t n, const uint8_t* raw, uint8_t* scratch,
                            uint32_t* out_len) {
  KernelScope _ks(ctx, "k_snappy_compress");
  if (n == 0) return;
  k_snappy_compress<<<(unsigned)ceil_div(n, kCompWarps), kCompWarps * 32, 0, ctx->stream>>>(frags, n, raw, scratch, out_len);
  HS_LAUNCH_CHECK(ctx);
}

}  // namespace hs
