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
