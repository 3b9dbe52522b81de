// Spatial tiling of a coordinate set for the output-stationary sparse convolution (conv_umma.cu).
//
// The reference processes rulebook pairs in whatever order the voxels arrive (indice.cu.h:171-208), so its
// gather/scatter kernels (reordering.cu.h:21-157) touch feature rows at random.  Here every coordinate set is
// sorted ONCE along a Morton (Z-order) curve; convolutions then walk the output rows in tiles of 128 that are
// compact surface patches, which turns the 9-27x re-reads of the gather into L1/L2 hits.
#include "common.cuh"

namespace wsis {

constexpr int kTile = 128;

__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// key = batch : morton(x>>s, y>>s, z>>s); `cbits` coordinate bits per axis after the shift
__global__ void morton_key_kernel(const int32_t *__restrict__ coords, int64_t N, int shift, int cbits,
                                  uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + i);
  uint32_t m = spread3((uint32_t)c.y >> shift) << 2 | spread3((uint32_t)c.z >> shift) << 1 | spread3((uint32_t)c.w >> shift);
  keys[i] = ((uint32_t)c.x << (3 * cbits)) | m;
  vals[i] = (uint32_t)i;
}

__global__ void pad_order_kernel(int32_t *order, int64_t N, int64_t n_pad) {
  int64_t i = N + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) order[i] = -1;
}

__global__ void iota_kernel(int32_t *order, int64_t N, int64_t n_pad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) order[i] = i < N ? (int32_t)i : -1;
}

// ------------------------------------------------------------------------------------------------
// Tile records: what one conv CTA needs to know about one tile of 128 destination rows, compacted.
//   [0, 16K)                 valid[K][4]  u32   bit r of valid[k] = tile slot r has a source row through offset k
//   [16K, 16K + hdr2)        start[K+1]   u16   entries of offset k are [start[k], start[k+1])
//   [hdr, hdr + 4P)          idx[P]       i32   source row of each entry (P = start[K] <= 128K)
//   [hdr + 4P, hdr + 5P)     slot[P]      u8    tile slot (accumulator lane) of each entry, ascending inside an offset
// Records live at a fixed stride (worst case P = 128K) so they are built in one pass without a size scan; a CTA
// copies only rec_bytes[t] = align16(hdr + 5P) of it.  On a surface sampled like a scan P is ~4-10 per row instead
// of K = 27, so the map traffic of a layer drops from 108 B to ~25-50 B per row and absent (row, offset) slots cost
// the conv kernel nothing.
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kTile) tile_record_kernel(const int32_t *__restrict__ map, int K, int flip,
                                                            const int32_t *__restrict__ order, uint8_t *__restrict__ recs,
                                                            int32_t *__restrict__ rec_bytes) {
  extern __shared__ int32_t s_map[];  // [K][128] transposed slice, then the K+1 prefix
  int32_t *s_cnt = s_map + K * kTile;
  const int64_t t = blockIdx.x;
  const int r = threadIdx.x, lane = r & 31, w = r >> 5;
  const int32_t dst = __ldg(order + t * kTile + r);
  for (int k = 0; k < K; ++k)
    s_map[k * kTile + r] = dst >= 0 ? __ldg(map + (int64_t)dst * K + (flip ? K - 1 - k : k)) : -1;
  __syncthreads();
  uint8_t *rec = recs + t * (int64_t)rec_stride_bytes(K);
  uint32_t *valid = reinterpret_cast<uint32_t *>(rec);
  uint16_t *start = reinterpret_cast<uint16_t *>(rec + 16 * K);
  // per-offset population (one warp ballot per 32 slots)
  for (int k = 0; k < K; ++k) {
    const uint32_t bal = __ballot_sync(0xffffffffu, s_map[k * kTile + r] >= 0);
    if (lane == 0) {
      valid[k * 4 + w] = bal;
      s_cnt[k * 4 + w] = __popc(bal);
    }
  }
  __syncthreads();
  if (r == 0) {  // exclusive prefix over (offset, warp) -- K*4 <= 128 values
    int run = 0;
    for (int i = 0; i < K * 4; ++i) {
      int c = s_cnt[i];
      s_cnt[i] = run;
      run += c;
    }
    s_cnt[K * 4] = run;
  }
  __syncthreads();
  const int P = s_cnt[K * 4];
  if (r <= K) start[r] = (uint16_t)(r < K ? s_cnt[r * 4] : P);
  int32_t *idx = reinterpret_cast<int32_t *>(rec + rec_hdr_bytes(K));
  uint8_t *slot = rec + rec_hdr_bytes(K) + 4 * P;
  for (int k = 0; k < K; ++k) {
    const int32_t v = s_map[k * kTile + r];
    const uint32_t bal = __ballot_sync(0xffffffffu, v >= 0);
    if (v >= 0) {
      const int pos = s_cnt[k * 4 + w] + __popc(bal & ((1u << lane) - 1u));
      idx[pos] = v;
      slot[pos] = (uint8_t)r;
    }
  }
  if (r == 0) rec_bytes[t] = (rec_hdr_bytes(K) + 5 * P + 15) & ~15;
}

static int ilog2_ceil(int64_t v) {
  int b = 0;
  while (((int64_t)1 << b) < v) ++b;
  return b;
}

}  // namespace wsis

using namespace wsis;

extern "C" {

int64_t wsis_tile_pad(int64_t n) { return ceil_div(n, kTile) * kTile; }

int64_t wsis_spatial_order_ws_bytes(int64_t N) {
  int64_t a = (N * 4 + 255) / 256 * 256;
  return 3 * a + wsis_sort_ws_bytes(N);
}

int wsis_spatial_order(const int32_t *coords, int64_t N, const int32_t spatial_shape[3], int batch_size,
                       int32_t *order, void *ws, wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  WSIS_CHECK(N >= 0 && N < ((int64_t)1 << 31), "spatial_order: N out of range");
  WSIS_CHECK(batch_size >= 1 && batch_size <= 65536, "spatial_order: bad batch size %d", batch_size);
  const int64_t n_pad = wsis_tile_pad(N);
  if (N == 0) return 0;
  int ext = std::max(spatial_shape[0], std::max(spatial_shape[1], spatial_shape[2]));
  WSIS_CHECK(ext >= 1 && ext <= 65536, "spatial_order: bad spatial shape");
  int bbits = ilog2_ceil(batch_size);
  int cbits = std::min(ilog2_ceil(ext), 10), shift = ilog2_ceil(ext) - cbits;
  while (3 * cbits + bbits > 32) {  // coarsen the cells until batch:morton fits one 32-bit key
    --cbits;
    ++shift;
  }
  int64_t a = (N * 4 + 255) / 256 * 256;
  char *p = reinterpret_cast<char *>(ws);
  uint32_t *k0 = reinterpret_cast<uint32_t *>(p), *v0 = reinterpret_cast<uint32_t *>(p + a),
           *k1 = reinterpret_cast<uint32_t *>(p + 2 * a);
  void *sort_ws = p + 3 * a;
  morton_key_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(coords, N, shift, cbits, k0, v0);
  WSIS_LAUNCH_OK();
  if (wsis_sort_pairs_u32(k0, v0, k1, reinterpret_cast<uint32_t *>(order), N, 0, 3 * cbits + bbits, sort_ws, stream))
    return 1;
  if (n_pad > N) {
    pad_order_kernel<<<1, kTile, 0, st>>>(order, N, n_pad);
    WSIS_LAUNCH_OK();
  }
  return 0;
}

int wsis_identity_order(int64_t N, int32_t *order, wsis_stream_t stream) {
  const int64_t n_pad = wsis_tile_pad(N);
  if (n_pad == 0) return 0;
  iota_kernel<<<(unsigned)ceil_div(n_pad, 256), 256, 0, as_stream(stream)>>>(order, N, n_pad);
  WSIS_LAUNCH_OK();
  return 0;
}

int64_t wsis_tile_record_stride(int K) { return rec_stride_bytes(K); }

int wsis_tile_records(const int32_t *map, int64_t n_rows, int K, int flip, const int32_t *order, void *records,
                      int32_t *rec_bytes, wsis_stream_t stream) {
  const int64_t n_tiles = wsis_tile_pad(n_rows) / kTile;
  if (n_tiles == 0) return 0;
  WSIS_CHECK(K >= 1 && K <= 32, "tile_records: kernel volume %d not in [1,32]", K);
  WSIS_CHECK((reinterpret_cast<uintptr_t>(records) & 15) == 0, "tile_records: records must be 16-byte aligned");
  size_t smem = ((size_t)K * kTile + 4 * K + 1) * sizeof(int32_t);
  tile_record_kernel<<<(unsigned)n_tiles, kTile, smem, as_stream(stream)>>>(map, K, flip, order, (uint8_t *)records,
                                                                             rec_bytes);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
