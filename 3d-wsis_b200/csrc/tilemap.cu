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
// Tile records: what one conv CTA needs to know about one tile of 128 destination rows.
//   record (records + t * rec_stride_bytes(K), all of it meaningful)
//     [0, 16K)                 valid[K][4]  u32   bit r of valid[k] = tile slot r has a source row through offset k
//     [16K, 16K + 80)          nU u32 | amask u32 | P u32 | npack u32 | members u16[32]
//                              A PACK is one or two active offsets whose valid-slot sets are disjoint (greedy first
//                              fit in ascending offset order): members[i] = k0 | k1 << 8 (0xFF = none).  The conv
//                              kernel assembles ONE operand block per pack (every slot takes the row of whichever
//                              member it has a neighbour through) and multiplies it once per member with that
//                              member's lane mask and weights: a surface tile has 27 active offsets but ~16 packs.
//     [16K + 80, ...+256K)     loc[K][128]  u16   row i < npack: index of slot r's source row, through the pack's
//                                                 member that reaches it, in the tile's list of DISTINCT source rows
//                                                 (0xFFFF where the slot has none); rows >= npack are unused
//   unique rows (uidx + t * 128K)  i32[nU]        the distinct source rows the tile reads; a surface patch of 128
//                                                 voxels reads ~160-230 distinct rows through ~480-1400 entries, so the
//                                                 conv kernel fetches (and converts) every source row ONCE per tile
//   meta[t] = {record bytes, nU, active-offset mask (1 if the tile has no entry at all), P}
// The dense [K][128] form is what the conv kernel's operand builders want: builder thread r owns tile slot r (= the
// TMEM lane of that row) and reads loc[k][r] with a conflict-free 16-bit load.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t tile_hash(int32_t v, uint32_t mask) {
  return ((uint32_t)v ^ ((uint32_t)v >> 13)) & mask;  // keeps runs of consecutive rows in order
}

__global__ void __launch_bounds__(kTile) tile_record_kernel(const int32_t *__restrict__ map, int K, int flip, int HT,
                                                            const int32_t *__restrict__ order, uint8_t *__restrict__ recs,
                                                            int32_t *__restrict__ uidx, int4 *__restrict__ meta,
                                                            int32_t *__restrict__ stats) {
  extern __shared__ int32_t s_map[];  // [K][128] transposed slice
  int32_t *s_cnt = s_map + K * kTile;                                   // [K] entries per offset
  int32_t *s_tab = s_cnt + K + 8;                                       // [HT] open-addressing set of source rows
  int32_t *s_first = s_tab + HT;                                        // [HT] first entry (k * 128 + slot) of the row
  uint32_t *s_bits = reinterpret_cast<uint32_t *>(s_first + HT);        // [4 K] bitmap of first entries
  int32_t *s_pre = reinterpret_cast<int32_t *>(s_bits + 4 * K);         // [4 K + 1] exclusive prefix of its popcounts
  uint32_t *s_vm = reinterpret_cast<uint32_t *>(s_pre + 4 * K + 4);     // [4 K] valid-slot masks of the offsets
  int32_t *s_packof = reinterpret_cast<int32_t *>(s_vm + 4 * K);        // [K] pack of each offset, [K] = npack
  uint16_t *s_pos = reinterpret_cast<uint16_t *>(s_packof + K + 4);     // [K][128] table slot of each entry
  const int64_t t = blockIdx.x;
  const int r = threadIdx.x, lane = r & 31, w = r >> 5;
  const int32_t dst = __ldg(order + t * kTile + r);
  for (int k = 0; k < K; ++k)
    s_map[k * kTile + r] = dst >= 0 ? __ldg(map + (int64_t)dst * K + (flip ? K - 1 - k : k)) : -1;
  for (int i = r; i < HT; i += kTile) {
    s_tab[i] = -1;
    s_first[i] = 0x7fffffff;
  }
  for (int i = r; i < 4 * K; i += kTile) s_bits[i] = 0u;
  if (r < K) s_cnt[r] = 0;
  __syncthreads();
  uint8_t *rec = recs + t * (int64_t)rec_stride_bytes(K);
  uint32_t *valid = reinterpret_cast<uint32_t *>(rec);
  // per-offset population (one warp ballot per 32 slots), insertion of the source rows into the set, and the FIRST
  // entry (in offset-major, slot-minor order) that reads each row
  for (int k = 0; k < K; ++k) {
    const int32_t v = s_map[k * kTile + r];
    const uint32_t bal = __ballot_sync(0xffffffffu, v >= 0);
    if (lane == 0) {
      valid[k * 4 + w] = bal;
      s_vm[k * 4 + w] = bal;
      if (bal) atomicAdd(s_cnt + k, __popc(bal));
    }
    if (v >= 0) {
      uint32_t h = tile_hash(v, HT - 1);
      while (true) {
        const int32_t prev = atomicCAS(s_tab + h, -1, v);
        if (prev == -1 || prev == v) break;
        h = (h + 1) & (HT - 1);
      }
      s_pos[k * kTile + r] = (uint16_t)h;
      atomicMin(s_first + h, k * kTile + r);
    }
  }
  __syncthreads();
  // Distinct rows are numbered in first-entry order: consecutive slots that read new rows through an offset get
  // consecutive indices, so the 32 lanes of a conv builder warp (consecutive slots, one offset) read row-cache rows
  // whose indices mostly differ in their low bits -> different shared-memory bank groups (conv_umma.cu, rc_swz).
  for (int k = 0; k < K; ++k) {
    const bool first = s_map[k * kTile + r] >= 0 && s_first[s_pos[k * kTile + r]] == k * kTile + r;
    const uint32_t bal = __ballot_sync(0xffffffffu, first);
    if (lane == 0) s_bits[k * 4 + w] = bal;
  }
  __syncthreads();
  if (w == 0) {  // exclusive prefix over the 4 K <= 128 bitmap words
    int run = 0;
    for (int i0 = 0; i0 < 4 * K; i0 += 32) {
      const int i = i0 + lane;
      const int c = i < 4 * K ? __popc(s_bits[i]) : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      if (i < 4 * K) s_pre[i] = run + inc - c;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_pre[4 * K] = run;
  } else if (w == 1) {
    // packs: lane i keeps pack i (mask of occupied slots, member count, members); offsets are placed first fit
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0, cnt = 0, mem = 0xFFFFu;
    int npack = 0;
    for (int k = 0; k < K; ++k) {
      const uint32_t v0 = s_vm[4 * k], v1 = s_vm[4 * k + 1], v2 = s_vm[4 * k + 2], v3 = s_vm[4 * k + 3];
      if ((v0 | v1 | v2 | v3) == 0) {
        if (lane == 0) s_packof[k] = -1;
        continue;
      }
      const bool fit = lane < npack && cnt < 2 && (((m0 & v0) | (m1 & v1) | (m2 & v2) | (m3 & v3)) == 0);
      const uint32_t bal = __ballot_sync(0xffffffffu, fit);
      const int tgt = bal ? __ffs(bal) - 1 : npack;
      if (!bal) ++npack;
      if (lane == tgt) {
        m0 |= v0, m1 |= v1, m2 |= v2, m3 |= v3;
        mem = cnt == 0 ? (0xFF00u | (uint32_t)k) : ((mem & 0xFFu) | ((uint32_t)k << 8));
        ++cnt;
      }
      if (lane == 0) s_packof[k] = tgt;
    }
    if (npack == 0) npack = 1;  // a tile without any entry still sends one all-lanes-off unit through the pipeline
    reinterpret_cast<uint16_t *>(rec + 16 * K + 16)[lane] = lane < npack ? (uint16_t)mem : (uint16_t)0xFFFFu;
    if (lane == 0) s_packof[K] = npack;
  }
  __syncthreads();
  const int nU = s_pre[4 * K];
  const int npack = s_packof[K];
  int32_t *u = uidx + t * (int64_t)(kTile * K);
  uint16_t *loc = reinterpret_cast<uint16_t *>(rec + rec_hdr_bytes(K));
  for (int i = 0; i < npack; ++i) loc[i * kTile + r] = 0xFFFFu;
  for (int k = 0; k < K; ++k) {
    const int32_t v = s_map[k * kTile + r];
    if (v >= 0) {
      const int key = s_first[s_pos[k * kTile + r]];  // first entry of this row
      const int rank = s_pre[key >> 5] + __popc(s_bits[key >> 5] & ((1u << (key & 31)) - 1u));
      if (key == k * kTile + r) u[rank] = v;
      loc[s_packof[k] * kTile + r] = (uint16_t)rank;  // the members of a pack reach disjoint slots
    }
  }
  if (r == 0) {
    uint32_t am = 0;
    int P = 0;
    for (int k = 0; k < K; ++k) {
      if (s_cnt[k] > 0) am |= 1u << k;
      P += s_cnt[k];
    }
    if (am == 0) am = 1u;
    *reinterpret_cast<uint4 *>(rec + 16 * K) = make_uint4((uint32_t)nU, am, (uint32_t)P, (uint32_t)npack);
    meta[t] = make_int4(rec_stride_bytes(K), nU, npack, P);
    atomicMax(stats, P);  // most entries / most distinct rows of a tile of the map
    atomicMax(stats + 1, nU);
  }
}

// Records of the identity rulebook (K = 1, map[r] = r, natural order): what a 1x1 convolution looks like to conv_umma.
// Same format as tile_record_kernel writes, without the hash set: slot r of tile t reads row t*128 + r.
__global__ void __launch_bounds__(kTile) tile_record_identity_kernel(int64_t n, uint8_t *__restrict__ recs,
                                                                     int32_t *__restrict__ uidx, int4 *__restrict__ meta,
                                                                     int32_t *__restrict__ order, int32_t *__restrict__ stats) {
  const int64_t t = blockIdx.x;
  const int r = threadIdx.x, lane = r & 31, w = r >> 5;
  const int64_t row = t * kTile + r;
  const bool valid = row < n;
  order[row] = valid ? (int32_t)row : -1;
  uint8_t *rec = recs + t * (int64_t)rec_stride_bytes(1);
  const uint32_t bal = __ballot_sync(0xffffffffu, valid);
  if (lane == 0) reinterpret_cast<uint32_t *>(rec)[w] = bal;
  const int nU = (int)min((int64_t)kTile, n - t * kTile);
  uidx[t * kTile + r] = valid ? (int32_t)row : 0;
  reinterpret_cast<uint16_t *>(rec + rec_hdr_bytes(1))[r] = valid ? (uint16_t)r : (uint16_t)0xFFFFu;
  if (r < 32) reinterpret_cast<uint16_t *>(rec + 16 + 16)[r] = r == 0 ? (uint16_t)0xFF00u : (uint16_t)0xFFFFu;
  if (r == 0) {
    *reinterpret_cast<uint4 *>(rec + 16) = make_uint4((uint32_t)nU, 1u, (uint32_t)nU, 1u);
    meta[t] = make_int4(rec_stride_bytes(1), nU, 1, nU);
    if (t == 0) stats[0] = stats[1] = nU;          // most entries / most distinct rows of a tile (tile 0 is the fullest)
  }
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

int wsis_tile_records_identity(int64_t n_rows, void *records, int32_t *uidx, int32_t *meta, int32_t *stats, int32_t *order,
                               wsis_stream_t stream) {
  const int64_t n_tiles = wsis_tile_pad(n_rows) / kTile;
  if (n_tiles == 0) return 0;
  tile_record_identity_kernel<<<(unsigned)n_tiles, kTile, 0, as_stream(stream)>>>(
      n_rows, (uint8_t *)records, uidx, reinterpret_cast<int4 *>(meta), order, stats);
  WSIS_LAUNCH_OK();
  return 0;
}

int64_t wsis_tile_record_stride(int K) { return rec_stride_bytes(K); }

int64_t wsis_tile_unique_stride(int K) { return (int64_t)kTile * K; }

static int tile_table_slots(int K) {
  // A tile has at most 128 K distinct source rows (every entry a different row: random maps); a surface patch has
  // ~200.  The table is sized for the worst case at a load factor <= 0.89 (K = 27: 4096 slots, 32 KB for table + first
  // entries) rather than <= 0.5: halving it takes the kernel from 2 to 4 resident CTAs per SM, and the linear probe is
  // only long in the adversarial case.
  int ht = 256;
  while (ht < kTile * K + kTile * K / 8) ht <<= 1;
  return ht;
}

int wsis_tile_records(const int32_t *map, int64_t n_rows, int K, int flip, const int32_t *order, void *records,
                      int32_t *uidx, int32_t *meta, int32_t *stats, wsis_stream_t stream) {
  const int64_t n_tiles = wsis_tile_pad(n_rows) / kTile;
  WSIS_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int32_t), as_stream(stream)));
  if (n_tiles == 0) return 0;
  WSIS_CHECK(K >= 1 && K <= 32, "tile_records: kernel volume %d not in [1,32]", K);
  WSIS_CHECK(((reinterpret_cast<uintptr_t>(records) | reinterpret_cast<uintptr_t>(uidx) |
               reinterpret_cast<uintptr_t>(meta)) & 15) == 0,
             "tile_records: records/uidx/meta must be 16-byte aligned");
  const int HT = tile_table_slots(K);
  size_t smem = ((size_t)K * kTile + K + 8 + 2 * HT + 8 * K + 8 + 4 * K + K + 4) * sizeof(int32_t) +
                (size_t)K * kTile * sizeof(uint16_t);
  // the opt-in is per device: set it on every call rather than caching it process-wide (multi-GPU processes)
  WSIS_CUDA(cudaFuncSetAttribute(tile_record_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tile_record_kernel<<<(unsigned)n_tiles, kTile, smem, as_stream(stream)>>>(
      map, K, flip, HT, order, (uint8_t *)records, uidx, reinterpret_cast<int4 *>(meta), stats);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
