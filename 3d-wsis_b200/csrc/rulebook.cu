// GPU hash-table rulebook ("indice pair") builder.
//
// Replaces getIndicePair<3> (include/spconv/spconv_ops.h:27-137) and its kernels
// (include/spconv/indice.cu.h:22-65,112-148 strided; :150-208 submanifold).  The reference looks voxels
// up in a dense int32 grid of batch*D*H*W cells that it allocates and fills with -1 on every call
// (spconv_ops.h:60-62: 61 MB for a 400x300x128 scene); here a 64-bit-key open-addressing hash table of
// 2N..4N slots (a few MB, L2 resident) is used instead, so memory and time scale with the number of active
// voxels, not with the bounding-box volume.
//
// Outputs are deterministic: outputs of a strided conv are numbered in the CPU reference's first-touch
// order (geometry.h:181-187) via an atomicMin on (input row, kernel offset) + a prefix sum, and the
// reference-format pair lists are emitted in ascending input-row order inside each offset
// (geometry.h:176-190) -- one valid instance of the GPU reference's atomics-dependent order
// (indice.cu.h:57,202).
#include <limits.h>

#include "common.cuh"

namespace wsis {

struct Geo {
  int ks[3], st[3], pad[3], dil[3], shape[3];
  int K;
};

__global__ void hash_insert_rows(const int32_t *__restrict__ coords, int64_t N, unsigned long long *keys,
                                 int32_t *vals, int64_t mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int4 c = reinterpret_cast<const int4 *>(coords)[i];
  int64_t s = hash_insert(keys, mask, pack_key(c.x, c.y, c.z, c.w));
  atomicMax(vals + s, (int32_t)i);  // duplicates: largest row wins (geometry.h:272-277 on the CPU)
}

// one thread per (input row, kernel offset): nbr_in[i,k] = row of the voxel at p + pad - k*dil, or -1
__global__ void subm_nbr_kernel(const int32_t *__restrict__ coords, int64_t N, Geo g,
                                const unsigned long long *__restrict__ keys, const int32_t *__restrict__ vals,
                                int64_t mask, int32_t *__restrict__ nbr_in) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * g.K) return;
  int64_t i = idx / g.K;
  int k = (int)(idx - i * g.K);
  int kz = k % g.ks[2], ky = (k / g.ks[2]) % g.ks[1], kx = k / (g.ks[2] * g.ks[1]);
  int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + i);
  int ox = c.y + g.pad[0] - kx * g.dil[0];
  int oy = c.z + g.pad[1] - ky * g.dil[1];
  int oz = c.w + g.pad[2] - kz * g.dil[2];
  int32_t r = -1;
  if (ox >= 0 && ox < g.shape[0] && oy >= 0 && oy < g.shape[1] && oz >= 0 && oz < g.shape[2]) {
    int64_t s = hash_find(keys, mask, pack_key(c.x, ox, oy, oz));
    if (s >= 0) r = __ldg(vals + s);
  }
  nbr_in[idx] = r;
}

// strided conv, pass A: insert every reachable output cell, remember its first toucher (i*K+k)
__global__ void conv_touch_kernel(const int32_t *__restrict__ coords, int64_t N, Geo g, unsigned long long *keys,
                                  int32_t *first, int64_t mask, int32_t *__restrict__ nbr_in) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * g.K) return;
  int64_t i = idx / g.K;
  int k = (int)(idx - i * g.K);
  int kk[3] = {k / (g.ks[2] * g.ks[1]), (k / g.ks[2]) % g.ks[1], k % g.ks[2]};
  int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + i);
  int p[3] = {c.y, c.z, c.w};
  int o[3];
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    int num = p[d] + g.pad[d] - kk[d] * g.dil[d];
    int q = num / g.st[d];
    ok = ok && num >= 0 && q * g.st[d] == num && q < g.shape[d];
    o[d] = q;
  }
  int32_t slot = -1;
  if (ok) {
    int64_t s = hash_insert(keys, mask, pack_key(c.x, o[0], o[1], o[2]));
    atomicMin(first + s, (int32_t)idx);
    slot = (int32_t)s;
  }
  nbr_in[idx] = slot;
}

__global__ void conv_flag_kernel(const int32_t *__restrict__ nbr_in, const int32_t *__restrict__ first, int64_t NK,
                                 int32_t *__restrict__ flag) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NK) return;
  int32_t s = nbr_in[idx];
  flag[idx] = (s >= 0 && first[s] == (int32_t)idx) ? 1 : 0;
}

// pass C: first touchers publish rank + coordinates of their output voxel
__global__ void conv_assign_kernel(const int32_t *__restrict__ nbr_in, const int32_t *__restrict__ first,
                                   const unsigned long long *__restrict__ keys, const int32_t *__restrict__ rank,
                                   int64_t NK, int32_t *__restrict__ slot_rank, int32_t *__restrict__ out_coords) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NK) return;
  int32_t s = nbr_in[idx];
  if (s < 0 || first[s] != (int32_t)idx) return;
  int32_t r = rank[idx];
  slot_rank[s] = r;
  unsigned long long key = keys[s];
  int4 oc = make_int4((int)(key >> 48) & 0xFFFF, (int)(key >> 32) & 0xFFFF, (int)(key >> 16) & 0xFFFF,
                      (int)key & 0xFFFF);
  reinterpret_cast<int4 *>(out_coords)[r] = oc;
}

// pass D: slot ids -> output rows, and the transposed map
__global__ void conv_link_kernel(int32_t *__restrict__ nbr_in, const int32_t *__restrict__ slot_rank, int64_t N,
                                 int K, int32_t *__restrict__ nbr_out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * K) return;
  int32_t s = nbr_in[idx];
  if (s < 0) return;
  int64_t i = idx / K;
  int k = (int)(idx - i * K);
  int32_t r = slot_rank[s];
  nbr_in[idx] = r;
  nbr_out[(int64_t)r * K + k] = (int32_t)i;
}

// ---- reference-format pairs -------------------------------------------------------------------
constexpr int kTileRows = 32;

// pos[k*N + i] = (nbr_in[i,k] >= 0), transposed through shared memory so both sides stay coalesced
__global__ void pairs_flag_kernel(const int32_t *__restrict__ nbr_in, int64_t N, int K, int32_t *__restrict__ pos) {
  extern __shared__ int32_t sm[];
  int64_t i0 = (int64_t)blockIdx.x * kTileRows;
  int rows = (int)min((int64_t)kTileRows, N - i0);
  for (int e = threadIdx.x; e < rows * K; e += blockDim.x) sm[e] = nbr_in[i0 * K + e];
  __syncthreads();
  for (int e = threadIdx.x; e < kTileRows * K; e += blockDim.x) {
    int k = e / kTileRows, r = e % kTileRows;
    if (r < rows) pos[(int64_t)k * N + i0 + r] = sm[r * K + k] >= 0 ? 1 : 0;
  }
}

__global__ void pairs_write_kernel(const int32_t *__restrict__ nbr_in, int64_t N, int K,
                                   const int32_t *__restrict__ pos, int32_t *__restrict__ pairs,
                                   int32_t *__restrict__ num) {
  extern __shared__ int32_t sm[];
  int64_t i0 = (int64_t)blockIdx.x * kTileRows;
  int rows = (int)min((int64_t)kTileRows, N - i0);
  for (int e = threadIdx.x; e < rows * K; e += blockDim.x) sm[e] = nbr_in[i0 * K + e];
  __syncthreads();
  for (int e = threadIdx.x; e < kTileRows * K; e += blockDim.x) {
    int k = e / kTileRows, r = e % kTileRows;
    if (r >= rows) continue;
    int32_t v = sm[r * K + k];
    if (v < 0) continue;
    int32_t j = pos[(int64_t)k * N + i0 + r] - __ldg(pos + (int64_t)k * N);
    pairs[((int64_t)k * 2 + 0) * N + j] = (int32_t)(i0 + r);
    pairs[((int64_t)k * 2 + 1) * N + j] = v;
  }
  if (blockIdx.x == 0)
    for (int k = threadIdx.x; k < K; k += blockDim.x) num[k] = pos[(int64_t)(k + 1) * N] - pos[(int64_t)k * N];
}

__global__ void nbr_from_pairs_kernel(const int32_t *__restrict__ pairs, const int32_t *__restrict__ num,
                                      int64_t stride, int K, int dst_side, int32_t *__restrict__ map) {
  int k = blockIdx.y;
  int32_t n = num[k];
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    int32_t i = pairs[((int64_t)k * 2 + 0) * stride + j];
    int32_t o = pairs[((int64_t)k * 2 + 1) * stride + j];
    if (dst_side)
      map[(int64_t)o * K + k] = i;
    else
      map[(int64_t)i * K + k] = o;
  }
}

static int fill_geo(Geo &g, const int32_t *ks, const int32_t *st, const int32_t *pad, const int32_t *dil,
                    const int32_t *shape) {
  g.K = 1;
  for (int d = 0; d < 3; ++d) {
    g.ks[d] = ks[d];
    g.st[d] = st ? st[d] : 1;
    g.pad[d] = pad ? pad[d] : ks[d] / 2;  // subm: padding := ksize/2 (spconv_ops.h:74-77)
    g.dil[d] = dil[d];
    g.shape[d] = shape[d];
    WSIS_CHECK(ks[d] >= 1 && g.st[d] >= 1 && g.dil[d] >= 1 && shape[d] >= 1 && shape[d] <= 65536,
               "rulebook: bad geometry in dim %d", d);
    WSIS_CHECK(g.st[d] == 1 || g.dil[d] == 1, "rulebook: stride>1 with dilation>1 is not supported (conv.py:80-81)");
    g.K *= ks[d];
  }
  WSIS_CHECK(g.K <= 256, "rulebook: kernel volume %d > 256 (spconv_ops.h:50)", g.K);
  return 0;
}

}  // namespace wsis

using namespace wsis;

extern "C" {

int64_t wsis_hash_slots(int64_t n) {
  int64_t s = 1024;
  while (s < 2 * n) s <<= 1;
  return s;
}

int wsis_rulebook_subm(const int32_t *coords, int64_t N, const int32_t ksize[3], const int32_t dilation[3],
                       const int32_t spatial_shape[3], uint64_t *hash_keys, int32_t *hash_vals, int64_t slots,
                       int32_t *nbr_in, wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  Geo g;
  if (fill_geo(g, ksize, nullptr, nullptr, dilation, spatial_shape)) return 1;
  WSIS_CHECK(slots >= 2 * N && (slots & (slots - 1)) == 0, "rulebook: slots must be a power of two >= 2N");
  WSIS_CHECK(N * g.K < ((int64_t)1 << 31), "rulebook: N*K overflows int32");
  if (N == 0) return 0;
  WSIS_CUDA(cudaMemsetAsync(hash_keys, 0xFF, slots * sizeof(uint64_t), st));
  WSIS_CUDA(cudaMemsetAsync(hash_vals, 0xFF, slots * sizeof(int32_t), st));
  hash_insert_rows<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(coords, N, (unsigned long long *)hash_keys, hash_vals,
                                                                slots - 1);
  WSIS_LAUNCH_OK();
  subm_nbr_kernel<<<(unsigned)ceil_div(N * g.K, 256), 256, 0, st>>>(
      coords, N, g, (const unsigned long long *)hash_keys, hash_vals, slots - 1, nbr_in);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_rulebook_conv_count(const int32_t *coords, int64_t N, const int32_t ksize[3], const int32_t stride[3],
                             const int32_t padding[3], const int32_t dilation[3], const int32_t out_shape[3],
                             uint64_t *hash_keys, int32_t *hash_vals, int64_t slots, int32_t *nbr_in,
                             int32_t *rank_ws, void *scan_ws, int32_t *n_out_dev, wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  Geo g;
  if (fill_geo(g, ksize, stride, padding, dilation, out_shape)) return 1;
  WSIS_CHECK((slots & (slots - 1)) == 0, "rulebook: slots must be a power of two");
  WSIS_CHECK(N * g.K < ((int64_t)1 << 31), "rulebook: N*K overflows int32");
  // an input reaches at most prod_d ceil(ks_d / stride_d) distinct outputs (taps with the right residue)
  int64_t tpi = 1;
  for (int d = 0; d < 3; ++d) tpi *= (g.ks[d] + g.st[d] - 1) / g.st[d];
  WSIS_CHECK(slots >= 2 * N * tpi, "rulebook: hash table too small (need >= %lld slots)", (long long)(2 * N * tpi));
  if (N == 0) {
    WSIS_CUDA(cudaMemsetAsync(n_out_dev, 0, sizeof(int32_t), st));
    return 0;
  }
  int64_t NK = N * g.K;
  WSIS_CUDA(cudaMemsetAsync(hash_keys, 0xFF, slots * sizeof(uint64_t), st));
  if (wsis_fill_i32(hash_vals, slots, INT_MAX, stream)) return 1;
  conv_touch_kernel<<<(unsigned)ceil_div(NK, 256), 256, 0, st>>>(coords, N, g, (unsigned long long *)hash_keys,
                                                                 hash_vals, slots - 1, nbr_in);
  WSIS_LAUNCH_OK();
  conv_flag_kernel<<<(unsigned)ceil_div(NK, 256), 256, 0, st>>>(nbr_in, hash_vals, NK, rank_ws);
  WSIS_LAUNCH_OK();
  if (wsis_exclusive_scan_i32(rank_ws, rank_ws, NK, scan_ws, stream)) return 1;
  WSIS_CUDA(cudaMemcpyAsync(n_out_dev, rank_ws + NK, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int wsis_rulebook_conv_fill(const int32_t *coords, int64_t N, int K, const uint64_t *hash_keys,
                            const int32_t *hash_vals, int64_t slots, int32_t *nbr_in, const int32_t *rank_ws,
                            int32_t *slot_rank, int64_t n_out, int32_t *out_coords, int32_t *nbr_out,
                            wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  (void)coords;
  (void)slots;
  if (N == 0 || n_out == 0) return 0;
  int64_t NK = N * K;
  WSIS_CUDA(cudaMemsetAsync(nbr_out, 0xFF, n_out * K * sizeof(int32_t), st));
  conv_assign_kernel<<<(unsigned)ceil_div(NK, 256), 256, 0, st>>>(nbr_in, hash_vals, (const unsigned long long *)hash_keys,
                                                                  rank_ws, NK, slot_rank, out_coords);
  WSIS_LAUNCH_OK();
  conv_link_kernel<<<(unsigned)ceil_div(NK, 256), 256, 0, st>>>(nbr_in, slot_rank, N, K, nbr_out);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_pairs_from_nbr(const int32_t *nbr_in, int64_t N, int K, int32_t *pairs, int32_t *num, int32_t *pos_ws,
                        void *scan_ws, wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  WSIS_CHECK(K >= 1 && K <= 256, "pairs: bad K");
  if (N == 0) {
    WSIS_CUDA(cudaMemsetAsync(num, 0, K * sizeof(int32_t), st));
    return 0;
  }
  size_t smem = (size_t)kTileRows * K * sizeof(int32_t);
  unsigned blocks = (unsigned)ceil_div(N, kTileRows);
  if (wsis_fill_i32(pairs, (int64_t)K * 2 * N, -1, stream)) return 1;
  pairs_flag_kernel<<<blocks, 256, smem, st>>>(nbr_in, N, K, pos_ws);
  WSIS_LAUNCH_OK();
  if (wsis_exclusive_scan_i32(pos_ws, pos_ws, (int64_t)K * N, scan_ws, stream)) return 1;
  pairs_write_kernel<<<blocks, 256, smem, st>>>(nbr_in, N, K, pos_ws, pairs, num);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_nbr_from_pairs(const int32_t *pairs, const int32_t *num, int64_t pair_stride, int K, int dst_side,
                        int32_t *map, wsis_stream_t stream) {
  if (pair_stride == 0) return 0;
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(pair_stride, 256), 1024), (unsigned)K);
  nbr_from_pairs_kernel<<<grid, 256, 0, as_stream(stream)>>>(pairs, num, pair_stride, K, dst_side, map);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
