// Voxelization (the `pointgroup_ops` contract) -- written from the call sites, the upstream source is not in
// the reference tree (README.md:37-41):
//   voxelization_idx(coords int64[N,4], batchsize, mode) -> (voxel_locs int64[M,4], p2v int32[N],
//                                                            v2p int32[M,1+maxActive])      scannetv2_dataset.py:449
//   voxelization(feats f32[N,C], v2p, mode=4 mean) -> f32[M,C]   (+ backward)             train_scannetv2.py:189
// Voxels are numbered in first-occurrence order over the point list and v2p lists points in ascending order,
// on the host path and on the device path alike, so both give identical tensors.
#include <limits.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace wsis {

// ------------------------------------------------------------------------------------------------
// host path (DataLoader workers: CUDA-free, fork-safe, no global state)
// ------------------------------------------------------------------------------------------------
struct HostSlot {
  int64_t b, x, y, z;
  int32_t voxel;
};

static inline uint64_t host_hash(const int64_t *c) {
  uint64_t h = mix64((uint64_t)c[0] + 0x9E3779B97F4A7C15ull);
  h = mix64(h ^ (uint64_t)c[1]);
  h = mix64(h ^ (uint64_t)c[2]);
  return mix64(h ^ (uint64_t)c[3]);
}

// ------------------------------------------------------------------------------------------------
// device path
// ------------------------------------------------------------------------------------------------
struct VoxWs {
  unsigned long long *keys;
  int32_t *first, *slot_of, *rank, *cnt, *off, *iota, *sorted_pts;
  uint32_t *sorted_keys;
  void *scan_ws, *sort_ws;
  int64_t slots, bytes;
};

static inline int64_t al(int64_t x) { return (x + 255) / 256 * 256; }

static VoxWs carve(void *ws, int64_t N) {
  VoxWs w;
  char *p = reinterpret_cast<char *>(ws);
  int64_t n1 = N > 0 ? N : 1;
  w.slots = wsis_hash_slots(n1);
  auto take = [&](int64_t bytes) {
    char *q = p;
    p += al(bytes);
    return q;
  };
  w.keys = (unsigned long long *)take(w.slots * 8);
  w.first = (int32_t *)take(w.slots * 4);
  w.slot_of = (int32_t *)take(n1 * 4);
  w.rank = (int32_t *)take((n1 + 1) * 4);
  w.cnt = (int32_t *)take((n1 + 1) * 4);
  w.off = (int32_t *)take((n1 + 2) * 4);
  w.iota = (int32_t *)take(n1 * 4);
  w.sorted_pts = (int32_t *)take(n1 * 4);
  w.sorted_keys = (uint32_t *)take(n1 * 4);
  w.scan_ws = take(wsis_scan_ws_bytes(n1 + 1));
  w.sort_ws = take(wsis_sort_ws_bytes(n1));
  w.bytes = p - reinterpret_cast<char *>(ws);
  return w;
}

__global__ void vox_touch_kernel(const int64_t *__restrict__ coords, int64_t N, unsigned long long *keys,
                                 int32_t *first, int64_t mask, int32_t *__restrict__ slot_of, int32_t *err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const longlong2 *c2 = reinterpret_cast<const longlong2 *>(coords) + i * 2;
  longlong2 a = __ldg(c2), b = __ldg(c2 + 1);
  if ((unsigned long long)a.x > 65535ull || (unsigned long long)a.y > 65535ull || (unsigned long long)b.x > 65535ull ||
      (unsigned long long)b.y > 65535ull)
    atomicExch(err, 1);
  int64_t s = hash_insert(keys, mask, pack_key((int)a.x, (int)a.y, (int)b.x, (int)b.y));
  atomicMin(first + s, (int32_t)i);
  slot_of[i] = (int32_t)s;
}

__global__ void vox_flag_kernel(const int32_t *__restrict__ slot_of, const int32_t *__restrict__ first, int64_t N,
                                int32_t *__restrict__ flag, int32_t *__restrict__ iota) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  flag[i] = first[slot_of[i]] == (int32_t)i;
  iota[i] = (int32_t)i;
}

__global__ void vox_assign_kernel(const int32_t *__restrict__ slot_of, const int32_t *__restrict__ first,
                                  const int32_t *__restrict__ rank, int64_t N, int32_t *__restrict__ p2v,
                                  int32_t *cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int32_t m = rank[first[slot_of[i]]];
  p2v[i] = m;
  atomicAdd(cnt + m, 1);
}

__global__ void vox_max_kernel(const int32_t *__restrict__ cnt, int64_t N, int32_t *out_max) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int v = i < N ? cnt[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out_max, v);
}

__global__ void vox_fill_kernel(const int64_t *__restrict__ coords, const int32_t *__restrict__ cnt,
                                const int32_t *__restrict__ off, const int32_t *__restrict__ sorted_pts, int64_t M,
                                int32_t stride, int64_t *__restrict__ voxel_locs, int32_t *__restrict__ v2p) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * stride) return;
  int64_t m = idx / stride;
  int j = (int)(idx - m * stride);
  int32_t n = cnt[m], o = off[m];
  int32_t v = 0;
  if (j == 0) {
    v = n;
    int32_t p0 = sorted_pts[o];
#pragma unroll
    for (int d = 0; d < 4; ++d) voxel_locs[m * 4 + d] = coords[(int64_t)p0 * 4 + d];
  } else if (j - 1 < n) {
    v = sorted_pts[o + j - 1];
  }
  v2p[idx] = v;
}

// one warp per voxel; lanes over channels (C is small: 6)
__global__ void vox_mean_fwd_kernel(const float *__restrict__ feats, const int32_t *__restrict__ v2p, int64_t M,
                                    int32_t stride, int C, float *__restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * C) return;
  int64_t m = idx / C;
  int c = (int)(idx - m * C);
  const int32_t *row = v2p + m * stride;
  int32_t n = __ldg(row);
  float w = n > 0 ? 1.0f / (float)n : 0.f;
  float acc = 0.f;
  for (int j = 0; j < n; ++j) acc += w * __ldg(feats + (int64_t)__ldg(row + 1 + j) * C + c);
  out[idx] = acc;
}

__global__ void vox_mean_bwd_kernel(const float *__restrict__ dout, const int32_t *__restrict__ v2p, int64_t M,
                                    int32_t stride, int C, float *__restrict__ dfeats) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * C) return;
  int64_t m = idx / C;
  int c = (int)(idx - m * C);
  const int32_t *row = v2p + m * stride;
  int32_t n = __ldg(row);
  float g = n > 0 ? (1.0f / (float)n) * dout[idx] : 0.f;
  for (int j = 0; j < n; ++j) dfeats[(int64_t)__ldg(row + 1 + j) * C + c] = g;
}

}  // namespace wsis

using namespace wsis;

extern "C" {

int64_t wsis_voxelize_idx_host(const int64_t *coords, int64_t N, int64_t *voxel_locs, int32_t *p2v, int32_t *v2p,
                               int32_t v2p_stride, int32_t *max_active) {
  if (N < 0 || N >= ((int64_t)1 << 31)) {
    set_error("voxelize_idx: N out of range");
    return -1;
  }
  size_t cap = 64;
  while (cap < (size_t)(2 * N + 1)) cap <<= 1;
  std::vector<HostSlot> tab(cap);
  for (auto &s : tab) s.voxel = -1;
  std::vector<int32_t> cnt;
  cnt.reserve((size_t)N / 2 + 1);
  const bool fill = voxel_locs != nullptr;
  int64_t M = 0;
  std::vector<int32_t> p2v_local;
  if (!p2v) p2v_local.resize((size_t)N);
  int32_t *pv = p2v ? p2v : p2v_local.data();
  for (int64_t i = 0; i < N; ++i) {
    const int64_t *c = coords + i * 4;
    size_t s = (size_t)host_hash(c) & (cap - 1);
    while (true) {
      HostSlot &h = tab[s];
      if (h.voxel < 0) {
        h.b = c[0]; h.x = c[1]; h.y = c[2]; h.z = c[3];
        h.voxel = (int32_t)M++;
        cnt.push_back(0);
        break;
      }
      if (h.b == c[0] && h.x == c[1] && h.y == c[2] && h.z == c[3]) break;
      s = (s + 1) & (cap - 1);
    }
    int32_t m = tab[s].voxel;
    pv[i] = m;
    cnt[(size_t)m]++;
  }
  int32_t ma = 0;
  for (int32_t c : cnt) ma = c > ma ? c : ma;
  if (max_active) *max_active = ma;
  if (fill) {
    if (v2p_stride < 1 + ma) {
      set_error("voxelize_idx: v2p_stride %d < 1+max_active %d", v2p_stride, 1 + ma);
      return -1;
    }
    memset(v2p, 0, sizeof(int32_t) * (size_t)M * v2p_stride);
    std::vector<int32_t> fillpos((size_t)M, 0);
    for (int64_t i = 0; i < N; ++i) {
      int32_t m = pv[i];
      int32_t k = fillpos[(size_t)m]++;
      if (k == 0) memcpy(voxel_locs + (int64_t)m * 4, coords + i * 4, sizeof(int64_t) * 4);
      v2p[(int64_t)m * v2p_stride + 1 + k] = (int32_t)i;
    }
    for (int64_t m = 0; m < M; ++m) v2p[m * v2p_stride] = cnt[(size_t)m];
  }
  return M;
}

int wsis_voxelize_idx_host_fill(const int64_t *coords, int64_t N, const int32_t *p2v, int64_t M, int64_t *voxel_locs,
                                int32_t *v2p, int32_t v2p_stride) {
  // second phase of the host routine when the caller kept the p2v of the counting call: one linear pass, no hashing
  if (N < 0 || M < 0 || v2p_stride < 1) {
    set_error("voxelize_idx_host_fill: bad sizes");
    return 1;
  }
  memset(v2p, 0, sizeof(int32_t) * (size_t)M * v2p_stride);
  for (int64_t i = 0; i < N; ++i) {
    const int32_t m = p2v[i];
    if (m < 0 || m >= M) {
      set_error("voxelize_idx_host_fill: p2v[%lld] = %d outside [0, %lld)", (long long)i, m, (long long)M);
      return 1;
    }
    int32_t *row = v2p + (int64_t)m * v2p_stride;
    const int32_t k = row[0]++;
    if (1 + k >= v2p_stride) {
      set_error("voxelize_idx_host_fill: v2p_stride %d too small", v2p_stride);
      return 1;
    }
    if (k == 0) memcpy(voxel_locs + (int64_t)m * 4, coords + i * 4, sizeof(int64_t) * 4);
    row[1 + k] = (int32_t)i;
  }
  return 0;
}

int64_t wsis_voxelize_ws_bytes(int64_t N) { return carve(nullptr, N).bytes; }

int wsis_voxelize_idx_count(const int64_t *coords, int64_t N, int32_t *p2v, int32_t *counts_dev, void *ws,
                            wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  WSIS_CHECK(N >= 0 && N < ((int64_t)1 << 31), "voxelize_idx: N out of range");
  WSIS_CUDA(cudaMemsetAsync(counts_dev, 0, 3 * sizeof(int32_t), st));
  if (N == 0) return 0;
  VoxWs w = carve(ws, N);
  unsigned blocks = (unsigned)ceil_div(N, 256);
  WSIS_CUDA(cudaMemsetAsync(w.keys, 0xFF, w.slots * 8, st));
  if (wsis_fill_i32(w.first, w.slots, INT_MAX, stream)) return 1;
  WSIS_CUDA(cudaMemsetAsync(w.cnt, 0, (N + 1) * 4, st));
  vox_touch_kernel<<<blocks, 256, 0, st>>>(coords, N, w.keys, w.first, w.slots - 1, w.slot_of, counts_dev + 2);
  WSIS_LAUNCH_OK();
  vox_flag_kernel<<<blocks, 256, 0, st>>>(w.slot_of, w.first, N, w.rank, w.iota);
  WSIS_LAUNCH_OK();
  if (wsis_exclusive_scan_i32(w.rank, w.rank, N, w.scan_ws, stream)) return 1;
  vox_assign_kernel<<<blocks, 256, 0, st>>>(w.slot_of, w.first, w.rank, N, p2v, w.cnt);
  WSIS_LAUNCH_OK();
  vox_max_kernel<<<blocks, 256, 0, st>>>(w.cnt, N, counts_dev + 1);
  WSIS_LAUNCH_OK();
  WSIS_CUDA(cudaMemcpyAsync(counts_dev, w.rank + N, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int wsis_voxelize_idx_fill(const int64_t *coords, int64_t N, const int32_t *p2v, int64_t M, int32_t max_active,
                           int64_t *voxel_locs, int32_t *v2p, void *ws, wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  if (N == 0 || M == 0) return 0;
  VoxWs w = carve(ws, N);
  int bits = 1;
  while (((int64_t)1 << bits) < M) ++bits;
  if (wsis_sort_pairs_u32((const uint32_t *)p2v, (const uint32_t *)w.iota, w.sorted_keys, (uint32_t *)w.sorted_pts, N, 0,
                          bits, w.sort_ws, stream))
    return 1;
  if (wsis_exclusive_scan_i32(w.cnt, w.off, M, w.scan_ws, stream)) return 1;
  int32_t stride = 1 + max_active;
  vox_fill_kernel<<<(unsigned)ceil_div(M * stride, 256), 256, 0, st>>>(coords, w.cnt, w.off, w.sorted_pts, M, stride,
                                                                       voxel_locs, v2p);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_voxelize_mean_fwd(const float *feats, const int32_t *v2p, int64_t M, int32_t v2p_stride, int C, float *out,
                           wsis_stream_t stream) {
  if (M == 0) return 0;
  vox_mean_fwd_kernel<<<(unsigned)ceil_div(M * C, 256), 256, 0, as_stream(stream)>>>(feats, v2p, M, v2p_stride, C, out);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_voxelize_mean_bwd(const float *dout, const int32_t *v2p, int64_t M, int32_t v2p_stride, int C, int64_t N,
                           float *dfeats, wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  WSIS_CUDA(cudaMemsetAsync(dfeats, 0, sizeof(float) * (size_t)N * C, st));
  if (M == 0) return 0;
  vox_mean_bwd_kernel<<<(unsigned)ceil_div(M * C, 256), 256, 0, st>>>(dout, v2p, M, v2p_stride, C, dfeats);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
