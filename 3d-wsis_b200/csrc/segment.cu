// Segmented reductions: the torch_scatter.scatter(dim=0, reduce=mean/sum/max) calls on the hot path
// (modules/model/backbone_3D_WSIS.py:188 superpoint pooling, :225,232,244 edge softmax; centres at
// train_scannetv2.py:177).  torch_scatter reduces with global atomics in arrival order; here a CSR of the
// (unsorted) segment ids is built once per scene with the stable radix sort, and one warp reduces each segment
// in a fixed order: no atomics, deterministic, every source row read once as a full 128-byte line.
#include "common.cuh"

namespace wsis {

static inline int64_t al(int64_t x) { return (x + 255) / 256 * 256; }

struct CsrWs {
  uint32_t *keys, *iota, *keys_sorted;
  int32_t *cnt;
  void *scan_ws, *sort_ws;
  int64_t bytes;
};

static CsrWs carve_csr(void *ws, int64_t N, int64_t S) {
  CsrWs w;
  char *p = reinterpret_cast<char *>(ws);
  int64_t n1 = N > 0 ? N : 1, s1 = S > 0 ? S : 1;
  auto take = [&](int64_t bytes) {
    char *q = p;
    p += al(bytes);
    return q;
  };
  w.keys = (uint32_t *)take(n1 * 4);
  w.iota = (uint32_t *)take(n1 * 4);
  w.keys_sorted = (uint32_t *)take(n1 * 4);
  w.cnt = (int32_t *)take((s1 + 1) * 4);
  w.scan_ws = take(wsis_scan_ws_bytes(s1 + 1));
  w.sort_ws = take(wsis_sort_ws_bytes(n1));
  w.bytes = p - reinterpret_cast<char *>(ws);
  return w;
}

__global__ void csr_prepare_kernel(const int64_t *__restrict__ ids, int64_t N, int64_t S, uint32_t *__restrict__ keys,
                                   uint32_t *__restrict__ iota, int32_t *cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int64_t s = ids[i];
  // out-of-range ids are clamped into the last bucket rather than corrupting memory
  if (s < 0) s = 0;
  if (s >= S) s = S - 1;
  keys[i] = (uint32_t)s;
  iota[i] = (uint32_t)i;
  atomicAdd(cnt + s, 1);
}

// one warp per segment
template <int REDUCE>
__global__ void segment_reduce_kernel(const float *__restrict__ src, const int32_t *__restrict__ gather,
                                      const int32_t *__restrict__ order, const int32_t *__restrict__ offsets,
                                      int64_t S, int C, float *__restrict__ out) {
  int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= S) return;
  int beg = __ldg(offsets + s), end = __ldg(offsets + s + 1);
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    int j = beg;
    // 4 independent row loads in flight, accumulated in row order
    for (; j + 4 <= end; j += 4) {
      int32_t r[4];
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        r[u] = __ldg(order + j + u);
        if (gather) r[u] = __ldg(gather + r[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(src + (int64_t)r[u] * C + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (REDUCE == 2)
          acc = (j + u == beg) ? v[u] : fmaxf(acc, v[u]);
        else
          acc += v[u];
      }
    }
    for (; j < end; ++j) {
      int32_t r = __ldg(order + j);
      if (gather) r = __ldg(gather + r);
      float v = __ldg(src + (int64_t)r * C + c);
      if (REDUCE == 2)
        acc = (j == beg) ? v : fmaxf(acc, v);
      else
        acc += v;
    }
    if (REDUCE == 1) acc = acc / (float)max(end - beg, 1);
    out[s * C + c] = acc;
  }
}

__global__ void gather_rows_kernel(const float *__restrict__ src, const int32_t *__restrict__ idx, int64_t n, int C,
                                   float *__restrict__ dst) {
  int64_t total = n * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = e / C;
    int c = (int)(e - i * C);
    dst[e] = __ldg(src + (int64_t)__ldg(idx + i) * C + c);
  }
}

__global__ void gather_rows_vec4_kernel(const float4 *__restrict__ src, const int32_t *__restrict__ idx, int64_t n,
                                        int C4, float4 *__restrict__ dst) {
  int64_t total = n * C4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = e / C4;
    int c = (int)(e - i * C4);
    dst[e] = __ldg(src + (int64_t)__ldg(idx + i) * C4 + c);
  }
}

}  // namespace wsis

using namespace wsis;

extern "C" {

int64_t wsis_segment_csr_ws_bytes(int64_t N, int64_t S) { return carve_csr(nullptr, N, S).bytes; }

int wsis_segment_csr(const int64_t *ids, int64_t N, int64_t S, int32_t *order, int32_t *offsets, void *ws,
                     wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  WSIS_CHECK(S >= 0 && S < ((int64_t)1 << 31) && N >= 0 && N < ((int64_t)1 << 31), "segment_csr: size out of range");
  CsrWs w = carve_csr(ws, N, S);
  WSIS_CUDA(cudaMemsetAsync(w.cnt, 0, (S + 1) * 4, st));
  if (N > 0) {
    WSIS_CHECK(S > 0, "segment_csr: S must be > 0 when N > 0");
    csr_prepare_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(ids, N, S, w.keys, w.iota, w.cnt);
    WSIS_LAUNCH_OK();
    int bits = 1;
    while (((int64_t)1 << bits) < S) ++bits;
    if (wsis_sort_pairs_u32(w.keys, w.iota, w.keys_sorted, (uint32_t *)order, N, 0, bits, w.sort_ws, stream)) return 1;
  }
  return wsis_exclusive_scan_i32(w.cnt, offsets, S, w.scan_ws, stream);
}

int wsis_segment_reduce(const float *src, const int32_t *gather, const int32_t *order, const int32_t *offsets,
                        int64_t S, int C, int reduce, float *out, wsis_stream_t stream) {
  if (S == 0) return 0;
  cudaStream_t st = as_stream(stream);
  unsigned blocks = (unsigned)ceil_div(S * 32, 256);
  switch (reduce) {
    case 0: segment_reduce_kernel<0><<<blocks, 256, 0, st>>>(src, gather, order, offsets, S, C, out); break;
    case 1: segment_reduce_kernel<1><<<blocks, 256, 0, st>>>(src, gather, order, offsets, S, C, out); break;
    case 2: segment_reduce_kernel<2><<<blocks, 256, 0, st>>>(src, gather, order, offsets, S, C, out); break;
    default: WSIS_CHECK(false, "segment_reduce: unknown reduce %d", reduce);
  }
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_gather_rows(const float *src, const int32_t *idx, int64_t n, int C, float *dst, wsis_stream_t stream) {
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  int64_t total = vec ? n * (C / 4) : n * C;
  unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
  if (vec)
    gather_rows_vec4_kernel<<<blocks, 256, 0, st>>>((const float4 *)src, idx, n, C / 4, (float4 *)dst);
  else
    gather_rows_kernel<<<blocks, 256, 0, st>>>(src, idx, n, C, dst);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
