// Exact-fp32 SIMT sparse convolution (output-stationary gather-GEMM), weight gradient, and the
// eval-BatchNorm+ReLU elementwise kernel.
//
// Replaces indiceConv<float> / indiceConvBackward<float> (include/spconv/spconv_ops.h:253-349, 351-433) and
// the gather / scatterAdd kernels (include/spconv/reordering.cu.h:21-157).  The reference runs, per kernel
// offset, gather -> cuBLAS mm -> scatter-add through two staging buffers that are written and re-read
// (~2*P*(Cin+Cout)*4 B per layer) behind a device->host sync per layer.  Here each CTA owns a tile of OUTPUT
// rows, loops over the K offsets reading the neighbour map, accumulates in registers and writes every output
// row exactly once: no staging buffers, no scatter, no atomics, no sync, deterministic.
//
// This fp32 FFMA path serves the 6->32 input conv, channel counts the tensor-core path does not take, and is
// the on-GPU cross-check of conv_umma.cu.  The tcgen05 path carries the UNet layers.
#include "common.cuh"

namespace wsis {

constexpr int kTM = 64;   // output rows per CTA
constexpr int kKC = 16;   // input channels per smem step

template <int MAXJ>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const float *__restrict__ src, const int32_t *__restrict__ map, int64_t n_dst, int K, int flip,
                 const float *__restrict__ W, int transpose_w, int Cin, int Cout,
                 const float *__restrict__ in_scale, const float *__restrict__ in_shift, int in_relu,
                 const float *__restrict__ residual, float *__restrict__ dst) {
  __shared__ float As[kTM][kKC + 1];
  __shared__ float Ws[kKC][16 * MAXJ];
  __shared__ int32_t s_idx[kTM];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t row0 = (int64_t)blockIdx.x * kTM;
  const int col0 = blockIdx.y * 16 * MAXJ;
  const int ncol = min(16 * MAXJ, Cout - col0);
  float acc[4][MAXJ];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) acc[r][j] = 0.f;

  for (int k = 0; k < K; ++k) {
    const int kk = flip ? K - 1 - k : k;
    int any = 0;
    if (tid < kTM) {
      int64_t r = row0 + tid;
      int32_t v = r < n_dst ? __ldg(map + r * K + kk) : -1;
      s_idx[tid] = v;
      any = v >= 0;
    }
    if (!__syncthreads_or(any)) continue;
    for (int c0 = 0; c0 < Cin; c0 += kKC) {
      // gathered rows (+ fused BN/ReLU prologue); absent neighbours stay exactly zero
#pragma unroll
      for (int e = tid; e < kTM * kKC; e += 256) {
        int r = e / kKC, c = e % kKC;
        int32_t idx = s_idx[r];
        float x = 0.f;
        if (idx >= 0 && c0 + c < Cin) {
          x = __ldg(src + (int64_t)idx * Cin + c0 + c);
          if (in_scale) x = fmaf(x, __ldg(in_scale + c0 + c), __ldg(in_shift + c0 + c));
          if (in_relu) x = fmaxf(x, 0.f);
        }
        As[r][c] = x;
      }
      for (int e = tid; e < kKC * 16 * MAXJ; e += 256) {
        int kc = e / (16 * MAXJ), co = e % (16 * MAXJ);
        float w = 0.f;
        if (c0 + kc < Cin && co < ncol) {
          int ci = c0 + kc, cg = col0 + co;
          w = transpose_w ? __ldg(W + ((int64_t)k * Cout + cg) * Cin + ci)
                          : __ldg(W + ((int64_t)k * Cin + ci) * Cout + cg);
        }
        Ws[kc][co] = w;
      }
      __syncthreads();
#pragma unroll
      for (int kc = 0; kc < kKC; ++kc) {
        float a[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = As[ty * 4 + r][kc];
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          float b = Ws[kc][tx + 16 * j];
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[r][j] = fmaf(a[r], b, acc[r][j]);
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int64_t row = row0 + ty * 4 + r;
    if (row >= n_dst) continue;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      int co = tx + 16 * j;
      if (co < ncol) {
        float v = acc[r][j];
        if (residual) v += __ldg(residual + row * Cout + col0 + co);
        dst[row * Cout + col0 + co] = v;
      }
    }
  }
}

// dW[k] (32x32 tile) += A^T . G over a chunk of rows; one atomicAdd per element per chunk
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float *__restrict__ src, const int32_t *__restrict__ map, int64_t n_dst, int K, int flip,
                  const float *__restrict__ g, int Cin, int Cout, const float *__restrict__ in_scale,
                  const float *__restrict__ in_shift, int in_relu, float *__restrict__ dW, int64_t rows_per_chunk,
                  int tiles_co) {
  __shared__ float As[32][33];
  __shared__ float Gs[32][33];
  __shared__ int32_t s_idx[32];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int k = blockIdx.y;
  const int kk = flip ? K - 1 - k : k;
  const int ci0 = (blockIdx.z / tiles_co) * 32, co0 = (blockIdx.z % tiles_co) * 32;
  const int64_t rbeg = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t rend = min(n_dst, rbeg + rows_per_chunk);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int64_t r0 = rbeg; r0 < rend; r0 += 32) {
    int any = 0;
    if (tid < 32) {
      int64_t r = r0 + tid;
      int32_t v = r < rend ? __ldg(map + r * K + kk) : -1;
      s_idx[tid] = v;
      any = v >= 0;
    }
    if (!__syncthreads_or(any)) continue;
    for (int e = tid; e < 32 * 32; e += 256) {
      int r = e >> 5, c = e & 31;
      int32_t idx = s_idx[r];
      float x = 0.f, gv = 0.f;
      if (idx >= 0) {
        if (ci0 + c < Cin) {
          x = __ldg(src + (int64_t)idx * Cin + ci0 + c);
          if (in_scale) x = fmaf(x, __ldg(in_scale + ci0 + c), __ldg(in_shift + ci0 + c));
          if (in_relu) x = fmaxf(x, 0.f);
        }
        if (co0 + c < Cout) gv = __ldg(g + (r0 + r) * Cout + co0 + c);
      }
      As[r][c] = x;
      Gs[r][c] = gv;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      float a0 = As[r][ty * 2], a1 = As[r][ty * 2 + 1];
      float b0 = Gs[r][tx * 2], b1 = Gs[r][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      int ci = ci0 + ty * 2 + a, co = co0 + tx * 2 + b;
      if (ci < Cin && co < Cout && acc[a][b] != 0.f) atomicAdd(dW + ((int64_t)k * Cin + ci) * Cout + co, acc[a][b]);
    }
}

__global__ void affine_relu_kernel(const float *__restrict__ x, int64_t n, int C, const float *__restrict__ scale,
                                   const float *__restrict__ shift, int relu, float *__restrict__ y) {
  int64_t total = n * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float v = fmaf(x[i], __ldg(scale + c), __ldg(shift + c));
    y[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

__global__ void affine_relu_vec4_kernel(const float4 *__restrict__ x, int64_t n4, int C4,
                                        const float4 *__restrict__ scale, const float4 *__restrict__ shift, int relu,
                                        float4 *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C4);
    float4 v = x[i], s = __ldg(scale + c), b = __ldg(shift + c);
    v.x = fmaf(v.x, s.x, b.x);
    v.y = fmaf(v.y, s.y, b.y);
    v.z = fmaf(v.z, s.z, b.z);
    v.w = fmaf(v.w, s.w, b.w);
    if (relu) {
      v.x = fmaxf(v.x, 0.f);
      v.y = fmaxf(v.y, 0.f);
      v.z = fmaxf(v.z, 0.f);
      v.w = fmaxf(v.w, 0.f);
    }
    y[i] = v;
  }
}

}  // namespace wsis

using namespace wsis;

extern "C" {

int wsis_conv_simt(const float *src, const int32_t *map, int64_t n_dst, int K, int flip, const float *W,
                   int transpose_w, int Cin, int Cout, const float *in_scale, const float *in_shift, int in_relu,
                   const float *residual, float *dst, wsis_stream_t stream) {
  WSIS_CHECK(Cin >= 1 && Cout >= 1 && K >= 1, "conv_simt: bad dims");
  WSIS_CHECK((in_scale == nullptr) == (in_shift == nullptr), "conv_simt: in_scale/in_shift must both be set");
  if (n_dst == 0) return 0;
  cudaStream_t st = as_stream(stream);
  unsigned gx = (unsigned)ceil_div(n_dst, kTM);
#define WSIS_SIMT(MJ)                                                                                          \
  conv_simt_kernel<MJ><<<dim3(gx, (unsigned)ceil_div(Cout, 16 * MJ)), 256, 0, st>>>(                          \
      src, map, n_dst, K, flip, W, transpose_w, Cin, Cout, in_scale, in_shift, in_relu, residual, dst)
  if (Cout <= 32)
    WSIS_SIMT(2);
  else if (Cout <= 64)
    WSIS_SIMT(4);
  else
    WSIS_SIMT(8);
#undef WSIS_SIMT
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_conv_wgrad(const float *src, const int32_t *map, int64_t n_dst, int K, int flip, const float *g, int Cin,
                    int Cout, const float *in_scale, const float *in_shift, int in_relu, float *dW,
                    wsis_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  WSIS_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)K * Cin * Cout, st));
  if (n_dst == 0) return 0;
  int tiles_ci = (Cin + 31) / 32, tiles_co = (Cout + 31) / 32;
  int64_t per = (int64_t)K * tiles_ci * tiles_co;
  int64_t chunks = std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_dst, 256), ceil_div((int64_t)sm_count() * 8, per)));
  int64_t rows_per_chunk = ceil_div(ceil_div(n_dst, chunks), 32) * 32;
  chunks = ceil_div(n_dst, rows_per_chunk);
  WSIS_CHECK(K <= 65535 && tiles_ci * tiles_co <= 65535, "wgrad: grid too large");
  conv_wgrad_kernel<<<dim3((unsigned)chunks, (unsigned)K, (unsigned)(tiles_ci * tiles_co)), 256, 0, st>>>(
      src, map, n_dst, K, flip, g, Cin, Cout, in_scale, in_shift, in_relu, dW, rows_per_chunk, tiles_co);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_affine_relu(const float *x, int64_t n, int C, const float *scale, const float *shift, int relu, float *y,
                     wsis_stream_t stream) {
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int64_t total = n * C;
  bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                                reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0;
  if (vec) {
    int64_t n4 = total / 4;
    unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n4, 256), (int64_t)sm_count() * 16);
    affine_relu_vec4_kernel<<<blocks, 256, 0, st>>>((const float4 *)x, n4, C / 4, (const float4 *)scale,
                                                    (const float4 *)shift, relu, (float4 *)y);
  } else {
    unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
    affine_relu_kernel<<<blocks, 256, 0, st>>>(x, n, C, scale, shift, relu, y);
  }
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
