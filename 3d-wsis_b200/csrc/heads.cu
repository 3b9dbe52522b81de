// Per-row MLP heads of the network in inference: Linear -> BatchNorm1d(eval) -> ReLU -> Linear
// (modules/model/backbone_3D_WSIS.py:57-62 `linear` over every POINT, :71-104 the superpoint heads), one kernel per head
// instead of two library GEMMs + a batch-norm + a ReLU launch, optionally fused with the voxel -> point gather (:179):
//
//   y[i, :] = W2 . relu(W1' . x[g(i), :] + t') + b2,     g(i) = gather ? gather[i] : i
//
// W1' = diag(scale) W1 and t' = scale*b1 + shift fold the eval-mode BatchNorm (host side, once per parameter version).
// One thread owns one row: the hidden vector lives in registers, the weights sit in shared memory transposed
// ([in][out]) so that every warp-wide read is a 16-byte broadcast.  HBM traffic is the row in and the scores out.
#include "common.cuh"

namespace wsis {

template <int CIN, int H, int COUTP>
__global__ void __launch_bounds__(128)
mlp_head_kernel(const float *__restrict__ src, const int32_t *__restrict__ gather, int64_t n, const float *__restrict__ w1t,
                const float *__restrict__ t1, const float *__restrict__ w2t, const float *__restrict__ b2, int Cout,
                float *__restrict__ out) {
  __shared__ __align__(16) float s_w1[CIN * H];
  __shared__ __align__(16) float s_w2[H * COUTP];
  __shared__ __align__(16) float s_t1[H];
  __shared__ __align__(16) float s_b2[COUTP];
  for (int i = threadIdx.x; i < CIN * H; i += blockDim.x) s_w1[i] = __ldg(w1t + i);
  for (int i = threadIdx.x; i < H * COUTP; i += blockDim.x) s_w2[i] = __ldg(w2t + i);
  for (int i = threadIdx.x; i < H; i += blockDim.x) s_t1[i] = __ldg(t1 + i);
  for (int i = threadIdx.x; i < COUTP; i += blockDim.x) s_b2[i] = __ldg(b2 + i);
  __syncthreads();
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = gather ? (int64_t)__ldg(gather + row) : row;
    const float4 *xp = reinterpret_cast<const float4 *>(src + r * CIN);
    float hid[H];
#pragma unroll
    for (int h = 0; h < H; ++h) hid[h] = s_t1[h];
#pragma unroll
    for (int i4 = 0; i4 < CIN / 4; ++i4) {
      const float4 xv = __ldg(xp + i4);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 *wr = reinterpret_cast<const float4 *>(s_w1 + (i4 * 4 + u) * H);
#pragma unroll
        for (int h4 = 0; h4 < H / 4; ++h4) {
          const float4 w = wr[h4];
          hid[h4 * 4 + 0] = fmaf(xs[u], w.x, hid[h4 * 4 + 0]);
          hid[h4 * 4 + 1] = fmaf(xs[u], w.y, hid[h4 * 4 + 1]);
          hid[h4 * 4 + 2] = fmaf(xs[u], w.z, hid[h4 * 4 + 2]);
          hid[h4 * 4 + 3] = fmaf(xs[u], w.w, hid[h4 * 4 + 3]);
        }
      }
    }
    float y[COUTP];
#pragma unroll
    for (int c = 0; c < COUTP; ++c) y[c] = s_b2[c];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float a = fmaxf(hid[h], 0.f);
      const float4 *wr = reinterpret_cast<const float4 *>(s_w2 + h * COUTP);
#pragma unroll
      for (int c4 = 0; c4 < COUTP / 4; ++c4) {
        const float4 w = wr[c4];
        y[c4 * 4 + 0] = fmaf(a, w.x, y[c4 * 4 + 0]);
        y[c4 * 4 + 1] = fmaf(a, w.y, y[c4 * 4 + 1]);
        y[c4 * 4 + 2] = fmaf(a, w.z, y[c4 * 4 + 2]);
        y[c4 * 4 + 3] = fmaf(a, w.w, y[c4 * 4 + 3]);
      }
    }
    float *op = out + row * Cout;
    if (Cout == COUTP && (Cout & 3) == 0) {
#pragma unroll
      for (int c4 = 0; c4 < COUTP / 4; ++c4)
        reinterpret_cast<float4 *>(op)[c4] = make_float4(y[c4 * 4], y[c4 * 4 + 1], y[c4 * 4 + 2], y[c4 * 4 + 3]);
    } else {
#pragma unroll
      for (int c = 0; c < COUTP; ++c)
        if (c < Cout) op[c] = y[c];
    }
  }
}

}  // namespace wsis

using namespace wsis;

extern "C" {

int wsis_mlp_head_coutp(int Cout) { return Cout <= 4 ? 4 : (Cout <= 8 ? 8 : (Cout <= 20 ? 20 : (Cout <= 32 ? 32 : -1))); }

int wsis_mlp_head(const float *src, const int32_t *gather, int64_t n, int Cin, int H, int Cout, const float *w1t,
                  const float *t1, const float *w2t, const float *b2, float *out, wsis_stream_t stream) {
  const int cp = wsis_mlp_head_coutp(Cout);
  WSIS_CHECK(cp > 0 && ((Cin == 32 && H == 32) || (Cin == 64 && H == 64)), "mlp_head: (Cin, H) must be (32,32) or (64,64), Cout <= 32");
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 128), (int64_t)sm_count() * 8);
#define WSIS_HEAD(CI, HH, CP)                                                                                     \
  mlp_head_kernel<CI, HH, CP><<<blocks, 128, 0, st>>>(src, gather, n, w1t, t1, w2t, b2, Cout, out)
  if (Cin == 32) {
    if (cp == 4) WSIS_HEAD(32, 32, 4); else if (cp == 8) WSIS_HEAD(32, 32, 8); else if (cp == 20) WSIS_HEAD(32, 32, 20); else WSIS_HEAD(32, 32, 32);
  } else {
    if (cp == 4) WSIS_HEAD(64, 64, 4); else if (cp == 8) WSIS_HEAD(64, 64, 8); else if (cp == 20) WSIS_HEAD(64, 64, 20); else WSIS_HEAD(64, 64, 32);
  }
#undef WSIS_HEAD
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
