// One step of the edge-conditioned GRU on the superpoint graph (ECC-GRU).
//
// The reference runs, per step (modules/model/spg_modules.py:152-185, 226-253; 7 steps, graphnet.py:77-92):
//   NNConv  m[t] = mean over edges (s -> t) of  h[s]^T . W_e          (W_e = the edge's 32x32 filter, [E,1024] fp32)
//   GRUCellEx   inp = sigmoid(ig(h)) * m ;  gi = LN(inp . Wih^T) ;  gh = LN(h . Whh^T)
//               r = sig(gi_r + bih_r + gh_r + bhh_r) ; z = sig(gi_z + bih_z + gh_z + bhh_z)
//               n = tanh(gi_n + bih_n + r * (gh_n + bhh_n)) ;  h' = n + z * (h - n)
// as a batched gemv, an index_add, a count, a divide and ~15 small elementwise / norm / GEMM launches.  Here one warp
// owns one target superpoint for the whole step: it streams the filters of its in-edges once (the only HBM-sized
// traffic: 4 KB per edge), keeps the message, the gates and the two 96-wide layer norms in registers, and writes the
// new hidden state once -- into the ping-pong state buffer and into its column block of the concatenated output
// (cat_all, spg_modules.py:183-185).  No atomics, fixed summation order (edges in the order given, channels ascending).
//
// Lane o owns channel o of every 32-vector and gate rows o, o+32, o+64 of the 96-vectors.  The GRU parameters live
// in shared memory TRANSPOSED ([in][out]) so that the 32 lanes read consecutive words.
#include "common.cuh"

namespace wsis {

constexpr int kF = 32;  // nfeat of the 3D-WSIS ECC network (backbone_3D_WSIS.py:67)

// packed parameters (floats): WigT[32][32] | big[32] | WihT[32][96] | WhhT[32][96] | bih[96] | bhh[96]
constexpr int kOffWig = 0, kOffBig = kOffWig + kF * kF, kOffWih = kOffBig + kF, kOffWhh = kOffWih + kF * 3 * kF,
              kOffBih = kOffWhh + kF * 3 * kF, kOffBhh = kOffBih + 3 * kF, kEccParams = kOffBhh + 3 * kF;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// un-affine layer norm over the 96 values held as 3 per lane (biased variance, like InstanceNorm1d / layer_norm)
__device__ __forceinline__ void layer_norm96(float &a, float &b, float &c, float eps) {
  const float mean = warp_sum(a + b + c) * (1.f / 96.f);
  a -= mean;
  b -= mean;
  c -= mean;
  const float var = warp_sum(a * a + b * b + c * c) * (1.f / 96.f);
  const float inv = rsqrtf(var + eps);
  a *= inv;
  b *= inv;
  c *= inv;
}

__global__ void __launch_bounds__(256)
ecc_gru_step_kernel(const float *__restrict__ h, const float *__restrict__ filters, const int64_t *__restrict__ src,
                    const int32_t *__restrict__ eorder, const int32_t *__restrict__ offsets, int64_t S, const float *__restrict__ params, int layernorm,
                    float eps, float *__restrict__ h_out, float *__restrict__ cat_out, int64_t cat_stride,
                    const float *__restrict__ msg) {
  __shared__ float sp[kEccParams];
  for (int i = threadIdx.x; i < kEccParams; i += blockDim.x) sp[i] = __ldg(params + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= S) return;
  const int beg = __ldg(offsets + t), end = __ldg(offsets + t + 1);
  // ---- NNConv: mean of h[s]^T W_e over the in-edges ----
  float m = 0.f;
  if (filters == nullptr) {                  // the messages were produced by ecc_umma.cu, one row per edge in CSR order
    for (int j = beg; j < end; ++j) m += __ldg(msg + (int64_t)j * kF + lane);
  } else
  for (int j = beg; j < end; ++j) {
    const int e = eorder != nullptr ? __ldg(eorder + j) : j;
    const int64_t s = __ldg(src + e);
    const float xs = __ldg(h + s * kF + lane);
    const float *we = filters + (int64_t)e * (kF * kF) + lane;
    float acc = 0.f;
#pragma unroll 8
    for (int i = 0; i < kF; ++i) acc = fmaf(__shfl_sync(0xffffffffu, xs, i), __ldg(we + i * kF), acc);
    m += acc;
  }
  m = m / (float)max(end - beg, 1);
  // ---- GRUCellEx ----
  const float hv = __ldg(h + t * kF + lane);
  float ig = sp[kOffBig + lane];
#pragma unroll 8
  for (int i = 0; i < kF; ++i) ig = fmaf(__shfl_sync(0xffffffffu, hv, i), sp[kOffWig + i * kF + lane], ig);
  const float inp = sigmoidf_(ig) * m;
  float gi0 = 0.f, gi1 = 0.f, gi2 = 0.f, gh0 = 0.f, gh1 = 0.f, gh2 = 0.f;
#pragma unroll 4
  for (int i = 0; i < kF; ++i) {
    const float xi = __shfl_sync(0xffffffffu, inp, i), hi = __shfl_sync(0xffffffffu, hv, i);
    const float *wi = sp + kOffWih + i * 3 * kF + lane, *wh = sp + kOffWhh + i * 3 * kF + lane;
    gi0 = fmaf(xi, wi[0], gi0);
    gi1 = fmaf(xi, wi[kF], gi1);
    gi2 = fmaf(xi, wi[2 * kF], gi2);
    gh0 = fmaf(hi, wh[0], gh0);
    gh1 = fmaf(hi, wh[kF], gh1);
    gh2 = fmaf(hi, wh[2 * kF], gh2);
  }
  if (layernorm) {
    layer_norm96(gi0, gi1, gi2, eps);
    layer_norm96(gh0, gh1, gh2, eps);
  }
  const float *bih = sp + kOffBih + lane, *bhh = sp + kOffBhh + lane;
  const float r = sigmoidf_(gi0 + bih[0] + gh0 + bhh[0]);
  const float z = sigmoidf_(gi1 + bih[kF] + gh1 + bhh[kF]);
  const float n = tanhf(gi2 + bih[2 * kF] + r * (gh2 + bhh[2 * kF]));
  const float hn = n + z * (hv - n);
  h_out[t * kF + lane] = hn;
  if (cat_out != nullptr) cat_out[t * cat_stride + lane] = hn;
}

}  // namespace wsis

using namespace wsis;

extern "C" {

int64_t wsis_ecc_gru_param_floats(void) { return kEccParams; }

int wsis_ecc_gru_step(const float *h, const float *filters, const int64_t *src, const int32_t *eorder,
                      const int32_t *offsets, int64_t S,
                      const float *params, int layernorm, float eps, float *h_out, float *cat_out, int64_t cat_stride,
                      wsis_stream_t stream) {
  WSIS_CHECK(S >= 0 && S < ((int64_t)1 << 31), "ecc_gru_step: S out of range");
  WSIS_CHECK(h != h_out, "ecc_gru_step: the state buffers must differ (other warps still read h)");
  if (S == 0) return 0;
  const int warps_per_block = 8;
  ecc_gru_step_kernel<<<(unsigned)ceil_div(S, warps_per_block), warps_per_block * 32, 0, as_stream(stream)>>>(
      h, filters, src, eorder, offsets, S, params, layernorm, eps, h_out, cat_out, cat_stride, nullptr);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_ecc_gru_step_msg(const float *h, const float *msg, const int32_t *offsets, int64_t S, const float *params,
                          int layernorm, float eps, float *h_out, float *cat_out, int64_t cat_stride,
                          wsis_stream_t stream) {
  WSIS_CHECK(S >= 0 && S < ((int64_t)1 << 31), "ecc_gru_step_msg: S out of range");
  WSIS_CHECK(h != h_out, "ecc_gru_step_msg: the state buffers must differ");
  if (S == 0) return 0;
  const int warps_per_block = 8;
  ecc_gru_step_kernel<<<(unsigned)ceil_div(S, warps_per_block), warps_per_block * 32, 0, as_stream(stream)>>>(
      h, nullptr, nullptr, nullptr, offsets, S, params, layernorm, eps, h_out, cat_out, cat_stride, msg);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
