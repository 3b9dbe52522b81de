// ECC-GRU without materialised edge filters (reference: modules/model/graphnet.py:21-39 filter
// network, spg_modules.py:152-185 NNConv message).
//
// The reference generates one 32x32 filter per edge, F_e = W4 . he_e + b4 (he_e = the 64-wide hidden of the filter
// network, [E,1024] fp32 = 267 MB per 4-scene batch), and every one of the 7 GRU steps streams all of them.  Here the
// filters never exist in memory: per step and per tile of 128 edges the tensor cores regenerate them into TENSOR
// MEMORY, D[128 edges][256 columns] = He_tile[128 x 64] . W4_q^T (bf16 hi/mid split operands, fp32 accumulate: the
// fp32 contract), and the epilogue contracts them on the spot with the edge's source state,
//     m_e[o] = sum_i h[src(e)][i] * (D[e][32 i + o] + b4[32 i + o]),
// so HBM sees 8 KB + 128 B per edge tile row instead of 4 KB per edge.  The mean over the in-edges and the GRU cell
// run in ecc.cu's step kernel on the [E,32] messages (edges are in target-CSR order: a target's messages are
// contiguous rows).
//
//   wsis_ecc_edge_mlp    once per forward: he = relu(bn(L3(relu(L2(relu(L1(edge features))))))) per edge, one thread
//                        per edge, written directly as the bf16 hi/mid K-major SWIZZLE_64B operand tiles of the MMA
//   wsis_ecc_messages    once per step: 2 CTAs of 128 threads per SM; per tile 4 quarters of {cp.async.bulk of the W4
//                        quarter (64 KB), 12 tcgen05.mma M=128 N=256 K=16, TMEM -> registers contraction}
#include <cuda_bf16.h>

#include "common.cuh"

namespace wsis {
namespace eccu {

constexpr int kF = 32, kH = 64, kNF = kF * kF;   // state width, filter-net hidden width, filter size
constexpr int kTile = 128;
constexpr int kABlk = 128 * 64;                  // [128 rows x 32 k] bf16, SWIZZLE_64B
constexpr int kATile = 4 * kABlk;                // kb(2) x {hi, mid}
constexpr int kQ = 256;                          // columns per quarter
constexpr int kWBlk = kQ * 64;                   // [256 rows x 32 k] bf16
constexpr int kWQuarter = 4 * kWBlk;             // kb(2) x {hi, mid} = 64 KB
// edge-MLP parameter pack (floats): W1T[13][32] | b1[32] | W2T[32][128] | b2[128] | W3T[128][64] (BN folded) | b3[64]
constexpr int kIn = 13, kH1 = 32, kH2 = 128;
constexpr int kOffW1 = 0, kOffB1 = kOffW1 + kIn * kH1, kOffW2 = kOffB1 + kH1, kOffB2 = kOffW2 + kH1 * kH2,
              kOffW3 = kOffB2 + kH2, kOffB3 = kOffW3 + kH2 * kH, kMlpParams = kOffB3 + kH;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ unsigned int *d_trap_word = nullptr;   // host-mapped diagnostics record (common.cuh trap_word_device)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t n = 0; n < (1u << 24) && !ok; ++n) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  if (!ok) {
    if (d_trap_word != nullptr) {
      volatile unsigned int *w = d_trap_word;
      w[1] = bar, w[2] = parity, w[3] = blockIdx.x, w[4] = threadIdx.x;
      w[0] = 2u;
      __threadfence_system();
    }
    __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major SWIZZLE_64B descriptor: 8-row groups at SBO = 512 B (same form as conv_umma.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__host__ __device__ __forceinline__ uint32_t sw64(uint32_t row, uint32_t c16) {
  return (row >> 3) * 512u + (row & 7u) * 64u + ((c16 ^ ((row & 7u) >> 1)) << 4);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ void split8(const float *v, uint4 &hi, uint4 &mid) {
  uint32_t h[4], m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
    m[i] = pack_bf16(v[2 * i] - __uint_as_float(h[i] << 16), v[2 * i + 1] - __uint_as_float(h[i] & 0xFFFF0000u));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  mid = make_uint4(m[0], m[1], m[2], m[3]);
}

// ---------------------------------------------------------------------------------------------------------------
// filter-network hidden layers, one thread per edge (position j of the target-CSR order)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
edge_mlp_kernel(const float *__restrict__ edgefeats, const int32_t *__restrict__ eorder, int64_t E,
                const float *__restrict__ params, uint8_t *__restrict__ he_packed) {
  extern __shared__ __align__(16) float sp[];
  for (int i = threadIdx.x; i < kMlpParams; i += blockDim.x) sp[i] = __ldg(params + i);
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // blockDim == kTile: one block = one tile
  const int r = threadIdx.x;
  float he[kH];
#pragma unroll
  for (int c = 0; c < kH; ++c) he[c] = 0.f;
  if (j < E) {
    const int64_t e = eorder ? (int64_t)__ldg(eorder + j) : j;
    float h1[kH1];
#pragma unroll
    for (int c = 0; c < kH1; ++c) h1[c] = sp[kOffB1 + c];
#pragma unroll
    for (int i = 0; i < kIn; ++i) {
      const float x = __ldg(edgefeats + e * kIn + i);
#pragma unroll
      for (int c = 0; c < kH1; ++c) h1[c] = fmaf(x, sp[kOffW1 + i * kH1 + c], h1[c]);
    }
#pragma unroll
    for (int c = 0; c < kH1; ++c) h1[c] = fmaxf(h1[c], 0.f);
#pragma unroll
    for (int c = 0; c < kH; ++c) he[c] = sp[kOffB3 + c];
#pragma unroll 2
    for (int k = 0; k < kH2; ++k) {                       // layer 2 output k, consumed by layer 3 at once
      float a = sp[kOffB2 + k];
#pragma unroll
      for (int i = 0; i < kH1; ++i) a = fmaf(h1[i], sp[kOffW2 + i * kH2 + k], a);
      a = fmaxf(a, 0.f);
      const float4 *w3 = reinterpret_cast<const float4 *>(sp + kOffW3 + k * kH);
#pragma unroll
      for (int c4 = 0; c4 < kH / 4; ++c4) {
        const float4 w = w3[c4];
        he[c4 * 4 + 0] = fmaf(a, w.x, he[c4 * 4 + 0]);
        he[c4 * 4 + 1] = fmaf(a, w.y, he[c4 * 4 + 1]);
        he[c4 * 4 + 2] = fmaf(a, w.z, he[c4 * 4 + 2]);
        he[c4 * 4 + 3] = fmaf(a, w.w, he[c4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < kH; ++c) he[c] = fmaxf(he[c], 0.f);
  }
  // tile layout: [kb][hi | mid][128 rows x 64 B swizzled]
  uint8_t *tile = he_packed + (int64_t)blockIdx.x * kATile;
#pragma unroll
  for (int kb = 0; kb < 2; ++kb)
#pragma unroll
    for (int c16 = 0; c16 < 4; ++c16) {
      uint4 hi, mid;
      split8(he + kb * 32 + c16 * 8, hi, mid);
      const uint32_t off = sw64(r, c16);
      *reinterpret_cast<uint4 *>(tile + (kb * 2 + 0) * kABlk + off) = hi;
      *reinterpret_cast<uint4 *>(tile + (kb * 2 + 1) * kABlk + off) = mid;
    }
}

// W4 fp32 [1024][64] -> [quarter(4)][kb(2)][hi | mid][256 rows x 64 B swizzled] bf16
__global__ void pack_w4_kernel(const float *__restrict__ w4, uint8_t *__restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;                     // filter element (row of W4)
  if (n >= kNF) return;
  const int q = n / kQ, r = n % kQ;
#pragma unroll
  for (int kb = 0; kb < 2; ++kb)
#pragma unroll
    for (int c16 = 0; c16 < 4; ++c16) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(w4 + (int64_t)n * kH + kb * 32 + c16 * 8 + i);
      uint4 hi, mid;
      split8(v, hi, mid);
      const uint32_t off = sw64(r, c16);
      uint8_t *base = out + (int64_t)q * kWQuarter;
      *reinterpret_cast<uint4 *>(base + (kb * 2 + 0) * kWBlk + off) = hi;
      *reinterpret_cast<uint4 *>(base + (kb * 2 + 1) * kWBlk + off) = mid;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// messages of one GRU step
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 2)
ecc_messages_kernel(const uint8_t *__restrict__ he_packed, const uint8_t *__restrict__ w4_packed,
                    const float *__restrict__ b4, const float *__restrict__ h, const int64_t *__restrict__ src,
                    const int32_t *__restrict__ eorder, int64_t E, int tiles, float *__restrict__ msg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sA = smem;                         // 32 KB
  uint8_t *sW = sA + kATile;                  // 64 KB
  float *s_b4 = reinterpret_cast<float *>(sW + kWQuarter);   // 4 KB
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_b4 + kNF);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t barL = smem_u32(s_bar), barM = smem_u32(s_bar + 1);
  if ((int)blockIdx.x >= tiles) return;
  for (int i = tid; i < kNF; i += blockDim.x) s_b4[i] = b4 ? __ldg(b4 + i) : 0.f;
  if (tid == 0) {
    mbar_init(barL, 1);
    mbar_init(barM, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t idesc = make_idesc(kQ);
  uint32_t phL = 0, phM = 0;

  // first loads: A of the first tile + W quarter 0
  if (tid == 0) {
    mbar_expect_tx(barL, kATile + kWQuarter);
    bulk_g2s(smem_u32(sA), he_packed + (int64_t)blockIdx.x * kATile, kATile, barL);
    bulk_g2s(smem_u32(sW), w4_packed, kWQuarter, barL);
  }
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t j = (int64_t)t * kTile + tid;
    float x[kF], m[kF];
#pragma unroll
    for (int i = 0; i < kF; ++i) x[i] = 0.f, m[i] = 0.f;
    if (j < E) {
      const int64_t e = eorder ? (int64_t)__ldg(eorder + j) : j;
      const float4 *hp = reinterpret_cast<const float4 *>(h + __ldg(src + e) * kF);
#pragma unroll
      for (int i4 = 0; i4 < kF / 4; ++i4) {
        const float4 v = __ldg(hp + i4);
        x[i4 * 4] = v.x, x[i4 * 4 + 1] = v.y, x[i4 * 4 + 2] = v.z, x[i4 * 4 + 3] = v.w;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {                          // unrolled: x[8 q + i] is a static register
      mbar_wait(barL, phL);                                // operands of this quarter have landed
      phL ^= 1;
      // Every thread must have OBSERVED this phase of barL before thread 0 may start the next one: the next loads need
      // nobody's participation, so a warp that is held up between the __syncthreads() below and this wait (co-resident
      // CTAs of another stream) would otherwise find the barrier two phases ahead, read the parity as "not yet" and wait
      // forever.  Warps 1-3 arrive (non-blocking, whole warps: named barriers count warps) here; warp 0 syncs on the
      // same barrier, converged, right before its lane 0 issues the next loads.
      if (warp != 0) asm volatile("bar.arrive 1, 128;" ::: "memory");
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          uint32_t acc = 0;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t ah = make_desc(smem_u32(sA + (kb * 2 + 0) * kABlk)), am = make_desc(smem_u32(sA + (kb * 2 + 1) * kABlk));
            const uint64_t wh = make_desc(smem_u32(sW + (kb * 2 + 0) * kWBlk)), wm = make_desc(smem_u32(sW + (kb * 2 + 1) * kWBlk));
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {               // K = 16 bf16 = 32 B inside the 64-byte row
              mma_ss(tmem_base, ah + 2 * ks, wh + 2 * ks, idesc, acc);
              acc = 1;
              mma_ss(tmem_base, ah + 2 * ks, wm + 2 * ks, idesc, 1);
              mma_ss(tmem_base, am + 2 * ks, wh + 2 * ks, idesc, 1);
            }
          }
          mma_commit(barM);
        }
        __syncwarp();
      }
      mbar_wait(barM, phM);                                // D complete; sA (last quarter) and sW are free again
      phM ^= 1;
      tc_fence_after();
      if (warp == 0) {                                     // all 4 warps have passed this quarter's barL wait
        __syncwarp();
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (tid == 0) {                                      // next operands stream in under the epilogue
        const int tn = t + gridDim.x;
        if (q < 3) {
          mbar_expect_tx(barL, kWQuarter);
          bulk_g2s(smem_u32(sW), w4_packed + (int64_t)(q + 1) * kWQuarter, kWQuarter, barL);
        } else if (tn < tiles) {
          mbar_expect_tx(barL, kATile + kWQuarter);
          bulk_g2s(smem_u32(sA), he_packed + (int64_t)tn * kATile, kATile, barL);
          bulk_g2s(smem_u32(sW), w4_packed, kWQuarter, barL);
        }
      }
      // epilogue: lane = edge row; columns c = 32 (i - 8 q) + o
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int cc = 0; cc < kQ / 16; ++cc) {
        float v[16];
        tmem_ld16(trow + cc * 16, v);
        const int o0 = (cc & 1) * 16;
        const float *bb = s_b4 + q * kQ + cc * 16;
        const float xi = x[q * 8 + (cc >> 1)];
#pragma unroll
        for (int k = 0; k < 16; ++k) m[o0 + k] = fmaf(xi, v[k] + bb[k], m[o0 + k]);
      }
      tc_fence_before();
      __syncthreads();                                     // every lane has read D before the next quarter overwrites it
    }
    if (j < E) {
      float4 *mp = reinterpret_cast<float4 *>(msg + j * kF);
#pragma unroll
      for (int i4 = 0; i4 < kF / 4; ++i4) mp[i4] = make_float4(m[i4 * 4], m[i4 * 4 + 1], m[i4 * 4 + 2], m[i4 * 4 + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace eccu
}  // namespace wsis

using namespace wsis;

extern "C" {

int64_t wsis_ecc_edge_mlp_param_floats(void) { return eccu::kMlpParams; }
int64_t wsis_ecc_he_bytes(int64_t E) { return ceil_div(std::max<int64_t>(E, 1), eccu::kTile) * eccu::kATile; }
int64_t wsis_ecc_w4_bytes(void) { return 4 * (int64_t)eccu::kWQuarter; }

int wsis_ecc_pack_w4(const float *w4, void *w4_packed, wsis_stream_t stream) {
  eccu::pack_w4_kernel<<<eccu::kNF / 128, 128, 0, as_stream(stream)>>>(w4, (uint8_t *)w4_packed);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_ecc_edge_mlp(const float *edgefeats, const int32_t *eorder, int64_t E, const float *params, void *he_packed,
                      wsis_stream_t stream) {
  if (E == 0) return 0;
  size_t smem = sizeof(float) * eccu::kMlpParams;
  WSIS_CUDA(cudaFuncSetAttribute(eccu::edge_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eccu::edge_mlp_kernel<<<(unsigned)ceil_div(E, eccu::kTile), eccu::kTile, smem, as_stream(stream)>>>(
      edgefeats, eorder, E, params, (uint8_t *)he_packed);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_ecc_messages(const void *he_packed, const void *w4_packed, const float *b4, const float *h, const int64_t *src,
                      const int32_t *eorder, int64_t E, float *msg, wsis_stream_t stream) {
  if (E == 0) return 0;
  {
    static bool trap_set = false;
    if (!trap_set) {
      unsigned int *tw = trap_word_device();
      cudaMemcpyToSymbol(eccu::d_trap_word, &tw, sizeof(tw));
      trap_set = true;
    }
  }
  int tiles = (int)ceil_div(E, eccu::kTile);
  size_t smem = eccu::kATile + eccu::kWQuarter + sizeof(float) * eccu::kNF + 64 + 1024;
  WSIS_CUDA(cudaFuncSetAttribute(eccu::ecc_messages_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = std::min(tiles, 2 * sm_count());
  eccu::ecc_messages_kernel<<<grid, 128, smem, as_stream(stream)>>>((const uint8_t *)he_packed, (const uint8_t *)w4_packed,
                                                                  b4, h, src, eorder, E, tiles, msg);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
