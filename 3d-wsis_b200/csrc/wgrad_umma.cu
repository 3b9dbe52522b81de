// Weight gradient of the sparse convolution on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces the per-offset gather -> cuBLAS mm(in^T, dout) loop of indiceConvBackward
// (include/spconv/spconv_ops.h:351-433):
//
//   dW[k][ci][co] = sum_r  prologue(src[map[r, k'], ci]) * g[r, co]            k' = flip ? K-1-k : k
//
// The contraction runs over the ROWS, so both MMA operands are "MN-major": the gathered rows [row][channel] and the
// gradient rows [row][co] go to shared memory in their natural row-major form (64-byte-swizzled blocks of 32
// channels x 128 rows) and the instruction descriptor's a_major / b_major bits say so.  Four kernel offsets x 32
// input channels are stacked along M (= 128 TMEM lanes); the accumulator D[4 offsets x 32 ci][co] of every offset
// group stays in TMEM for the whole launch (ceil(K/4) groups x N columns <= 512) and is added to dW once at the end.
//
//   grid = (row chunks, Cin/32, Cout/N), N = 64 when Cout % 64 == 0 else 32; one CTA per SM, 256 threads.
//   per tile of 128 destination rows:  the tile's K neighbour indices and its gradient block go to shared memory once,
//   then per offset group: all threads gather + convert (fused BatchNorm+ReLU prologue, fp32 -> bf16 hi [+ mid]) the
//   group's 4 x 128 rows into one of two operand buffers, one elected lane issues the 8 K-steps (x3 for the fp32
//   contract: hi.hi + hi.mid + mid.hi) and commits to the buffer's mbarrier; the gather of group i+1 overlaps the
//   MMAs of group i.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace wsis {
namespace wgrad {

constexpr int kThreads = 512;
constexpr int kTile = 128;
constexpr int kBlk = 128 * 64;             // bytes of one [128 rows x 32 ch] bf16 block (SWIZZLE_64B, MN-major)
constexpr int kABuf = 4 * kBlk;            // 4 offsets stacked along M
constexpr int kMaxGroups = 8;
constexpr int kRowsPerWarp = kTile / (kThreads / 32);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t n = 0; n < (1u << 24) && !ok; ++n) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  if (!ok) __trap();   // a protocol bug must not hang the device
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// MN-major SWIZZLE_64B descriptor (cute::UMMA::SmemDescriptor): 32-element (64 B) groups along M/N at stride LBO,
// 8-row groups along K at stride SBO = 512 B; version 1; layout type SWIZZLE_64B = 4.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(kBlk >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
         (4ull << 61);
}
// kind::f16: D = f32, A = B = bf16, both MN-major (bits 15, 16), N >> 3, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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

// byte offset of (row, 16-byte chunk c16 in [0,4)) inside a [128 rows x 64 B] SWIZZLE_64B block
__device__ __forceinline__ uint32_t sw64(uint32_t row, uint32_t c16) {
  return (row >> 3) * 512u + (row & 7u) * 64u + ((c16 ^ ((row & 7u) >> 1)) << 4);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

// 8 consecutive channels -> one 16-byte chunk of hi and (SPLIT) one of mid = bf16(v - hi)
template <bool SPLIT>
__device__ __forceinline__ void convert8(const float (&v)[8], uint4 &hi, uint4 &mid) {
  uint32_t h[4], m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16(v[2 * i], v[2 * i + 1]);             // one F2FP.BF16.F32.PACK_AB
    if (SPLIT) {
      float r0 = v[2 * i] - __uint_as_float(h[i] << 16), r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xFFFF0000u);
      m[i] = pack_bf16(r0, r1);
    }
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  if (SPLIT) mid = make_uint4(m[0], m[1], m[2], m[3]);
}

struct Params {
  const float *src;
  const int32_t *map;
  const int32_t *order;   // destination rows in tile order (-1 padded) or NULL = identity
  const float *g;
  const float *scale, *shift;
  float *dW;
  int64_t n_dst;
  int K, flip, Cin, Cout, relu, N, tiles, tiles_per_cta, ngroups;
  int debug;   // timing experiments (WSIS_WGRAD_DEBUG): 1 = no MMAs, 2 = no global gather loads, 4 = no atomics, 8 = no G loads
};

// Shared memory: [A hi 2 x 32 KB][A mid 2 x 32 KB][G hi 2 x (N/32) x 8 KB][G mid ...][idx 128 x K][rows 128][barriers]
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1) wgrad_umma_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzled blocks: 1024-byte aligned
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = p.N / 32;
  uint8_t *sAh = smem;
  uint8_t *sAm = sAh + 2 * kABuf;
  uint8_t *sGh = sAm + (SPLIT ? 2 * kABuf : 0);
  uint8_t *sGm = sGh + 2 * nb * kBlk;
  int32_t *s_idx = reinterpret_cast<int32_t *>(sGm + (SPLIT ? 2 * nb * kBlk : 0));
  int32_t *s_row = s_idx + kTile * p.K;
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_row + kTile);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 4);

  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(p.tiles, t_begin + p.tiles_per_cta);
  if (t_begin >= t_end) return;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.z * p.N;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(smem_u32(s_bar + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t idesc = make_idesc(p.N);

  // per-thread slice of the prologue: the thread converts channels [ch0, ch0 + 16) of its entries
  const int half = tid & 1;
  float sc[16], sh[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    sc[i] = p.scale ? __ldg(p.scale + ci0 + half * 16 + i) : 1.f;
    sh[i] = p.shift ? __ldg(p.shift + ci0 + half * 16 + i) : 0.f;
  }

  int32_t nx_row[kRowsPerWarp], nx_idx[kRowsPerWarp];
  auto fetch_tile = [&](int tt) {
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
      const int64_t slot = (int64_t)tt * kTile + warp + i * (kThreads / 32);
      nx_row[i] = p.order ? __ldg(p.order + slot) : (slot < p.n_dst ? (int32_t)slot : -1);
    }
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i)
      nx_idx[i] = (nx_row[i] >= 0 && lane < p.K && !(p.debug & 16))
                      ? __ldg(p.map + (int64_t)nx_row[i] * p.K + (p.flip ? p.K - 1 - lane : lane))
                      : -1;
  };
  fetch_tile(t_begin);

  uint32_t iter = 0;
  for (int t = t_begin; t < t_end; ++t) {
    const int gb = (t - t_begin) & 1;
    // ---- tile set-up: destination rows, neighbour indices, gradient block ----
    // The rows and neighbour indices of a tile are fetched one tile ahead into registers (warp w owns rows w, w + 16,
    // ...; lane k owns offset k) and only parked in shared memory here, so their global-load latency is off the path.
    __syncthreads();                                      // everybody is done with s_idx / s_row of the previous tile
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
      const int r = warp + i * (kThreads / 32);
      if (lane == 0) s_row[r] = nx_row[i];
      if (lane < p.K) s_idx[r * p.K + lane] = nx_idx[i];
    }
    __syncthreads();
    if (t + 1 < t_end) fetch_tile(t + 1);
    // G[gb]: rows x N channels, 8 channels (one 16-byte chunk) per work item
    const int c8bits = p.N == 64 ? 3 : 2;
    for (int e = tid; e < (kTile << c8bits); e += kThreads) {
      int r = e >> c8bits, c8 = e & ((1 << c8bits) - 1);
      int32_t row = s_row[r];
      uint4 hi = make_uint4(0, 0, 0, 0), mid = make_uint4(0, 0, 0, 0);
      if (row >= 0 && !(p.debug & 8)) {
        const float4 *gp = reinterpret_cast<const float4 *>(p.g + (int64_t)row * p.Cout + co0 + c8 * 8);
        float4 a = __ldg(gp), b = __ldg(gp + 1);
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        convert8<SPLIT>(v, hi, mid);
      }
      uint32_t off = (uint32_t)(gb * nb + (c8 >> 2)) * kBlk + sw64(r, c8 & 3);
      *reinterpret_cast<uint4 *>(sGh + off) = hi;
      if (SPLIT) *reinterpret_cast<uint4 *>(sGm + off) = mid;
    }
    __syncthreads();                                      // s_idx complete

    for (int grp = 0; grp < p.ngroups; ++grp, ++iter) {
      const int ab = iter & 1;
      if (iter >= 2) {                                    // the MMAs that read this operand buffer have completed
        mbar_wait(smem_u32(s_bar + ab), ((iter >> 1) - 1) & 1);
        tc_fence_after();
      }
      // ---- gather: 4 offsets x 128 rows x 32 channels; 2 threads per (row, offset) entry, 16 channels each ----
      // 2 threads per (offset j, row r) entry, 16 channels each; a thread owns entries (j, r) and (j + 2, r)
      const int r = (tid >> 1) & 127, j0 = tid >> 8;
      int32_t idx2[2];
      float4 raw[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {                       // all loads of the thread are issued before any is used
        const int k = grp * 4 + j0 + 2 * u;
        idx2[u] = (k < p.K) ? s_idx[r * p.K + k] : -1;
        if (p.debug & 2) idx2[u] = -1;
        if (idx2[u] >= 0) {
          const float4 *sp = reinterpret_cast<const float4 *>(p.src + (int64_t)idx2[u] * p.Cin + ci0 + half * 16);
          raw[u][0] = __ldg(sp), raw[u][1] = __ldg(sp + 1), raw[u][2] = __ldg(sp + 2), raw[u][3] = __ldg(sp + 3);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint4 h0 = make_uint4(0, 0, 0, 0), m0 = h0, h1 = h0, m1 = h0;
        if (idx2[u] >= 0) {                               // entries without a neighbour are plain zero stores
          float4 a = raw[u][0], b = raw[u][1], c = raw[u][2], d = raw[u][3];
          float lo8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}, hi8[8] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float y0 = fmaf(lo8[i], sc[i], sh[i]), y1 = fmaf(hi8[i], sc[8 + i], sh[8 + i]);
            lo8[i] = p.relu ? fmaxf(y0, 0.f) : y0;
            hi8[i] = p.relu ? fmaxf(y1, 0.f) : y1;
          }
          convert8<SPLIT>(lo8, h0, m0);
          convert8<SPLIT>(hi8, h1, m1);
        }
        const uint32_t base = (uint32_t)ab * kABuf + (uint32_t)(j0 + 2 * u) * kBlk;
        const uint32_t o0 = base + sw64(r, half * 2), o1 = base + sw64(r, half * 2 + 1);
        if (p.debug & 32) continue;
        *reinterpret_cast<uint4 *>(sAh + o0) = h0;
        *reinterpret_cast<uint4 *>(sAh + o1) = h1;
        if (SPLIT) {
          *reinterpret_cast<uint4 *>(sAm + o0) = m0;
          *reinterpret_cast<uint4 *>(sAm + o1) = m1;
        }
      }
      if (!(p.debug & 64)) proxy_fence();                 // generic-proxy writes -> visible to the tensor core
      __syncthreads();
      // ---- issue: one elected lane of warp 0 ----
      if (warp == 0) {
        tc_fence_after();
        if (elect_one() && !(p.debug & 1)) {
          const uint32_t d = tmem_base + (uint32_t)(grp * p.N);
          const uint64_t ah = make_desc(smem_u32(sAh + ab * kABuf)), am = make_desc(smem_u32(sAm + ab * kABuf));
          const uint64_t gh = make_desc(smem_u32(sGh + gb * nb * kBlk)), gm = make_desc(smem_u32(sGm + gb * nb * kBlk));
          const uint32_t first = (t == t_begin) ? 0u : 1u;
#pragma unroll 1
          for (int ks = 0; ks < 8; ++ks) {                // K = 16 rows per MMA = two 8-row groups = 1024 B
            const uint64_t adv = (uint64_t)(ks * (1024 >> 4));
            mma_ss(d, ah + adv, gh + adv, idesc, (ks == 0) ? first : 1u);
            if (SPLIT) {
              mma_ss(d, ah + adv, gm + adv, idesc, 1u);
              mma_ss(d, am + adv, gh + adv, idesc, 1u);
            }
          }
          mma_commit(smem_u32(s_bar + ab));
        } else if ((p.debug & 1) && elect_one()) {
          mma_commit(smem_u32(s_bar + ab));
        }
        __syncwarp();
      }
    }
  }
  // ---- drain: every MMA has completed, then TMEM -> dW ----
  if (warp == 0) {
    if (elect_one()) mma_commit(smem_u32(s_bar + 2));
    __syncwarp();
  }
  mbar_wait(smem_u32(s_bar + 2), 0);
  tc_fence_after();
  {
    const int q = warp & 3;                               // this warp's TMEM lane quarter = offset j of every group
    for (int grp = warp >> 2; grp < p.ngroups; grp += kThreads / 128) {
      const int k = grp * 4 + q;
      if (k >= p.K) continue;
      float *out = p.dW + ((int64_t)k * p.Cin + ci0 + lane) * p.Cout + co0;
      for (int c = 0; c < p.N; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * p.N + c), v);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (v[i] != 0.f && !(p.debug & 4)) atomicAdd(out + c + i, v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static size_t smem_bytes(int K, int N, bool split) {
  size_t nb = N / 32;
  size_t b = 2 * kABuf + 2 * nb * kBlk;
  if (split) b *= 2;
  return b + sizeof(int32_t) * (kTile * K + kTile) + 64 + 1024;   // + barriers/TMEM pointer + alignment slack
}

}  // namespace wgrad
}  // namespace wsis

using namespace wsis;

extern "C" {

int wsis_conv_wgrad_umma_supported(int K, int Cin, int Cout) {
  return (K >= 5 && K <= 4 * wgrad::kMaxGroups && Cin % 32 == 0 && Cout % 32 == 0 && Cin >= 32 && Cout >= 32) ? 1 : 0;
}

int wsis_conv_wgrad_umma(const float *src, const int32_t *map, const int32_t *order, int64_t n_dst, int K, int flip,
                         const float *g, int Cin, int Cout, const float *in_scale, const float *in_shift, int in_relu,
                         int precision, float *dW, wsis_stream_t stream) {
  WSIS_CHECK(wsis_conv_wgrad_umma_supported(K, Cin, Cout), "wgrad_umma: needs K <= 32, Cin %% 32 == 0, Cout %% 32 == 0");
  WSIS_CHECK(precision == 1 || precision == 3, "wgrad_umma: precision 1 (bf16) or 3 (bf16x3 split)");
  WSIS_CHECK((in_scale == nullptr) == (in_shift == nullptr), "wgrad_umma: in_scale/in_shift must both be set");
  cudaStream_t st = as_stream(stream);
  WSIS_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)K * Cin * Cout, st));
  if (n_dst == 0) return 0;
  wgrad::Params p;
  p.src = src, p.map = map, p.order = order, p.g = g, p.scale = in_scale, p.shift = in_shift, p.dW = dW;
  p.n_dst = n_dst, p.K = K, p.flip = flip, p.Cin = Cin, p.Cout = Cout, p.relu = in_relu;
  p.ngroups = (K + 3) / 4;
  {
    const char *dbg = getenv("WSIS_WGRAD_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  p.N = (Cout % 64 == 0 && p.ngroups * 64 <= 512) ? 64 : 32;
  p.tiles = (int)ceil_div(n_dst, wgrad::kTile);
  int blocks_c = (Cin / 32) * (Cout / p.N);
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(p.tiles, (sm_count() + blocks_c - 1) / blocks_c));
  p.tiles_per_cta = (p.tiles + chunks - 1) / chunks;
  chunks = (p.tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  const bool split = precision == 3;
  size_t smem = wgrad::smem_bytes(K, p.N, split);
  auto kern = split ? wgrad::wgrad_umma_kernel<true> : wgrad::wgrad_umma_kernel<false>;
  WSIS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((unsigned)chunks, (unsigned)(Cin / 32), (unsigned)(Cout / p.N)), wgrad::kThreads, smem, st>>>(p);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
