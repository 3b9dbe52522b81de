// Output-stationary sparse convolution on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces the reference's per-offset gather -> cuBLAS mm -> scatter-add loop
// (include/spconv/spconv_ops.h:296-344; kernels include/spconv/reordering.cu.h:21-157).
//
//   dst[r,:] = residual[r,:] + sum_k  prologue(src[map[r,k'],:]) . W[k]        k' = flip ? K-1-k : k
//
// One persistent CTA per SM; a CTA owns tiles of 128 OUTPUT rows (= the 128 TMEM lanes of one accumulator):
//   warps 0-3   epilogue : tcgen05.ld accumulator -> registers -> (+residual) -> global, once per tile
//   warps 4-11  loaders  : read the tile's slice of the neighbour map, gather the fp32 source rows straight
//                          from global/L2 with 16-byte loads, apply the fused eval-BatchNorm+ReLU prologue,
//                          split fp32 -> bf16 hi (+ bf16 mid for the 1e-4 path) and store them in the
//                          128B-swizzle-64 K-major layout the UMMA descriptors expect
//   warp 12     MMA      : one elected thread issues tcgen05.mma (M=128, N=Cout, K=16) into TMEM
//   warp 13     weights  : cp.async.bulk (1-D TMA) of the pre-swizzled weight block of the unit into smem
// The pipeline unit is (kernel offset k, 32-channel block kb); units flow through an NSTAGE mbarrier ring
// (full: loaders + bulk-copy tx bytes -> MMA; empty: tcgen05.commit -> loaders / weight warp).  The
// accumulator is double buffered in TMEM so the epilogue of tile t overlaps the MMAs of tile t+1.
// There is no scatter and there are no atomics: every output row is written exactly once.
//
// Precision: precision==1 uses bf16 operands (fp32 accumulate).  precision==3 splits both operands into
// bf16 hi + bf16 mid and issues hi.hi + hi.mid + mid.hi (error ~2^-17 per product, fp32 accumulate in TMEM)
// to honour the reference's fp32 contract (1e-4) while staying on the tensor pipe.
#include <cuda_bf16.h>

#include "common.cuh"

namespace wsis {
namespace umma {

constexpr int kTileM = 128;
constexpr int kKB = 32;                  // channels per pipeline unit (64 bytes of bf16 = one swizzle-64 row)
constexpr int kABlockBytes = kTileM * 64;  // 8 KB
constexpr int kEpiWarps = 4, kLoadWarps = 8;
constexpr int kThreads = (kEpiWarps + kLoadWarps + 2) * 32;  // 448
constexpr int kLoaderThreads = kLoadWarps * 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (8 rows * 64 B = 512)
// | version=1 [46,48) | layout_type SWIZZLE_64B = 4 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, K-major both, N>>3, M>>4
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
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

// byte offset of (row, 16-byte chunk c16 in [0,4)) inside a [rows x 64 B] K-major SWIZZLE_64B block
__host__ __device__ __forceinline__ uint32_t sw64(uint32_t row, uint32_t c16) {
  return (row >> 3) * 512u + (row & 7u) * 64u + ((c16 ^ ((row & 7u) >> 1)) << 4);
}

struct Params {
  const float *src;
  const int32_t *map;
  const uint8_t *packed;
  const float *in_scale, *in_shift, *residual;
  float *dst;
  int64_t n_dst;
  int K, flip, Cin, Cout, in_relu, nstage, tmem_cols;
  int64_t num_tiles;
};

struct Pipe {
  uint32_t stage = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t n) {
    if (++stage == n) {
      stage = 0;
      phase ^= 1;
    }
  }
};

template <int NS>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t *sm = smem_raw + pad;
  const int KB = p.Cin / kKB;
  const uint32_t a_bytes = NS * kABlockBytes;
  const uint32_t b_block = (uint32_t)p.Cout * 64u;
  const uint32_t b_bytes = NS * b_block;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t *tail = sm + (size_t)p.nstage * stage_bytes;
  int32_t *s_nbr = reinterpret_cast<int32_t *>(tail);
  float *s_scale = reinterpret_cast<float *>(tail + ((kTileM * p.K * 4 + 15) & ~15));
  float *s_shift = s_scale + p.Cin;
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_shift + p.Cin);  // Cin % 32 == 0 keeps 8-byte alignment
  // bars: full[nstage], empty[nstage], acc_full[2], acc_empty[2]
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (p.nstage + s); };
  auto accf_bar = [&](uint32_t s) { return bar0 + 8u * (2 * p.nstage + s); };
  auto acce_bar = [&](uint32_t s) { return bar0 + 8u * (2 * p.nstage + 2 + s); };
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.nstage + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int c = tid; c < p.Cin; c += kThreads) {
    s_scale[c] = p.in_scale ? p.in_scale[c] : 1.f;
    s_shift[c] = p.in_shift ? p.in_shift[c] : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < p.nstage; ++s) {
      mbar_init(full_bar(s), kLoadWarps + 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accf_bar(s), 1);
      mbar_init(acce_bar(s), kEpiWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps + kLoadWarps) {  // MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int units = p.K * KB;

  if (warp < kEpiWarps) {
    // ===================== epilogue =====================
    uint32_t as = 0, aph = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(accf_bar(as), aph);
      tc_fence_after();
      const int64_t row = tile * kTileM + warp * 32 + lane;
      const bool live = row < p.n_dst;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * (uint32_t)p.Cout;
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (live) {
          float4 *o = reinterpret_cast<float4 *>(p.dst + row * p.Cout + c0);
          if (p.residual) {
            const float4 *rs = reinterpret_cast<const float4 *>(p.residual + row * p.Cout + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 r4 = __ldg(rs + q);
              v[4 * q] += r4.x;
              v[4 * q + 1] += r4.y;
              v[4 * q + 2] += r4.z;
              v[4 * q + 3] += r4.w;
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(acce_bar(as));
      if (++as == 2) {
        as = 0;
        aph ^= 1;
      }
    }
  } else if (warp < kEpiWarps + kLoadWarps) {
    // ===================== loaders =====================
    const int lt = tid - kEpiWarps * 32;
    const int chunk = lt & 7;   // 4 fp32 channels = 8 bytes of bf16
    const int rbase = lt >> 3;  // rows rbase + 32*j
    const uint32_t c16 = chunk >> 1, sub = (chunk & 1) * 8;
    Pipe pipe;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      asm volatile("bar.sync 1, %0;" ::"n"(kLoaderThreads) : "memory");
      {
        const int64_t r0 = tile * kTileM;
        const int64_t lim = (min((int64_t)kTileM, p.n_dst - r0)) * p.K;
        const int32_t *mp = p.map + r0 * p.K;
        for (int e = lt; e < kTileM * p.K; e += kLoaderThreads) s_nbr[e] = e < lim ? __ldg(mp + e) : -1;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kLoaderThreads) : "memory");

      float4 cur[4], nxt[4];
      auto fetch = [&](int unit, float4 *v) {
        const int k = unit / KB, kb = unit - k * KB;
        const int kk = p.flip ? p.K - 1 - k : k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int32_t idx = s_nbr[(rbase + 32 * j) * p.K + kk];
          if (idx >= 0) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(p.src + (int64_t)idx * p.Cin + kb * kKB) + chunk);
            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + kb * kKB + chunk * 4);
            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + kb * kKB + chunk * 4);
            float4 y = make_float4(fmaf(x.x, sc.x, sh.x), fmaf(x.y, sc.y, sh.y), fmaf(x.z, sc.z, sh.z),
                                   fmaf(x.w, sc.w, sh.w));
            if (p.in_relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
            v[j] = y;
          } else {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      fetch(0, cur);
      for (int unit = 0; unit < units; ++unit) {
        if (unit + 1 < units) fetch(unit + 1, nxt);
        mbar_wait(empty_bar(pipe.stage), pipe.phase ^ 1);
        uint8_t *abase = sm + (size_t)pipe.stage * stage_bytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = sw64(rbase + 32 * j, c16) + sub;
          const float4 y = cur[j];
          __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
          uint2 hv = make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
          *reinterpret_cast<uint2 *>(abase + off) = hv;
          if (NS == 2) {
            float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            __nv_bfloat162 m0 = __floats2bfloat162_rn(y.x - f0.x, y.y - f0.y);
            __nv_bfloat162 m1 = __floats2bfloat162_rn(y.z - f1.x, y.w - f1.y);
            uint2 mv = make_uint2(*reinterpret_cast<uint32_t *>(&m0), *reinterpret_cast<uint32_t *>(&m1));
            *reinterpret_cast<uint2 *>(abase + kABlockBytes + off) = mv;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(pipe.stage));
        pipe.advance(p.nstage);
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
      }
    }
  } else if (warp == kEpiWarps + kLoadWarps) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.Cout);
      Pipe pipe;
      uint32_t as = 0, aph = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(acce_bar(as), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)p.Cout;
        uint32_t acc = 0;
        for (int unit = 0; unit < units; ++unit) {
          mbar_wait(full_bar(pipe.stage), pipe.phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sm + (size_t)pipe.stage * stage_bytes);
          const uint32_t b0 = a0 + a_bytes;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ah = make_desc(a0 + ks * 32), bh = make_desc(b0 + ks * 32);
            mma_bf16(d_tmem, ah, bh, idesc, acc);
            acc = 1;
            if (NS == 2) {
              const uint64_t am = make_desc(a0 + kABlockBytes + ks * 32), bm = make_desc(b0 + b_block + ks * 32);
              mma_bf16(d_tmem, ah, bm, idesc, 1);
              mma_bf16(d_tmem, am, bh, idesc, 1);
            }
          }
          mma_commit(empty_bar(pipe.stage));
          pipe.advance(p.nstage);
        }
        mma_commit(accf_bar(as));
        if (++as == 2) {
          as = 0;
          aph ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== weight producer =====================
    if (lane == 0) {
      Pipe pipe;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int unit = 0; unit < units; ++unit) {
          mbar_wait(empty_bar(pipe.stage), pipe.phase ^ 1);
          mbar_expect_tx(full_bar(pipe.stage), b_bytes);
          bulk_g2s(smem_u32(sm + (size_t)pipe.stage * stage_bytes + a_bytes), p.packed + (size_t)unit * b_bytes, b_bytes,
                   full_bar(pipe.stage));
          pipe.advance(p.nstage);
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + kLoadWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// W[K, Cin_w, Cout_w] fp32 -> [K][KB][NS][N x 64 B swizzled] bf16 (hi, mid)
__global__ void pack_weights_kernel(const float *__restrict__ W, int K, int Cin, int Cout, int transpose_w, int NS,
                                    uint8_t *__restrict__ packed) {
  const int KB = Cin / kKB;
  const int64_t total = (int64_t)K * KB * Cout * kKB;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int kk = (int)(e % kKB);
    int n = (int)((e / kKB) % Cout);
    int kb = (int)((e / ((int64_t)kKB * Cout)) % KB);
    int k = (int)(e / ((int64_t)kKB * Cout * KB));
    int c = kb * kKB + kk;
    float w = transpose_w ? W[((int64_t)k * Cout + n) * Cin + c] : W[((int64_t)k * Cin + c) * Cout + n];
    __nv_bfloat16 hi = __float2bfloat16_rn(w);
    size_t block = ((size_t)k * KB + kb) * NS * (size_t)Cout * 64;
    uint32_t off = sw64((uint32_t)n, (uint32_t)(kk >> 3)) + (kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16 *>(packed + block + off) = hi;
    if (NS == 2) {
      __nv_bfloat16 mid = __float2bfloat16_rn(w - __bfloat162float(hi));
      *reinterpret_cast<__nv_bfloat16 *>(packed + block + (size_t)Cout * 64 + off) = mid;
    }
  }
}

}  // namespace umma
}  // namespace wsis

using namespace wsis;
using namespace wsis::umma;

extern "C" {

int wsis_conv_umma_supported(int Cin, int Cout) {
  return Cin >= 32 && Cin % 32 == 0 && Cout >= 16 && Cout % 16 == 0 && Cout <= 256;
}

int64_t wsis_conv_pack_bytes(int K, int Cin, int Cout, int precision) {
  int NS = precision == 3 ? 2 : 1;
  return (int64_t)K * (Cin / kKB) * NS * Cout * 64;
}

int wsis_conv_pack_weights(const float *W, int K, int Cin, int Cout, int transpose_w, int precision, void *packed,
                           wsis_stream_t stream) {
  WSIS_CHECK(wsis_conv_umma_supported(Cin, Cout), "pack_weights: unsupported Cin=%d Cout=%d", Cin, Cout);
  WSIS_CHECK(precision == 1 || precision == 3, "pack_weights: precision must be 1 or 3");
  WSIS_CHECK((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "pack_weights: packed must be 16-byte aligned");
  int64_t total = (int64_t)K * Cin * Cout;
  unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 8);
  pack_weights_kernel<<<blocks, 256, 0, as_stream(stream)>>>(W, K, Cin, Cout, transpose_w, precision == 3 ? 2 : 1,
                                                             (uint8_t *)packed);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_conv_umma(const float *src, const int32_t *map, int64_t n_dst, int K, int flip, const void *packed, int Cin,
                   int Cout, int precision, const float *in_scale, const float *in_shift, int in_relu,
                   const float *residual, float *dst, wsis_stream_t stream) {
  WSIS_CHECK(wsis_conv_umma_supported(Cin, Cout), "conv_umma: unsupported Cin=%d Cout=%d", Cin, Cout);
  WSIS_CHECK(precision == 1 || precision == 3, "conv_umma: precision must be 1 or 3");
  WSIS_CHECK((in_scale == nullptr) == (in_shift == nullptr), "conv_umma: in_scale/in_shift must both be set");
  WSIS_CHECK(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) |
               reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(packed)) & 15) == 0,
             "conv_umma: src/dst/residual/packed must be 16-byte aligned");
  if (n_dst == 0) return 0;
  const int NS = precision == 3 ? 2 : 1;
  Params p;
  p.src = src;
  p.map = map;
  p.packed = (const uint8_t *)packed;
  p.in_scale = in_scale;
  p.in_shift = in_shift;
  p.residual = residual;
  p.dst = dst;
  p.n_dst = n_dst;
  p.K = K;
  p.flip = flip;
  p.Cin = Cin;
  p.Cout = Cout;
  p.in_relu = in_relu;
  p.num_tiles = ceil_div(n_dst, kTileM);
  int cols = 32;
  while (cols < 2 * Cout) cols <<= 1;
  p.tmem_cols = cols;
  const int64_t stage_bytes = (int64_t)NS * kABlockBytes + (int64_t)NS * Cout * 64;
  const int64_t fixed = 1024 /*align*/ + ((kTileM * K * 4 + 15) & ~15) + 2 * (Cin + 1) * 4 + (2 * 8 + 4) * 8 + 64;
  const int64_t budget = 220 * 1024;
  int nstage = (int)std::min<int64_t>(8, (budget - fixed) / stage_bytes);
  WSIS_CHECK(nstage >= 2, "conv_umma: shared memory budget too small for Cin=%d Cout=%d K=%d", Cin, Cout, K);
  p.nstage = nstage;
  const int64_t smem = fixed + nstage * stage_bytes;
  auto kern = NS == 2 ? conv_umma_kernel<2> : conv_umma_kernel<1>;
  static int64_t smem_set[2] = {0, 0};
  if (smem > smem_set[NS - 1]) {
    WSIS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[NS - 1] = smem;
  }
  unsigned grid = (unsigned)std::min<int64_t>(p.num_tiles, sm_count());
  kern<<<grid, kThreads, smem, as_stream(stream)>>>(p);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
