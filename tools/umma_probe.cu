// tcgen05.mma cost probe for the shapes the sparse conv issues (M=128, N=Cout, K=16, bf16 hi/mid split), sm_100a.
//
// Questions it answers (numbers land in profiles/, DESIGN.md 4 quotes them):
//   * what does ONE short MMA cost when A comes from shared memory (SS) vs from tensor memory (TS)?
//     An M=128 x K=16 bf16 A operand is 4 KB of shared-memory reads per MMA whatever N is, so for N <= 64 the
//     SS form is bound by the operand fetch, not by the N/2-cycle tensor floor.
//   * does the disable-output-lane mask, a second issuer, a second accumulate chain or a commit per unit change it?
//   * is the A-in-TMEM layout what the conv kernel assumes (lane = row, 32-bit column j = bf16 elements 2j, 2j+1)?
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/umma_probe tools/umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t uni(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint4 off) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %5, %6, %7}, p;\n}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(off.x), "r"(off.y), "r"(off.z), "r"(off.w)
      : "memory");
}
__device__ __forceinline__ void mma_ss_nomask(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint4 off) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%4, %5, %6, %7}, p;\n}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(off.x), "r"(off.y), "r"(off.z), "r"(off.w)
      : "memory");
}
__device__ __forceinline__ void mma_ts_nomask(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__host__ __device__ __forceinline__ uint32_t sw64(uint32_t row, uint32_t c16) {
  return (row >> 3) * 512u + (row & 7u) * 64u + ((c16 ^ ((row & 7u) >> 1)) << 4);
}

// test operands, exactly representable in bf16
__host__ __device__ inline float a_hi(int m, int c) { return (float)((m * 3 + c * 5) % 17 - 8) * 0.25f; }
__host__ __device__ inline float a_mid(int m, int c) { return (float)((m + c) % 5 - 2) * (1.f / 1024.f); }
__host__ __device__ inline float b_hi(int n, int c) { return (float)((n * 7 + c * 3) % 13 - 6) * 0.5f; }
__host__ __device__ inline float b_mid(int n, int c) { return (float)((n + 2 * c) % 7 - 3) * (1.f / 512.f); }

struct Cfg {
  int ts, masked, N, nchain, nissuer, units, commit_each, st_probe;
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}

__global__ void __launch_bounds__(256, 1) probe_kernel(Cfg c, long long *cycles, float *dout) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t *sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *sAh = sm, *sAm = sm + 8192, *sBh = sm + 16384, *sBm = sBh + 256 * 64;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sBm + 256 * 64);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = c.N;
  // operands in shared memory (A for the SS form, B for both)
  for (int e = tid; e < 128 * 32; e += blockDim.x) {
    const int m = e >> 5, ch = e & 31;
    const uint32_t off = sw64(m, ch >> 3) + (ch & 7) * 2;
    *reinterpret_cast<__nv_bfloat16 *>(sAh + off) = __float2bfloat16_rn(a_hi(m, ch));
    *reinterpret_cast<__nv_bfloat16 *>(sAm + off) = __float2bfloat16_rn(a_mid(m, ch));
  }
  for (int e = tid; e < N * 32; e += blockDim.x) {
    const int n = e >> 5, ch = e & 31;
    const uint32_t off = sw64(n, ch >> 3) + (ch & 7) * 2;
    *reinterpret_cast<__nv_bfloat16 *>(sBh + off) = __float2bfloat16_rn(b_hi(n, ch));
    *reinterpret_cast<__nv_bfloat16 *>(sBm + off) = __float2bfloat16_rn(b_mid(n, ch));
  }
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *s_tmem;
  // TMEM map: accumulator 0 at column 0, accumulator 1 at column 256 - (N..) , A hi at 448, A mid at 464
  const uint32_t colD0 = 0, colD1 = 224, colA = 448;
  if (warp < 4) {
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 448; c0 += 16)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
                       taddr + c0),
                   "r"(0u)
                   : "memory");
    // A operand of the TS form: lane = row m, 32-bit column j of a 16-column group = channels (2j, 2j+1)
    const int m = warp * 32 + lane;
    uint32_t h[16], md[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      h[j] = pack2(a_hi(m, 2 * j), a_hi(m, 2 * j + 1));
      md[j] = pack2(a_mid(m, 2 * j), a_mid(m, 2 * j + 1));
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            taddr + colA),
        "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]), "r"(h[8]), "r"(h[9]),
        "r"(h[10]), "r"(h[11]), "r"(h[12]), "r"(h[13]), "r"(h[14]), "r"(h[15])
        : "memory");
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            taddr + colA + 16),
        "r"(md[0]), "r"(md[1]), "r"(md[2]), "r"(md[3]), "r"(md[4]), "r"(md[5]), "r"(md[6]), "r"(md[7]), "r"(md[8]),
        "r"(md[9]), "r"(md[10]), "r"(md[11]), "r"(md[12]), "r"(md[13]), "r"(md[14]), "r"(md[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (c.st_probe) {  // cost of the builders' TMEM stores: 4 warps x (units) x two x16 stores
      __syncwarp();
      const long long t0 = clock64();
      for (int u = 0; u < c.units; ++u) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                taddr + colA),
            "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]), "r"(h[8]), "r"(h[9]),
            "r"(h[10]), "r"(h[11]), "r"(h[12]), "r"(h[13]), "r"(h[14]), "r"(h[15])
            : "memory");
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                taddr + colA + 16),
            "r"(md[0]), "r"(md[1]), "r"(md[2]), "r"(md[3]), "r"(md[4]), "r"(md[5]), "r"(md[6]), "r"(md[7]), "r"(md[8]),
            "r"(md[9]), "r"(md[10]), "r"(md[11]), "r"(md[12]), "r"(md[13]), "r"(md[14]), "r"(md[15])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      const long long t1 = clock64();
      if (tid == 0) cycles[gridDim.x + blockIdx.x] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp >= 4 && warp < 4 + c.nissuer) {
    const uint32_t mi = uni((uint32_t)(warp - 4));
    const uint32_t idesc = make_idesc(N);
    const uint64_t adh = make_desc(smem_u32(sAh)), adm = make_desc(smem_u32(sAm));
    const uint64_t bdh = make_desc(smem_u32(sBh)), bdm = make_desc(smem_u32(sBm));
    const uint32_t tb = uni(tbase);
    // with two issuers each owns one accumulator; with one issuer and two chains it alternates
    const uint32_t d0 = tb + (mi ? colD1 : colD0);
    const uint32_t d1 = (c.nchain > 1 && c.nissuer == 1) ? tb + colD1 : d0;
    const uint32_t ah = tb + colA, am = tb + colA + 16;
    const uint4 off = c.masked ? make_uint4(0xDB6DB6DBu, 0x6DB6DB6Du, 0xB6DB6DB6u, 0xDB6DB6DBu) : make_uint4(0, 0, 0, 0);
    const uint32_t fin = smem_u32(bars + mi), scratch = smem_u32(bars + 4 + mi);
    __syncwarp();
    const long long t0 = clock64();
    for (int u = 0; u < c.units; ++u) {
      if (elect_one()) {
        if (c.ts) {
          if (c.masked) {
            mma_ts(d0, ah, bdh, idesc, off);
            mma_ts(d1, ah, bdm, idesc, off);
            mma_ts(d0, am, bdh, idesc, off);
            mma_ts(d1, ah + 8, bdh + 2, idesc, off);
            mma_ts(d0, ah + 8, bdm + 2, idesc, off);
            mma_ts(d1, am + 8, bdh + 2, idesc, off);
          } else {
            mma_ts_nomask(d0, ah, bdh, idesc);
            mma_ts_nomask(d1, ah, bdm, idesc);
            mma_ts_nomask(d0, am, bdh, idesc);
            mma_ts_nomask(d1, ah + 8, bdh + 2, idesc);
            mma_ts_nomask(d0, ah + 8, bdm + 2, idesc);
            mma_ts_nomask(d1, am + 8, bdh + 2, idesc);
          }
        } else {
          if (c.masked) {
            mma_ss(d0, adh, bdh, idesc, off);
            mma_ss(d1, adh, bdm, idesc, off);
            mma_ss(d0, adm, bdh, idesc, off);
            mma_ss(d1, adh + 2, bdh + 2, idesc, off);
            mma_ss(d0, adh + 2, bdm + 2, idesc, off);
            mma_ss(d1, adm + 2, bdh + 2, idesc, off);
          } else {
            mma_ss_nomask(d0, adh, bdh, idesc);
            mma_ss_nomask(d1, adh, bdm, idesc);
            mma_ss_nomask(d0, adm, bdh, idesc);
            mma_ss_nomask(d1, adh + 2, bdh + 2, idesc);
            mma_ss_nomask(d0, adh + 2, bdm + 2, idesc);
            mma_ss_nomask(d1, adm + 2, bdh + 2, idesc);
          }
        }
        if (c.commit_each) mma_commit(scratch);
      }
      __syncwarp();
    }
    const long long t1 = clock64();  // issue time only
    if (elect_one()) mma_commit(fin);
    __syncwarp();
    mbar_wait(fin, 0);
    const long long t2 = clock64();
    if (lane == 0 && mi == 0) {
      cycles[blockIdx.x] = t2 - t0;
      cycles[2 * gridDim.x + blockIdx.x] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4 && blockIdx.x == 0 && dout != nullptr) {
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16], q[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
          "tcgen05.wait::ld.sync.aligned;\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr + colD0 + c0)
          : "memory");
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
          "tcgen05.wait::ld.sync.aligned;\n"
          : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
            "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
          : "r"(taddr + colD1 + c0)
          : "memory");
      const bool two = (c.nchain > 1 || c.nissuer > 1);
      for (int i = 0; i < 16; ++i)
        dout[(size_t)(warp * 32 + lane) * 256 + c0 + i] = __uint_as_float(r[i]) + (two ? __uint_as_float(q[i]) : 0.f);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

// ---- hand-off latency: two warps bounce a token through two mbarriers; mode 0: both sides arrive with mbarrier.arrive,
// mode 1: the second warp answers with tcgen05.commit (no MMA outstanding), mode 2: commit after one short MMA
__global__ void __launch_bounds__(256, 1) pingpong_kernel(int mode, int rounds, long long *cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t *sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t *sB = sm + 16384;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm + 16384 + 16384);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 4096; e += blockDim.x) reinterpret_cast<uint32_t *>(sB)[e] = 0;
  if (tid == 0) {
    mbar_init(smem_u32(bars + 0), mode >= 3 ? 4 : 1);
    mbar_init(smem_u32(bars + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = *s_tmem;
  const uint32_t b0 = smem_u32(bars + 0), b1 = smem_u32(bars + 1);
  if (mode >= 3) {
    // modes 3..5: warps 0-3 are a builder group (b0 counts their 4 arrivals), warp 4 the issuer.
    //   3: fences only; 4: + tcgen05.wait::st; 5: + a 32-column tcgen05.st per round
    if (warp < 4) {
      const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + 384;
      __syncwarp();
      const long long t0 = clock64();
      for (int r = 0; r < rounds; ++r) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mode >= 5)
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,"
              "%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"((uint32_t)r)
              : "memory");
        if (mode >= 4) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0) : "memory");
        mbar_wait(b1, r & 1);
      }
      const long long t1 = clock64();
      if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    } else if (warp == 4) {
      for (int r = 0; r < rounds; ++r) {
        mbar_wait(b0, r & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) mma_commit(b1);
        __syncwarp();
      }
    }
  } else if (warp == 0) {
    __syncwarp();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0) : "memory");
      mbar_wait(b1, r & 1);
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
  } else if (warp == 4) {
    const uint32_t idesc = make_idesc(32);
    const uint64_t bd = make_desc(smem_u32(sB));
    for (int r = 0; r < rounds; ++r) {
      mbar_wait(b0, r & 1);
      if (mode == 0) {
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory");
      } else {
        if (elect_one()) {
          if (mode == 2) mma_ts_nomask(tb, tb + 448, bd, idesc);
          mma_commit(b1);
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

static double median(std::vector<long long> v) {
  std::sort(v.begin(), v.end());
  return (double)v[v.size() / 2];
}

int main() {
  int dev_sms = 0;
  CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = dev_sms;
  long long *d_cyc;
  float *d_out;
  CK(cudaMalloc(&d_cyc, sizeof(long long) * grid * 3));
  CK(cudaMalloc(&d_out, sizeof(float) * 128 * 256));
  const int smem = 8192 * 2 + 256 * 64 * 2 + 256 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(pingpong_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<long long> h(grid * 3);
  std::vector<float> hout(128 * 256);

  // ---- correctness of both operand forms (one unit = hi.hi + hi.mid + mid.hi over 32 channels) ----
  for (int ts = 0; ts < 2; ++ts)
    for (int masked = 0; masked < 2; ++masked)
      for (int N : {32, 64, 160}) {
        Cfg c{ts, masked, N, 1, 1, 1, 0, 0};
        CK(cudaMemset(d_out, 0, sizeof(float) * 128 * 256));
        probe_kernel<<<grid, 256, smem>>>(c, d_cyc, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hout.data(), d_out, sizeof(float) * 128 * 256, cudaMemcpyDeviceToHost));
        const uint32_t off[4] = {0xDB6DB6DBu, 0x6DB6DB6Du, 0xB6DB6DB6u, 0xDB6DB6DBu};
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int ch = 0; ch < 32; ++ch)
              ref += (double)a_hi(m, ch) * b_hi(n, ch) + (double)a_hi(m, ch) * b_mid(n, ch) +
                     (double)a_mid(m, ch) * b_hi(n, ch);
            if (masked && ((off[m >> 5] >> (m & 31)) & 1)) ref = 0;
            maxerr = std::max(maxerr, fabs(ref - (double)hout[m * 256 + n]));
          }
        printf("{\"check\": \"%s%s\", \"N\": %d, \"max_abs_err\": %.3g, \"ok\": %s}\n", ts ? "TS" : "SS",
               masked ? "+mask" : "", N, maxerr, maxerr < 1e-3 ? "true" : "false");
      }

  // ---- cost per MMA ----
  const int units = 400;
  for (int N : {32, 64, 96, 128, 160})
    for (int ts = 0; ts < 2; ++ts)
      for (int masked = 0; masked < 2; ++masked)
        for (int variant = 0; variant < 4; ++variant) {
          // variant 0: 1 issuer 1 chain; 1: 1 issuer 2 chains; 2: 2 issuers; 3: 1 issuer, 1 chain, commit per unit
          Cfg c{ts, masked, N, variant == 1 ? 2 : 1, variant == 2 ? 2 : 1, units, variant == 3, 0};
          if (variant == 1 && 224 + N > 448) continue;
          if (variant == 2 && 224 + N > 448) continue;
          probe_kernel<<<grid, 256, smem>>>(c, d_cyc, nullptr);  // warm
          probe_kernel<<<grid, 256, smem>>>(c, d_cyc, nullptr);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * grid * 3, cudaMemcpyDeviceToHost));
          std::vector<long long> tot(h.begin(), h.begin() + grid), iss(h.begin() + 2 * grid, h.end());
          const double mm = (double)units * 6 * (variant == 2 ? 2 : 1);
          printf(
              "{\"N\": %d, \"form\": \"%s\", \"masked\": %d, \"variant\": \"%s\", \"cyc_per_mma\": %.1f, "
              "\"issue_cyc_per_mma_per_issuer\": %.1f, \"tensor_floor_cyc\": %.0f}\n",
              N, ts ? "TS" : "SS", masked,
              variant == 0 ? "1 issuer" : variant == 1 ? "1 issuer, 2 chains" : variant == 2 ? "2 issuers" : "commit per unit",
              median(tot) / mm, median(iss) / (units * 6.0), N / 2.0);
        }
  // ---- TMEM store cost (what a builder pays to park one unit's A rows: 32 channels hi + mid = 2 x16 stores) ----
  {
    Cfg c{1, 0, 64, 1, 1, units, 0, 1};
    probe_kernel<<<grid, 256, smem>>>(c, d_cyc, nullptr);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * grid * 3, cudaMemcpyDeviceToHost));
    std::vector<long long> st(h.begin() + grid, h.begin() + 2 * grid);
    printf("{\"tmem_store\": \"2 x tcgen05.st.32x32b.x16 + wait::st, 4 warps\", \"cyc_per_unit\": %.1f}\n",
           median(st) / units);
  }
  // ---- hand-off round trips ----
  for (int mode = 0; mode < 6; ++mode) {
    const int rounds = 2000;
    pingpong_kernel<<<grid, 256, smem>>>(mode, rounds, d_cyc);
    pingpong_kernel<<<grid, 256, smem>>>(mode, rounds, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    std::vector<long long> rt(h.begin(), h.begin() + grid);
    printf("{\"pingpong\": \"%s\", \"cyc_per_round_trip\": %.1f}\n",
           mode == 0 ? "arrive <-> arrive" : mode == 1 ? "arrive <-> tcgen05.commit (idle pipe)" : mode == 2 ? "arrive <-> one N=32 MMA + commit" : mode == 3 ? "4 warps {fences, arrive} <-> commit" : mode == 4 ? "4 warps {fences, wait::st, arrive} <-> commit" : "4 warps {tcgen05.st.x32, wait::st, fences, arrive} <-> commit",
           median(rt) / rounds);
  }
  return 0;
}
