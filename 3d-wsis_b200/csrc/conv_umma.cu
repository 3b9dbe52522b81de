// Output-stationary sparse convolution on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces the reference's per-offset gather -> cuBLAS mm -> scatter-add loop
// (include/spconv/spconv_ops.h:296-344; kernels include/spconv/reordering.cu.h:21-157).
//
//   dst[r,:] = residual[r,:] + sum_k  prologue(src[nbr[r,k],:]) . W[k]
//
// Work is organised in TILES of 128 output rows taken in a spatially sorted (Morton) order, built once per
// coordinate set by tilemap.cu: `order[t*128+i]` is the output row of tile slot i and `pmap[t*128+i, k]` the
// source row feeding it through kernel offset k (-1 = none).  A tile is a compact surface patch, so the rows
// it gathers are shared by ~9 offsets and by the neighbouring tiles: the gather is served by L1/L2 and HBM
// sees each feature row about once.
//
// One persistent CTA per SM; a CTA owns tiles (= the 128 TMEM lanes of one accumulator):
//   warps 0-3   epilogue : tcgen05.ld accumulator -> registers -> (+residual) -> global, once per tile
//   warps 4-11  loaders  : find the kernel offsets that have any neighbour in the tile (the others cost
//                          nothing), gather the fp32 source rows with 16-byte loads kept DEPTH units in
//                          flight, apply the fused eval-BatchNorm+ReLU prologue, split fp32 -> bf16 hi (+ bf16
//                          mid for the 1e-4 path) and store them in the 64B-swizzled K-major layout the UMMA
//                          descriptors expect; one elected loader also issues the cp.async.bulk (1-D TMA) of
//                          the unit's pre-swizzled weight block
//   warp 12     MMA      : one elected thread issues tcgen05.mma (M=128, N=Cout, K=16) into TMEM
//   warp 13     map      : cp.async.bulk of the next tile's neighbour-map slice (double buffered)
// The pipeline unit is (active kernel offset k, 32-channel block kb); units flow through an NSTAGE mbarrier
// ring (full: loaders + bulk-copy tx bytes -> MMA; empty: tcgen05.commit -> loaders).  The accumulator is
// double buffered in TMEM so the epilogue of tile t overlaps the MMAs of tile t+1.
// There is no scatter and there are no atomics: every output row is written exactly once.
//
// Precision: precision==1 uses bf16 operands (fp32 accumulate).  precision==3 splits both operands into
// bf16 hi + bf16 mid and issues hi.hi + hi.mid + mid.hi (error ~2^-17 per product, fp32 accumulate in TMEM)
// to honour the reference's fp32 contract (1e-4) while staying on the tensor pipe.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace wsis {
namespace umma {

constexpr int kTileM = 128;
constexpr int kKB = 32;                  // channels per pipeline unit (64 bytes of bf16 = one swizzle-64 row)
constexpr int kABlockBytes = kTileM * 64;  // 8 KB
constexpr int kEpiWarps = 4, kLoadWarps = 8;
constexpr int kMmaWarps = 2;
constexpr int kThreads = (kEpiWarps + kLoadWarps + kMmaWarps + 1) * 32;  // 480
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
// try_wait suspends the thread in hardware until the phase completes or the time hint expires, so waiting
// warps do not burn issue slots the loader warps need
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
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

// D[lane] += A[lane] . B for every accumulator lane whose bit in `off` (4 x 32 lanes) is CLEAR; lanes with the bit
// set are not written at all (the PTX disable-output-lane vector), so their A rows may hold stale data
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint4 off) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, 1, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %5, %6, %7}, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(off.x), "r"(off.y), "r"(off.z), "r"(off.w)
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

__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr),
      "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of (row, 16-byte chunk c16 in [0,4)) inside a [rows x 64 B] K-major SWIZZLE_64B block
__host__ __device__ __forceinline__ uint32_t sw64(uint32_t row, uint32_t c16) {
  return (row >> 3) * 512u + (row & 7u) * 64u + ((c16 ^ ((row & 7u) >> 1)) << 4);
}

struct Params {
  const float *src;
  const uint8_t *recs;       // tile records at stride rec_stride_bytes(K) (tilemap.cu)
  const int32_t *rec_bytes;  // [num_tiles] bytes of each record actually used
  const int32_t *order;      // [num_tiles*128] destination row of each tile slot (-1 = padding)
  const uint8_t *packed;
  const float *in_scale, *in_shift, *residual;
  float *dst;
  int K, Cin, Cout, in_relu, nstage, nrec, nlw, nacc, nmma, tmem_cols;
  int64_t num_tiles;
};

constexpr uint32_t kHdrLast = 1u;

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
  const uint32_t rec_stride = (uint32_t)rec_stride_bytes(p.K);
  const uint32_t hdr_bytes = (uint32_t)rec_hdr_bytes(p.K);
  uint8_t *tail = sm + (size_t)p.nstage * stage_bytes;
  uint8_t *s_rec = tail;  // [nrec][rec_stride], filled by bulk copies
  float *s_scale = reinterpret_cast<float *>(tail + (size_t)p.nrec * rec_stride);
  float *s_shift = s_scale + p.Cin;
  uint4 *s_smask = reinterpret_cast<uint4 *>(s_shift + p.Cin);   // [16] valid-slot mask of the unit in each stage
  uint32_t *s_hdr = reinterpret_cast<uint32_t *>(s_smask + 16);  // [16]
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_hdr + 16);
  // bars: full[nstage], empty[nstage], acc_full[2], acc_empty[2], rec_full[nrec], rec_empty[nrec]
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (p.nstage + s); };
  auto accf_bar = [&](uint32_t s) { return bar0 + 8u * (2 * p.nstage + s); };
  auto acce_bar = [&](uint32_t s) { return bar0 + 8u * (2 * p.nstage + 2 + s); };
  auto recf_bar = [&](uint32_t s) { return bar0 + 8u * (2 * p.nstage + 4 + s); };
  auto rece_bar = [&](uint32_t s) { return bar0 + 8u * (2 * p.nstage + 4 + p.nrec + s); };
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * p.nstage + 4 + 2 * p.nrec);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int c = tid; c < p.Cin; c += kThreads) {
    s_scale[c] = p.in_scale ? p.in_scale[c] : 1.f;
    s_shift[c] = p.in_shift ? p.in_shift[c] : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < p.nstage; ++s) {
      mbar_init(full_bar(s), 2);  // the owning loader warp: expect_tx arrive (weights) + arrive (A rows written)
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accf_bar(s), p.nmma);
      mbar_init(acce_bar(s), kEpiWarps * 32);
    }
    for (int s = 0; s < p.nrec; ++s) {
      mbar_init(recf_bar(s), 1);
      mbar_init(rece_bar(s), kLoadWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps + kLoadWarps) {  // the first MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (warp < kEpiWarps) {
    // Accumulators start at zero and every MMA accumulates: an MMA only touches the lanes (tile slots) that have a
    // neighbour through its offset, so no unit can be the one that "initialises" a tile.
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 2 * p.nacc * p.Cout; c0 += 16) tmem_zero16(taddr + c0);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kEpiWarps) {
    // ===================== epilogue: TMEM -> registers -> (+residual) -> global, once per tile =====================
    uint32_t as = 0, aph = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int32_t rid = __ldg(p.order + tile * kTileM + warp * 32 + lane);
      mbar_wait(accf_bar(as), aph);
      tc_fence_after();
      const int64_t row = rid;
      const bool live = rid >= 0;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * (uint32_t)(p.nacc * p.Cout);
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        tmem_zero16(taddr + c0);  // hand the accumulator back cleared
        for (int a = 1; a < p.nacc; ++a) {  // partial sums of the independent MMA chains
          float t[16];
          tmem_ld16(taddr + a * p.Cout + c0, t);
          tmem_zero16(taddr + a * p.Cout + c0);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += t[i];
        }
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
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(acce_bar(as));
      if (++as == 2) {
        as = 0;
        aph ^= 1;
      }
    }
  } else if (warp < kEpiWarps + kLoadWarps) {
    // ===================== loaders: one WARP per pipeline unit =====================
    // The CTA runs nmma independent PIPELINES (one MMA issuer, its own stage ring and its own loader warps each).
    // Unit g of this CTA (counted across its tiles) belongs to pipeline g % nmma; inside a pipeline the j-th unit
    // lives in ring stage j % nsp and is gathered by the pipeline's loader warp j % nlp, so several units are
    // gathered concurrently and the per-unit protocol cost (barrier wait, fence, arrive) is paid by one warp, not
    // by all.  nlp <= nsp: a warp moves from unit j to j + nlp, and a parity wait on the stage's empty barrier is
    // only unambiguous while that is at most one ring generation ahead of the (in-order) commits of the issuer.
    // 8 lanes cover one source row slice (32 channels = 128 B), 4 rows per load instruction, up to 8 load
    // instructions in flight per lane.
    const int lw = warp - kEpiWarps;
    const int chunk = lane & 7;  // 4 fp32 channels = 8 bytes of bf16
    const int sub4 = lane >> 3;  // entry within a group of 4
    const uint32_t c16 = chunk >> 1, sub = (chunk & 1) * 8;
    uint32_t it = 0;
    uint32_t g0 = 0;  // units of this CTA before the current tile
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t rb = it % p.nrec;
      mbar_wait(recf_bar(rb), (it / p.nrec) & 1);
      const uint8_t *rec = s_rec + (size_t)rb * rec_stride;
      const uint16_t *start = reinterpret_cast<const uint16_t *>(rec + 16 * p.K);
      const int P = start[p.K];
      const int32_t *ridx = reinterpret_cast<const int32_t *>(rec + hdr_bytes);
      const uint8_t *rslot = rec + hdr_bytes + 4 * P;
      // active offsets of the tile (those with at least one entry); a tile without any entry still sends one
      // all-lanes-off unit through the pipeline so that its rows are written (zeros + residual)
      uint32_t mask = __ballot_sync(0xffffffffu, lane < p.K && start[lane + 1] > start[lane]);
      if (mask == 0) mask = 1u;
      const uint32_t nreal = (uint32_t)__popc(mask) * KB;
      // every MMA issuer gets the same number of units per tile: pad with empty units (no lanes, no weights)
      const uint32_t nunits = (nreal + p.nmma - 1) / p.nmma * p.nmma;
      // my units: pipeline pi = lw % nmma, loader slot pc = lw / nmma (slots >= nlp only keep the record barriers
      // moving); the tile's t-th unit of the pipeline is its j = j0 + t -th overall
      const uint32_t P_ = (uint32_t)p.nmma, nlp = (uint32_t)p.nlw, nsp = (uint32_t)p.nstage / P_;
      const uint32_t pi = (uint32_t)lw % P_, pc = (uint32_t)lw / P_;
      const uint32_t j0 = g0 / P_, per_tile = nunits / P_;
      for (uint32_t t = pc < nlp ? (pc + nlp - j0 % nlp) % nlp : per_tile; t < per_tile; t += nlp) {
        const uint32_t u = pi + P_ * t, j = j0 + t;
        const uint32_t stage = pi * nsp + j % nsp, phase = (j / nsp) & 1;
        uint8_t *abase = sm + (size_t)stage * stage_bytes;
        const uint32_t hdr = (u + p.nmma >= nunits) ? kHdrLast : 0u;  // the last unit of each issuer
        if (u >= nreal) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (lane == 0) {
            s_hdr[stage] = hdr;
            s_smask[stage] = make_uint4(0u, 0u, 0u, 0u);
            mbar_arrive(full_bar(stage));
            mbar_arrive(full_bar(stage));
          }
          continue;
        }
        const int ak = (int)(u / KB), kb = (int)(u - ak * KB);
        const int k = __fns(mask, 0, ak + 1);  // ak-th active offset
        const int s0 = start[k], n = start[k + 1] - s0;
        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + kb * kKB + chunk * 4);
        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + kb * kKB + chunk * 4);
        const float *srcc = p.src + kb * kKB + chunk * 4;
        bool acquired = false;
        auto acquire = [&]() {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (lane == 0) {
            s_hdr[stage] = hdr;
            s_smask[stage] = *reinterpret_cast<const uint4 *>(rec + 16 * k);
            mbar_expect_tx(full_bar(stage), b_bytes);
            bulk_g2s(smem_u32(abase + a_bytes), p.packed + (size_t)(k * KB + kb) * b_bytes, b_bytes, full_bar(stage));
          }
        };
        for (int e0 = 0; e0 < n; e0 += 32) {
          const int nq = min(8, (n - e0 + 3) >> 2);  // load instructions of this batch (warp-uniform)
          float4 v[8];
          uint32_t slots[2] = {0, 0};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q >= nq) break;
            const int e = e0 + q * 4 + sub4;
            if (e < n) {
              const int32_t idx = ridx[s0 + e];
              slots[q >> 2] |= (uint32_t)rslot[s0 + e] << (8 * (q & 3));
              v[q] = __ldg(reinterpret_cast<const float4 *>(srcc + (int64_t)idx * p.Cin));
            }
          }
          if (!acquired) {  // the first batch of loads is in flight while the stage drains
            acquired = true;
            acquire();
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q >= nq) break;
            const int e = e0 + q * 4 + sub4;
            if (e < n) {
              const uint32_t slot = (slots[q >> 2] >> (8 * (q & 3))) & 0xffu;
              const uint32_t off = sw64(slot, c16) + sub;
              const float4 x = v[q];
              float4 y = make_float4(fmaf(x.x, sc.x, sh.x), fmaf(x.y, sc.y, sh.y), fmaf(x.z, sc.z, sh.z),
                                     fmaf(x.w, sc.w, sh.w));
              if (p.in_relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
              __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
              uint2 hv = make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
              *reinterpret_cast<uint2 *>(abase + off) = hv;
              if (NS == 2) {
                // bf16 -> fp32 is a 16-bit shift: the residual y - hi is exact in fp32
                const float fx = __uint_as_float(hv.x << 16), fy = __uint_as_float(hv.x & 0xffff0000u);
                const float fz = __uint_as_float(hv.y << 16), fw = __uint_as_float(hv.y & 0xffff0000u);
                __nv_bfloat162 m0 = __floats2bfloat162_rn(y.x - fx, y.y - fy);
                __nv_bfloat162 m1 = __floats2bfloat162_rn(y.z - fz, y.w - fw);
                uint2 mv = make_uint2(*reinterpret_cast<uint32_t *>(&m0), *reinterpret_cast<uint32_t *>(&m1));
                *reinterpret_cast<uint2 *>(abase + kABlockBytes + off) = mv;
              }
            }
          }
        }
        if (!acquired) acquire();  // a unit without entries (the tile has none at all)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(stage));
      }
      g0 += nunits;
      __syncwarp();
      if (lane == 0) mbar_arrive(rece_bar(rb));  // this warp no longer reads s_rec[rb]
    }
  } else if (warp < kEpiWarps + kLoadWarps + kMmaWarps) {
    // ===================== MMA issuers =====================
    // Issuer i (one elected thread of warp 12+i) takes the units g = i (mod nmma) and accumulates into its own TMEM
    // accumulators: M=128 x N=Cout x K=16 MMAs are short (N/2 cycles), so a single dependent accumulate chain issued
    // by a single thread is bound by the MMA pipeline latency and the per-unit barrier handling, not by the tensor
    // pipe.  Independent chains (nacc accumulators, summed by the epilogue) and two issuers hide both.
    const int mi = warp - (kEpiWarps + kLoadWarps);
    if (lane == 0 && mi < p.nmma) {
      const uint32_t idesc = make_idesc(p.Cout);
      const uint32_t nsp = (uint32_t)(p.nstage / p.nmma);
      const uint32_t sm_base = smem_u32(sm);
      const uint64_t desc0 = make_desc(0);
      uint32_t j = 0;  // units this issuer has consumed
      uint32_t as = 0, aph = 0;
      const uint32_t per = (uint32_t)(p.nacc / p.nmma);  // accumulators of this issuer: mi, mi + nmma, ...
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(acce_bar(as), aph ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + as * (uint32_t)(p.nacc * p.Cout);
        const uint32_t d0 = d_base + mi * p.Cout, d1 = per > 1 ? d0 + p.nmma * p.Cout : d0;
        while (true) {
          const uint32_t stage = mi * nsp + j % nsp, phase = (j / nsp) & 1;
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t hdr = *reinterpret_cast<volatile uint32_t *>(s_hdr + stage);
          uint4 vm;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(vm.x), "=r"(vm.y), "=r"(vm.z), "=r"(vm.w)
                       : "r"(smem_u32(s_smask + stage))
                       : "memory");
          const uint4 off = make_uint4(~vm.x, ~vm.y, ~vm.z, ~vm.w);
          const uint32_t a0 = sm_base + stage * stage_bytes;
          const uint64_t ad = desc0 + (a0 >> 4), bd = desc0 + ((a0 + a_bytes) >> 4);
          if ((vm.x | vm.y | vm.z | vm.w) != 0) {
            // accumulators of this issuer alternate between consecutive MMAs (independent chains)
            mma_bf16(d0, ad, bd, idesc, off);
            if (NS == 2) {
              mma_bf16(d1, ad, bd + (b_block >> 4), idesc, off);
              mma_bf16(d0, ad + (kABlockBytes >> 4), bd, idesc, off);
            }
            mma_bf16(d1, ad + 2, bd + 2, idesc, off);
            if (NS == 2) {
              mma_bf16(d0, ad + 2, bd + (b_block >> 4) + 2, idesc, off);
              mma_bf16(d1, ad + (kABlockBytes >> 4) + 2, bd + 2, idesc, off);
            }
          }
          mma_commit(empty_bar(stage));
          ++j;
          if (hdr & kHdrLast) break;
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
    // ===================== record producer: 1-D bulk copy of each tile's record, nrec-1 tiles ahead ===============
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t rb = it % p.nrec;
        const uint32_t bytes = (uint32_t)__ldg(p.rec_bytes + tile);
        mbar_wait(rece_bar(rb), ((it / p.nrec) & 1) ^ 1);
        mbar_expect_tx(recf_bar(rb), bytes);
        bulk_g2s(smem_u32(s_rec + (size_t)rb * rec_stride), p.recs + tile * (int64_t)rec_stride, bytes, recf_bar(rb));
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

int wsis_conv_umma(const float *src, const void *records, const int32_t *rec_bytes, const int32_t *order,
                   int64_t num_tiles, int K, const void *packed, int Cin, int Cout, int precision,
                   const float *in_scale, const float *in_shift, int in_relu, const float *residual, float *dst,
                   wsis_stream_t stream) {
  WSIS_CHECK(wsis_conv_umma_supported(Cin, Cout), "conv_umma: unsupported Cin=%d Cout=%d", Cin, Cout);
  WSIS_CHECK(K >= 1 && K <= 32, "conv_umma: kernel volume %d not in [1,32]", K);
  WSIS_CHECK(precision == 1 || precision == 3, "conv_umma: precision must be 1 or 3");
  WSIS_CHECK((in_scale == nullptr) == (in_shift == nullptr), "conv_umma: in_scale/in_shift must both be set");
  WSIS_CHECK(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) |
               reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(packed) |
               reinterpret_cast<uintptr_t>(records)) & 15) == 0,
             "conv_umma: src/dst/residual/packed/records must be 16-byte aligned");
  if (num_tiles == 0) return 0;
  const int NS = precision == 3 ? 2 : 1;
  Params p;
  p.src = src;
  p.recs = (const uint8_t *)records;
  p.rec_bytes = rec_bytes;
  p.order = order;
  p.packed = (const uint8_t *)packed;
  p.in_scale = in_scale;
  p.in_shift = in_shift;
  p.residual = residual;
  p.dst = dst;
  p.K = K;
  p.Cin = Cin;
  p.Cout = Cout;
  p.in_relu = in_relu;
  p.num_tiles = num_tiles;
  // independent accumulate chains per accumulator buffer: as many as the 512 TMEM columns allow, up to 4
  int nacc = 4;
  while (nacc > 1 && 2 * nacc * Cout > 512) nacc >>= 1;
  p.nacc = nacc;
  p.nmma = nacc >= 2 ? kMmaWarps : 1;
  int cols = 32;
  while (cols < 2 * nacc * Cout) cols <<= 1;
  p.tmem_cols = cols;
  const int64_t stage_bytes = (int64_t)NS * kABlockBytes + (int64_t)NS * Cout * 64;
  const int64_t rec_stride = rec_stride_bytes(K);
  const int64_t misc = 1024 /*align*/ + 2 * Cin * 4 + 16 * 16 + 16 * 4 + (2 * 16 + 4 + 2 * 3) * 8 + 64;
  const int64_t budget = 226 * 1024;
  // three record buffers keep the record copies two tiles ahead; fall back to two when the stages need the room
  int nrec = 3;
  int nstage = (int)std::min<int64_t>(16, (budget - misc - nrec * rec_stride) / stage_bytes);
  if (nstage < kLoadWarps) {
    nrec = 2;
    nstage = (int)std::min<int64_t>(16, (budget - misc - nrec * rec_stride) / stage_bytes);
  }
  WSIS_CHECK(nstage >= 2, "conv_umma: shared memory budget exceeded for Cin=%d Cout=%d K=%d", Cin, Cout, K);
  p.nstage = nstage / p.nmma * p.nmma;
  p.nrec = nrec;
  p.nlw = std::min(kLoadWarps / p.nmma, p.nstage / p.nmma);  // loader warps per pipeline
  const int64_t smem = misc + nrec * rec_stride + nstage * stage_bytes;
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
