// Output-stationary sparse convolution on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces the reference's per-offset gather -> cuBLAS mm -> scatter-add loop
// (include/spconv/spconv_ops.h:296-344; kernels include/spconv/reordering.cu.h:21-157).
//
//   dst[r,:] = residual[r,:] + sum_k  prologue(src[nbr[r,k],:]) . W[k]
//
// Work is organised in TILES of 128 output rows taken in a spatially sorted (Morton) order, built once per
// coordinate set by tilemap.cu: `order[t*128+i]` is the output row of tile slot i; the tile's RECORD lists, per
// kernel offset k, the entries (tile slot, index into the tile's list of DISTINCT source rows).  A tile is a
// compact surface patch: ~160-230 distinct rows feed its ~480-1400 entries, and neighbouring tiles share their
// halos through L2, so HBM sees each feature row about once (ncu: DRAM bytes <= algorithmic bytes).
//
// One persistent CTA per SM; a CTA owns tiles (= the 128 TMEM lanes of one accumulator).  640 threads:
//   warps 0-3    epilogue  : tcgen05.ld accumulator -> registers -> (+residual) -> global, once per tile
//   warps 4-7    gatherers : fetch every distinct source row of the tile ONCE per 32-channel block (16-byte loads,
//                            8 in flight per lane), fused eval-BatchNorm+ReLU prologue, fp32 -> bf16 hi (+ bf16
//                            mid for the 1e-4 path), park the rows in a shared-memory ROW CACHE (nrc buffers)
//   warps 8-15   builders  : two warps per pipeline unit (active offset k, 32-channel block kb): copy the unit's
//                            entries row cache -> 64B-swizzled K-major A block (shared -> shared); rows are
//                            prefetched into registers before the stage-free wait
//   warps 16-17  issuers   : warp-uniform loop, one elect.sync lane issues tcgen05.mma (M=128, N=Cout, K=16) with
//                            the unit's disable-output-lane mask, then one tcgen05.commit per unit
//   warp 18      records   : cp.async.bulk of the next tile's record (+ its first kRcap unique rows)
//   warp 19      weights   : cp.async.bulk of each unit's pre-swizzled weight block into the unit's stage
// Units flow through a ring of 2^lna stages {A block(s), weight block(s)} with one barrier pair per stage
// (full: 2 builder arrives + the weight copy's expect_tx and bytes; empty: tcgen05.commit).  The accumulator is
// double buffered in TMEM so the epilogue of tile t overlaps the MMAs of tile t+1; each issuer has its own
// accumulator (summed by the epilogue).  There is no scatter and there are no atomics: every output row is
// written exactly once.  DESIGN.md 4 has the measured cost model and what bounds the kernel today.
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
constexpr int kKB = 32;                    // channels per pipeline unit (64 bytes of bf16 = one swizzle-64 row)
constexpr int kABlockBytes = kTileM * 64;  // 8 KB
constexpr int kEpiWarps = 4, kGatherWarps = 4, kBuildWarps = 8, kMmaWarps = 2, kProdWarps = 2;
constexpr int kThreads = (kEpiWarps + kGatherWarps + kBuildWarps + kMmaWarps + kProdWarps) * 32;  // 640
constexpr int kMaxRec = 4;  // tile records in flight (p.nrec <= kMaxRec)
constexpr int kRcap = 256;  // rows of one row-cache buffer (a surface tile reads ~160-230 distinct rows)

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
// Diagnostics (timing experiments through WSIS_CONV_DEBUG and the CTA-0 timeline) are compiled in only with
// -DWSIS_CONV_DIAG=1: in the product build they cost nothing on the single-warp critical paths.
#ifndef WSIS_CONV_DIAG
#define WSIS_CONV_DIAG 0
#endif
constexpr bool kDiag = WSIS_CONV_DIAG != 0;
#define WSIS_DBG(bit) (kDiag && (p.dbg & (bit)))

// try_wait blocks in hardware for a short, implementation-defined time; NO suspend-time hint: with a hint ptxas emits
// NANOSLEEP.SYNCS after a failed probe and the wake-up latency of a sleeping warp (on both sides of every stage hand-off)
// dominated the pipeline's round trip.  A wait that has not completed after ~4 s is a protocol bug: report which
// barrier and trap instead of hanging the device.
__device__ __noinline__ void mbar_deadlock(uint32_t bar, uint32_t parity) {
  printf("wsis conv_umma: barrier wait timed out: block %d warp %d lane %d barrier@%u parity %u\n", (int)blockIdx.x,
         (int)(threadIdx.x >> 5), (int)(threadIdx.x & 31), bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, fails = 0;
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
    // every failed try_wait has already blocked for a hardware time quantum: 2^22 of them in a row is seconds
    if (!ok && ++fails == (1u << 22)) mbar_deadlock(bar, parity);
  } while (!ok);
}

// One lane of a fully active warp.  Code under this predicate is known to be single-lane, and values made uniform
// with uni() live in uniform registers: ptxas then feeds tcgen05.mma / tcgen05.commit / cp.async.bulk their
// uniform-register operands directly instead of wrapping every instruction in an ELECT + R2UR.BROADCAST waterfall
// loop (which costs ~150 cycles per MMA and made the tensor pipe issue-bound at 13 % busy).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t uni(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
// the same for a lane that runs alone in divergent code (membermask = that lane)
__device__ __forceinline__ bool elect_lane(uint32_t lane_mask) {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, %1;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred)
      : "r"(lane_mask));
  return pred != 0;
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
  const uint8_t *recs;    // entry records at stride rec_stride_bytes(K) (tilemap.cu)
  const int32_t *uidx;    // [num_tiles][128 K] distinct source rows of each tile
  const int4 *meta;       // [num_tiles] {record bytes, nU, active-offset mask, P}
  const int32_t *order;   // [num_tiles*128] destination row of each tile slot (-1 = padding)
  const uint8_t *packed;
  const float *in_scale, *in_shift, *residual;
  float *dst;
  int K, Cin, Cout, KB, in_relu, vec4;
  unsigned long long *tl;  // optional timeline buffer (wsis_conv_debug_timeline): [0] = event count, then (time, code)
  int tl_cap;
  int dbg;  // WSIS_CONV_DEBUG bits (timing experiments only; results are wrong): 1 no MMA, 2 no build, 4 no gather, 8 no weight copy
  int lna, nrc, nrec, nb, nacc, nmma, tmem_cols;  // 1 << lna pipeline stages, nrc row-cache buffers, nrec record
                                                  // buffers, nb stage-owning builder pairs, TMEM accumulators, issuers
  int rec_main;  // shared-memory bytes reserved for one entry record (>= the largest meta[t].x of this tile map)
  int64_t num_tiles;
};

// timeline event of CTA 0 (diagnostics only): every role appends to its own 512-entry region (no atomics, stores
// are fire-and-forget); code = role << 24 | tile iteration << 16 | event << 12 | unit
__device__ __forceinline__ void tl_event(const Params &p, uint32_t role, uint32_t it, uint32_t ev, uint32_t unit) {
  if (kDiag && p.tl != nullptr && blockIdx.x == 0 && it < 4) {
    unsigned long long now;
    now = (unsigned long long)clock64();  // SM-local cycle counter: all roles of the CTA share it
    // slot inside the role's region derived from (it, ev, unit): at most 4 tiles x 4 events x 32 units
    const uint32_t slot = (it * 4 + ev) * 32 + (unit & 31u);
    p.tl[2 * (role * 512 + slot)] = now;
    p.tl[2 * (role * 512 + slot) + 1] = (role << 24) | (it << 16) | (ev << 12) | (unit & 0xfffu);
  }
}

// relu?(x * sc + sh) -> bf16 hi (part 0) or bf16 mid = bf16(y - hi) (part 1), two values per 32-bit word
__device__ __forceinline__ uint32_t split2(float a, float b, int part) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  uint32_t hv = *reinterpret_cast<uint32_t *>(&h);
  if (part == 0) return hv;
  // bf16 -> fp32 is a 16-bit shift: the residual y - hi is exact in fp32
  __nv_bfloat162 m = __floats2bfloat162_rn(a - __uint_as_float(hv << 16), b - __uint_as_float(hv & 0xffff0000u));
  return *reinterpret_cast<uint32_t *>(&m);
}

__device__ __forceinline__ float4 load_row4(const Params &p, int32_t row, int c0) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const float *s = p.src + (int64_t)row * p.Cin + c0;
  if (p.vec4) {
    if (c0 < p.Cin) v = __ldg(reinterpret_cast<const float4 *>(s));
  } else {  // narrow or unaligned rows (the 6-channel input layer): guarded scalar loads, zero padding
    if (c0 < p.Cin) v.x = __ldg(s);
    if (c0 + 1 < p.Cin) v.y = __ldg(s + 1);
    if (c0 + 2 < p.Cin) v.z = __ldg(s + 2);
    if (c0 + 3 < p.Cin) v.w = __ldg(s + 3);
  }
  return v;
}

__device__ __forceinline__ float4 prologue4(float4 x, float4 sc, float4 sh, int relu) {
  float4 y = make_float4(fmaf(x.x, sc.x, sh.x), fmaf(x.y, sc.y, sh.y), fmaf(x.z, sc.z, sh.z), fmaf(x.w, sc.w, sh.w));
  if (relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
  return y;
}

// A tile that reads more distinct rows than a row-cache buffer holds fetches the overflow rows directly: 8 channels
// [c0, c0 + 8) of unique row `loc` of the tile -> prologue -> 8 bf16 (hi or mid).  Rare; kept out of line.
__device__ __noinline__ uint4 fetch_direct(const Params &p, int64_t tile, uint32_t loc, int c0, int part,
                                           const float *s_scale, const float *s_shift) {
  const int32_t row = __ldg(p.uidx + tile * (int64_t)(kTileM * p.K) + loc);
  const float4 y0 = prologue4(load_row4(p, row, c0), *reinterpret_cast<const float4 *>(s_scale + c0),
                              *reinterpret_cast<const float4 *>(s_shift + c0), p.in_relu);
  const float4 y1 = prologue4(load_row4(p, row, c0 + 4), *reinterpret_cast<const float4 *>(s_scale + c0 + 4),
                              *reinterpret_cast<const float4 *>(s_shift + c0 + 4), p.in_relu);
  return make_uint4(split2(y0.x, y0.y, part), split2(y0.z, y0.w, part), split2(y1.x, y1.y, part),
                    split2(y1.z, y1.w, part));
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int NS>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t *sm = smem_raw + pad;
  const int KB = p.KB;
  const uint32_t na = 1u << p.lna;  // pipeline stages: A operand block(s) + the unit's weight block(s), one barrier pair
  constexpr uint32_t a_stage = NS * kABlockBytes;
  const uint32_t b_block = (uint32_t)p.Cout * 64u;
  const uint32_t w_stage = NS * b_block;
  constexpr uint32_t row_b = NS * 64;  // one row-cache row: 32 channels of bf16 hi (+ 32 of bf16 mid)
  constexpr uint32_t rc_buf = kRcap * row_b;
  const uint32_t rec_stride = (uint32_t)rec_stride_bytes(p.K);
  const uint32_t rec_main = (uint32_t)p.rec_main;
  const uint32_t rec_buf = rec_main + kRcap * 4;  // entry record + the first kRcap unique rows
  const uint32_t hdr_bytes = (uint32_t)rec_hdr_bytes(p.K);
  uint8_t *s_a = sm;                                  // [na][NS][128 x 64 B swizzled]  A operands
  uint8_t *s_w = s_a + (size_t)na * a_stage;          // [na][NS][Cout x 64 B swizzled] weight blocks (bulk copies)
  uint8_t *s_rc = s_w + (size_t)na * w_stage;         // [nrc][kRcap][row_b]            converted source rows
  uint8_t *s_rec = s_rc + (size_t)p.nrc * rc_buf;     // [p.nrec][rec_buf]               tile records (bulk copies)
  float *s_scale = reinterpret_cast<float *>(s_rec + (size_t)p.nrec * rec_buf);
  float *s_shift = s_scale + KB * kKB;
  uint4 *s_smask = reinterpret_cast<uint4 *>(s_shift + KB * kKB);  // [na] valid-slot mask of the unit in each A stage
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_smask + na);
  const uint32_t bar0 = smem_u32(bars);
  auto afull_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto aempty_bar = [&](uint32_t s) { return bar0 + 8u * (na + s); };
  const uint32_t bar1 = bar0 + 8u * (2 * na);
  auto rcf_bar = [&](uint32_t s) { return bar1 + 8u * s; };
  auto rce_bar = [&](uint32_t s) { return bar1 + 8u * (p.nrc + s); };
  const uint32_t bar2 = bar1 + 8u * (2 * p.nrc);
  auto recf_bar = [&](uint32_t s) { return bar2 + 8u * s; };
  auto rece_bar = [&](uint32_t s) { return bar2 + 8u * (p.nrec + s); };
  auto accf_bar = [&](uint32_t s) { return bar2 + 8u * (2 * p.nrec + s); };
  auto acce_bar = [&](uint32_t s) { return bar2 + 8u * (2 * p.nrec + 2 + s); };
  const uint32_t nbars = 2 * na + 2 * p.nrc + 2 * p.nrec + 4;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + nbars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int c = tid; c < KB * kKB; c += kThreads) {  // channels padded up to the unit width contribute exact zeros
    s_scale[c] = c < p.Cin ? (p.in_scale ? p.in_scale[c] : 1.f) : 0.f;
    s_shift[c] = c < p.Cin ? (p.in_shift ? p.in_shift[c] : 0.f) : 0.f;
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < na; ++s) {
      // full: the two builder warps of the unit (each assembles half of its entries) + the weight producer's
      // expect_tx arrive (+ the bytes of its bulk copy); empty: ONE tcgen05.commit per unit releases the operand block
      // and the weight block together (every tcgen05 instruction costs the issuing thread ~75 cycles, so the
      // per-unit protocol is one wait, the MMAs and one commit)
      mbar_init(afull_bar(s), 3);
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < p.nrc; ++s) {
      mbar_init(rcf_bar(s), kGatherWarps);
      mbar_init(rce_bar(s), 2 * p.nb);
    }
    for (int s = 0; s < p.nrec; ++s) {
      mbar_init(recf_bar(s), 1);
      mbar_init(rece_bar(s), kGatherWarps + 2 * p.nb);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accf_bar(s), p.nmma);
      mbar_init(acce_bar(s), kEpiWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int kMmaWarp0 = kEpiWarps + kGatherWarps + kBuildWarps;
  if (warp == kMmaWarp0) {  // the first MMA warp owns the TMEM allocation
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
    // Everything that does not depend on the accumulator is in flight before the wait: the tile's destination rows
    // (fetched one tile ahead) and the residual of the first 16-column chunk; inside the tile the residual of chunk
    // c + 1 is fetched while chunk c is read out of TMEM.
    uint32_t as = 0, aph = 0;
    int32_t rid = __ldg(p.order + (int64_t)blockIdx.x * kTileM + warp * 32 + lane);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t row = rid;
      const bool live = rid >= 0;
      const int64_t nxt = tile + gridDim.x;
      if (nxt < p.num_tiles) rid = __ldg(p.order + nxt * kTileM + warp * 32 + lane);
      const bool has_res = live && p.residual != nullptr;
      const float4 *rs = reinterpret_cast<const float4 *>(p.residual + (has_res ? row * p.Cout : 0));
      float4 r4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) r4[q] = has_res ? __ldg(rs + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait(accf_bar(as), aph);
      tc_fence_after();
      if (tid == 0) tl_event(p, 0, (uint32_t)((tile - blockIdx.x) / gridDim.x), 0, 0);
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * (uint32_t)(p.nacc * p.Cout);
      for (int c0 = 0; c0 < p.Cout && !WSIS_DBG(32); c0 += 16) {
        float4 rn[4];
        const bool more = c0 + 16 < p.Cout;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          rn[q] = (has_res && more) ? __ldg(rs + (c0 + 16) / 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
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
#pragma unroll
          for (int q = 0; q < 4; ++q)
            o[q] = make_float4(v[4 * q] + r4[q].x, v[4 * q + 1] + r4[q].y, v[4 * q + 2] + r4[q].z,
                               v[4 * q + 3] + r4[q].w);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) r4[q] = rn[q];
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(acce_bar(as));
      if (tid == 0) tl_event(p, 0, (uint32_t)((tile - blockIdx.x) / gridDim.x), 1, 0);
      if (++as == 2) {
        as = 0;
        aph ^= 1;
      }
    }
  } else if (warp < kEpiWarps + kGatherWarps) {
    // ===================== gatherers: every distinct source row of a tile is fetched ONCE =====================
    // One PASS = (tile, 32-channel block kb): the 128-byte slices of the tile's unique rows are loaded with
    // 16-byte loads (8 lanes per row, 8 rows per lane in flight), the fused eval-BatchNorm+ReLU prologue is applied,
    // the result is split fp32 -> bf16 hi (+ bf16 mid) and parked in a row-cache buffer.  The builders then
    // assemble the per-offset A operands from shared memory, so HBM/L2 see each row once per tile instead of once
    // per (row, offset) entry, and the loads of up to nrc passes are in flight ahead of the tensor pipe.
    const int gt = (warp - kEpiWarps) * 32 + lane;
    const int rsub = gt >> 3, chunk = gt & 7;
    uint32_t it = 0, q = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t rb = it % p.nrec;
      mbar_wait(recf_bar(rb), (it / p.nrec) & 1);
      if (gt == 0) tl_event(p, 1, it, 0, 0);
      const uint8_t *rec = s_rec + (size_t)rb * rec_buf;
      // distinct rows of the tile (the overflow beyond a row-cache buffer is fetched directly by the builders)
      const int ng = min((int)reinterpret_cast<const uint16_t *>(rec + 16 * p.K)[p.K + 1], kRcap);
      const int32_t *uidx = reinterpret_cast<const int32_t *>(rec + rec_main);
      for (int kb = 0; kb < KB; ++kb, ++q) {
        const uint32_t slot = q % p.nrc;
        const int c0 = kb * kKB + chunk * 4;
        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c0);
        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c0);
        uint8_t *rcb = s_rc + (size_t)slot * rc_buf;
        bool waited = false;
        constexpr int kSweep = kGatherWarps * 4;  // rows per load instruction of the gather warps
        constexpr int kInFlight = 8;              // 16-byte loads in flight per gather lane
        for (int u0 = 0; u0 < ng; u0 += kInFlight * kSweep) {
          int32_t idx[kInFlight];
          float4 v[kInFlight];
#pragma unroll
          for (int i = 0; i < kInFlight; ++i) {
            const int u = u0 + i * kSweep + rsub;
            if (u < ng) idx[i] = uidx[u];
          }
#pragma unroll
          for (int i = 0; i < kInFlight; ++i) {
            const int u = u0 + i * kSweep + rsub;
            if (u < ng && !WSIS_DBG(4)) v[i] = load_row4(p, idx[i], c0);
          }
          if (gt == 0) tl_event(p, 1, it, 1, (uint32_t)kb);
          if (!waited) {  // the loads are in flight while the buffer drains
            mbar_wait(rce_bar(slot), ((q / p.nrc) & 1) ^ 1);
            waited = true;
          }
          if (gt == 0) tl_event(p, 1, it, 2, (uint32_t)kb);
#pragma unroll
          for (int i = 0; i < kInFlight; ++i) {
            const int u = u0 + i * kSweep + rsub;
            if (u < ng) {
              const float4 y = prologue4(v[i], sc, sh, p.in_relu);
              if (NS == 2) {
                // hi half and mid half of a row swap places on odd rows (bank spread); 16-byte chunk c of the
                // logical row [hi 0..3 | mid 4..7] lives at chunk c ^ ((u & 1) << 2)
                const uint32_t x4 = (uint32_t)(u & 1) << 2;
                uint8_t *r = rcb + (size_t)u * row_b + (chunk & 1) * 8;
                *reinterpret_cast<uint2 *>(r + (((chunk >> 1) ^ x4) << 4)) =
                    make_uint2(split2(y.x, y.y, 0), split2(y.z, y.w, 0));
                *reinterpret_cast<uint2 *>(r + ((((chunk >> 1) + 4) ^ x4) << 4)) =
                    make_uint2(split2(y.x, y.y, 1), split2(y.z, y.w, 1));
              } else {
                *reinterpret_cast<uint2 *>(rcb + (size_t)u * row_b + chunk * 8) =
                    make_uint2(split2(y.x, y.y, 0), split2(y.z, y.w, 0));
              }
            }
          }
        }
        if (!waited) mbar_wait(rce_bar(slot), ((q / p.nrc) & 1) ^ 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(rcf_bar(slot));
        if (gt == 0) tl_event(p, 1, it, 3, (uint32_t)kb);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(rece_bar(rb));  // this warp no longer reads s_rec[rb]
    }
  } else if (warp < kMmaWarp0) {
    // ===================== builders: one WARP per pipeline unit, shared memory -> shared memory =====================
    // Unit = (32-channel block kb, active kernel offset k) of a tile; the CTA's j-th unit lives in A stage j % na and
    // is assembled by builder warp j % nb: the entries of offset k copy their converted row from the row cache into
    // the 64B-swizzled K-major block the UMMA descriptor expects, at the row of their tile slot.
    // nb <= na (and both even when two issuers split the units): a builder moves from unit j to j + nb, and a parity
    // wait on a stage's empty barrier is only unambiguous while that is at most one ring generation ahead of the
    // (per-issuer in-order) commits.
    // Two warps share a unit (even / odd groups of its entries): dense neighbourhoods have 50-128 entries per unit and
    // one warp per unit made the builders, not the issuers, the slowest stage on the coarser U-Net levels.
    const uint32_t nb = (uint32_t)p.nb;
    const int bw = warp - (kEpiWarps + kGatherWarps);
    const int b = bw % (int)nb, half = bw / (int)nb;
    constexpr int LPE = 4 * NS;    // lanes per entry (16 bytes each: 4 hi chunks [+ 4 mid chunks])
    constexpr int EPI = 32 / LPE;  // entries per warp instruction
    const int e_in = lane / LPE, l = lane % LPE, part = l >> 2, c16 = l & 3;
    const uint32_t s_a32 = smem_u32(s_a), s_rc32 = smem_u32(s_rc);
    uint32_t it = 0, q = 0, j0 = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles && half < 2; tile += gridDim.x, ++it) {
      const uint32_t rb = it % p.nrec;
      mbar_wait(recf_bar(rb), (it / p.nrec) & 1);
      const uint8_t *rec = s_rec + (size_t)rb * rec_buf;
      const uint16_t *start = reinterpret_cast<const uint16_t *>(rec + 16 * p.K);
      const int P = start[p.K];
      const uint16_t *eloc = reinterpret_cast<const uint16_t *>(rec + hdr_bytes);
      const uint8_t *eslot = rec + hdr_bytes + 2 * P;
      // active offsets of the tile (those with at least one entry); a tile without any entry still sends one
      // all-lanes-off unit through the pipeline so that its rows are written (zeros + residual)
      uint32_t mask = __ballot_sync(0xffffffffu, lane < p.K && start[lane + 1] > start[lane]);
      if (mask == 0) mask = 1u;
      const uint32_t nact = (uint32_t)__popc(mask);
      for (int kb = 0; kb < KB; ++kb, ++q) {
        const uint32_t slot = q % p.nrc;
        mbar_wait(rcf_bar(slot), (q / p.nrc) & 1);
        const uint32_t rcb = s_rc32 + slot * rc_buf;
        const uint32_t jb = j0 + (uint32_t)kb * nact;
        uint32_t ak = ((uint32_t)b + nb - jb % nb) % nb;
        uint32_t mm = mask;  // lowest set bit of mm = the ak-th active offset
        for (uint32_t i = 0; i < ak; ++i) mm &= mm - 1;
        for (; ak < nact; ak += nb) {
          const uint32_t j = jb + ak;
          const uint32_t stage = j & (na - 1), phase = (j >> p.lna) & 1;
          const int k = __ffs(mm) - 1;
          for (uint32_t i = 0; i < nb; ++i) mm &= mm - 1;
          const int s0 = start[k], n = start[k + 1] - s0;
          // Before the stage is free: fetch this lane's share of the unit (entry -> row-cache row -> registers), so
          // that the stage's critical path (commit -> builder -> issuer) is only store + fence + arrive.
          constexpr int PF = 4;
          uint4 v[PF];
          uint32_t off[PF];
          const uint32_t xl = (uint32_t)l;
#pragma unroll
          for (int i = 0; i < PF; ++i) {
            const int e = (2 * i + half) * EPI + e_in;
            if (e < n && !WSIS_DBG(2)) {
              const uint32_t loc = eloc[s0 + e];
              off[i] = sw64(eslot[s0 + e], c16);
              if (loc < kRcap)
                v[i] = lds128(rcb + loc * row_b + ((NS == 2 ? (xl ^ ((loc & 1u) << 2)) : xl) << 4));
              else
                v[i] = fetch_direct(p, tile, loc, kb * kKB + c16 * 8, part, s_scale, s_shift);
            }
          }
          if (lane == 0) tl_event(p, 2 + b, it, 0, j - j0);
          mbar_wait(aempty_bar(stage), phase ^ 1);
          if (lane == 0) tl_event(p, 2 + b, it, 1, j - j0);
          if (lane == 0 && half == 0) s_smask[stage] = *reinterpret_cast<const uint4 *>(rec + 16 * k);
          const uint32_t abase = s_a32 + stage * a_stage + (uint32_t)part * kABlockBytes;
#pragma unroll
          for (int i = 0; i < PF; ++i)
            if ((2 * i + half) * EPI + e_in < n && !WSIS_DBG(2)) sts128(abase + off[i], v[i]);
          // long units (dense neighbourhoods, 50-128 entries): the same batches of PF loads then PF stores
          for (int g0 = 2 * PF; g0 * EPI < n; g0 += 2 * PF) {
#pragma unroll
            for (int i = 0; i < PF; ++i) {
              const int e = (g0 + 2 * i + half) * EPI + e_in;
              if (e < n) {
                const uint32_t loc = eloc[s0 + e];
                off[i] = sw64(eslot[s0 + e], c16);
                if (loc < kRcap)
                  v[i] = lds128(rcb + loc * row_b + ((NS == 2 ? (xl ^ ((loc & 1u) << 2)) : xl) << 4));
                else
                  v[i] = fetch_direct(p, tile, loc, kb * kKB + c16 * 8, part, s_scale, s_shift);
              }
            }
#pragma unroll
            for (int i = 0; i < PF; ++i)
              if ((g0 + 2 * i + half) * EPI + e_in < n) sts128(abase + off[i], v[i]);
          }
          if (!WSIS_DBG(128)) fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(afull_bar(stage));
          if (lane == 0) tl_event(p, 2 + b, it, 2, j - j0);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(rce_bar(slot));  // this warp no longer reads row-cache buffer `slot`
      }
      j0 += nact * (uint32_t)KB;
      __syncwarp();
      if (lane == 0) mbar_arrive(rece_bar(rb));  // this warp no longer reads s_rec[rb]
    }
  } else if (warp < kMmaWarp0 + kMmaWarps) {
    // ===================== MMA issuers =====================
    // Issuer i (one elected thread of warp kMmaWarp0 + i) takes the units j = i (mod nmma) and accumulates into its
    // own TMEM accumulators: M=128 x N=Cout x K=16 MMAs are short (N/2 cycles), so a single dependent accumulate
    // chain issued by a single thread is bound by the MMA pipeline latency and the per-unit barrier handling, not by
    // the tensor pipe.  Independent chains (nacc accumulators, summed by the epilogue) and two issuers hide both.
    // The whole warp runs the (warp-uniform) loop; one elected lane issues.
    const uint32_t mi = uni((uint32_t)(warp - kMmaWarp0));
    if (mi < (uint32_t)p.nmma) {
      const uint32_t nmma = (uint32_t)p.nmma;
      const uint32_t tbase = uni(tmem_base);
      const uint32_t idesc = make_idesc(p.Cout);
      const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w), m_base = smem_u32(s_smask);
      const uint64_t desc0 = make_desc(0);
      uint32_t j0 = 0;
      uint32_t as = 0, aph = 0;
      const uint32_t per = (uint32_t)p.nacc / nmma;  // accumulators of this issuer: mi, mi + nmma, ...
      uint32_t itx = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++itx) {
        const uint32_t n = uni((uint32_t)__popc((uint32_t)__ldg(&p.meta[tile].z))) * (uint32_t)KB;
        mbar_wait(acce_bar(as), aph ^ 1);
        tc_fence_after();
        const uint32_t d_base = tbase + as * (uint32_t)(p.nacc * p.Cout);
        const uint32_t d0 = d_base + mi * p.Cout, d1 = per > 1 ? d0 + nmma * p.Cout : d0;
        for (uint32_t u = (mi - j0) & (nmma - 1); u < n; u += nmma) {
          const uint32_t j = j0 + u;
          const uint32_t sa = j & (na - 1);
          mbar_wait(afull_bar(sa), (j >> p.lna) & 1);  // operand rows (generic proxy, fenced by the builders) + weights
          if (lane == 0) tl_event(p, 16 + mi, itx, 0, u);
          const uint64_t ad = desc0 + ((a_base + sa * a_stage) >> 4), bd = desc0 + ((w_base + sa * w_stage) >> 4);
          if (elect_one()) {
            const uint4 vm = lds128(m_base + sa * 16);  // valid-slot mask of the unit; the MMA takes its complement
            const uint4 off = make_uint4(~vm.x, ~vm.y, ~vm.z, ~vm.w);
            if ((vm.x | vm.y | vm.z | vm.w) != 0 && !WSIS_DBG(1)) {
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
            mma_commit(aempty_bar(sa));
          }
          __syncwarp();
          if (lane == 0) tl_event(p, 16 + mi, itx, 2, u);
        }
        j0 += n;
        if (elect_one()) mma_commit(accf_bar(as));
        __syncwarp();
        if (++as == 2) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else if (warp == kMmaWarp0 + kMmaWarps) {
    // ===================== record producer: bulk copies of each tile's record, p.nrec - 1 tiles ahead ===============
    uint32_t it = 0;
    const uint32_t rec0 = smem_u32(s_rec);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t rb = it % p.nrec;
      const int4 m = __ldg(p.meta + tile);
      const uint32_t nbytes = uni((uint32_t)m.x);
      const uint32_t ub = (min(uni((uint32_t)m.y), (uint32_t)kRcap) * 4u + 15u) & ~15u;
      mbar_wait(rece_bar(rb), ((it / p.nrec) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(recf_bar(rb), nbytes + ub);
        const uint32_t dst = rec0 + rb * rec_buf;
        bulk_g2s(dst, p.recs + tile * (int64_t)rec_stride, nbytes, recf_bar(rb));
        if (ub) bulk_g2s(dst + rec_main, p.uidx + tile * (int64_t)(kTileM * p.K), ub, recf_bar(rb));
      }
      __syncwarp();
    }
  } else {
    // ===================== weight producer: bulk copy of each unit's pre-swizzled weight block into its stage ======
    uint32_t j = 0;
    const uint32_t w_base = smem_u32(s_w);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const uint32_t mask = uni((uint32_t)__ldg(&p.meta[tile].z));
      for (int kb = 0; kb < KB; ++kb) {
        for (uint32_t mm = mask; mm; mm &= mm - 1, ++j) {
          const int k = __ffs(mm) - 1;
          const uint32_t sw = j & (na - 1);
          mbar_wait(aempty_bar(sw), ((j >> p.lna) & 1) ^ 1);
          if (lane == 0) tl_event(p, 24, 0, 0, j);
          if (elect_one()) {
            if WSIS_DBG(8) {
              mbar_arrive(afull_bar(sw));
            } else {
              mbar_expect_tx(afull_bar(sw), w_stage);
              bulk_g2s(w_base + sw * w_stage, p.packed + (size_t)(k * KB + kb) * w_stage, w_stage, afull_bar(sw));
            }
          }
          __syncwarp();
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// W[K, Cin_w, Cout_w] fp32 -> [K][KB][NS][N x 64 B swizzled] bf16 (hi, mid); KB = ceil(Cin / 32), the channels
// beyond Cin are zero
__global__ void pack_weights_kernel(const float *__restrict__ W, int K, int Cin, int Cout, int transpose_w, int NS,
                                    uint8_t *__restrict__ packed) {
  const int KB = (Cin + kKB - 1) / kKB;
  const int64_t total = (int64_t)K * KB * Cout * kKB;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int kk = (int)(e % kKB);
    int n = (int)((e / kKB) % Cout);
    int kb = (int)((e / ((int64_t)kKB * Cout)) % KB);
    int k = (int)(e / ((int64_t)kKB * Cout * KB));
    int c = kb * kKB + kk;
    float w = c >= Cin ? 0.f : transpose_w ? W[((int64_t)k * Cout + n) * Cin + c] : W[((int64_t)k * Cin + c) * Cout + n];
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

// Launch plan of a layer shape: TMEM accumulators / issuers and the largest pipeline (stages, row-cache buffers) that
// fits the 227 KB of dynamic shared memory.  Pure host arithmetic (also exported as wsis_conv_umma_plan so that the
// CPU test suite can check that every layer shape of the model has a plan).
static int plan_launch(wsis::umma::Params &p, int K, int Cin, int Cout, int NS, int max_record_bytes, int64_t *smem_out) {
  using namespace wsis::umma;
  p.KB = (Cin + kKB - 1) / kKB;
  // one accumulate chain (TMEM accumulator) per issuer, as far as the 512 TMEM columns allow
  int nacc = kMmaWarps;
  while (nacc > 1 && 2 * nacc * Cout > 512) nacc >>= 1;
  p.nacc = nacc;
  p.nmma = std::min(nacc, kMmaWarps);  // one accumulate chain per issuer (two when nacc > nmma)
  int cols = 32;
  while (cols < 2 * nacc * Cout) cols <<= 1;
  p.tmem_cols = cols;
  // shared memory: 2^lna pipeline stages (operand + weight blocks), nrc row-cache buffers, nrec record buffers;
  // shrink in this order of preference until the layer fits
  const int64_t a_stage = (int64_t)NS * kABlockBytes, w_stage = (int64_t)NS * Cout * 64;
  // the record buffers are sized for the largest record of THIS tile map when the caller knows it (0 = worst case)
  WSIS_CHECK(max_record_bytes >= 0 && max_record_bytes % 16 == 0 && max_record_bytes <= rec_stride_bytes(K),
             "conv_umma: max_record_bytes %d must be a multiple of 16 in [0, %d]", max_record_bytes, rec_stride_bytes(K));
  p.rec_main = max_record_bytes ? max_record_bytes : rec_stride_bytes(K);
  const int64_t rc_buf = (int64_t)kRcap * NS * 64, rec_buf = p.rec_main + kRcap * 4;
  static const int pref[][2] = {{3, 3}, {3, 2}, {2, 3}, {2, 2}, {1, 2}, {1, 1}, {0, 1}};  // {lna, nrc}
  const int64_t budget = 227 * 1024;
  p.nrec = 2;  // more record buffers did not pay (profiles/README.md); the kernel supports up to kMaxRec
  int64_t smem = 0;
  bool fit = false;
  for (auto &c : pref) {
    const int na = 1 << c[0], nrc = c[1];
    const int64_t misc = 1024 /*align*/ + 2 * p.KB * kKB * 4 + na * 16 + (2 * na + 2 * nrc + 2 * p.nrec + 4) * 8 + 64;
    smem = misc + na * (a_stage + w_stage) + nrc * rc_buf + p.nrec * rec_buf;
    if (smem <= budget) {
      p.lna = c[0];
      p.nrc = nrc;
      p.nb = std::min(kBuildWarps / 2, na);  // stage owners; two warps each
      p.nmma = std::min(p.nmma, na);  // every A/W stage belongs to exactly one issuer
      fit = true;
      break;
    }
  }
  WSIS_CHECK(fit, "conv_umma: shared memory budget exceeded for Cin=%d Cout=%d K=%d", Cin, Cout, K);
  *smem_out = smem;
  return 0;
}

static unsigned long long *g_tl = nullptr;
static int g_tl_cap = 0;

extern "C" {

int wsis_conv_debug_timeline(void *buf, int capacity) {
  g_tl = reinterpret_cast<unsigned long long *>(buf);
  g_tl_cap = capacity;
  return 0;
}

int wsis_conv_umma_supported(int Cin, int Cout) {
  return Cin >= 1 && Cin <= 1024 && Cout >= 16 && Cout % 16 == 0 && Cout <= 256;
}

int64_t wsis_conv_pack_bytes(int K, int Cin, int Cout, int precision) {
  int NS = precision == 3 ? 2 : 1;
  return (int64_t)K * ((Cin + kKB - 1) / kKB) * NS * Cout * 64;
}

int wsis_conv_pack_weights(const float *W, int K, int Cin, int Cout, int transpose_w, int precision, void *packed,
                           wsis_stream_t stream) {
  WSIS_CHECK(wsis_conv_umma_supported(Cin, Cout), "pack_weights: unsupported Cin=%d Cout=%d", Cin, Cout);
  WSIS_CHECK(precision == 1 || precision == 3, "pack_weights: precision must be 1 or 3");
  WSIS_CHECK((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "pack_weights: packed must be 16-byte aligned");
  int64_t total = (int64_t)K * ((Cin + kKB - 1) / kKB) * kKB * Cout;
  unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 8);
  pack_weights_kernel<<<blocks, 256, 0, as_stream(stream)>>>(W, K, Cin, Cout, transpose_w, precision == 3 ? 2 : 1,
                                                             (uint8_t *)packed);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_conv_umma_plan(int K, int Cin, int Cout, int precision, int max_record_bytes, int32_t *plan) {
  WSIS_CHECK(wsis_conv_umma_supported(Cin, Cout), "conv_umma_plan: unsupported Cin=%d Cout=%d", Cin, Cout);
  WSIS_CHECK(K >= 1 && K <= 32, "conv_umma_plan: kernel volume %d not in [1,32]", K);
  WSIS_CHECK(precision == 1 || precision == 3, "conv_umma_plan: precision must be 1 or 3");
  Params p;
  int64_t smem = 0;
  if (plan_launch(p, K, Cin, Cout, precision == 3 ? 2 : 1, max_record_bytes, &smem)) return 1;
  const int out[8] = {(int)smem, 1 << p.lna, p.nrc, p.nrec, p.nb, p.nacc, p.nmma, p.tmem_cols};
  for (int i = 0; i < 8; ++i) plan[i] = out[i];
  return 0;
}

int wsis_conv_umma(const float *src, const void *records, const int32_t *uidx, const int32_t *meta,
                   const int32_t *order, int64_t num_tiles, int K, int max_record_bytes, const void *packed, int Cin,
                   int Cout, int precision,
                   const float *in_scale, const float *in_shift, int in_relu, const float *residual, float *dst,
                   wsis_stream_t stream) {
  WSIS_CHECK(wsis_conv_umma_supported(Cin, Cout), "conv_umma: unsupported Cin=%d Cout=%d", Cin, Cout);
  WSIS_CHECK(K >= 1 && K <= 32, "conv_umma: kernel volume %d not in [1,32]", K);
  WSIS_CHECK(precision == 1 || precision == 3, "conv_umma: precision must be 1 or 3");
  WSIS_CHECK((in_scale == nullptr) == (in_shift == nullptr), "conv_umma: in_scale/in_shift must both be set");
  WSIS_CHECK(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(residual) |
               reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(records) |
               reinterpret_cast<uintptr_t>(uidx) | reinterpret_cast<uintptr_t>(meta)) & 15) == 0,
             "conv_umma: dst/residual/packed/records/uidx/meta must be 16-byte aligned");
  if (num_tiles == 0) return 0;
  const int NS = precision == 3 ? 2 : 1;
  Params p;
  p.src = src;
  p.recs = (const uint8_t *)records;
  p.uidx = uidx;
  p.meta = reinterpret_cast<const int4 *>(meta);
  p.order = order;
  p.packed = (const uint8_t *)packed;
  p.in_scale = in_scale;
  p.in_shift = in_shift;
  p.residual = residual;
  p.dst = dst;
  p.K = K;
  p.Cin = Cin;
  p.Cout = Cout;
  p.KB = (Cin + kKB - 1) / kKB;
  p.in_relu = in_relu;
  p.vec4 = (Cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  p.num_tiles = num_tiles;
  p.tl = g_tl;
  p.tl_cap = g_tl_cap;
  {
    const char *d = getenv("WSIS_CONV_DEBUG");
    p.dbg = d ? atoi(d) : 0;
  }
  int64_t smem = 0;
  if (plan_launch(p, K, Cin, Cout, NS, max_record_bytes, &smem)) return 1;
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
