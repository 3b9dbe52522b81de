// Output-stationary sparse convolution on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces the reference's per-offset gather -> cuBLAS mm -> scatter-add loop
// (include/spconv/spconv_ops.h:296-344; kernels include/spconv/reordering.cu.h:21-157).
//
//   dst[r,:] = residual[r,:] + sum_k  prologue(src[nbr[r,k],:]) . W[k]
//
// Work is organised in TILES of 128 output rows taken in a spatially sorted (Morton) order, built once per
// coordinate set by tilemap.cu: `order[t*128+i]` is the output row of tile slot i; the tile's RECORD holds, per
// kernel offset k, the valid-slot mask and for every slot the index of its source row in the tile's list of
// DISTINCT source rows.  A tile is a compact surface patch: ~160-230 distinct rows feed its ~480-1400 entries, and
// neighbouring tiles share their halos through L2, so HBM sees each feature row about once.
//
// v4: the gathered operand lives in TENSOR MEMORY.  An M=128 x K=16 bf16 A operand read from shared memory costs
// 4 KB of shared-memory bandwidth per MMA whatever N is, which bounds the SS form at 32 + N/4 cycles per MMA
// (measured, tools/umma_probe.cu -> profiles/r02_umma_probe.jsonl: 40 / 48 cycles at N = 32 / 64 against a tensor
// floor of N/2 = 16 / 32).  With A in TMEM (`tcgen05.mma [d], [a_tmem], b_desc`) two issuers reach the N/2 floor,
// shared memory carries only the weights, and the operand builders need no swizzled shared-memory staging at all:
// builder thread r owns tile slot r = TMEM lane r, reads its neighbour's converted row from the row cache and parks
// it with one tcgen05.st.
//
// One persistent CTA per SM; a CTA owns tiles (= the 128 TMEM lanes of the accumulators).  736 threads:
//   warps 0-3    epilogue  : tcgen05.ld accumulators -> registers -> (+residual) -> global, once per tile
//   warps 4-7    gatherers : fetch every distinct source row of the tile ONCE per channel block (16-byte loads,
//                            8 in flight per lane), fused eval-BatchNorm+ReLU prologue, fp32 -> bf16 hi (+ bf16
//                            mid for the 1e-4 path), park the rows in a shared-memory ROW CACHE (nrc buffers)
//   warps 8-15   builders  : two groups of four warps; a group owns a pipeline unit (active offset k, channel
//                            block kb): thread r loads loc[k][r], copies its 128-byte row-cache row into registers
//                            and stores it to lane r of the unit's TMEM operand slot (tcgen05.st.32x32b.x32)
//   warps 16-19  issuers   : nmma of them; warp-uniform loop, one elect.sync lane issues the unit's
//                            tcgen05.mma.kind::f16 (M=128, N=Cout, K=16, A from TMEM, B = weights in shared memory)
//                            with the unit's disable-output-lane mask, then ONE tcgen05.commit per unit
//   warp 20      records   : cp.async.bulk of the next tile records (+ their first kRcap unique rows)
//   warps 21-22  weights   : cp.async.bulk of each unit's pre-swizzled weight block into the unit's stage; layers
//                            whose whole packed weight fits (e.g. 32->32: 108 KB) keep it resident instead
// Units flow through a ring of `na` stages {TMEM operand slot, weight block} with one barrier pair per stage
// (full: 4 builder-warp arrives [+ the weight copy's expect_tx and bytes]; empty: tcgen05.commit).  Every issuer
// has its own accumulator (summed by the epilogue); accumulators are double buffered in TMEM so the epilogue of
// tile t overlaps the MMAs of tile t+1.  There is no scatter and there are no atomics: every output row is
// written exactly once.  DESIGN.md 4 has the measured cost model.
//
// Precision: precision==3 splits both operands into bf16 hi + bf16 mid and issues hi.hi + hi.mid + mid.hi (error
// ~2^-17 per product, fp32 accumulate in TMEM) to honour the reference's fp32 contract (1e-4) while staying on the
// tensor pipe; a unit is 32 channels (128 B of hi+mid per row).  precision==1 uses bf16 operands (1e-2 contract);
// a unit is 64 channels (again 128 B per row).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

#ifndef WSIS_KO
#define WSIS_KO 0
#endif

namespace wsis {
namespace umma {

constexpr int kTileM = 128;
constexpr int kEpiWarps = 4, kGatherWarps = 4, kBuildGroups = 2, kBuildWarps = 4 * kBuildGroups, kMmaWarps = 4,
              kRecWarps = 1, kWgtWarps = 2;
constexpr int kGatherWarp0 = kEpiWarps, kBuildWarp0 = kGatherWarp0 + kGatherWarps, kMmaWarp0 = kBuildWarp0 + kBuildWarps,
              kRecWarp0 = kMmaWarp0 + kMmaWarps, kWgtWarp0 = kRecWarp0 + kRecWarps;
constexpr int kThreads = 768;  // six warpgroups; the last warp of the producers' group is idle
static_assert((kWgtWarp0 + kWgtWarps) * 32 <= kThreads && kMmaWarp0 % 4 == 0 && kRecWarp0 % 4 == 0, "roles are laid out in whole warpgroups");
static_assert(kBuildWarp0 % 4 == 0, "a builder warp must own TMEM lanes 32 * (warp % 4)");
constexpr int kRcap = 256;    // rows of one row-cache buffer (a surface tile reads ~160-230 distinct rows)
constexpr int kRowB = 128;    // bytes of one converted row: 32 ch x (hi, mid) or 64 ch x bf16
constexpr int kRowStride = 144;  // row-cache row pitch: 128 + 16 bytes of padding (see rc layout below)
constexpr int kRcBuf = kRcap * kRowStride;
constexpr int kSlotCols = 32;  // TMEM columns of one operand slot (128 B per lane)
constexpr int kUlSlots = 4;                 // ring of unique-row lists (what the gatherers read), prefetched this deep
constexpr int kUlBytes = 16 + kRcap * 4;    // {nU, pad} + the first kRcap unique rows of a tile
constexpr int kResidentBytes = 112 * 1024;  // packed weights up to this size stay in shared memory for the whole launch

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

// try_wait blocks in hardware for a short, implementation-defined time; NO suspend-time hint: with a hint ptxas emits
// NANOSLEEP.SYNCS after a failed probe and the wake-up latency of a sleeping warp (on both sides of every stage hand-off)
// dominated the pipeline's round trip.  A wait that has not completed after ~4 s is a protocol bug: report which
// barrier and trap instead of hanging the device.
__device__ unsigned int *d_trap_word = nullptr;   // host-mapped diagnostics record (common.cuh trap_word_device)
__device__ __forceinline__ void mbar_deadlock(uint32_t bar, uint32_t parity) {
  if (d_trap_word != nullptr) {
    volatile unsigned int *w = d_trap_word;
    w[1] = bar, w[2] = parity, w[3] = blockIdx.x, w[4] = threadIdx.x;
    w[0] = 1u;
    __threadfence_system();
  }
  // no printf here: a device-side call would force the ABI on the kernel, and setmaxnreg needs a call-free kernel.
  // The trap surfaces as a launch failure on the next CUDA call; WSIS_CONV_DIAG builds report which barrier.
  (void)bar;
  (void)parity;
  __trap();
}
// The wait loop is written in PTX so that one failed probe costs ~3 issue slots (try_wait + branch, the watchdog count
// every fourth probe): the C++ form of this loop compiled to ~20 instructions per probe, and with ~20 warps of a CTA
// parked on barriers the spinning alone used 60 % of the SM's issue slots (ncu: 153 M warp instructions, 6.6 M probes
// per launch) and starved the warps doing the work.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .u32 n;\n"
      "mov.u32 n, 0;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "@p bra LAB_DONE;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "@p bra LAB_DONE;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "@p bra LAB_DONE;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "@p bra LAB_DONE;\n"
      "add.u32 n, n, 1;\n"
      "setp.lt.u32 q, n, 1048576;\n"  // every failed try_wait has already blocked for a hardware time quantum
      "@q bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  if (!ok) mbar_deadlock(bar, parity);
}
// Waits that are not on the unit pipeline's critical path (the epilogue waits several microseconds for a tile, the
// record producer for a free buffer) let the warp sleep between probes: a sleeping warp costs no issue slots.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t ns = 600) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .u32 n;\n"
      "mov.u32 n, 0;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "@p bra LAB_DONE;\n"
      "nanosleep.u32 %3;\n"
      "add.u32 n, n, 1;\n"
      "setp.lt.u32 q, n, 4194304;\n"
      "@q bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  if (!ok) mbar_deadlock(bar, parity);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// In-order issue means a warp stalls at the first instruction that needs the probe's result (~90 cycles even when the
// barrier has already flipped).  These two forms put independent shared-memory loads between the first probe and the
// branch that consumes it, so the probe's latency and the loads' latency overlap.
//   rows: v[0..7] <- 8 x 16 bytes at rbase + ((c ^ f) << 4) when `has` (other lanes keep their registers)
__device__ __forceinline__ void mbar_wait_and_load_row(uint32_t bar, uint32_t parity, uint4 (&v)[8], uint32_t rbase,
                                                       uint32_t has, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p, q, h;\n"
      ".reg .u32 n;\n"
      "setp.ne.u32 h, %36, 0;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%33], %34;\n"
      "@h ld.shared.v4.u32 {%1,%2,%3,%4}, [%35];\n"
      "@h ld.shared.v4.u32 {%5,%6,%7,%8}, [%35+16];\n"
      "@h ld.shared.v4.u32 {%9,%10,%11,%12}, [%35+32];\n"
      "@h ld.shared.v4.u32 {%13,%14,%15,%16}, [%35+48];\n"
      "@h ld.shared.v4.u32 {%17,%18,%19,%20}, [%35+64];\n"
      "@h ld.shared.v4.u32 {%21,%22,%23,%24}, [%35+80];\n"
      "@h ld.shared.v4.u32 {%25,%26,%27,%28}, [%35+96];\n"
      "@h ld.shared.v4.u32 {%29,%30,%31,%32}, [%35+112];\n"
      "mov.u32 n, 0;\n"
      "@p bra LAB_DONE;\n"
      "LAB_WAIT:\n"
      "nanosleep.u32 %37;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%33], %34;\n"
      "@p bra LAB_DONE;\n"
      "add.u32 n, n, 1;\n"
      "setp.lt.u32 q, n, 2097152;\n"
      "@q bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok), "+r"(v[0].x), "+r"(v[0].y), "+r"(v[0].z), "+r"(v[0].w), "+r"(v[1].x), "+r"(v[1].y), "+r"(v[1].z),
        "+r"(v[1].w), "+r"(v[2].x), "+r"(v[2].y), "+r"(v[2].z), "+r"(v[2].w), "+r"(v[3].x), "+r"(v[3].y), "+r"(v[3].z),
        "+r"(v[3].w), "+r"(v[4].x), "+r"(v[4].y), "+r"(v[4].z), "+r"(v[4].w), "+r"(v[5].x), "+r"(v[5].y), "+r"(v[5].z),
        "+r"(v[5].w), "+r"(v[6].x), "+r"(v[6].y), "+r"(v[6].z), "+r"(v[6].w), "+r"(v[7].x), "+r"(v[7].y), "+r"(v[7].z),
        "+r"(v[7].w)
      : "r"(bar), "r"(parity), "r"(rbase), "r"(has), "r"(ns)
      : "memory");
  if (!ok) mbar_deadlock(bar, parity);
}
__device__ __forceinline__ void load_row(uint4 (&v)[8], uint32_t rbase, uint32_t has) {
  asm volatile(
      "{\n"
      ".reg .pred h;\n"
      "setp.ne.u32 h, %33, 0;\n"
      "@h ld.shared.v4.u32 {%0,%1,%2,%3}, [%32];\n"
      "@h ld.shared.v4.u32 {%4,%5,%6,%7}, [%32+16];\n"
      "@h ld.shared.v4.u32 {%8,%9,%10,%11}, [%32+32];\n"
      "@h ld.shared.v4.u32 {%12,%13,%14,%15}, [%32+48];\n"
      "@h ld.shared.v4.u32 {%16,%17,%18,%19}, [%32+64];\n"
      "@h ld.shared.v4.u32 {%20,%21,%22,%23}, [%32+80];\n"
      "@h ld.shared.v4.u32 {%24,%25,%26,%27}, [%32+96];\n"
      "@h ld.shared.v4.u32 {%28,%29,%30,%31}, [%32+112];\n"
      "}\n"
      : "+r"(v[0].x), "+r"(v[0].y), "+r"(v[0].z), "+r"(v[0].w), "+r"(v[1].x), "+r"(v[1].y), "+r"(v[1].z), "+r"(v[1].w),
        "+r"(v[2].x), "+r"(v[2].y), "+r"(v[2].z), "+r"(v[2].w), "+r"(v[3].x), "+r"(v[3].y), "+r"(v[3].z), "+r"(v[3].w),
        "+r"(v[4].x), "+r"(v[4].y), "+r"(v[4].z), "+r"(v[4].w), "+r"(v[5].x), "+r"(v[5].y), "+r"(v[5].z), "+r"(v[5].w),
        "+r"(v[6].x), "+r"(v[6].y), "+r"(v[6].z), "+r"(v[6].w), "+r"(v[7].x), "+r"(v[7].y), "+r"(v[7].z), "+r"(v[7].w)
      : "r"(rbase), "r"(has)
      : "memory");
}
//   mask: m <- 16 bytes at maddr (the issuer's disable-output-lane mask of the unit)
__device__ __forceinline__ void mbar_wait_and_load16(uint32_t bar, uint32_t parity, uint4 &m, uint32_t maddr, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .u32 n;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%5], %6;\n"
      "ld.shared.v4.u32 {%1,%2,%3,%4}, [%7];\n"
      "mov.u32 n, 0;\n"
      "@p bra LAB_DONE;\n"
      "LAB_WAIT:\n"
      "nanosleep.u32 %8;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%5], %6;\n"
      "@p bra LAB_DONE;\n"
      "add.u32 n, n, 1;\n"
      "setp.lt.u32 q, n, 2097152;\n"
      "@q bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok), "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w)
      : "r"(bar), "r"(parity), "r"(maddr), "r"(ns)
      : "memory");
  if (!ok) mbar_deadlock(bar, parity);
}

// Diagnostics build of the kernel (template parameter DIAG): every role accumulates the cycles it spends in each kind
// of wait; CTA 0 writes them to p.diag[warp * 8 + i] (i = 0: whole role loop, 1..: the role's waits in code order).
template <bool DIAG, bool RELAXED = false>
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long &acc, uint32_t ns) {
  const long long t0 = DIAG ? clock64() : 0;
  if (RELAXED)
    mbar_wait_relaxed(bar, parity);
  else if (ns)
    mbar_wait_relaxed(bar, parity, ns);
  else
    mbar_wait(bar, parity);
  if (DIAG) acc += clock64() - t0;
}

// One lane of a fully active warp.  Code under this predicate is known to be single-lane, and values made uniform
// with uni() live in uniform registers: ptxas then feeds tcgen05.mma / tcgen05.commit / cp.async.bulk their
// uniform-register operands directly instead of wrapping every instruction in an ELECT + R2UR.BROADCAST waterfall
// loop (which costs ~150 cycles per MMA).
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
// set are not written at all (the PTX disable-output-lane vector), so their A rows may hold stale data.
// A: TMEM, lane = row, 8 columns = 16 bf16 (column j holds elements 2j, 2j+1); B: shared memory, K-major.
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint4 off) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, 1, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%4, %5, %6, %7}, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(off.x), "r"(off.y), "r"(off.z), "r"(off.w)
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
// one operand row (128 B = 32 columns) into this thread's TMEM lane
__device__ __forceinline__ void tmem_st_row(uint32_t taddr, const uint4 *v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "r"(v[0].x), "r"(v[0].y), "r"(v[0].z), "r"(v[0].w), "r"(v[1].x), "r"(v[1].y), "r"(v[1].z), "r"(v[1].w),
      "r"(v[2].x), "r"(v[2].y), "r"(v[2].z), "r"(v[2].w), "r"(v[3].x), "r"(v[3].y), "r"(v[3].z), "r"(v[3].w),
      "r"(v[4].x), "r"(v[4].y), "r"(v[4].z), "r"(v[4].w), "r"(v[5].x), "r"(v[5].y), "r"(v[5].z), "r"(v[5].w),
      "r"(v[6].x), "r"(v[6].y), "r"(v[6].z), "r"(v[6].w), "r"(v[7].x), "r"(v[7].y), "r"(v[7].z), "r"(v[7].w)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of (row, 16-byte chunk c16 in [0,4)) inside a [rows x 64 B] K-major SWIZZLE_64B block
__host__ __device__ __forceinline__ uint32_t sw64(uint32_t row, uint32_t c16) {
  return (row >> 3) * 512u + (row & 7u) * 64u + ((c16 ^ ((row & 7u) >> 1)) << 4);
}
// Row cache: row u is 8 chunks of 16 B at pitch 144 B, so chunk c of row u sits in shared-memory bank group
// (u + c) mod 8.  Builder lanes read the SAME chunk of DIFFERENT rows: rows whose indices differ in their low three bits
// land in different bank groups, and all eight chunks of a row are base + immediate offsets (no per-chunk address math).

struct Params {
  const float *src;
  const uint8_t *recs;    // tile records at stride rec_stride_bytes(K) (tilemap.cu)
  const int32_t *uidx;    // [num_tiles][128 K] distinct source rows of each tile
  const int4 *meta;       // [num_tiles] {record bytes, nU, active-offset mask, P}
  const int32_t *order;   // [num_tiles*128] destination row of each tile slot (-1 = padding)
  const uint8_t *packed;
  const float *in_scale, *in_shift, *residual;
  float *dst;
  int K, Cin, Cout, KB, in_relu, vec4;
  int us;  // units (offset x channel block) per pipeline stage, 1..4
  int lna, nrc, nrec, nacc, nmma, nbuf, resident, nwp;  // 1 << lna pipeline stages, nrc row-cache buffers, nrec record
                                                        // buffers, accumulators (= issuers), accumulator buffers (1|2),
                                                        // weights resident in shared memory, weight-producer warps
  int64_t num_tiles;
  int dbg;          // timing experiments (WSIS_CONV_DEBUG, results are wrong): 1 no MMA, 2 no row-cache reads, 4 no
                    // TMEM operand store, 8 no gather loads, 16 no epilogue read-out
  int wait_ns;      // experiment: > 0 = every pipeline wait sleeps up to this long between probes
  long long *diag;  // optional (wsis_conv_debug_stats): per warp of CTA 0, cycles in the role loop and in its waits
};

// relu?(x * sc + sh) -> bf16 hi (part 0) or bf16 mid = bf16(y - hi) (part 1), two values per 32-bit word
__device__ __forceinline__ uint32_t split2(float a, float b, int part) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  uint32_t hv = *reinterpret_cast<uint32_t *>(&h);
  if (part == 0) return hv;
  // bf16 -> fp32 is a 16-bit shift: the residual y - hi is exact in fp32
  __nv_bfloat162 m = __floats2bfloat162_rn(a - __uint_as_float(hv << 16), b - __uint_as_float(hv & 0xffff0000u));
  return *reinterpret_cast<uint32_t *>(&m);
}

__device__ __forceinline__ float4 load_row4(const float *src, int Cin, int vec4, int32_t row, int c0) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const float *s = src + (int64_t)row * Cin + c0;
  if (vec4) {
    if (c0 < Cin) v = __ldg(reinterpret_cast<const float4 *>(s));
  } else {  // narrow or unaligned rows (the 6-channel input layer): guarded scalar loads, zero padding
    if (c0 < Cin) v.x = __ldg(s);
    if (c0 + 1 < Cin) v.y = __ldg(s + 1);
    if (c0 + 2 < Cin) v.z = __ldg(s + 2);
    if (c0 + 3 < Cin) v.w = __ldg(s + 3);
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
__device__ __forceinline__ uint4 fetch_direct(const float *src, int Cin, int vec4, int in_relu, const int32_t *urows,
                                           uint32_t loc, int c0, int part, const float *s_scale, const float *s_shift) {
  const int32_t row = __ldg(urows + loc);
  const float4 y0 = prologue4(load_row4(src, Cin, vec4, row, c0), *reinterpret_cast<const float4 *>(s_scale + c0),
                              *reinterpret_cast<const float4 *>(s_shift + c0), in_relu);
  const float4 y1 = prologue4(load_row4(src, Cin, vec4, row, c0 + 4), *reinterpret_cast<const float4 *>(s_scale + c0 + 4),
                              *reinterpret_cast<const float4 *>(s_shift + c0 + 4), in_relu);
  return make_uint4(split2(y0.x, y0.y, part), split2(y0.z, y0.w, part), split2(y1.x, y1.y, part),
                    split2(y1.z, y1.w, part));
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// NS = 2: fp32 contract (hi + mid, 32 channels per unit); NS = 1: bf16 operands (64 channels per unit)
template <int NS, int US, bool DIAG>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t *sm = smem_raw + pad;
  constexpr int CPU = NS == 2 ? 32 : 64;  // channels per unit
  const int KB = p.KB;
  const uint32_t na = 1u << p.lna;  // pipeline stages: TMEM operand slot + the unit's weight block, one barrier pair
  const uint32_t b_block = (uint32_t)p.Cout * 64u;  // one [Cout x 32 ch] K-major swizzled weight block
  const uint32_t w_stage = 2 * b_block;             // (hi, mid) of 32 channels, or two 32-channel halves of bf16
  const uint32_t rec_main = (uint32_t)rec_stride_bytes(p.K);
  const uint32_t rec_buf = rec_main;
  constexpr uint32_t us = US;  // units per stage
  // timing experiments: run-time switches in the diagnostics build; the product build can be compiled with a fixed set
  // (-DWSIS_KO=bits, tools/gpu_knockout.sh) so that the experiment runs on product-quality code
  const int dbg = DIAG ? p.dbg : WSIS_KO;
  const uint32_t w_bytes = p.resident ? (uint32_t)(p.K * KB) * w_stage : na * us * 2 * w_stage;  // two members per unit
  uint8_t *s_w = sm;                                   // weight blocks: ring of na stages, or the whole packed weight
  uint8_t *s_rc = s_w + w_bytes;                       // [nrc][kRcap][128 B]   converted source rows
  uint8_t *s_rec = s_rc + (size_t)p.nrc * kRcBuf;      // [nrec][rec_buf]       tile records (bulk copies)
  uint8_t *s_ul = s_rec + (size_t)p.nrec * rec_buf;    // [kUlSlots][kUlBytes]  unique-row lists (bulk copies)
  float *s_scale = reinterpret_cast<float *>(s_ul + kUlSlots * kUlBytes);
  float *s_shift = s_scale + KB * CPU;
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_shift + KB * CPU);
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
  const uint32_t wres_bar = bar2 + 8u * (2 * p.nrec + 4);
  auto ulf_bar = [&](uint32_t s) { return wres_bar + 8u * (1 + s); };
  auto ule_bar = [&](uint32_t s) { return wres_bar + 8u * (1 + kUlSlots + s); };
  auto rawf_bar = [&](uint32_t s) { return wres_bar + 8u * (1 + 2 * kUlSlots + s); };  // raw rows landed (bulk copies)
  const uint32_t nbars = 2 * na + 3 * p.nrc + 2 * p.nrec + 5 + 2 * kUlSlots;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + nbars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long dg[4] = {0, 0, 0, 0};
  const long long dg_t0 = DIAG ? clock64() : 0;
  int tl_n = 0;  // timeline of CTA 0 (diagnostics build): p.diag[256 + warp * 512 + n] = cycles since start << 8 | code
#define TL(code)                                                                                   \
  do {                                                                                             \
    if (DIAG && p.diag != nullptr && blockIdx.x == 0 && lane == 0 && tl_n < 512)                   \
      p.diag[256 + warp * 512 + tl_n++] = ((clock64() - dg_t0) << 8) | (long long)((code) & 0xff); \
  } while (0)
#define WAIT(i, bar, parity) mbar_wait_t<DIAG>(bar, parity, dg[i], (uint32_t)p.wait_ns)
#define WAIT_RELAXED(i, bar, parity) mbar_wait_t<DIAG, true>(bar, parity, dg[i], 0u)

  for (int c = tid; c < KB * CPU; c += kThreads) {  // channels padded up to the unit width contribute exact zeros
    s_scale[c] = c < p.Cin ? (p.in_scale ? p.in_scale[c] : 1.f) : 0.f;
    s_shift[c] = c < p.Cin ? (p.in_shift ? p.in_shift[c] : 0.f) : 0.f;
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < na; ++s) {
      // full: the four warps of the builder group that owns the unit (+ the weight producer's expect_tx arrive and
      // the bytes of its bulk copy); empty: ONE tcgen05.commit per unit releases the operand slot and the weight block
      mbar_init(afull_bar(s), 4 + (p.resident ? 0 : 1));
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < p.nrc; ++s) {
      mbar_init(rcf_bar(s), kGatherWarps);
      mbar_init(rce_bar(s), kBuildWarps);
      mbar_init(rawf_bar(s), kGatherWarps);
    }
    for (int s = 0; s < p.nrec; ++s) {
      mbar_init(recf_bar(s), 1);
      mbar_init(rece_bar(s), kBuildWarps + p.nmma);
    }
    for (int s = 0; s < kUlSlots; ++s) {
      mbar_init(ulf_bar(s), 1);
      mbar_init(ule_bar(s), kGatherWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accf_bar(s), p.nmma);
      mbar_init(acce_bar(s), kEpiWarps * 32);
    }
    mbar_init(wres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp0) {  // the first MMA warp owns the TMEM allocation (all 512 columns: one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t acc_cols = (uint32_t)(p.nacc * p.Cout);       // one accumulator buffer
  const uint32_t colA = tmem_base + (uint32_t)p.nbuf * acc_cols;  // operand slots follow the accumulators
  if (warp < kEpiWarps) {
    // Accumulators start at zero and every MMA accumulates: an MMA only touches the lanes (tile slots) that have a
    // neighbour through its offset, so no unit can be the one that "initialises" a tile.
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (uint32_t c0 = 0; c0 < (uint32_t)p.nbuf * acc_cols; c0 += 16) tmem_zero16(taddr + c0);
    tmem_wait_st();
  }
  if (p.resident && warp == kWgtWarp0) {  // the whole packed weight, once
    if (elect_one()) {
      mbar_expect_tx(wres_bar, w_bytes);
      for (uint32_t o = 0; o < w_bytes; o += 16384u)
        bulk_g2s(smem_u32(s_w) + o, p.packed + o, min(16384u, w_bytes - o), wres_bar);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kEpiWarps) {
    // ===================== epilogue: TMEM -> registers -> (+residual) -> global, once per tile =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;" ::: "memory");
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
      WAIT(0, accf_bar(as), aph);
      TL(1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * acc_cols;
      for (int c0 = 0; c0 < p.Cout && !(dbg & 16); c0 += 16) {
        float4 rn[4];
        const bool more = c0 + 16 < p.Cout;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          rn[q] = (has_res && more) ? __ldg(rs + (c0 + 16) / 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float v[16];
        tmem_ld16(taddr + c0, v);
        tmem_zero16(taddr + c0);  // hand the accumulator back cleared
        for (int a = 1; a < p.nacc; ++a) {  // partial sums of the issuers' independent chains
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
      TL(2);
      if (++as == (uint32_t)p.nbuf) {
        as = 0;
        aph ^= 1;
      }
    }
  } else if (warp < kBuildWarp0) {
    // ===================== gatherers: every distinct source row of a tile is fetched ONCE =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;" ::: "memory");
    // One PASS = (tile, channel block kb): the slices of the tile's unique rows are loaded with 16-byte loads
    // (8 rows per lane in flight), the fused eval-BatchNorm+ReLU prologue is applied, the result is split
    // fp32 -> bf16 hi (+ bf16 mid) and parked in a row-cache buffer, from which the builders assemble the per-offset
    // operands: HBM/L2 see each row once per tile instead of once per (row, offset) entry, and the loads of up to nrc
    // passes are in flight ahead of the tensor pipe.
    constexpr int LPR = CPU / 4;  // lanes per row: each lane converts 4 channels
    const int gt = (warp - kGatherWarp0) * 32 + lane;
    const int rsub = gt / LPR, chunk = gt % LPR;
    constexpr int kInFlight = 8;  // 16-byte register loads in flight per gather lane (the non-bulk path)
    const bool bulk_rows = p.vec4 && (p.Cin % CPU) == 0;
    uint32_t it = 0, q = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      // the tile's list of distinct rows comes through its own small ring, several tiles ahead of the (large) record
      // the builders and issuers hold: the gather of tile t+1 does not wait for the record ring to turn over
      const uint32_t ub = it % kUlSlots;
      WAIT_RELAXED(0, ulf_bar(ub), (it / kUlSlots) & 1);
      TL(3);
      const uint8_t *ul = s_ul + (size_t)ub * kUlBytes;
      // distinct rows of the tile (the overflow beyond a row-cache buffer is fetched directly by the builders)
      const int ng = min((int)*reinterpret_cast<const uint32_t *>(ul), kRcap);
      const int32_t *uidx = reinterpret_cast<const int32_t *>(ul + 16);
      for (int kb = 0; kb < KB; ++kb, ++q) {
        const uint32_t slot = q % p.nrc;
        const int c0 = kb * CPU + chunk * 4;
        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c0);
        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c0);
        uint8_t *rcb = s_rc + (size_t)slot * kRcBuf;
        bool waited = false;
        constexpr int kSweep = kGatherWarps * 32 / LPR;  // rows per load instruction of the gather warps
        if (NS == 2 && bulk_rows) {
          // ---- fp32 contract, 16-byte aligned rows of whole 32-channel blocks: the raw 128-byte row slices are
          // copied global -> row cache by cp.async (ALL rows of the pass in flight at once, no registers held, one
          // round trip of latency per pass instead of one per batch of 8 register loads; 128-byte cp.async.bulk
          // copies were slower: ~46 cycles of the copy engine each) and converted in place: a row's eight lanes read
          // their 16 bytes of fp32, then write 8 bytes of hi and 8 bytes of mid ----
          WAIT(1, rce_bar(slot), ((q / p.nrc) & 1) ^ 1);
          waited = true;
          if (!(dbg & 8)) {
            // every lane copies ITS 16-byte chunk of each of its rows and later converts exactly those bytes, so the
            // per-thread completion of cp.async is all the ordering the conversion needs
            for (int u = rsub; u < ng; u += kSweep) {
              const float *g = p.src + (int64_t)uidx[u] * p.Cin + c0;
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(rcb) + (uint32_t)u * kRowStride +
                                                                              (uint32_t)chunk * 16u),
                           "l"(g)
                           : "memory");
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
          }
          for (int u0 = 0; u0 < ng && !(dbg & 32); u0 += 4 * kSweep) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int u = u0 + i * kSweep + rsub;
              if (u < ng) v[i] = *reinterpret_cast<const float4 *>(rcb + (size_t)u * kRowStride + chunk * 16);
            }
            __syncwarp();  // every lane of a row has read its 16 raw bytes before any of them overwrites the row
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int u = u0 + i * kSweep + rsub;
              if (u < ng) {
                const float4 y = prologue4(v[i], sc, sh, p.in_relu);
                uint8_t *r = rcb + (size_t)u * kRowStride + chunk * 8;
                *reinterpret_cast<uint2 *>(r) = make_uint2(split2(y.x, y.y, 0), split2(y.z, y.w, 0));
                *reinterpret_cast<uint2 *>(r + 64) = make_uint2(split2(y.x, y.y, 1), split2(y.z, y.w, 1));
              }
            }
          }
        } else
        for (int u0 = 0; u0 < ng && !(dbg & 32); u0 += kInFlight * kSweep) {
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
            if (u < ng && !(dbg & 8)) v[i] = load_row4(p.src, p.Cin, p.vec4, idx[i], c0);
          }
          if (!waited) {  // the loads are in flight while the buffer drains
            WAIT(1, rce_bar(slot), ((q / p.nrc) & 1) ^ 1);
            waited = true;
          }
#pragma unroll
          for (int i = 0; i < kInFlight; ++i) {
            const int u = u0 + i * kSweep + rsub;
            if (u < ng) {
              const float4 y = prologue4(v[i], sc, sh, p.in_relu);
              uint8_t *r = rcb + (size_t)u * kRowStride + chunk * 8;  // 16-byte chunk (chunk >> 1), half (chunk & 1)
              *reinterpret_cast<uint2 *>(r) = make_uint2(split2(y.x, y.y, 0), split2(y.z, y.w, 0));
              if (NS == 2) *reinterpret_cast<uint2 *>(r + 64) = make_uint2(split2(y.x, y.y, 1), split2(y.z, y.w, 1));
            }
          }
        }
        if (!waited) WAIT(1, rce_bar(slot), ((q / p.nrc) & 1) ^ 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(rcf_bar(slot));
        TL(4);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(ule_bar(ub));  // this warp no longer reads the list
    }
  } else if (warp < kMmaWarp0) {
    // ===================== builders: thread r owns tile slot r = TMEM lane r =====================
    // Unit = (channel block kb, active kernel offset k) of a tile; a STAGE is up to `us` consecutive active offsets of
    // one (tile, kb) pass, assembled by builder group (stage counter mod groups): every thread whose slot has a
    // neighbour through the unit's offset copies that neighbour's converted row (row cache, shared memory) into
    // registers and all 128 threads store their registers to the unit's TMEM operand slot.  Lanes without a neighbour
    // store stale registers: the MMA never reads them into a live accumulator lane (disable-output-lane mask).
    // One barrier round trip, one fence and one arrive per stage; the row indices of the group's NEXT stage are fetched
    // while the current one is assembled.
    // The builders hold a 32-register operand row plus the next stage's row indices: they take the registers the
    // issuers and producers give back (setmaxnreg moves them between whole warpgroups).
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
    const int bw = warp - kBuildWarp0;
    const uint32_t g = (uint32_t)bw >> 2, w4 = (uint32_t)bw & 3u;
    const uint32_t slot_r = w4 * 32 + (uint32_t)lane;
    const uint32_t s_rc32 = smem_u32(s_rc), s_rec32 = smem_u32(s_rec);
    const uint32_t tlane = colA + ((w4 * 32u) << 16);
    const uint32_t hdr_off = 16u * (uint32_t)p.K;
    uint32_t it = 0, q = 0, Q = 0;  // Q = the CTA's stage counter at the start of the current pass
    uint4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t rb = it % p.nrec;
      WAIT(0, recf_bar(rb), (it / p.nrec) & 1);
      const uint32_t rec32 = s_rec32 + rb * rec_buf;
      const uint32_t nact = lds32(rec32 + hdr_off + 12);  // packs of the tile (units of one pass)
      const bool big = lds32(rec32 + hdr_off) > (uint32_t)kRcap;  // more distinct rows than a row-cache buffer holds
      const uint32_t locs32 = rec32 + hdr_off + 80 + slot_r * 2;  // loc row of the i-th pack at + 256 i
      const uint32_t nq = (nact + us - 1) / us;                    // stages of one pass
      for (int kb = 0; kb < KB; ++kb, ++q) {
        const uint32_t slot = q % p.nrc;
        WAIT(1, rcf_bar(slot), (q / p.nrc) & 1);
        TL(5);
        const uint32_t rcb = s_rc32 + slot * kRcBuf;
        uint32_t i = (g - Q) & (kBuildGroups - 1);
        if (!big) {
          // ---- the hot loop: every row of the tile is in the row cache ----
          uint32_t loc0 = 0xFFFFu, loc1 = 0xFFFFu, loc2 = 0xFFFFu, loc3 = 0xFFFFu;
          if (i < nq) {
            const uint32_t r0 = i * us, la = locs32 + 256u * r0;
            loc0 = lds_u16(la);
            if (us > 1 && r0 + 1 < nact) loc1 = lds_u16(la + 256);
            if (us > 2 && r0 + 2 < nact) loc2 = lds_u16(la + 512);
            if (us > 3 && r0 + 3 < nact) loc3 = lds_u16(la + 768);
          }
          for (; i < nq; i += kBuildGroups) {
            const uint32_t Qi = Q + i, stage = Qi & (na - 1), phase = (Qi >> p.lna) & 1;
            const uint32_t t0 = tlane + stage * us * kSlotCols;
            const uint32_t nu = min(us, nact - i * us);
            // unit 0: probe the stage's empty barrier, fetch this lane's row under the probe, then consume the probe
            const long long tw0 = DIAG ? clock64() : 0;
            mbar_wait_and_load_row(aempty_bar(stage), phase ^ 1, v, rcb + loc0 * kRowStride, loc0 != 0xFFFFu && !(dbg & 2),
                                   (uint32_t)p.wait_ns);
            if (DIAG) dg[2] += clock64() - tw0;
            TL(6);
            tc_fence_after();
            const long long ts0 = DIAG ? clock64() : 0;
            if (!(dbg & 4)) tmem_st_row(t0, v);
            if (nu > 1) {
              load_row(v, rcb + loc1 * kRowStride, loc1 != 0xFFFFu && !(dbg & 2));
              if (!(dbg & 4)) tmem_st_row(t0 + kSlotCols, v);
            }
            if (nu > 2) {
              load_row(v, rcb + loc2 * kRowStride, loc2 != 0xFFFFu && !(dbg & 2));
              if (!(dbg & 4)) tmem_st_row(t0 + 2 * kSlotCols, v);
            }
            if (nu > 3) {
              load_row(v, rcb + loc3 * kRowStride, loc3 != 0xFFFFu && !(dbg & 2));
              if (!(dbg & 4)) tmem_st_row(t0 + 3 * kSlotCols, v);
            }
            if (DIAG) dg[3] += clock64() - ts0;
            // row indices of this group's next stage (consumed one iteration later)
            const uint32_t rn = (i + kBuildGroups) * us, la = locs32 + 256u * rn;
            loc0 = loc1 = loc2 = loc3 = 0xFFFFu;
            if (rn < nact) loc0 = lds_u16(la);
            if (us > 1 && rn + 1 < nact) loc1 = lds_u16(la + 256);
            if (us > 2 && rn + 2 < nact) loc2 = lds_u16(la + 512);
            if (us > 3 && rn + 3 < nact) loc3 = lds_u16(la + 768);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(afull_bar(stage));
            TL(7);
          }
        } else {
          // ---- tiles that read more distinct rows than the row cache holds (dense blobs, random maps, the 2x2x2 strided
          // convolutions): rows beyond the cache are fetched from global memory and converted in place; kept out of the
          // hot loop ----
          const int32_t *urows = p.uidx + tile * (int64_t)(kTileM * p.K);
          for (; i < nq; i += kBuildGroups) {
            const uint32_t Qi = Q + i, stage = Qi & (na - 1), phase = (Qi >> p.lna) & 1;
            const uint32_t t0 = tlane + stage * us * kSlotCols;
            const uint32_t nu = min(us, nact - i * us);
            WAIT(2, aempty_bar(stage), phase ^ 1);
            tc_fence_after();
            for (uint32_t u = 0; u < nu; ++u) {
              const uint32_t l = lds_u16(locs32 + 256u * (i * us + u));
              if (l < (uint32_t)kRcap) {  // still in the row cache
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) v[c] = lds128(rcb + l * kRowStride + 16 * c);
              } else if (l != 0xFFFFu) {
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c)
                  v[c] = fetch_direct(p.src, p.Cin, p.vec4, p.in_relu, urows, l,
                                      kb * CPU + (NS == 2 ? (int)(c & 3u) : (int)c) * 8, NS == 2 ? (int)(c >> 2) : 0, s_scale,
                                      s_shift);
              }
              __syncwarp();
              tmem_st_row(t0 + u * kSlotCols, v);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(afull_bar(stage));
          }
        }
        Q += nq;
        __syncwarp();
        if (lane == 0) mbar_arrive(rce_bar(slot));  // this warp no longer reads row-cache buffer `slot`
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(rece_bar(rb));  // this warp no longer reads s_rec[rb]
    }
  } else if (warp < kRecWarp0) {
    // ===================== MMA issuers =====================
    // Issuer i (one elected thread of warp kMmaWarp0 + i) takes the units j = i (mod nmma) and accumulates into its
    // own TMEM accumulator.  One thread issues a short MMA every ~39 cycles at best (tools/umma_probe.cu), so
    // N <= 32 needs four issuers and N <= 64 two to keep the tensor pipe at its N/2-cycle floor.
    // The whole warp runs the (warp-uniform) loop; one elected lane issues.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    const uint32_t mi = uni((uint32_t)(warp - kMmaWarp0));
    if (mi < (uint32_t)p.nmma) {
      const uint32_t nmma = (uint32_t)p.nmma;
      const uint32_t idesc = make_idesc(p.Cout);
      const uint32_t w_base = smem_u32(s_w);
      const uint64_t desc0 = make_desc(0);
      const uint32_t a_base = uni(colA), d_base = uni(tmem_base) + mi * (uint32_t)p.Cout;
      uint32_t it = 0, Q = 0, as = 0, aph = 0;
      const uint32_t hdr_off = 16u * (uint32_t)p.K;
      if (p.resident) mbar_wait(wres_bar, 0);
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t rb = it % p.nrec;
        WAIT(0, recf_bar(rb), (it / p.nrec) & 1);
        const uint32_t rec32 = smem_u32(s_rec) + rb * rec_buf;
        const uint32_t nact = uni(lds32(rec32 + hdr_off + 12));               // packs of the tile
        const uint32_t kl = lds_u16(rec32 + hdr_off + 16 + 2u * (uint32_t)lane);  // lane l: members of pack l
        const uint32_t nq = (nact + us - 1) / us;
        WAIT(1, acce_bar(as), aph ^ 1);
        TL(8);
        tc_fence_after();
        const uint32_t d = d_base + as * acc_cols;
        for (int kb = 0; kb < KB; ++kb) {
          // 16-channel K steps of this unit that hold real channels
          const int ksteps = min(CPU / 16, (p.Cin - kb * CPU + 15) / 16);
          // stage Q of the CTA belongs to issuer Q mod nmma, who issues all of its units into its own accumulator.
          // A unit is a PACK: one operand block, multiplied once per member offset with that member's lane mask and
          // weight block.
          for (uint32_t i = (mi - Q) & (nmma - 1); i < nq; i += nmma) {
            const uint32_t Qi = Q + i, stage = Qi & (na - 1);
            const uint32_t nu = min(us, nact - i * us);
            const uint32_t mem0 = __shfl_sync(0xffffffffu, kl, (int)((i * us) & 31u));
            // operand rows (tcgen05.st, fenced by the builders) + weights; the valid-slot mask of the first member is
            // fetched under the probe (the MMA takes its complement)
            uint4 vm;
            const long long tw0 = DIAG ? clock64() : 0;
            mbar_wait_and_load16(afull_bar(stage), (Qi >> p.lna) & 1, vm, rec32 + 16 * (mem0 & 0xFFu) * ((mem0 & 0xFFu) != 0xFFu),
                                 (uint32_t)p.wait_ns);
            if (DIAG) dg[2] += clock64() - tw0;
            TL(9);
            tc_fence_after();
            const long long ti0 = DIAG ? clock64() : 0;
            for (uint32_t u = 0; u < nu; ++u) {
              const uint32_t mem = u == 0 ? mem0 : __shfl_sync(0xffffffffu, kl, (int)((i * us + u) & 31u));
              if (elect_one()) {
                const uint32_t a = a_base + (stage * us + u) * kSlotCols;
#pragma unroll
                for (uint32_t m = 0; m < 2; ++m) {
                  const uint32_t k = (mem >> (8 * m)) & 0xFFu;
                  if (k != 0xFFu) {
                    if (u || m) vm = lds128(rec32 + 16 * k);
                    const uint4 off = make_uint4(~vm.x, ~vm.y, ~vm.z, ~vm.w);
                    const uint64_t bd = desc0 + ((w_base + (p.resident ? (uint32_t)(k * KB + kb)
                                                                       : (stage * us + u) * 2 + m) * w_stage) >> 4);
                    if (!(dbg & 1)) {
                      if (NS == 2) {
                        const uint64_t bm = bd + (b_block >> 4);  // mid block
                        mma_ts(d, a, bd, idesc, off);
                        mma_ts(d, a, bm, idesc, off);
                        mma_ts(d, a + 16, bd, idesc, off);
                        if (ksteps > 1) {
                          mma_ts(d, a + 8, bd + 2, idesc, off);
                          mma_ts(d, a + 8, bm + 2, idesc, off);
                          mma_ts(d, a + 24, bd + 2, idesc, off);
                        }
                      } else {
                        mma_ts(d, a, bd, idesc, off);
                        if (ksteps > 1) mma_ts(d, a + 8, bd + 2, idesc, off);
                        if (ksteps > 2) mma_ts(d, a + 16, bd + (b_block >> 4), idesc, off);
                        if (ksteps > 3) mma_ts(d, a + 24, bd + (b_block >> 4) + 2, idesc, off);
                      }
                    }
                  }
                }
                if (u + 1 == nu) mma_commit(aempty_bar(stage));
              }
              __syncwarp();
            }
            if (DIAG) dg[3] += clock64() - ti0;
            TL(10);
          }
          Q += nq;
        }
        if (elect_one()) mma_commit(accf_bar(as));
        __syncwarp();
        TL(11);
        if (lane == 0) mbar_arrive(rece_bar(rb));  // this warp no longer reads s_rec[rb]
        if (++as == (uint32_t)p.nbuf) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else if (warp == kRecWarp0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    // ===================== record producer: bulk copies of each tile's record, p.nrec - 1 tiles ahead ===============
    uint32_t it = 0;
    const uint32_t rec0 = smem_u32(s_rec);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t rb = it % p.nrec;
      WAIT_RELAXED(0, rece_bar(rb), ((it / p.nrec) & 1) ^ 1);
      TL(12);
      if (elect_one()) {
        mbar_expect_tx(recf_bar(rb), rec_main);
        bulk_g2s(rec0 + rb * rec_buf, p.recs + tile * (int64_t)rec_main, rec_main, recf_bar(rb));
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    if (warp == kWgtWarp0 + kWgtWarps) {
      // ===================== list producer: {nU} + the first kRcap unique rows of each tile, kUlSlots tiles ahead =====
      uint32_t it = 0;
      const uint32_t ul0 = smem_u32(s_ul);
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t ub = it % kUlSlots;
        const uint32_t rows = (min(uni((uint32_t)__ldg(&p.meta[tile].y)), (uint32_t)kRcap) * 4u + 15u) & ~15u;
        WAIT_RELAXED(0, ule_bar(ub), ((it / kUlSlots) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(ulf_bar(ub), 16 + rows);
          bulk_g2s(ul0 + ub * kUlBytes, p.recs + tile * (int64_t)rec_main + 16 * p.K, 16, ulf_bar(ub));  // header: nU first
          if (rows) bulk_g2s(ul0 + ub * kUlBytes + 16, p.uidx + tile * (int64_t)(kTileM * p.K), rows, ulf_bar(ub));
        }
        __syncwarp();
      }
    } else if (!p.resident && warp < kWgtWarp0 + kWgtWarps) {
    // ===================== weight producers: bulk copy of each unit's pre-swizzled weight block into its stage =====
    const uint32_t wi = uni((uint32_t)(warp - kWgtWarp0));
    if (wi < (uint32_t)p.nwp) {
      uint32_t Q = 0;
      const uint32_t w_base = smem_u32(s_w);
      const uint32_t nwp = (uint32_t)p.nwp;  // a power of two
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const uint32_t nact = uni((uint32_t)__ldg(&p.meta[tile].z));  // packs of the tile
        // lane l: members of pack l (the record's header, read from global memory: the producers run ahead of the ring)
        const uint32_t kl = __ldg(reinterpret_cast<const uint16_t *>(p.recs + tile * (int64_t)rec_main + 16 * p.K + 16) + lane);
        const uint32_t nq = (nact + us - 1) / us;
        for (int kb = 0; kb < KB; ++kb) {
          for (uint32_t i = (wi - Q) & (nwp - 1); i < nq; i += nwp) {
            const uint32_t Qi = Q + i, sw = Qi & (na - 1);
            const uint32_t nu = min(us, nact - i * us);
            uint32_t mu[4];
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) mu[u] = __shfl_sync(0xffffffffu, kl, (int)((i * us + u) & 31u));
            WAIT(0, aempty_bar(sw), ((Qi >> p.lna) & 1) ^ 1);
            if (elect_one()) {
              uint32_t nblk = 0;
#pragma unroll
              for (uint32_t u = 0; u < 4; ++u)
                if (u < us && u < nu) nblk += ((mu[u] & 0xFFu) != 0xFFu) + ((mu[u] >> 8) != 0xFFu);
              if (nblk == 0) {
                mbar_arrive(afull_bar(sw));  // a tile without entries: the dummy unit has no weights
              } else {
                mbar_expect_tx(afull_bar(sw), nblk * w_stage);
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u)
                  if (u < us && u < nu) {
#pragma unroll
                    for (uint32_t m = 0; m < 2; ++m) {
                      const uint32_t k = (mu[u] >> (8 * m)) & 0xFFu;
                      if (k != 0xFFu)
                        bulk_g2s(w_base + ((sw * us + u) * 2 + m) * w_stage, p.packed + (size_t)(k * KB + kb) * w_stage,
                                 w_stage, afull_bar(sw));
                    }
                  }
              }
            }
            __syncwarp();
          }
          Q += nq;
        }
      }
    }
    }
  }

#undef TL
#undef WAIT
#undef WAIT_RELAXED
  if (DIAG && p.diag != nullptr && blockIdx.x == 0 && lane == 0) {
    p.diag[warp * 8] = clock64() - dg_t0;
    for (int i = 0; i < 4; ++i) p.diag[warp * 8 + 1 + i] = dg[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// W[K, Cin_w, Cout_w] fp32 -> [K][KB][2][Cout x 64 B swizzled] bf16.  NS = 2: KB = ceil(Cin / 32), block 0 = hi and
// block 1 = mid = bf16(w - hi) of channels [32 kb, 32 kb + 32).  NS = 1: KB = ceil(Cin / 64), block b = channels
// [64 kb + 32 b, 64 kb + 32 b + 32).  Channels beyond Cin are zero.
__global__ void pack_weights_kernel(const float *__restrict__ W, int K, int Cin, int Cout, int transpose_w, int NS,
                                    uint8_t *__restrict__ packed) {
  const int cpu = NS == 2 ? 32 : 64;
  const int KB = (Cin + cpu - 1) / cpu;
  const int64_t total = (int64_t)K * KB * Cout * cpu;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int kk = (int)(e % cpu);
    int n = (int)((e / cpu) % Cout);
    int kb = (int)((e / ((int64_t)cpu * Cout)) % KB);
    int k = (int)(e / ((int64_t)cpu * Cout * KB));
    int c = kb * cpu + kk;
    float w = c >= Cin ? 0.f : transpose_w ? W[((int64_t)k * Cout + n) * Cin + c] : W[((int64_t)k * Cin + c) * Cout + n];
    __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const size_t unit = ((size_t)k * KB + kb) * 2 * (size_t)Cout * 64;
    const int k32 = kk & 31;
    uint32_t off = sw64((uint32_t)n, (uint32_t)(k32 >> 3)) + (k32 & 7) * 2;
    if (NS == 2) {
      *reinterpret_cast<__nv_bfloat16 *>(packed + unit + off) = hi;
      __nv_bfloat16 mid = __float2bfloat16_rn(w - __bfloat162float(hi));
      *reinterpret_cast<__nv_bfloat16 *>(packed + unit + (size_t)Cout * 64 + off) = mid;
    } else {
      *reinterpret_cast<__nv_bfloat16 *>(packed + unit + (size_t)(kk >> 5) * Cout * 64 + off) = hi;
    }
  }
}

}  // namespace umma
}  // namespace wsis

using namespace wsis;
using namespace wsis::umma;

// Launch plan of a layer shape: issuers / TMEM accumulators and the largest pipeline (stages, row-cache buffers,
// record buffers) that fits the 512 TMEM columns and the 227 KB of dynamic shared memory.  Pure host arithmetic (also
// exported as wsis_conv_umma_plan so that the CPU test suite can check that every layer shape of the model has a plan).
static int plan_launch(wsis::umma::Params &p, int K, int Cin, int Cout, int NS, int64_t *smem_out) {
  using namespace wsis::umma;
  const int cpu = NS == 2 ? 32 : 64;
  p.KB = (Cin + cpu - 1) / cpu;
  // one thread issues a short MMA every ~39 cycles: N/2-cycle MMAs need 4 (N <= 32) or 2 (N <= 64) issuers
  p.nmma = Cout <= 32 ? 4 : Cout <= 64 ? 2 : 1;
  if (const char *e = getenv("WSIS_CONV_NMMA")) p.nmma = std::max(1, std::min(p.nmma, atoi(e)));
  p.nacc = p.nmma;
  // accumulators are double buffered when that leaves at least four operand slots
  p.nbuf = (2 * p.nacc * Cout + 4 * kSlotCols <= 512) ? 2 : 1;
  const int slots = (512 - p.nbuf * p.nacc * Cout) / kSlotCols;
  WSIS_CHECK(slots >= 2, "conv_umma: no TMEM left for operand slots at Cout=%d", Cout);
  const int64_t w_stage = 2 * (int64_t)Cout * 64, w_all = (int64_t)K * p.KB * w_stage;
  p.resident = w_all <= kResidentBytes;
  p.nwp = kWgtWarps;
  const int64_t rec_buf = rec_stride_bytes(K);
  // The ring holds at most `slots` operand slots; its throughput is slots / (round trip of a slot: TMEM store, hand-off,
  // MMA issue + execution, hand-off back), so all of them are used: stages x units per stage = slots (a power of two,
  // stages a multiple of the issuers and of the builder groups so that every stage has one owner of each kind).
  // One unit per stage unless WSIS_CONV_US says otherwise: measured best (profiles/README.md).
  int us_want = 1;
  if (const char *e = getenv("WSIS_CONV_US")) us_want = std::max(1, std::min(4, atoi(e)));
  int pool = slots >= 8 ? 8 : slots >= 4 ? 4 : 2;
  static const int pref[][2] = {{3, 3}, {2, 3}, {3, 2}, {2, 2}, {1, 2}, {1, 1}};  // {row-cache buffers, record buffers}
  const int64_t budget = 227 * 1024;
  int64_t smem = 0;
  bool fit = false;
  for (; pool >= 2 && !fit; pool >>= 1) {
    int us_ = std::min(us_want, pool / std::max(p.nmma, kBuildGroups));
    if (us_ < 1) continue;  // fewer slots than issuers / builder groups
    if (us_ == 3) us_ = 2;
    const int na = pool / us_;
    for (auto &c : pref) {
      const int64_t misc = 1024 /*align*/ + 2 * p.KB * cpu * 4 + (2 * na + 3 * c[0] + 2 * c[1] + 5 + 2 * kUlSlots) * 8 + 64 +
                           kUlSlots * kUlBytes;
      smem = misc + (p.resident ? w_all : (int64_t)pool * 2 * w_stage) + c[0] * (int64_t)kRcBuf + c[1] * rec_buf;
      if (smem <= budget) {
        p.us = us_;
        p.lna = na == 8 ? 3 : na == 4 ? 2 : na == 2 ? 1 : 0;
        p.nrc = c[0];
        p.nrec = c[1];
        fit = true;
        break;
      }
    }
  }
  WSIS_CHECK(fit, "conv_umma: shared memory budget exceeded for Cin=%d Cout=%d K=%d", Cin, Cout, K);
  *smem_out = smem;
  return 0;
}

static long long *g_diag = nullptr;

extern "C" {

int wsis_conv_debug_stats(void *buf) {
  g_diag = reinterpret_cast<long long *>(buf);
  return 0;
}

int wsis_conv_umma_supported(int Cin, int Cout) {
  return Cin >= 1 && Cin <= 1024 && Cout >= 16 && Cout % 16 == 0 && Cout <= 256;
}

int64_t wsis_conv_pack_bytes(int K, int Cin, int Cout, int precision) {
  const int cpu = precision == 3 ? 32 : 64;
  return (int64_t)K * ((Cin + cpu - 1) / cpu) * 2 * Cout * 64;
}

int wsis_conv_pack_weights(const float *W, int K, int Cin, int Cout, int transpose_w, int precision, void *packed,
                           wsis_stream_t stream) {
  WSIS_CHECK(wsis_conv_umma_supported(Cin, Cout), "pack_weights: unsupported Cin=%d Cout=%d", Cin, Cout);
  WSIS_CHECK(precision == 1 || precision == 3, "pack_weights: precision must be 1 or 3");
  WSIS_CHECK((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "pack_weights: packed must be 16-byte aligned");
  const int cpu = precision == 3 ? 32 : 64;
  int64_t total = (int64_t)K * ((Cin + cpu - 1) / cpu) * cpu * Cout;
  if (precision == 1)  // the unused half block of an odd 32-channel tail is read by no MMA, but keep the image defined
    WSIS_CUDA(cudaMemsetAsync(packed, 0, (size_t)wsis_conv_pack_bytes(K, Cin, Cout, precision), as_stream(stream)));
  unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 8);
  pack_weights_kernel<<<blocks, 256, 0, as_stream(stream)>>>(W, K, Cin, Cout, transpose_w, precision == 3 ? 2 : 1,
                                                             (uint8_t *)packed);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_conv_umma_plan(int K, int Cin, int Cout, int precision, int32_t *plan) {
  WSIS_CHECK(wsis_conv_umma_supported(Cin, Cout), "conv_umma_plan: unsupported Cin=%d Cout=%d", Cin, Cout);
  WSIS_CHECK(K >= 1 && K <= 32, "conv_umma_plan: kernel volume %d not in [1,32]", K);
  WSIS_CHECK(precision == 1 || precision == 3, "conv_umma_plan: precision must be 1 or 3");
  Params p;
  int64_t smem = 0;
  if (plan_launch(p, K, Cin, Cout, precision == 3 ? 2 : 1, &smem)) return 1;
  const int out[11] = {(int)smem, 1 << p.lna, p.nrc, p.nrec, kBuildGroups, p.nacc, p.nmma, p.nbuf, p.resident, p.nwp, p.us};
  for (int i = 0; i < 11; ++i) plan[i] = out[i];
  return 0;
}

int wsis_conv_umma(const float *src, const void *records, const int32_t *uidx, const int32_t *meta,
                   const int32_t *order, int64_t num_tiles, int K, const void *packed, int Cin, int Cout, int precision,
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
  p.in_relu = in_relu;
  p.vec4 = (Cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  p.num_tiles = num_tiles;
  int64_t smem = 0;
  {
    static bool trap_set = false;           // diagnostics only: one record per process
    if (!trap_set) {
      unsigned int *tw = trap_word_device();
      cudaMemcpyToSymbol(wsis::umma::d_trap_word, &tw, sizeof(tw));
      trap_set = true;
    }
  }
  if (plan_launch(p, K, Cin, Cout, NS, &smem)) return 1;
  p.diag = g_diag;
  {
    const char *e = getenv("WSIS_CONV_WAIT_NS");
    p.wait_ns = e ? atoi(e) : 0;
    e = getenv("WSIS_CONV_DEBUG");
    p.dbg = e ? atoi(e) : 0;
  }
  void (*kern)(const Params) = nullptr;
  const bool diag = g_diag != nullptr || p.dbg != 0;  // WSIS_CONV_DEBUG experiments run on the diagnostics build
#define WSIS_PICK(NS_, US_) (diag ? conv_umma_kernel<NS_, US_, true> : conv_umma_kernel<NS_, US_, false>)
  if (NS == 2)
    kern = p.us == 1 ? WSIS_PICK(2, 1) : p.us == 2 ? WSIS_PICK(2, 2) : WSIS_PICK(2, 4);
  else
    kern = p.us == 1 ? WSIS_PICK(1, 1) : p.us == 2 ? WSIS_PICK(1, 2) : WSIS_PICK(1, 4);
#undef WSIS_PICK
  // the opt-in is per device: set it on every launch rather than caching it process-wide (multi-GPU processes)
  WSIS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  unsigned grid = (unsigned)std::min<int64_t>(p.num_tiles, sm_count());
  kern<<<grid, kThreads, smem, as_stream(stream)>>>(p);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
