// Training-step kernels around the sparse convs (BASELINE.json configs[2]; reference: train_scannetv2.py:200-252,
// 734-738): batch-statistics BatchNorm whose normalisation rides in the conv's gather prologue (only the column
// reductions and the backward elementwise pass are separate kernels), and one AdamW launch over the flat parameter
// buffer with the ECC gradient clamp (train_scannetv2.py:246-249) and the 1/world gradient average folded in.
//
// BatchNorm1d (train) over x f32[N, C] (the reference: torch.nn.BatchNorm1d / SyncBatchNorm, sparse_unet3d.py:135-171):
//   forward   sums[c] = (sum x, sum x^2), count          -> [all-reduce across ranks when the statistics are synced]
//             finalize: mean, invstd, scale = gamma*invstd, shift = beta - mean*scale, running stats
//             y = relu(x*scale + shift) is NOT materialised: the next conv applies it while gathering rows
//   backward  da = gradient w.r.t. y (the conv's dgrad);  dy = da * [x*scale+shift > 0]
//             sums[c] = (sum dy, sum dy*xhat)            -> [all-reduce]
//             dx = scale * (dy - mean(dy) - xhat * mean(dy*xhat)),  dgamma = sum dy*xhat (local), dbeta = sum dy (local)
// All reductions are two-stage with a fixed order (block partials in fp32, combined in fp64): deterministic.
#include <math.h>

#include "common.cuh"

namespace wsis {

constexpr int kBnThreads = 256;

// mode 0: (x, x^2); mode 1: (dy, dy*xhat) with dy = da * relu-mask
template <int MODE>
__global__ void __launch_bounds__(kBnThreads)
bn_colsum_kernel(const float *__restrict__ x, const float *__restrict__ da, int64_t N, int C,
                 const float *__restrict__ stat /* mean | invstd | scale | shift, [4][C] */, int relu,
                 float *__restrict__ partial /* [gridDim.x][3C + 1]: sum1[C] | sum2[C] | pivot[C] | rows */) {
  extern __shared__ float sm[];  // [rows_per_pass][2][C]
  const int tpr = (C + 3) / 4;                  // threads per row (4 channels each)
  const int rpp = kBnThreads / tpr;             // rows per pass
  const int tr = threadIdx.x / tpr, tq = threadIdx.x % tpr;
  const int c0 = tq * 4;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float mean[4], invstd[4], scale[4], shift[4];
  const bool active = tr < rpp;
  if (MODE == 1 && active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = min(c0 + j, C - 1);
      mean[j] = stat[c], invstd[j] = stat[C + c], scale[j] = stat[2 * C + c], shift[j] = stat[3 * C + c];
    }
  }
  const bool vec = (C % 4 == 0);
  // MODE 0 accumulates around a per-block pivot (the block's first row) so that sum (x-p)^2 does not cancel in fp32
  // when |mean| >> std; the combine kernel undoes the shift in fp64
  float piv[4] = {0.f, 0.f, 0.f, 0.f};
  if (MODE == 0 && active) {
    int64_t pr = (int64_t)blockIdx.x * rpp;
    if (pr < N) {
#pragma unroll
      for (int j = 0; j < 4; ++j) piv[j] = c0 + j < C ? x[pr * C + c0 + j] : 0.f;
    }
  }
  if (active) {
    for (int64_t r = (int64_t)blockIdx.x * rpp + tr; r < N; r += (int64_t)gridDim.x * rpp) {
      float xv[4], gv[4];
      if (vec) {
        float4 t = *reinterpret_cast<const float4 *>(x + r * C + c0);
        xv[0] = t.x, xv[1] = t.y, xv[2] = t.z, xv[3] = t.w;
        if (MODE == 1) {
          float4 g = *reinterpret_cast<const float4 *>(da + r * C + c0);
          gv[0] = g.x, gv[1] = g.y, gv[2] = g.z, gv[3] = g.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xv[j] = c0 + j < C ? x[r * C + c0 + j] : 0.f;
          if (MODE == 1) gv[j] = c0 + j < C ? da[r * C + c0 + j] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE == 0) {
          float d = xv[j] - piv[j];
          s1[j] += d;
          s2[j] = fmaf(d, d, s2[j]);
        } else {
          float dy = (relu && fmaf(xv[j], scale[j], shift[j]) <= 0.f) ? 0.f : gv[j];
          s1[j] += dy;
          s2[j] = fmaf(dy, (xv[j] - mean[j]) * invstd[j], s2[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c0 + j < C) {
        sm[(tr * 2 + 0) * C + c0 + j] = s1[j];
        sm[(tr * 2 + 1) * C + c0 + j] = s2[j];
      }
  }
  __syncthreads();
  float *pout = partial + (int64_t)blockIdx.x * (3 * C + 1);
  for (int i = threadIdx.x; i < 2 * C; i += kBnThreads) {
    float acc = 0.f;
    for (int r = 0; r < rpp; ++r) acc += sm[r * 2 * C + i];
    pout[i] = acc;
  }
  if (MODE == 0) {
    const int64_t r0 = (int64_t)blockIdx.x * rpp;
    for (int c = threadIdx.x; c < C; c += kBnThreads) pout[2 * C + c] = r0 < N ? x[r0 * C + c] : 0.f;
    if (threadIdx.x == 0) {                  // rows r0 + j + k * gridDim.x * rpp < N, j < rpp (< 2^24: exact in fp32)
      int64_t nb = 0;
      for (int j = 0; j < rpp && r0 + j < N; ++j) nb += (N - (r0 + j) - 1) / ((int64_t)gridDim.x * rpp) + 1;
      pout[3 * C] = (float)nb;
    }
  }
}

// sums[0..2C) = sum over blocks of the partials (fp64, fixed order: one warp per column, lanes stride over the blocks,
// butterfly reduction); sums[2C] = count (forward only); the backward form also writes the LOCAL parameter gradients
// dgamma = sum dy*xhat, dbeta = sum dy.
// `x` != NULL (forward statistics): block b accumulated around the pivot p_b = x[b*rpp, c] over n_b rows, so
//   sum x = sum_b (s1_b + n_b p_b),  sum x^2 = sum_b (s2_b + 2 p_b s1_b + n_b p_b^2).
__global__ void __launch_bounds__(256)
bn_combine_kernel(const float *__restrict__ partial, int nblocks, int C, double count, int write_count,
                  double *__restrict__ sums, float *__restrict__ dgamma, float *__restrict__ dbeta,
                  const float *__restrict__ x, int64_t N, int rpp) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // column of [sum1 | sum2]
  if (i < 2 * C) {
    const int c = i < C ? i : i - C;
    double acc = 0.0;
    for (int b = lane; b < nblocks; b += 32) {
      const float *pb = partial + (int64_t)b * (3 * C + 1);
      if (x) {
        const double nb = (double)pb[3 * C], p = (double)pb[2 * C + c], s1 = (double)pb[c], s2 = (double)pb[C + c];
        acc += i < C ? s1 + nb * p : s2 + 2.0 * p * s1 + nb * p * p;
      } else {
        acc += (double)pb[i];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      sums[i] = acc;
      if (dbeta && i < C) dbeta[i] = (float)acc;
      if (dgamma && i >= C) dgamma[i - C] = (float)acc;
    }
  }
  if (write_count && blockIdx.x == 0 && threadIdx.x == 0) sums[2 * C] = count;
}

__global__ void bn_finalize_kernel(const double *__restrict__ sums, int C, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float eps, float momentum,
                                   float *__restrict__ running_mean, float *__restrict__ running_var,
                                   float *__restrict__ stat) {
  const double n = sums[2 * C];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double mean = sums[c] / n;
    double var = sums[C + c] / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float scale = g * invstd;
    stat[c] = (float)mean;
    stat[C + c] = invstd;
    stat[2 * C + c] = scale;
    stat[3 * C + c] = b - (float)mean * scale;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
      double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

// thread (row slot, channel quad) like bn_colsum_kernel: the per-channel constants are computed once per thread
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const float *__restrict__ x, const float *__restrict__ da, int64_t N, int C,
                    const float *__restrict__ stat, int relu, const double *__restrict__ sums,
                    const double *__restrict__ count, const float *__restrict__ extra, float *__restrict__ dx) {
  const int tpr = (C + 3) / 4, rpp = kBnThreads / tpr;
  const int tr = threadIdx.x / tpr, c0 = (threadIdx.x % tpr) * 4;
  if (tr >= rpp) return;
  const double n = *count;
  float mean[4], invstd[4], scale[4], shift[4], m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = min(c0 + j, C - 1);
    mean[j] = stat[c], invstd[j] = stat[C + c], scale[j] = stat[2 * C + c], shift[j] = stat[3 * C + c];
    m1[j] = (float)(sums[c] / n), m2[j] = (float)(sums[C + c] / n);
  }
  const bool vec = (C % 4 == 0);
  for (int64_t r = (int64_t)blockIdx.x * rpp + tr; r < N; r += (int64_t)gridDim.x * rpp) {
    float xv[4], gv[4], ev[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
    if (vec) {
      const float4 t = *reinterpret_cast<const float4 *>(x + r * C + c0), g = *reinterpret_cast<const float4 *>(da + r * C + c0);
      xv[0] = t.x, xv[1] = t.y, xv[2] = t.z, xv[3] = t.w, gv[0] = g.x, gv[1] = g.y, gv[2] = g.z, gv[3] = g.w;
      if (extra) {
        const float4 e = *reinterpret_cast<const float4 *>(extra + r * C + c0);
        ev[0] = e.x, ev[1] = e.y, ev[2] = e.z, ev[3] = e.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool in = c0 + j < C;
        xv[j] = in ? x[r * C + c0 + j] : 0.f;
        gv[j] = in ? da[r * C + c0 + j] : 0.f;
        if (extra && in) ev[j] = extra[r * C + c0 + j];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dy = (relu && fmaf(xv[j], scale[j], shift[j]) <= 0.f) ? 0.f : gv[j];
      const float xhat = (xv[j] - mean[j]) * invstd[j];
      o[j] = scale[j] * (dy - m1[j] - xhat * m2[j]) + ev[j];
    }
    if (vec) {
      *reinterpret_cast<float4 *>(dx + r * C + c0) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + j < C) dx[r * C + c0 + j] = o[j];
    }
  }
}

// torch.optim.AdamW (decoupled weight decay, bias-corrected), one launch over the flat buffers.
__global__ void __launch_bounds__(256)
adamw_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
             int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt,
             float grad_scale, int64_t clamp_begin, int64_t clamp_end) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    if (i >= clamp_begin && i < clamp_end) gi = fminf(fmaxf(gi, -1.f), 1.f);
    float pi = p[i] * (1.f - lr * weight_decay);
    float mi = beta1 * m[i] + (1.f - beta1) * gi;
    float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Synchronised statistics: combine + all-reduce over NVLink peer memory + finalize in ONE kernel (one block).
//
// torch.nn.SyncBatchNorm (train_scannetv2.py:736) all-reduces 2C+1 numbers per BatchNorm and direction: with NCCL that
// is ~110 BatchNorms x 2 latency-bound collectives of a few hundred bytes per step.  Here every rank owns a SYMMETRIC
// buffer (torch symmetric memory: the same allocation mapped into every peer), laid out as
//     double data[2][world][slot]  |  uint32 flag[2][world]            (2 = parity of the call's sequence number)
// and the block that has just combined the local column sums (a) stores them into slot [parity][rank] of EVERY peer's
// buffer with plain NVLink stores, (b) fences and release-stores the call's sequence number into flag [parity][rank] of
// every peer, (c) acquire-spins until its own flags hold the sequence number for every rank, (d) adds the world's
// contributions in rank order (bitwise the same result on every rank) and goes on to the finalize step.  Two parities
// are enough: a peer cannot start call k+2 before this rank has contributed to k+1, i.e. finished reading call k.
// A wait that does not complete within ~2 s is a protocol error (ranks out of step): trap instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------------------------
struct PeerComm {
  const unsigned long long *peers;   // device array [world] of the peers' buffer base addresses (own included)
  int world, rank, slot;             // slot = doubles per contribution (>= 2C+1)
  unsigned int seq;                  // 1, 2, 3, ... identical on every rank
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// s_sums[0..n) (shared) holds this rank's contribution on entry and the world's sum on exit
__device__ void peer_allreduce(double *s_sums, int n, const PeerComm &pc) {
  if (pc.world <= 1) return;
  const int par = (int)(pc.seq & 1u);
  const size_t flag_off = (size_t)2 * pc.world * pc.slot;           // in doubles
  for (int r = 0; r < pc.world; ++r) {
    double *dst = reinterpret_cast<double *>(pc.peers[r]) + ((size_t)par * pc.world + pc.rank) * pc.slot;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = s_sums[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < pc.world) {
    unsigned int *pf = reinterpret_cast<unsigned int *>(reinterpret_cast<double *>(pc.peers[threadIdx.x]) + flag_off) +
                       par * pc.world + pc.rank;
    st_release_sys(pf, pc.seq);
    const unsigned int *mf = reinterpret_cast<const unsigned int *>(reinterpret_cast<double *>(pc.peers[pc.rank]) + flag_off) +
                             par * pc.world + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(mf) != pc.seq) {
      if (clock64() - t0 > 4000000000ll) __trap();
    }
    __threadfence_system();
  }
  __syncthreads();
  const double *mine = reinterpret_cast<const double *>(pc.peers[pc.rank]) + (size_t)par * pc.world * pc.slot;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < pc.world; ++r) acc += __ldcg(mine + (size_t)r * pc.slot + i);
    s_sums[i] = acc;
  }
  __syncthreads();
}

// one block of 1024 threads: local combine (warp per column) -> peer all-reduce -> forward: finalize / backward: sums out
__global__ void __launch_bounds__(1024)
bn_sync_kernel(const float *__restrict__ partial, int nblocks, int C, double count, int forward, PeerComm pc,
               double *__restrict__ sums_out, float *__restrict__ dgamma, float *__restrict__ dbeta,
               const float *__restrict__ gamma, const float *__restrict__ beta, float eps, float momentum,
               float *__restrict__ running_mean, float *__restrict__ running_var, float *__restrict__ stat) {
  extern __shared__ double s_sums[];          // [2C+1] sums, then [slices][2C] scratch
  // thread (slice, column): consecutive threads read consecutive columns of one partial row (coalesced); a column's
  // blocks are split over `slices` threads and added in slice order (fixed order: deterministic)
  const int ncol = 2 * C, slices = max(1, (int)blockDim.x / ncol);
  double *s_part = s_sums + (2 * C + 2);
  const int i = threadIdx.x % ncol, sl = threadIdx.x / ncol;
  if (sl < slices) {
    const int c = i < C ? i : i - C;
    double acc = 0.0;
    for (int b = sl; b < nblocks; b += slices) {
      const float *pb = partial + (int64_t)b * (3 * C + 1);
      if (forward) {
        const double nb = (double)pb[3 * C], p = (double)pb[2 * C + c], s1 = (double)pb[c], s2 = (double)pb[C + c];
        acc += i < C ? s1 + nb * p : s2 + 2.0 * p * s1 + nb * p * p;
      } else {
        acc += (double)pb[i];
      }
    }
    s_part[sl * ncol + i] = acc;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < ncol; j += blockDim.x) {
    double acc = 0.0;
    for (int s2 = 0; s2 < slices; ++s2) acc += s_part[s2 * ncol + j];
    s_sums[j] = acc;
    if (!forward) {                           // parameter gradients are the LOCAL sums (the gradient bucket is all-reduced later)
      if (dbeta && j < C) dbeta[j] = (float)acc;
      if (dgamma && j >= C) dgamma[j - C] = (float)acc;
    }
  }
  if (threadIdx.x == 0 && forward) s_sums[2 * C] = count;
  __syncthreads();
  const int n = forward ? 2 * C + 1 : 2 * C;
  peer_allreduce(s_sums, n, pc);
  for (int i = threadIdx.x; i < n; i += blockDim.x) sums_out[i] = s_sums[i];
  if (forward && stat) {
    const double tot = s_sums[2 * C];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double mean = s_sums[c] / tot;
      double var = s_sums[C + c] / tot - mean * mean;
      var = var < 0.0 ? 0.0 : var;
      float invstd = (float)(1.0 / sqrt(var + (double)eps));
      float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
      float scale = g * invstd;
      stat[c] = (float)mean;
      stat[C + c] = invstd;
      stat[2 * C + c] = scale;
      stat[3 * C + c] = b - (float)mean * scale;
      if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      if (running_var) {
        double unbiased = tot > 1.0 ? var * tot / (tot - 1.0) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
  }
}

static int bn_grid(int64_t N, int C, int *rpp_out) {
  int tpr = (C + 3) / 4;
  int rpp = kBnThreads / tpr;
  *rpp_out = rpp;
  int64_t blocks = std::min<int64_t>(std::min<int64_t>(ceil_div(N, (int64_t)rpp * 4), (int64_t)sm_count() * 2), 1024);
  return (int)std::max<int64_t>(blocks, 1);
}

}  // namespace wsis

using namespace wsis;

extern "C" {

int64_t wsis_bn_ws_bytes(int64_t N, int C) {
  int rpp;
  return (int64_t)bn_grid(N, C, &rpp) * (3 * C + 1) * sizeof(float);
}

int wsis_bn_stats(const float *x, int64_t N, int C, void *ws, double *sums, wsis_stream_t stream) {
  WSIS_CHECK(C >= 1 && C <= 1024, "bn_stats: 1 <= C <= 1024");
  cudaStream_t st = as_stream(stream);
  int rpp, blocks = bn_grid(N, C, &rpp);
  bn_colsum_kernel<0><<<blocks, kBnThreads, sizeof(float) * rpp * 2 * C, st>>>(x, nullptr, N, C, nullptr, 0, (float *)ws);
  WSIS_LAUNCH_OK();
  bn_combine_kernel<<<(2 * C + 7) / 8, 256, 0, st>>>((const float *)ws, blocks, C, (double)N, 1, sums, nullptr, nullptr, x, N, rpp);
  WSIS_LAUNCH_OK();
  return 0;
}

// Forward statistics + (world > 1: all-reduce over peer memory) + finalize as partial kernel + ONE block.
// peers: device uint64[world] of the symmetric buffers' base addresses (NULL when world == 1); seq counts the
// synchronised calls (same on all ranks, starts at 1); slot = doubles per contribution (>= 2C+1).
int wsis_bn_forward_sync(const float *x, int64_t N, int C, void *ws, double *sums, const float *gamma, const float *beta,
                         float eps, float momentum, float *running_mean, float *running_var, float *stat,
                         const void *peers, int world, int rank, int64_t seq, int slot, wsis_stream_t stream) {
  WSIS_CHECK(C >= 1 && C <= 512, "bn_forward_sync: 1 <= C <= 512");
  WSIS_CHECK(world == 1 || (peers != nullptr && slot >= 2 * C + 1 && seq >= 1), "bn_forward_sync: bad peer arguments");
  cudaStream_t st = as_stream(stream);
  int rpp, blocks = bn_grid(N, C, &rpp);
  bn_colsum_kernel<0><<<blocks, kBnThreads, sizeof(float) * rpp * 2 * C, st>>>(x, nullptr, N, C, nullptr, 0, (float *)ws);
  WSIS_LAUNCH_OK();
  PeerComm pc{(const unsigned long long *)peers, world, rank, slot, (unsigned int)seq};
  bn_sync_kernel<<<1, 1024, sizeof(double) * (2 * C + 2 + 1024 + 2 * C), st>>>((const float *)ws, blocks, C, (double)N, 1, pc, sums, nullptr,
                                                                nullptr, gamma, beta, eps, momentum, running_mean,
                                                                running_var, stat);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_bn_bwd_reduce_sync(const float *x, const float *da, int64_t N, int C, const float *stat, int relu, void *ws,
                            double *sums, float *dgamma, float *dbeta, const void *peers, int world, int rank,
                            int64_t seq, int slot, wsis_stream_t stream) {
  WSIS_CHECK(C >= 1 && C <= 512, "bn_bwd_reduce_sync: 1 <= C <= 512");
  WSIS_CHECK(world == 1 || (peers != nullptr && slot >= 2 * C + 1 && seq >= 1), "bn_bwd_reduce_sync: bad peer arguments");
  cudaStream_t st = as_stream(stream);
  int rpp, blocks = bn_grid(N, C, &rpp);
  bn_colsum_kernel<1><<<blocks, kBnThreads, sizeof(float) * rpp * 2 * C, st>>>(x, da, N, C, stat, relu, (float *)ws);
  WSIS_LAUNCH_OK();
  PeerComm pc{(const unsigned long long *)peers, world, rank, slot, (unsigned int)seq};
  bn_sync_kernel<<<1, 1024, sizeof(double) * (2 * C + 2 + 1024 + 2 * C), st>>>((const float *)ws, blocks, C, 0.0, 0, pc, sums, dgamma, dbeta,
                                                                nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, nullptr);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_bn_finalize(const double *sums, int C, const float *gamma, const float *beta, float eps, float momentum,
                     float *running_mean, float *running_var, float *stat, wsis_stream_t stream) {
  bn_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(sums, C, gamma, beta, eps, momentum, running_mean, running_var, stat);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_bn_bwd_reduce(const float *x, const float *da, int64_t N, int C, const float *stat, int relu, void *ws,
                       double *sums, float *dgamma, float *dbeta, wsis_stream_t stream) {
  WSIS_CHECK(C >= 1 && C <= 1024, "bn_bwd_reduce: 1 <= C <= 1024");
  cudaStream_t st = as_stream(stream);
  int rpp, blocks = bn_grid(N, C, &rpp);
  bn_colsum_kernel<1><<<blocks, kBnThreads, sizeof(float) * rpp * 2 * C, st>>>(x, da, N, C, stat, relu, (float *)ws);
  WSIS_LAUNCH_OK();
  bn_combine_kernel<<<(2 * C + 7) / 8, 256, 0, st>>>((const float *)ws, blocks, C, 0.0, 0, sums, dgamma, dbeta, nullptr, 0, 0);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_bn_bwd_apply(const float *x, const float *da, int64_t N, int C, const float *stat, int relu,
                      const double *sums, const double *count, const float *extra, float *dx, wsis_stream_t stream) {
  if (N == 0) return 0;
  int rpp;
  bn_grid(N, C, &rpp);
  unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(N, rpp), (int64_t)sm_count() * 8));
  bn_bwd_apply_kernel<<<blocks, kBnThreads, 0, as_stream(stream)>>>(x, da, N, C, stat, relu, sums, count, extra, dx);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_adamw_step(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int64_t step, float grad_scale, int64_t clamp_begin,
                    int64_t clamp_end, wsis_stream_t stream) {
  if (n == 0) return 0;
  WSIS_CHECK(step >= 1, "adamw: step counts from 1");
  float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 8);
  adamw_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt,
                                                      grad_scale, clamp_begin, clamp_end);
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Point-level semantic loss (losses_3D_WSIS.py:52-64): cross entropy (mean over labelled points) + multi-class dice
// on the softmax of the labelled points, forward and backward in two passes over the [N, C] scores instead of ~25
// elementwise / reduction launches that each stream [N, C].
//   fwd: per point lse, p = softmax; block partials of {ce, n, I_c = sum p_c [y=c], Q_c = sum p_c^2, T_c = #[y=c]}
//   fin: loss = ce/n + mean_c (1 - (2 I_c + eps)/(Q_c + T_c + 1e-4 + eps)); coefficients a_c, b_c of dL/dp_c = a_c [y=c] + b_c p_c
//   bwd: dlogit_k = g * { (p_k - [k=y])/n + p_k (gp_k - sum_j p_j gp_j) }, gp_j = a_j [y=j] + b_j p_j   (0 for ignored points)
// ------------------------------------------------------------------------------------------------------------------
namespace wsis {

constexpr int kLossMaxC = 32;
constexpr int kLossThreads = 256;

template <int CP>   // CP = classes padded to a multiple of 4
__device__ __forceinline__ void load_logits(const float *__restrict__ row, int C, float (&l)[CP]) {
  if (C == CP) {
#pragma unroll
    for (int c4 = 0; c4 < CP / 4; ++c4) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + c4);
      l[c4 * 4] = v.x, l[c4 * 4 + 1] = v.y, l[c4 * 4 + 2] = v.z, l[c4 * 4 + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < CP; ++c) l[c] = c < C ? __ldg(row + c) : -INFINITY;
  }
}

template <int CP>
__global__ void __launch_bounds__(kLossThreads)
ce_dice_fwd_kernel(const float *__restrict__ scores, const int64_t *__restrict__ labels, int64_t N, int C, int ignore,
                   float *__restrict__ partial /* [grid][2 + 3 CP] */) {
  float ce = 0.f, cnt = 0.f, I[CP], Q[CP], T[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) I[c] = Q[c] = T[c] = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)__ldg(labels + i);
    if (y == ignore) continue;
    float l[CP];
    load_logits<CP>(scores + i * C, C, l);
    float mx = l[0];
#pragma unroll
    for (int c = 1; c < CP; ++c) mx = fmaxf(mx, l[c]);
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      l[c] = __expf(l[c] - mx);
      se += l[c];
    }
    const float inv = 1.f / se;
    cnt += 1.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      const float p = l[c] * inv;
      Q[c] = fmaf(p, p, Q[c]);
      if (c == y) {
        I[c] += p;
        T[c] += 1.f;
        ce -= __logf(p);
      }
    }
  }
  __shared__ float s_red[kLossThreads / 32][2 + 3 * CP];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  auto red = [&](float v, int slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[w][slot] = v;
  };
  red(ce, 0);
  red(cnt, 1);
#pragma unroll
  for (int c = 0; c < CP; ++c) {
    red(I[c], 2 + c);
    red(Q[c], 2 + CP + c);
    red(T[c], 2 + 2 * CP + c);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < 2 + 3 * CP; s += blockDim.x) {
    float acc = 0.f;
    for (int ww = 0; ww < kLossThreads / 32; ++ww) acc += s_red[ww][s];
    partial[(int64_t)blockIdx.x * (2 + 3 * CP) + s] = acc;
  }
}

// out[0] = loss, out[1] = n, out[2 + c] = a_c, out[2 + CP + c] = b_c (floats)
__global__ void ce_dice_finalize_kernel(const float *__restrict__ partial, int nblocks, int C, int CP, int dice,
                                        float *__restrict__ out) {
  __shared__ double s[2 + 3 * kLossMaxC];
  const int slots = 2 + 3 * CP;
  for (int i = threadIdx.x; i < slots; i += blockDim.x) {
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += (double)partial[(int64_t)b * slots + i];
    s[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = s[1];
    double loss = n > 0 ? s[0] / n : 0.0 / 0.0;            // CrossEntropyLoss over zero labelled points is NaN in torch too
    if (dice) {
      double d = 0.0;
      for (int c = 0; c < C; ++c) {
        const double I = s[2 + c], Q = s[2 + CP + c], T = s[2 + 2 * CP + c];
        const double num = 2.0 * I + 1e-5, den = Q + T + 1e-4 + 1e-5;
        d += 1.0 - num / den;
        out[2 + c] = (float)(-2.0 / (C * den));
        out[2 + CP + c] = (float)(2.0 * num / (C * den * den));
      }
      loss += d / C;
    } else {
      for (int c = 0; c < CP; ++c) out[2 + c] = out[2 + CP + c] = 0.f;
    }
    out[0] = (float)loss;
    out[1] = (float)n;
  }
}

template <int CP>
__global__ void __launch_bounds__(kLossThreads)
ce_dice_bwd_kernel(const float *__restrict__ scores, const int64_t *__restrict__ labels, int64_t N, int C, int ignore,
                   const float *__restrict__ fin, const float *__restrict__ gout, float *__restrict__ dscores) {
  __shared__ float s_a[CP], s_b[CP];
  if (threadIdx.x < CP) {
    s_a[threadIdx.x] = fin[2 + threadIdx.x];
    s_b[threadIdx.x] = fin[2 + CP + threadIdx.x];
  }
  __syncthreads();
  const float g = gout ? __ldg(gout) : 1.f, invn = 1.f / fin[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)__ldg(labels + i);
    float l[CP];
    if (y == ignore) {
#pragma unroll
      for (int c = 0; c < CP; ++c) l[c] = 0.f;
    } else {
      load_logits<CP>(scores + i * C, C, l);
      float mx = l[0];
#pragma unroll
      for (int c = 1; c < CP; ++c) mx = fmaxf(mx, l[c]);
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        l[c] = __expf(l[c] - mx);
        se += l[c];
      }
      const float inv = 1.f / se;
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        l[c] *= inv;                                          // p_c
        dot = fmaf(l[c], (c == y ? s_a[c] : 0.f) + s_b[c] * l[c], dot);
      }
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const float gp = (c == y ? s_a[c] : 0.f) + s_b[c] * l[c];
        l[c] = g * ((l[c] - (c == y ? 1.f : 0.f)) * invn + l[c] * (gp - dot));
      }
    }
    float *o = dscores + i * C;
    if (C == CP) {
#pragma unroll
      for (int c4 = 0; c4 < CP / 4; ++c4)
        reinterpret_cast<float4 *>(o)[c4] = make_float4(l[c4 * 4], l[c4 * 4 + 1], l[c4 * 4 + 2], l[c4 * 4 + 3]);
    } else {
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (c < C) o[c] = l[c];
    }
  }
}

static int loss_grid(int64_t N) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(N, kLossThreads), (int64_t)sm_count() * 4));
}

}  // namespace wsis

extern "C" {

int64_t wsis_ce_dice_ws_bytes(int64_t N, int C) {
  int cp = (C + 3) / 4 * 4;
  return (int64_t)wsis::loss_grid(N) * (2 + 3 * cp) * sizeof(float);
}

int wsis_ce_dice_fwd(const float *scores, const int64_t *labels, int64_t N, int C, int ignore_label, int dice, void *ws,
                     float *fin, wsis_stream_t stream) {
  using namespace wsis;
  WSIS_CHECK(C >= 1 && C <= kLossMaxC, "ce_dice: 1 <= classes <= 32");
  cudaStream_t st = as_stream(stream);
  const int cp = (C + 3) / 4 * 4, grid = loss_grid(N);
  switch (cp) {
#define WSIS_CE(CPV) case CPV: ce_dice_fwd_kernel<CPV><<<grid, kLossThreads, 0, st>>>(scores, labels, N, C, ignore_label, (float *)ws); break;
    WSIS_CE(4) WSIS_CE(8) WSIS_CE(12) WSIS_CE(16) WSIS_CE(20) WSIS_CE(24) WSIS_CE(28) WSIS_CE(32)
#undef WSIS_CE
  }
  WSIS_LAUNCH_OK();
  ce_dice_finalize_kernel<<<1, 128, 0, st>>>((const float *)ws, grid, C, cp, dice, fin);
  WSIS_LAUNCH_OK();
  return 0;
}

int wsis_ce_dice_bwd(const float *scores, const int64_t *labels, int64_t N, int C, int ignore_label, const float *fin,
                     const float *grad_out, float *dscores, wsis_stream_t stream) {
  using namespace wsis;
  WSIS_CHECK(C >= 1 && C <= kLossMaxC, "ce_dice: 1 <= classes <= 32");
  if (N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int cp = (C + 3) / 4 * 4, grid = loss_grid(N);
  switch (cp) {
#define WSIS_CE(CPV) case CPV: ce_dice_bwd_kernel<CPV><<<grid, kLossThreads, 0, st>>>(scores, labels, N, C, ignore_label, fin, grad_out, dscores); break;
    WSIS_CE(4) WSIS_CE(8) WSIS_CE(12) WSIS_CE(16) WSIS_CE(20) WSIS_CE(24) WSIS_CE(28) WSIS_CE(32)
#undef WSIS_CE
  }
  WSIS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
