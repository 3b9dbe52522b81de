// Instance clustering on the superpoint graph, on the device (reference: test_scannetv2.py:281-455
// `clustering_in_graph`, host restatement + golden pinning: wsis_b200/cluster.py, tests/golden/make_golden_cluster.py).
//
// The algorithm is sequential by construction (the visited flags of one breadth-first merge decide what the next seed
// may claim; fragments are absorbed in order and change the radius of the instance that absorbs them), but it walks a
// graph of a few thousand superpoints -- what is heavy in the reference is the per-superpoint N-point masks.  Here the
// point-sized work is parallel kernels (per-superpoint aggregates: wsis_segment_reduce; distinct voxels per group: one
// hash insert per point; the final point -> instance table) and the graph walk is ONE WARP per scene:
//   cluster_bfs_kernel      seeds in ascending id, neighbours 32 at a time, ordered append by ballot: exactly the
//                           reference's queue order
//   cluster_voxels_kernel   #distinct voxels touched by each group's points (voxelization_idx(...).shape[0], :373-377)
//   cluster_merge_kernel    group aggregates, primary / fragment split, in-order absorption of the fragments,
//                           confidences, labels, superpoint -> instance table
//   cluster_points_kernel   instance id per point (the dense [I, N] masks are (point_inst[None] == arange(I)[:, None]))
// Scalar types follow what the reference's numpy expressions evaluate to under NumPy 2 (float32 distances and
// thresholds, float64 instance centres / radii / confidences); sums over a group run in member order, so values that
// the reference obtains from a pairwise float32 mean can differ in the last bit (never in a golden decision).
#include <math.h>

#include "common.cuh"

namespace wsis {

struct ClusterWs {     // all arrays sized by S unless noted
  int32_t *visited, *members, *group_start /*S+1*/, *group_of, *voxels, *counters /*8*/;
  int32_t *kind /*0 none,1 primary,2 fragment*/, *inst_of_group, *prim_list;
  double *g_expocc, *g_size, *g_n, *g_cx /*3S*/, *p_rset;
  unsigned long long *vkeys; /* hash slots */
};

__device__ __forceinline__ float dist3(const float *a, const float *b) {
  const float dx = __fsub_rn(a[0], b[0]), dy = __fsub_rn(a[1], b[1]), dz = __fsub_rn(a[2], b[2]);
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

// one warp; members[] receives the groups back to back in discovery order
__global__ void cluster_bfs_kernel(int S, const int32_t *__restrict__ nbr_off, const int32_t *__restrict__ nbr,
                                   const int32_t *__restrict__ sem, const float *__restrict__ centre /*[S,3] instance centres*/,
                                   const float *__restrict__ size, const int32_t *__restrict__ class_valid, int n_class,
                                   int32_t *__restrict__ visited, int32_t *__restrict__ members,
                                   int32_t *__restrict__ group_start, int32_t *__restrict__ group_of,
                                   int32_t *__restrict__ counters) {
  const int lane = threadIdx.x;
  int tail = 0, ngroups = 0;
  for (int i = lane; i < S; i += 32) visited[i] = 0, group_of[i] = -1;
  __syncwarp();
  for (int seed = 0; seed < S; ++seed) {
    const int label = sem[seed];
    if (label < 0 || label >= n_class || !class_valid[label] || visited[seed]) continue;
    const float thr = __fmul_rn(0.25f, size[seed]);
    int head = tail;
    if (lane == 0) {
      visited[seed] = 1;
      members[tail] = seed;
      group_start[ngroups] = tail;
      group_of[seed] = ngroups;
    }
    ++tail;
    __syncwarp();
    while (head < tail) {
      const int cur = members[head++];
      const int beg = nbr_off[cur], end = nbr_off[cur + 1];
      for (int j0 = beg; j0 < end; j0 += 32) {
        const int j = j0 + lane;
        int nb = -1;
        bool ok = false;
        if (j < end) {
          nb = nbr[j];
          ok = sem[nb] == label && !visited[nb] && dist3(centre + 3 * cur, centre + 3 * nb) < thr;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int pos = tail + __popc(bal & ((1u << lane) - 1u));
          members[pos] = nb;
          visited[nb] = 1;
          group_of[nb] = ngroups;
        }
        tail += __popc(bal);
        __syncwarp();
      }
    }
    ++ngroups;
  }
  if (lane == 0) {
    group_start[ngroups] = tail;
    counters[0] = ngroups;
  }
}

// distinct (group, voxel) pairs: voxel = trunc(xyz * scale) like (xyz * 50).long() (:373-375)
__global__ void cluster_voxels_kernel(const float *__restrict__ xyz, const int64_t *__restrict__ superpoint, int64_t N,
                                      const int32_t *__restrict__ group_of, float scale, unsigned long long *__restrict__ keys,
                                      int64_t mask, int32_t *__restrict__ voxels) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int g = group_of[superpoint[i]];
  if (g < 0) return;
  const long long vx = (long long)__fmul_rn(xyz[3 * i], scale), vy = (long long)__fmul_rn(xyz[3 * i + 1], scale),
                  vz = (long long)__fmul_rn(xyz[3 * i + 2], scale);
  // 16 bits of group id + 3 x 16 bits of voxel coordinate (offset so that negative coordinates stay distinct)
  const unsigned long long key = ((unsigned long long)(unsigned)g << 48) | ((unsigned long long)((vx + 32768) & 0xFFFF) << 32) |
                                 ((unsigned long long)((vy + 32768) & 0xFFFF) << 16) | (unsigned long long)((vz + 32768) & 0xFFFF);
  int64_t s = (int64_t)(mix64(key) & (unsigned long long)mask);
  while (true) {
    const unsigned long long prev = atomicCAS(keys + s, kEmptyKey, key);
    if (prev == kEmptyKey) {
      atomicAdd(voxels + g, 1);
      return;
    }
    if (prev == key) return;
    s = (s + 1) & mask;
  }
}

// one thread: the graph-sized sequential part (:375-457)
__global__ void cluster_merge_kernel(int S, const int32_t *__restrict__ members, const int32_t *__restrict__ group_start,
                                     const int32_t *__restrict__ sem, const float *__restrict__ centre,
                                     const int32_t *__restrict__ count, const float *__restrict__ occ,
                                     const float *__restrict__ size, const int32_t *__restrict__ voxels,
                                     const int32_t *__restrict__ ind2label, ClusterWs w, double *__restrict__ conf,
                                     int32_t *__restrict__ label_id, int32_t *__restrict__ inst_of_sp) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int G = w.counters[0];
  int nprim = 0;
  for (int s = 0; s < S; ++s) inst_of_sp[s] = -1;
  for (int g = 0; g < G; ++g) {
    double e = 0.0, sz = 0.0, n = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    float ef = 0.f, szf = 0.f;                      // float32 sums like np.exp(occ[group]).mean() / np.mean(size[group])
    const int b = group_start[g], en = group_start[g + 1];
    for (int j = b; j < en; ++j) {
      const int s = members[j];
      ef += expf(occ[s]);
      szf += size[s];
      n += (double)count[s];
      cx += (double)centre[3 * s] * (double)count[s];
      cy += (double)centre[3 * s + 1] * (double)count[s];
      cz += (double)centre[3 * s + 2] * (double)count[s];
    }
    e = (double)ef, sz = (double)szf;
    w.g_expocc[g] = e, w.g_size[g] = sz, w.g_n[g] = n;
    w.g_cx[3 * g] = cx, w.g_cx[3 * g + 1] = cy, w.g_cx[3 * g + 2] = cz;
    const float occ_mean = ef / (float)(en - b);
    if ((float)voxels[g] < 0.3f * occ_mean) {       // fragment (:381-389)
      w.kind[g] = 2;
    } else {                                        // primary (:390-414)
      w.kind[g] = 1;
      const double gs = (double)(szf / (float)(en - b));
      w.p_rset[g] = fmax(fmax(0.01 * sqrt(n), 0.02 * sqrt((double)occ_mean)), gs);
      w.prim_list[nprim] = g;
      w.inst_of_group[g] = nprim++;
    }
  }
  // fragments in discovery order (:418-446); a primary's aggregates grow as it absorbs fragments.  members of a group
  // are kept as a linked chain through inst_of_group of the absorbed fragment (-2 - primary index)
  if (nprim > 0) {
    for (int f = 0; f < G; ++f) {
      if (w.kind[f] != 2) continue;
      const int cls = sem[members[group_start[f]]];
      const double fn = w.g_n[f];
      const double fx = w.g_cx[3 * f] / fn, fy = w.g_cx[3 * f + 1] / fn, fz = w.g_cx[3 * f + 2] / fn;
      int index = -1;
      double dmin = INFINITY;
      for (int i = 0; i < nprim; ++i) {
        const int g = w.prim_list[i];
        if (sem[members[group_start[g]]] != cls) continue;
        const double pn = w.g_n[g];
        const double dx = fx - w.g_cx[3 * g] / pn, dy = fy - w.g_cx[3 * g + 1] / pn, dz = fz - w.g_cx[3 * g + 2] / pn;
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        if (d < dmin) dmin = d, index = i;
      }
      if (index < 0) continue;                       // (the reference looks at primaries[-1] with dis_min = inf: no merge)
      const int g = w.prim_list[index];
      if (dmin < w.p_rset[g]) {
        // merged aggregates: counts of superpoints ride in g_size's companion below (kind reused: members counted via w.voxels? no)
        w.g_expocc[g] += w.g_expocc[f];
        w.g_size[g] += w.g_size[f];
        w.g_n[g] += w.g_n[f];
        w.g_cx[3 * g] += w.g_cx[3 * f], w.g_cx[3 * g + 1] += w.g_cx[3 * f + 1], w.g_cx[3 * g + 2] += w.g_cx[3 * f + 2];
        w.visited[g] += w.visited[f];                // visited[] is re-used as the superpoint count of a group (set below)
        const double m = (double)w.visited[g];
        const double occ_mean = (double)(float)(w.g_expocc[g] / m), size_mean = (double)(float)(w.g_size[g] / m);
        w.p_rset[g] = fmax(fmax(0.02 * sqrt(occ_mean), 0.01 * sqrt(w.g_n[g])), fmax(w.p_rset[g], size_mean));
        w.inst_of_group[f] = index;                  // the fragment's superpoints now belong to this instance
        w.kind[f] = 3;
      }
    }
  }
  for (int g = 0; g < G; ++g) {
    if (w.kind[g] != 1 && w.kind[g] != 3) continue;
    const int inst = w.inst_of_group[g];
    for (int j = group_start[g]; j < group_start[g + 1]; ++j) inst_of_sp[members[j]] = inst;
  }
  for (int i = 0; i < nprim; ++i) {
    const int g = w.prim_list[i];
    const double occ_mean = (double)(float)(w.g_expocc[g] / (double)w.visited[g]);
    conf[i] = fmin(w.g_n[g] / occ_mean, 1.0);
    label_id[i] = ind2label[sem[members[group_start[g]]]];
  }
  w.counters[1] = nprim;
}

__global__ void cluster_group_sizes_kernel(int S, const int32_t *__restrict__ group_start, const int32_t *__restrict__ counters,
                                           int32_t *__restrict__ nsp) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < counters[0]) nsp[g] = group_start[g + 1] - group_start[g];
}

__global__ void cluster_points_kernel(const int64_t *__restrict__ superpoint, int64_t N, const int32_t *__restrict__ inst_of_sp,
                                      int32_t *__restrict__ point_inst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) point_inst[i] = inst_of_sp[superpoint[i]];
}

static ClusterWs carve(void *ws, int64_t S, int64_t slots, int64_t *bytes) {
  char *p = reinterpret_cast<char *>(ws);
  auto take = [&](int64_t n) {
    char *r = p;
    p += (n + 255) / 256 * 256;
    return r;
  };
  ClusterWs w;
  w.visited = (int32_t *)take(4 * S), w.members = (int32_t *)take(4 * S), w.group_start = (int32_t *)take(4 * (S + 1));
  w.group_of = (int32_t *)take(4 * S), w.voxels = (int32_t *)take(4 * S), w.counters = (int32_t *)take(32);
  w.kind = (int32_t *)take(4 * S), w.inst_of_group = (int32_t *)take(4 * S), w.prim_list = (int32_t *)take(4 * S);
  w.g_expocc = (double *)take(8 * S), w.g_size = (double *)take(8 * S), w.g_n = (double *)take(8 * S);
  w.g_cx = (double *)take(24 * S), w.p_rset = (double *)take(8 * S);
  w.vkeys = (unsigned long long *)take(8 * slots);
  *bytes = p - reinterpret_cast<char *>(ws);
  return w;
}

static int64_t cluster_slots(int64_t N) {
  int64_t s = 1024;
  while (s < 2 * N) s <<= 1;
  return s;
}

}  // namespace wsis

using namespace wsis;

extern "C" {

int64_t wsis_cluster_ws_bytes(int64_t N, int64_t S) {
  int64_t b;
  carve(nullptr, std::max<int64_t>(S, 1), cluster_slots(N), &b);
  return b;
}

int wsis_cluster(const float *xyz, const int64_t *superpoint, int64_t N, int64_t S, const int32_t *nbr_off,
                 const int32_t *nbr, const int32_t *sem, const float *centre, const int32_t *count, const float *occ,
                 const float *size, const int32_t *class_valid, const int32_t *ind2label, int n_class, float voxel_scale,
                 void *ws, double *conf, int32_t *label_id, int32_t *inst_of_sp, int32_t *point_inst, int32_t *n_inst,
                 wsis_stream_t stream) {
  WSIS_CHECK(S >= 1 && S < 65536, "cluster: 1 <= S < 65536 superpoints per scene");
  cudaStream_t st = as_stream(stream);
  int64_t bytes, slots = cluster_slots(N);
  ClusterWs w = carve(ws, S, slots, &bytes);
  WSIS_CUDA(cudaMemsetAsync(w.voxels, 0, 4 * S, st));
  WSIS_CUDA(cudaMemsetAsync(w.kind, 0, 4 * S, st));
  WSIS_CUDA(cudaMemsetAsync(w.vkeys, 0xFF, 8 * slots, st));
  cluster_bfs_kernel<<<1, 32, 0, st>>>((int)S, nbr_off, nbr, sem, centre, size, class_valid, n_class, w.visited, w.members,
                                       w.group_start, w.group_of, w.counters);
  WSIS_LAUNCH_OK();
  if (N > 0) {
    cluster_voxels_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(xyz, superpoint, N, w.group_of, voxel_scale, w.vkeys,
                                                                       slots - 1, w.voxels);
    WSIS_LAUNCH_OK();
  }
  // visited[] is free after the BFS: it becomes the number of superpoints of each group
  cluster_group_sizes_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, st>>>((int)S, w.group_start, w.counters, w.visited);
  WSIS_LAUNCH_OK();
  cluster_merge_kernel<<<1, 32, 0, st>>>((int)S, w.members, w.group_start, sem, centre, count, occ, size, w.voxels, ind2label,
                                         w, conf, label_id, inst_of_sp);
  WSIS_LAUNCH_OK();
  if (N > 0) {
    cluster_points_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(superpoint, N, inst_of_sp, point_inst);
    WSIS_LAUNCH_OK();
  }
  WSIS_CUDA(cudaMemcpyAsync(n_inst, w.counters + 1, 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
