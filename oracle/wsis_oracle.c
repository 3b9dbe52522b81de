/*
 * wsis_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the 3D-WSIS scene-level hot path, written from the
 * reference's algorithm (file:line citations are into /root/reference).  It is the checker that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs compare the
 * CUDA path against.  Nothing in the product package (3d-wsis_b200/) may import, link or execute
 * this file: the product path fails loudly when its CUDA library is missing.
 *
 * Pinning (see tests/test_cpu.py):
 *   - rulebooks and sparse convs are checked against the UNMODIFIED reference spconv CPU path
 *     compiled from the reference's own sources (oracle/_ref/libspconv_ref.so, oracle/Makefile) and
 *     against the dense nn.Conv3d equivalence that the reference's own test uses
 *     (modules/lib/spconv/test/test_conv.py:325-382, fixture test_utils.py:141-190);
 *   - voxelization (pointgroup_ops) is a third-party dependency whose source is NOT in the
 *     reference tree (README.md:37-41, no version pinned): "parity unpinned" -- the restatement
 *     follows the call-site contract (train_scannetv2.py:149-151,189; scannetv2_dataset.py:449);
 *   - scatter / edge attention restate torch_scatter semantics used at
 *     modules/model/backbone_3D_WSIS.py:188,209-249 ("parity unpinned", library absent) and are
 *     cross-checked against torch index_add_/amax formulations in the tests.
 *
 * All index outputs are deterministic and follow the reference CPU order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t i32;
typedef int64_t i64;

/* ------------------------------------------------------------------------------------------- */
/* geometry: restates getValidOutPos, include/spconv/geometry.h:22-86 (C integer division,    */
/* enumeration with the last spatial dim fastest, descending output coordinate).              */
/* out: [<=K][4] = (o_x, o_y, o_z, kernel_offset).  Returns the number of valid outputs.       */
/* ------------------------------------------------------------------------------------------- */
static int valid_out_pos(const i32 *p, const i32 *ks, const i32 *st, const i32 *pad,
                         const i32 *dil, const i32 *oshape, i32 *out) {
  i32 lowers[3], uppers[3], counter[3], csize[3];
  int npts = 1, cnt = 0;
  for (int i = 0; i < 3; ++i) {
    lowers[i] = (p[i] - (ks[i] - 1) * dil[i] - 1 + st[i] + pad[i]) / st[i];
    uppers[i] = (p[i] + pad[i]) / st[i];
  }
  for (int i = 0; i < 3; ++i) {
    csize[i] = (uppers[i] - lowers[i]) / dil[i] + 1;
    npts *= csize[i];
    counter[i] = 0;
  }
  for (int n = 0; n < npts; ++n) {
    int valid = 1;
    i32 m = 1, offset = 0;
    for (int j = 2; j >= 0; --j) {
      i32 val = uppers[j] - counter[j] * dil[j];
      out[cnt * 4 + j] = val;
      if (val < 0 || val > oshape[j] - 1) valid = 0;
      offset += m * (p[j] - val * st[j] + pad[j]) / dil[j];
      m *= ks[j];
    }
    out[cnt * 4 + 3] = offset;
    if (valid) ++cnt;
    counter[2] += 1;
    for (int c = 2; c >= 0; --c) {
      if (counter[c] == csize[c] && c > 0) {
        counter[c - 1] += 1;
        counter[c] = 0;
      }
    }
  }
  return cnt;
}

static inline i64 row_idx(const i32 *p, const i32 *shape) {
  return ((i64)p[0] * shape[1] + p[1]) * shape[2] + p[2];
}

/* Submanifold rulebook.  Restates getIndicePairsSubM, geometry.h:246-293, as driven by
 * getIndicePair<3> (spconv_ops.h:71-83: stride:=1, padding:=ksize/2).
 * coords int32[N,4]=(b,x,y,z); pairs int32[K,2,N] must be pre-filled with -1; num int32[K] zeroed.
 * Returns 0, or -1 when the dense grid cannot be allocated. */
int orc_rulebook_subm(const i32 *coords, i64 N, i32 batch, const i32 *shape, const i32 *ks,
                      const i32 *dil, i32 *pairs, i32 *num) {
  i32 st[3] = {1, 1, 1}, pad[3] = {ks[0] / 2, ks[1] / 2, ks[2] / 2};
  i64 vol = (i64)shape[0] * shape[1] * shape[2];
  int K = ks[0] * ks[1] * ks[2];
  i32 *grid = (i32 *)malloc(sizeof(i32) * (size_t)(vol * batch));
  if (!grid) return -1;
  memset(grid, 0xff, sizeof(i32) * (size_t)(vol * batch));
  i32 *vp = (i32 *)malloc(sizeof(i32) * 4 * (size_t)K);
  for (i64 j = 0; j < N; ++j) grid[row_idx(coords + j * 4 + 1, shape) + vol * coords[j * 4]] = (i32)j;
  for (i64 j = 0; j < N; ++j) {
    int n = valid_out_pos(coords + j * 4 + 1, ks, st, pad, dil, shape, vp);
    for (int i = 0; i < n; ++i) {
      i32 off = vp[i * 4 + 3];
      i64 idx = row_idx(vp + i * 4, shape) + vol * coords[j * 4];
      if (grid[idx] > -1) {
        pairs[((i64)off * 2 + 0) * N + num[off]] = (i32)j;
        pairs[((i64)off * 2 + 1) * N + num[off]] = grid[idx];
        num[off]++;
      }
    }
  }
  free(vp);
  free(grid);
  return 0;
}

/* Strided ("regular") sparse conv rulebook.  Restates getIndicePairsConv, geometry.h:145-194
 * (CPU reference: output voxels numbered in first-touch order).
 * out_coords int32[N*K,4] (over-allocated like spconv_ops.h:104-106); returns the number of
 * active outputs, or -1 on allocation failure. */
i64 orc_rulebook_conv(const i32 *coords, i64 N, i32 batch, const i32 *oshape, const i32 *ks,
                      const i32 *st, const i32 *pad, const i32 *dil, i32 *out_coords, i32 *pairs,
                      i32 *num) {
  i64 vol = (i64)oshape[0] * oshape[1] * oshape[2];
  int K = ks[0] * ks[1] * ks[2];
  i32 *grid = (i32 *)malloc(sizeof(i32) * (size_t)(vol * batch));
  if (!grid) return -1;
  memset(grid, 0xff, sizeof(i32) * (size_t)(vol * batch));
  i32 *vp = (i32 *)malloc(sizeof(i32) * 4 * (size_t)K);
  i64 nact = 0;
  for (i64 j = 0; j < N; ++j) {
    i32 b = coords[j * 4];
    int n = valid_out_pos(coords + j * 4 + 1, ks, st, pad, dil, oshape, vp);
    for (int i = 0; i < n; ++i) {
      i32 off = vp[i * 4 + 3];
      i64 idx = row_idx(vp + i * 4, oshape) + vol * b;
      if (grid[idx] == -1) {
        out_coords[nact * 4 + 0] = b;
        out_coords[nact * 4 + 1] = vp[i * 4 + 0];
        out_coords[nact * 4 + 2] = vp[i * 4 + 1];
        out_coords[nact * 4 + 3] = vp[i * 4 + 2];
        grid[idx] = (i32)nact++;
      }
      pairs[((i64)off * 2 + 0) * N + num[off]] = (i32)j;
      pairs[((i64)off * 2 + 1) * N + num[off]] = grid[idx];
      num[off]++;
    }
  }
  free(vp);
  free(grid);
  return nact;
}

/* ------------------------------------------------------------------------------------------- */
/* gather-GEMM-scatter.  Restates indiceConv<float>, spconv_ops.h:253-349: per kernel offset,   */
/* gather rows (reordering.cc gather), multiply by filters[k] (torch::mm_out), scatter-add.    */
/* The subM centre shortcut (:289-292) is arithmetically the same pair loop, so it is not      */
/* special-cased.  filters: [K, Cin, Cout] row-major (conv.py:98-99 viewed as :288).           */
/* ------------------------------------------------------------------------------------------- */
void orc_indice_conv_fwd(const float *feat, const float *filt, const i32 *pairs, const i32 *num,
                         i64 npairs_stride, int K, int Cin, int Cout, i64 n_out, int inverse,
                         float *out) {
  memset(out, 0, sizeof(float) * (size_t)(n_out * Cout));
  float *buf = (float *)malloc(sizeof(float) * (size_t)Cout);
  for (int k = 0; k < K; ++k) {
    const i32 *pin = pairs + ((i64)k * 2 + (inverse ? 1 : 0)) * npairs_stride;
    const i32 *pout = pairs + ((i64)k * 2 + (inverse ? 0 : 1)) * npairs_stride;
    const float *w = filt + (i64)k * Cin * Cout;
    for (i32 j = 0; j < num[k]; ++j) {
      const float *x = feat + (i64)pin[j] * Cin;
      for (int c = 0; c < Cout; ++c) buf[c] = 0.f;
      for (int ci = 0; ci < Cin; ++ci) {
        float xv = x[ci];
        const float *wr = w + (i64)ci * Cout;
        for (int c = 0; c < Cout; ++c) buf[c] += xv * wr[c];
      }
      float *o = out + (i64)pout[j] * Cout;
      for (int c = 0; c < Cout; ++c) o[c] += buf[c];
    }
  }
  free(buf);
}

/* Restates indiceConvBackward<float>, spconv_ops.h:351-433:
 * dW[k] = gather(feat)^T . gather(dout);  din += gather(dout) . W[k]^T (scatter-add on the in side). */
void orc_indice_conv_bwd(const float *feat, const float *filt, const float *dout, const i32 *pairs,
                         const i32 *num, i64 npairs_stride, int K, int Cin, int Cout, i64 n_in,
                         int inverse, float *din, float *dfilt) {
  memset(din, 0, sizeof(float) * (size_t)(n_in * Cin));
  memset(dfilt, 0, sizeof(float) * (size_t)((i64)K * Cin * Cout));
  for (int k = 0; k < K; ++k) {
    const i32 *pin = pairs + ((i64)k * 2 + (inverse ? 1 : 0)) * npairs_stride;
    const i32 *pout = pairs + ((i64)k * 2 + (inverse ? 0 : 1)) * npairs_stride;
    const float *w = filt + (i64)k * Cin * Cout;
    float *dw = dfilt + (i64)k * Cin * Cout;
    for (i32 j = 0; j < num[k]; ++j) {
      const float *x = feat + (i64)pin[j] * Cin;
      const float *g = dout + (i64)pout[j] * Cout;
      float *dx = din + (i64)pin[j] * Cin;
      for (int ci = 0; ci < Cin; ++ci) {
        float xv = x[ci], acc = 0.f;
        const float *wr = w + (i64)ci * Cout;
        float *dwr = dw + (i64)ci * Cout;
        for (int c = 0; c < Cout; ++c) {
          dwr[c] += xv * g[c];
          acc += g[c] * wr[c];
        }
        dx[ci] += acc;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------- */
/* pointgroup_ops.voxelization_idx -- THIRD-PARTY, source absent, "parity unpinned".            */
/* Call-site contract: scannetv2_dataset.py:449; shapes train_scannetv2.py:149-151.            */
/* Semantics restated (upstream PointGroup lib/pointgroup_ops, as recalled in SURVEY.md §8c):  */
/* voxels are numbered in first-occurrence order over the point list; voxel_locs[m] are the     */
/* coords of its first point; v2p[m] = [count, p0, p1, ...] in point order, zero padded.        */
/* Two-phase: pass v2p == NULL to obtain M (return value) and *max_active; then call again.     */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
  i64 key[4];
  i32 val;
  int used;
} vslot;

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

i64 orc_voxelization_idx(const i64 *coords, i64 N, i64 *voxel_locs, i32 *p2v, i32 *v2p,
                         i32 v2p_stride, i32 *max_active) {
  size_t cap = 16;
  while (cap < (size_t)(2 * N + 1)) cap <<= 1;
  vslot *tab = (vslot *)calloc(cap, sizeof(vslot));
  i32 *cnt = (i32 *)calloc((size_t)(N > 0 ? N : 1), sizeof(i32));
  i64 M = 0;
  for (i64 i = 0; i < N; ++i) {
    const i64 *c = coords + i * 4;
    uint64_t h = mix64((uint64_t)c[0] * 0x9E3779B97F4A7C15ULL ^ mix64((uint64_t)c[1] ^ mix64((uint64_t)c[2] ^ mix64((uint64_t)c[3]))));
    size_t s = (size_t)h & (cap - 1);
    for (;;) {
      if (!tab[s].used) {
        tab[s].used = 1;
        memcpy(tab[s].key, c, sizeof(i64) * 4);
        tab[s].val = (i32)M;
        if (voxel_locs) memcpy(voxel_locs + M * 4, c, sizeof(i64) * 4);
        ++M;
        break;
      }
      if (tab[s].key[0] == c[0] && tab[s].key[1] == c[1] && tab[s].key[2] == c[2] && tab[s].key[3] == c[3]) break;
      s = (s + 1) & (cap - 1);
    }
    i32 m = tab[s].val;
    if (p2v) p2v[i] = m;
    if (v2p) {
      i32 k = cnt[m];
      if (1 + k < v2p_stride) v2p[(i64)m * v2p_stride + 1 + k] = (i32)i;
    }
    cnt[m]++;
  }
  i32 ma = 0;
  for (i64 m = 0; m < M; ++m) {
    if (cnt[m] > ma) ma = cnt[m];
    if (v2p) v2p[m * v2p_stride] = cnt[m];
  }
  if (max_active) *max_active = ma;
  free(cnt);
  free(tab);
  return M;
}

/* pointgroup_ops.voxelization(feats, v2p, mode=4): per-voxel mean (train_scannetv2.py:189).
 * mean = sum_i (1/count) * feat[p_i], multiply-then-accumulate in point order. */
void orc_voxelization_fwd(const float *feats, const i32 *v2p, i64 M, i32 v2p_stride, int C, float *out) {
  for (i64 m = 0; m < M; ++m) {
    const i32 *row = v2p + m * v2p_stride;
    i32 n = row[0];
    float w = n > 0 ? 1.0f / (float)n : 0.f;
    float *o = out + m * C;
    for (int c = 0; c < C; ++c) o[c] = 0.f;
    for (i32 j = 0; j < n; ++j) {
      const float *f = feats + (i64)row[1 + j] * C;
      for (int c = 0; c < C; ++c) o[c] += w * f[c];
    }
  }
}

/* voxelization_backward: d_feats[p_i] = d_out[m] / count (each point belongs to one voxel). */
void orc_voxelization_bwd(const float *dout, const i32 *v2p, i64 M, i32 v2p_stride, int C, i64 N, float *dfeats) {
  memset(dfeats, 0, sizeof(float) * (size_t)(N * C));
  for (i64 m = 0; m < M; ++m) {
    const i32 *row = v2p + m * v2p_stride;
    i32 n = row[0];
    float w = n > 0 ? 1.0f / (float)n : 0.f;
    for (i32 j = 0; j < n; ++j) {
      float *f = dfeats + (i64)row[1 + j] * C;
      for (int c = 0; c < C; ++c) f[c] = w * dout[m * C + c];
    }
  }
}

/* ------------------------------------------------------------------------------------------- */
/* torch_scatter.scatter(src, index, dim=0, reduce=...) restatement (backbone_3D_WSIS.py:188,   */
/* 225,232,244; train_scannetv2.py:177).  reduce: 0=sum, 1=mean (sum / max(count,1)), 2=max     */
/* (empty segments stay 0, torch_scatter 2.x behaviour).  out: [S,C] with S = max(index)+1.    */
/* ------------------------------------------------------------------------------------------- */
void orc_scatter(const float *src, const i64 *index, i64 N, int C, i64 S, int reduce, float *out) {
  i32 *cnt = (i32 *)calloc((size_t)(S > 0 ? S : 1), sizeof(i32));
  memset(out, 0, sizeof(float) * (size_t)(S * C));
  for (i64 i = 0; i < N; ++i) {
    i64 s = index[i];
    float *o = out + s * C;
    const float *x = src + i * C;
    if (reduce == 2) {
      if (cnt[s] == 0) for (int c = 0; c < C; ++c) o[c] = x[c];
      else for (int c = 0; c < C; ++c) o[c] = x[c] > o[c] ? x[c] : o[c];
    } else {
      for (int c = 0; c < C; ++c) o[c] += x[c];
    }
    cnt[s]++;
  }
  if (reduce == 1)
    for (i64 s = 0; s < S; ++s) {
      float d = (float)(cnt[s] > 0 ? cnt[s] : 1);
      for (int c = 0; c < C; ++c) out[s * C + c] /= d;
    }
  free(cnt);
}

/* ------------------------------------------------------------------------------------------- */
/* Edge attention ("affinity"), restates backbone_3D_WSIS.py:209-249:                           */
/*   pos_e  = fc_position(c_u - c_v)            (Linear(3,16) -> ReLU -> Linear(16,1))           */
/*   a_e    = (q_u . k_v) / sqrt(D) * pos_e                                                     */
/*   aff_e  = softmax over edges sharing u (max-subtracted)                                     */
/*   res_u  = sum_e aff_e * v_v ;  sp_feat = ecc + res                                          */
/* w1:[16,3] b1:[16] w2:[1,16] b2:[1]  (torch Linear layout).  Float32 throughout like torch.   */
/* ------------------------------------------------------------------------------------------- */
void orc_edge_attention(const float *q, const float *k, const float *v, const float *ecc,
                        const float *centers, const i64 *eu, const i64 *ev, i64 S, i64 E, int D,
                        const float *w1, const float *b1, const float *w2, const float *b2,
                        float *aff_out, float *sp_feat) {
  float *logit = (float *)malloc(sizeof(float) * (size_t)(E > 0 ? E : 1));
  float *mx = (float *)malloc(sizeof(float) * (size_t)(S > 0 ? S : 1));
  float *sum = (float *)calloc((size_t)(S > 0 ? S : 1), sizeof(float));
  char *seen = (char *)calloc((size_t)(S > 0 ? S : 1), 1);
  float inv = 1.0f / sqrtf((float)D);
  for (i64 e = 0; e < E; ++e) {
    i64 u = eu[e], w = ev[e];
    float d[3] = {centers[u * 3] - centers[w * 3], centers[u * 3 + 1] - centers[w * 3 + 1],
                  centers[u * 3 + 2] - centers[w * 3 + 2]};
    float pos = b2[0];
    for (int h = 0; h < 16; ++h) {
      float t = b1[h] + w1[h * 3] * d[0] + w1[h * 3 + 1] * d[1] + w1[h * 3 + 2] * d[2];
      if (t < 0.f) t = 0.f;
      pos += w2[h] * t;
    }
    float dot = 0.f;
    for (int c = 0; c < D; ++c) dot += q[u * D + c] * k[w * D + c];
    float a = dot * inv * pos;
    logit[e] = a;
    if (!seen[u] || a > mx[u]) mx[u] = a;
    seen[u] = 1;
  }
  for (i64 e = 0; e < E; ++e) {
    logit[e] = expf(logit[e] - mx[eu[e]]);
    sum[eu[e]] += logit[e];
  }
  for (i64 i = 0; i < S * D; ++i) sp_feat[i] = ecc[i];
  for (i64 e = 0; e < E; ++e) {
    i64 u = eu[e], w = ev[e];
    float a = logit[e] / sum[u];
    aff_out[e] = a;
    for (int c = 0; c < D; ++c) sp_feat[u * D + c] += a * v[w * D + c];
  }
  free(logit); free(mx); free(sum); free(seen);
}
