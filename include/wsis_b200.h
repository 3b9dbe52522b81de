/*
 * wsis_b200.h -- C ABI of the B200-native 3D-WSIS scene-level hot path.
 *
 * This is the drop-in boundary: what the reference binds through
 *   - the 12 `torch.ops.spconv.*` operators registered at modules/lib/spconv/src/spconv/all.cc:19-34
 *     (signatures: include/spconv/spconv_ops.h:27-33, 253-256, 351-355), and
 *   - the external `pointgroup_ops` extension (call sites modules/datasets/scannetv2_dataset.py:449,
 *     train_scannetv2.py:189) and `torch_scatter.scatter` (modules/model/backbone_3D_WSIS.py:188,225-244),
 * is exported here as plain `extern "C"` functions over raw device pointers, sizes and a CUDA stream.
 * No torch types cross this boundary.  INTEGRATION.md shows the reference-side ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns and allocates all outputs and workspaces (the reference allocates inside the op,
 *     spconv_ops.h:55-62,284-287); `*_ws_bytes` helpers give workspace sizes;
 *   - every call is asynchronous on `stream` (no device->host sync inside, unlike spconv_ops.h:264),
 *     except where the name says `_host`/`_sync`;
 *   - return value: 0 = ok, non-zero = error; `wsis_last_error()` returns a thread-local message.  The
 *     Python host raises RuntimeError, matching TV_ASSERT_RT_ERR (include/tensorview/tensorview.h:70-101);
 *   - coords are int32 [N,4] = (batch, x, y, z) like spconv indices; each coordinate must be in [0, 65535].
 *   - a "neighbour map" `map[n_dst, K]` (int32, -1 = none) is the output-stationary form of the rulebook:
 *       conv:   dst[r] = sum_k src[ map[r, flip ? K-1-k : k] ] * W[k]
 *     For a rulebook with pairs (k, in=i, out=o) (the reference's indicePairs[k,0/1,:]):
 *       nbr_in [i,k] = o   and   nbr_out[o,k] = i.
 *     submanifold conv: nbr_out[o,k] == nbr_in[o,K-1-k], so only nbr_in is stored and flip=1 is used.
 */
#ifndef WSIS_B200_H_
#define WSIS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *wsis_stream_t; /* cudaStream_t */

/* ---------------------------------------------------------------------------------------------- */
/* library                                                                                        */
/* ---------------------------------------------------------------------------------------------- */
int wsis_version(void);
const char *wsis_last_error(void);
/* sm_count / compute capability of the current device; fails when no CUDA device is usable. */
int wsis_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t wsis_launch_count(void);
/* diagnostics: word i (0..15) of the host-mapped record a kernel fills before it traps on a barrier wait that never
 * completed ([0] kernel id: 1 conv_umma, 2 ecc_messages, 3 wgrad_umma; [1] barrier address; [2] parity; [3] CTA); the
 * record survives the context failure.  -1 when unavailable. */
int64_t wsis_debug_trap_word(int i);

/* ---------------------------------------------------------------------------------------------- */
/* primitives: fill, exclusive scan, stable LSD radix sort (hand-written; replace torch::_unique's  */
/* thrust sort at spconv_ops.h:126 and torch_scatter's atomics)                                   */
/* ---------------------------------------------------------------------------------------------- */
int wsis_fill_i32(int32_t *dst, int64_t n, int32_t value, wsis_stream_t stream);
int64_t wsis_scan_ws_bytes(int64_t n);
/* out[0..n] (n+1 entries): out[i] = sum_{j<i} in[j]; out[n] = total.  in may alias out[0..n-1]. */
int wsis_exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, void *ws, wsis_stream_t stream);
int64_t wsis_sort_ws_bytes(int64_t n);
/* stable sort of (key,val) by key bits [begin_bit,end_bit); result is written to keys_out/vals_out. */
int wsis_sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                        int64_t n, int begin_bit, int end_bit, void *ws, wsis_stream_t stream);

/* ---------------------------------------------------------------------------------------------- */
/* rulebook (indice pairs): GPU hash table instead of the reference's dense batch*D*H*W grid        */
/* (spconv_ops.h:60-62; indice.cu.h:150-208 subm, :22-65,112-148 strided)                          */
/* ---------------------------------------------------------------------------------------------- */
/* power-of-two slot count for n keys (load factor <= 0.5) */
int64_t wsis_hash_slots(int64_t n);

/* Submanifold rulebook (replaces getIndicePair<3> subm branch, spconv_ops.h:86-102).
 * hash_keys uint64[slots], hash_vals int32[slots]: workspaces, (re)initialised by the call.
 * nbr_in int32[N,K] out.  Duplicate coords: the largest row index wins (CPU reference, geometry.h:272-277). */
int wsis_rulebook_subm(const int32_t *coords, int64_t N, const int32_t ksize[3], const int32_t dilation[3],
                       const int32_t spatial_shape[3], uint64_t *hash_keys, int32_t *hash_vals, int64_t slots,
                       int32_t *nbr_in, wsis_stream_t stream);

/* Strided ("regular") sparse conv rulebook (replaces getIndicePair<3> conv branch, spconv_ops.h:103-136).
 * Phase 1 (count): builds the output hash set, numbers the outputs in the CPU reference's first-touch order
 *   (geometry.h:181-187) and writes the output count to n_out_dev[0].
 *   nbr_in int32[N,K] receives hash slot ids (scratch form);  rank_ws int32[N*K+1]; slot_rank int32[slots].
 * Phase 2 (fill), after the host has read n_out and allocated: out_coords int32[n_out,4], nbr_out int32[n_out,K]
 *   (pre-filled with -1 by the call), nbr_in rewritten to output row ids. */
int wsis_rulebook_conv_count(const int32_t *coords, int64_t N, const int32_t ksize[3], const int32_t stride[3],
                             const int32_t padding[3], const int32_t dilation[3], const int32_t out_shape[3],
                             uint64_t *hash_keys, int32_t *hash_vals, int64_t slots, int32_t *nbr_in,
                             int32_t *rank_ws, void *scan_ws, int32_t *n_out_dev, wsis_stream_t stream);
int wsis_rulebook_conv_fill(const int32_t *coords, int64_t N, int K, const uint64_t *hash_keys,
                            const int32_t *hash_vals, int64_t slots, int32_t *nbr_in, const int32_t *rank_ws,
                            int32_t *slot_rank, int64_t n_out, int32_t *out_coords, int32_t *nbr_out,
                            wsis_stream_t stream);

/* Reference-format rulebook from nbr_in: pairs int32[K,2,N] (-1 padded) and num int32[K], in exactly the
 * CPU reference's order (ascending input row inside each offset; geometry.h:176-190,281-289), which is a
 * valid instance of the GPU reference's atomics-ordered output (indice.cu.h:57,202).
 * pos_ws int32[K*N+1]; scan_ws from wsis_scan_ws_bytes(K*N). */
int wsis_pairs_from_nbr(const int32_t *nbr_in, int64_t N, int K, int32_t *pairs, int32_t *num, int32_t *pos_ws,
                        void *scan_ws, wsis_stream_t stream);
/* Inverse direction, for callers that hand us reference-format pairs (ops.indice_conv drop-in):
 * map int32[n_dst,K] must be pre-filled with -1; dst_side = 1 builds nbr_out (keyed by pairs[k,1,:]),
 * dst_side = 0 builds nbr_in (keyed by pairs[k,0,:]). */
int wsis_nbr_from_pairs(const int32_t *pairs, const int32_t *num, int64_t pair_stride, int K, int dst_side,
                        int32_t *map, wsis_stream_t stream);

/* ---------------------------------------------------------------------------------------------- */
/* sparse convolution = gather -> per-offset contraction -> (no scatter: output-stationary)          */
/* replaces indiceConv<T> / indiceConvBackward<T>, spconv_ops.h:253-349, 351-433                    */
/* ---------------------------------------------------------------------------------------------- */
/* Optional fused prologue on the gathered rows: x' = relu?(x * in_scale[c] + in_shift[c]) (eval-mode
 * BatchNorm1d + ReLU that precede every conv in ResidualBlock/UBlock, sparse_unet3d.py:163-172,258-298);
 * missing neighbours stay exactly zero.  Optional fused epilogue: dst += residual (identity branch).
 * Pass NULL to disable either.  W is the reference layout [K, Cin_w, Cout_w] (conv.py:98-99);
 * transpose_w = 1 contracts with W[k]^T (dgrad), i.e. Cin = Cout_w and Cout = Cin_w. */

/* exact-fp32 SIMT path (any Cin/Cout; used for widths the tensor-core kernel does not take, for
 * precision="simt" and as the GPU-side cross-check of the tensor-core path) */
int wsis_conv_simt(const float *src, const int32_t *map, int64_t n_dst, int K, int flip, const float *W,
                   int transpose_w, int Cin, int Cout, const float *in_scale, const float *in_shift, int in_relu,
                   const float *residual, float *dst, wsis_stream_t stream);

/* Spatial tiling (no reference equivalent: the reference walks pairs in arrival order, indice.cu.h:171-208).
 * wsis_tile_pad(n) = n rounded up to a multiple of 128.  wsis_spatial_order sorts a coordinate set along a
 * Morton curve (batch-major): order int32[wsis_tile_pad(N)] = row ids in curve order, padded with -1.
 * wsis_identity_order gives the trivial order for callers without coordinates.
 * wsis_tile_records turns a neighbour map into per-tile RECORDS, the only form of the rulebook the tensor-core
 * kernel reads: for tile t (destination rows order[128t .. 128t+127]) and `m[r,k] = map[order[..], flip ? K-1-k : k]`
 *   records + t * wsis_tile_record_stride(K)   (stride = 16 K + 80 + 256 K bytes):
 *     valid[K][4] u32 | {nU u32, amask u32, P u32, npack u32, members u16[32]} | loc[K][128] u16
 *     valid[k] = tile slots that have a source row through offset k; amask = offsets with at least one entry (1 if the
 *     tile has none at all); P = entries of the tile.  A PACK is one or two active offsets with disjoint valid-slot
 *     sets (greedy first fit in ascending offset order): members[i] = k0 | k1 << 8 (0xFF = none) for i < npack.
 *     loc row i < npack: loc[i][r] = index, in the tile's list of DISTINCT source rows, of the row slot r reads through
 *     the member of pack i that reaches it (0xFFFF where neither does)
 *   uidx + t * wsis_tile_unique_stride(K): int32[nU] the distinct source rows of the tile (any order)
 *   meta[t] = int32[4] {record bytes, nU, npack, P}
 *   stats = int32[2] {largest P, largest nU} over the tiles
 * records: uint8[num_tiles * stride], uidx: int32[num_tiles * unique_stride], meta: int32[num_tiles * 4], all 16-byte
 * aligned.  K <= 32. */
int64_t wsis_tile_pad(int64_t n);
int64_t wsis_spatial_order_ws_bytes(int64_t N);
int wsis_spatial_order(const int32_t *coords, int64_t N, const int32_t spatial_shape[3], int batch_size,
                       int32_t *order, void *ws, wsis_stream_t stream);
int wsis_identity_order(int64_t N, int32_t *order, wsis_stream_t stream);
int64_t wsis_tile_record_stride(int K);
int64_t wsis_tile_unique_stride(int K);
int wsis_tile_records(const int32_t *map, int64_t n_rows, int K, int flip, const int32_t *order, void *records,
                      int32_t *uidx, int32_t *meta, int32_t *stats, wsis_stream_t stream);
/* Records of the identity rulebook (K = 1, map[r] = r, natural row order) -- a 1x1 convolution / a Linear over rows on
 * wsis_conv_umma -- written directly (no hash set); `order` int32[wsis_tile_pad(n_rows)] receives the identity order. */
int wsis_tile_records_identity(int64_t n_rows, void *records, int32_t *uidx, int32_t *meta, int32_t *stats, int32_t *order,
                               wsis_stream_t stream);

/* tcgen05 tensor-core path over tile records (num_tiles = wsis_tile_pad(n_dst)/128).  precision: 1 = bf16 operands
 * (1e-2 contract), 3 = bf16x3 split operands with fp32 accumulation in TMEM (1e-4 contract).
 * Requires Cout % 16 == 0, 16 <= Cout <= 256, K <= 32; any Cin (channels are zero-padded to a multiple of the unit
 * width inside the kernel and in the packed weights; rows are read with 16-byte loads when Cin % 4 == 0 and src is
 * aligned). */
int wsis_conv_umma_supported(int Cin, int Cout);
int64_t wsis_conv_pack_bytes(int K, int Cin, int Cout, int precision);
int wsis_conv_pack_weights(const float *W, int K, int Cin, int Cout, int transpose_w, int precision, void *packed,
                           wsis_stream_t stream);
/* Launch plan of a layer shape (pure host arithmetic, no device needed): plan = int32[11] {dynamic shared-memory
 * bytes, pipeline stages, row-cache buffers, record buffers, builder groups, TMEM accumulators, MMA issuers,
 * accumulator buffers (1|2), weights resident in shared memory (0|1), weight-producer warps, units per stage}.
 * Fails when no pipeline fits the 227 KB of shared memory / 512 TMEM columns. */
int wsis_conv_umma_plan(int K, int Cin, int Cout, int precision, int32_t *plan);
int wsis_conv_umma(const float *src, const void *records, const int32_t *uidx, const int32_t *meta,
                   const int32_t *order, int64_t num_tiles, int K, const void *packed, int Cin, int Cout, int precision,
                   const float *in_scale, const float *in_shift, int in_relu, const float *residual, float *dst,
                   wsis_stream_t stream);

/* Diagnostics: when buf != NULL, every following wsis_conv_umma launch runs the instrumented build of the kernel and
 * CTA 0 writes, for each of its 23 warps w, buf[8 w + 0] = cycles in the role loop and buf[8 w + 1..4] = cycles spent in
 * the role's barrier waits (epilogue: accumulator full; gatherers: record, row-cache free; builders: record, row cache
 * full, operand slot free; issuers: record, accumulator free, stage full; record producer: buffer free; weight
 * producers: stage free).  buf = int64[23 * 8] device memory.  NULL switches back to the product kernel. */
int wsis_conv_debug_stats(void *buf);

/* dW[k] = sum_r prologue(src[map[r,k']])^T . g[r]   (fp32, dW is zeroed by the call). */
int wsis_conv_wgrad(const float *src, const int32_t *map, int64_t n_dst, int K, int flip, const float *g, int Cin,
                    int Cout, const float *in_scale, const float *in_shift, int in_relu, float *dW,
                    wsis_stream_t stream);

/* fused eval-BatchNorm(+ReLU) on a feature matrix, in place allowed: y = relu?(x*scale[c]+shift[c]) */
int wsis_affine_relu(const float *x, int64_t n, int C, const float *scale, const float *shift, int relu, float *y,
                     wsis_stream_t stream);

/* ---------------------------------------------------------------------------------------------- */
/* voxelization (pointgroup_ops; third-party, source not in the reference tree)                     */
/* ---------------------------------------------------------------------------------------------- */
/* Host (CUDA-free, fork-safe) voxelization_idx for DataLoader workers (scannetv2_dataset.py:449).
 * Two-phase: call with voxel_locs_host == NULL to get M (return value, <0 on error) and *max_active;
 * then with buffers voxel_locs int64[M,4], p2v int32[N], v2p int32[M,1+max_active] (zero-filled by the call). */
int64_t wsis_voxelize_idx_host(const int64_t *coords_host, int64_t N, int64_t *voxel_locs_host, int32_t *p2v_host,
                               int32_t *v2p_host, int32_t v2p_stride, int32_t *max_active_host);
/* Second phase without re-hashing: p2v = the point -> voxel map the counting call (voxel_locs = NULL, p2v != NULL) wrote,
 * M = its return value, v2p_stride >= 1 + max_active.  Fills voxel_locs int64[M,4] and v2p int32[M, v2p_stride]. */
int wsis_voxelize_idx_host_fill(const int64_t *coords, int64_t N, const int32_t *p2v, int64_t M, int64_t *voxel_locs,
                                int32_t *v2p, int32_t v2p_stride);

/* Device voxelization_idx (same numbering: first-occurrence order; v2p lists in ascending point order).
 * Phase 1: p2v int32[N] out; counts_dev int32[3]: [0] = M, [1] = max_active, [2] = 1 if a coordinate was
 *   outside [0,65535] (result invalid).
 *   ws: wsis_voxelize_ws_bytes(N).  Phase 2 (after the host read M, max_active): voxel_locs int64[M,4],
 *   v2p int32[M, 1+max_active] (zero-filled by the call). */
int64_t wsis_voxelize_ws_bytes(int64_t N);
int wsis_voxelize_idx_count(const int64_t *coords, int64_t N, int32_t *p2v, int32_t *counts_dev, void *ws,
                            wsis_stream_t stream);
int wsis_voxelize_idx_fill(const int64_t *coords, int64_t N, const int32_t *p2v, int64_t M, int32_t max_active,
                           int64_t *voxel_locs, int32_t *v2p, void *ws, wsis_stream_t stream);

/* voxelization(feats, v2p, mode=4 mean) forward / backward (train_scannetv2.py:189). */
int wsis_voxelize_mean_fwd(const float *feats, const int32_t *v2p, int64_t M, int32_t v2p_stride, int C, float *out,
                           wsis_stream_t stream);
int wsis_voxelize_mean_bwd(const float *dout, const int32_t *v2p, int64_t M, int32_t v2p_stride, int C, int64_t N,
                           float *dfeats, wsis_stream_t stream);

/* ---------------------------------------------------------------------------------------------- */
/* segmented reductions (torch_scatter.scatter(dim=0) at backbone_3D_WSIS.py:188,225,232,244)        */
/* ---------------------------------------------------------------------------------------------- */
/* CSR of an unsorted segment-id vector: order int32[N] = rows sorted stably by id, offsets int32[S+1]. */
int64_t wsis_segment_csr_ws_bytes(int64_t N, int64_t S);
int wsis_segment_csr(const int64_t *ids, int64_t N, int64_t S, int32_t *order, int32_t *offsets, void *ws,
                     wsis_stream_t stream);
/* out[s,:] = reduce_{j in seg s} src[ gather ? gather[order[j]] : order[j], :]; one warp per segment, fixed
 * summation order, no atomics.  reduce: 0 = sum, 1 = mean (sum / max(count,1)), 2 = max (empty -> 0).
 * `gather` (int32[N] or NULL) fuses the voxel->point gather output.features[p2v] (backbone_3D_WSIS.py:179)
 * with the superpoint pooling (:188). */
int wsis_segment_reduce(const float *src, const int32_t *gather, const int32_t *order, const int32_t *offsets,
                        int64_t S, int C, int reduce, float *out, wsis_stream_t stream);
/* dst[i,:] = src[idx[i],:]  (voxel->point gather, backbone_3D_WSIS.py:179) */
int wsis_gather_rows(const float *src, const int32_t *idx, int64_t n, int C, float *dst, wsis_stream_t stream);

/* Per-row MLP head in inference: Linear -> BatchNorm1d(eval) -> ReLU -> Linear (backbone_3D_WSIS.py:57-62 over every
 * point, :71-104 the superpoint heads) as one kernel, optionally fused with the voxel -> point gather (:179):
 *   out[i, 0..Cout) = W2 . relu(W1' . src[gather ? gather[i] : i, :] + t1) + b2
 * w1t float[Cin, H] = (diag(bn_scale) W1)^T, t1 float[H] = bn_scale*b1 + bn_shift, w2t float[H, coutp] = W2^T zero-padded
 * to coutp = wsis_mlp_head_coutp(Cout) columns, b2 float[coutp].  (Cin, H) in {(32,32), (64,64)}, Cout <= 32. */
int wsis_mlp_head_coutp(int Cout);
int wsis_mlp_head(const float *src, const int32_t *gather, int64_t n, int Cin, int H, int Cout, const float *w1t,
                  const float *t1, const float *w2t, const float *b2, float *out, wsis_stream_t stream);

/* ---------------------------------------------------------------------------------------------- */
/* inter-superpoint affinity (edge attention) and random-walk label propagation                     */
/* ---------------------------------------------------------------------------------------------- */
/* backbone_3D_WSIS.py:209-249 fused: position MLP, scaled dot product, segment softmax over edge_u,
 * weighted aggregation of v and the residual add.  Edges are given as a CSR over u:
 * eorder int32[E] (edge ids sorted by u), eoffsets int32[S+1] (from wsis_segment_csr on edge_u).
 * pos_mlp: float[16*3 + 16 + 16 + 1] = fc_position {w1[16,3], b1[16], w2[16], b2}.  D must be 64.
 * affinity float[E] (original edge order) and sp_feat float[S,D] = ecc + sum_e a_e v_v are written. */
int wsis_edge_attention(const float *q, const float *k, const float *v, const float *ecc, const float *centers,
                        const int64_t *edge_u, const int64_t *edge_v, const int32_t *eorder,
                        const int32_t *eoffsets, int64_t S, int64_t E, int D, const float *pos_mlp,
                        float *affinity, float *sp_feat, wsis_stream_t stream);

/* One step of the edge-conditioned GRU (ECC-GRU: modules/model/spg_modules.py:152-185 NNConv with mean aggregation,
 * :226-253 GRUCellEx with input gate and un-affine layer norms), nfeat = 32, one warp per target superpoint:
 *   m[t] = mean_{j in [offsets[t], offsets[t+1])} h[src[e]]^T . filters[e], e = eorder[j]  (filters float[E, 32*32],
 *          [in][out]; eorder = NULL means e = j)
 *   h_out[t] = GRUCellEx(m[t], h[t]);  cat_out[t*cat_stride + 0..31] = h_out[t] when cat_out != NULL.
 * (eorder int32[E], offsets int32[S+1]) = CSR of the edges over their TARGET superpoint (wsis_segment_csr on
 * edge_index[1]); src int64[E] = edge_index[0].  params float[wsis_ecc_gru_param_floats()] =
 *   ig.weight^T [32][32] | ig.bias [32] | weight_ih^T [32][96] | weight_hh^T [32][96] | bias_ih [96] | bias_hh [96].
 * h and h_out must be different buffers. */
int64_t wsis_ecc_gru_param_floats(void);
int wsis_ecc_gru_step(const float *h, const float *filters, const int64_t *src, const int32_t *eorder,
                      const int32_t *offsets, int64_t S,
                      const float *params, int layernorm, float eps, float *h_out, float *cat_out, int64_t cat_stride,
                      wsis_stream_t stream);

/* ECC-GRU without materialised edge filters (graphnet.py:21-39, spg_modules.py:152-185; csrc/ecc_umma.cu): the
 * 32x32 filter of every edge, F_e = W4 . he_e + b4, is regenerated by the tensor cores inside every step and
 * contracted with the source state on the spot, m_e = h[src(e)]^T F_e; [E,1024] never exists in memory.
 *   wsis_ecc_edge_mlp: once per forward.  he = ReLU(BN(L3(ReLU(L2(ReLU(L1(edgefeats))))))) (widths 13-32-128-64) for the
 *     edges in target-CSR order (edge j = eorder[j], eorder NULL = identity), written as the MMA's operand tiles into
 *     he_packed (wsis_ecc_he_bytes(E) bytes).  params float[wsis_ecc_edge_mlp_param_floats()] =
 *     W1^T[13][32] | b1[32] | W2^T[32][128] | b2[128] | (diag(bn_scale) W3)^T[128][64] | bn_scale*b3+bn_shift [64].
 *   wsis_ecc_pack_w4: W4 float[1024][64] (the last Linear's weight) -> w4_packed (wsis_ecc_w4_bytes() bytes).
 *   wsis_ecc_messages: one GRU step's messages msg float[E,32] (row j = edge eorder[j]); b4 float[1024] or NULL.
 *   wsis_ecc_gru_step_msg: wsis_ecc_gru_step with the mean taken over the message rows [offsets[t], offsets[t+1]). */
int64_t wsis_ecc_edge_mlp_param_floats(void);
int64_t wsis_ecc_he_bytes(int64_t E);
int64_t wsis_ecc_w4_bytes(void);
int wsis_ecc_pack_w4(const float *w4, void *w4_packed, wsis_stream_t stream);
int wsis_ecc_edge_mlp(const float *edgefeats, const int32_t *eorder, int64_t E, const float *params, void *he_packed,
                      wsis_stream_t stream);
int wsis_ecc_messages(const void *he_packed, const void *w4_packed, const float *b4, const float *h, const int64_t *src,
                      const int32_t *eorder, int64_t E, float *msg, wsis_stream_t stream);
int wsis_ecc_gru_step_msg(const float *h, const float *msg, const int32_t *offsets, int64_t S, const float *params,
                          int layernorm, float eps, float *h_out, float *cat_out, int64_t cat_stride,
                          wsis_stream_t stream);

/* Instance clustering on the superpoint graph of ONE scene (test_scannetv2.py:281-455 clustering_in_graph), on the
 * device.  xyz float[N,3] = xyz_origin; superpoint int64[N] in [0,S); (nbr_off int32[S+1], nbr int32[...]) = adjacency
 * lists in ascending neighbour id (igraph neighbors(mode='all')); sem int32[S] = argmax class per superpoint; centre
 * float[S,3] = superpoint centre + predicted offset (:305); count int32[S] = points per superpoint; occ / size float[S] =
 * pred_sp_occupancy / pred_sp_ins_size; class_valid int32[n_class] (1 = instance class), ind2label int32[n_class].
 * Outputs: conf double[<=S], label_id int32[<=S], inst_of_sp int32[S] (-1 = none), point_inst int32[N] (the dense
 * masks of the reference are point_inst[None,:] == arange(I)[:,None]), n_inst int32[1] (device).
 * ws: wsis_cluster_ws_bytes(N, S).  S < 65536. */
int64_t wsis_cluster_ws_bytes(int64_t N, int64_t S);
int wsis_cluster(const float *xyz, const int64_t *superpoint, int64_t N, int64_t S, const int32_t *nbr_off,
                 const int32_t *nbr, const int32_t *sem, const float *centre, const int32_t *count, const float *occ,
                 const float *size, const int32_t *class_valid, const int32_t *ind2label, int n_class, float voxel_scale,
                 void *ws, double *conf, int32_t *label_id, int32_t *inst_of_sp, int32_t *point_inst, int32_t *n_inst,
                 wsis_stream_t stream);

/* Random-walk label propagation (modules/datasets/scannetv2_dataset.py:664-735 + the dense fill at
 * train_scannetv2.py:565-570), float64 like the reference, exploiting that the transition matrix is
 * adjacency-masked and that only seed rows of T^(it+1) are read (:714-715).
 *   edges (u,v) int64[E] with affinity float[E]; adjacency = the same edge set (+ identity);
 *   seed_label int32[S] (-100 = unlabeled, else class id = vs['semantic_label']);
 *   pred int32[S], conf float[S]: argmax / max of softmax(sp_semantic_scores).
 *   (eorder,eoffsets) = CSR of the edges over edge_u, (torder,toffsets) = CSR over edge_v (wsis_segment_csr).
 *   The edge list must be simple (no duplicate (u,v), no self loops), as produced by the reference's
 *   preprocessing (data/ScanNetV2/prepare_data_inst_ScanNetV2.py:191-231).
 * Outputs: pseudo int32[S] (-100 or the seed superpoint id), score double[S].
 * ws: wsis_random_walk_ws_bytes(S, class_num). */
int64_t wsis_random_walk_ws_bytes(int64_t S, int class_num);
int wsis_random_walk(const int64_t *edge_u, const int64_t *edge_v, const float *affinity, const int32_t *eorder,
                     const int32_t *eoffsets, const int32_t *torder, const int32_t *toffsets, int64_t S, int64_t E,
                     const int32_t *seed_label, const int32_t *pred, const float *conf, int class_num, int iterations,
                     int32_t *pseudo, double *score, void *ws, wsis_stream_t stream);

/* ---------------------------------------------------------------------------------------------- */
/* training step (train_scannetv2.py:200-252): batch-statistics BatchNorm around the sparse convs,  */
/* fused AdamW                                                                                     */
/* ---------------------------------------------------------------------------------------------- */
/* Weight gradient on the tensor cores (tcgen05, both operands MN-major): the contraction over the rows of
 * indiceConvBackward's mm(in^T, dout) (spconv_ops.h:395-415), output-stationary like wsis_conv_umma:
 *   dW[k][ci][co] = sum_r prologue(src[map[r, flip ? K-1-k : k], ci]) * g[r, co]
 * `order` (int32[tile_pad(n_dst)] from wsis_spatial_order, or NULL) only changes the order in which destination
 * rows are visited (locality); precision 3 = bf16x3 split operands (fp32 contract), 1 = bf16 operands.
 * Needs 5 <= K <= 32, Cin % 32 == 0, Cout % 32 == 0 (wsis_conv_wgrad_umma_supported); otherwise use wsis_conv_wgrad. */
int wsis_conv_wgrad_umma_supported(int K, int Cin, int Cout);
int wsis_conv_wgrad_umma(const float *src, const int32_t *map, const int32_t *order, int64_t n_dst, int K, int flip,
                         const float *g, int Cin, int Cout, const float *in_scale, const float *in_shift, int in_relu,
                         int precision, float *dW, wsis_stream_t stream);

/* BatchNorm1d in training mode over x float[N,C] (torch.nn.BatchNorm1d / SyncBatchNorm of
 * sparse_unet3d.py:135-171, train_scannetv2.py:736).  The normalisation itself is applied by the consumer
 * (wsis_conv_umma's in_scale/in_shift/in_relu prologue, or wsis_affine_relu); these calls are the reductions.
 *   wsis_bn_stats:      sums double[2C+1] = (sum x | sum x^2 | N).  With synchronised statistics the caller
 *                       all-reduces `sums` across ranks before wsis_bn_finalize.
 *   wsis_bn_finalize:   stat float[4,C] = mean | invstd | scale = gamma*invstd | shift = beta - mean*scale;
 *                       running_mean/var (NULL = skip) updated with `momentum` (unbiased variance).
 *   wsis_bn_bwd_reduce: da = gradient w.r.t. y = relu?(x*scale+shift); dy = da*[y > 0] when relu;
 *                       sums double[2C] = (sum dy | sum dy*xhat); dgamma/dbeta float[C] = the LOCAL sums
 *                       (NULL = skip).  All-reduce `sums` for synchronised statistics.
 *   wsis_bn_bwd_apply:  dx = scale*(dy - sums[c]/count - xhat*sums[C+c]/count) (+ extra[i] when extra != NULL);
 *                       `count` points at the (global) row count, i.e. element 2C of the forward sums.
 * ws: wsis_bn_ws_bytes(N, C).  Reductions are two-stage in a fixed order (deterministic). */
int64_t wsis_bn_ws_bytes(int64_t N, int C);
/* Synchronised statistics with the collective INSIDE the kernel (torch.nn.SyncBatchNorm, train_scannetv2.py:736):
 * the block that combines the local column sums stores them into every peer's symmetric buffer over NVLink, signals
 * with a release-store of the call's sequence number, acquire-waits for all ranks, adds the contributions in rank
 * order and (forward) finalizes -- one kernel instead of combine + NCCL all-reduce + finalize.
 *   peers: device uint64[world] = base address of every rank's symmetric buffer as mapped in THIS process (own
 *          included); each buffer holds double data[2][world][slot] followed by uint32 flag[2][world], zero-filled
 *          before the first call.  seq = 1, 2, 3, ... (the same on every rank); slot >= 2C+1.  world == 1: no peers.
 *   wsis_bn_forward_sync    = wsis_bn_stats (+ all-reduce) + wsis_bn_finalize;  sums double[2C+1] also written.
 *   wsis_bn_bwd_reduce_sync = wsis_bn_bwd_reduce (+ all-reduce of sums; dgamma / dbeta stay local). */
int wsis_bn_forward_sync(const float *x, int64_t N, int C, void *ws, double *sums, const float *gamma, const float *beta,
                         float eps, float momentum, float *running_mean, float *running_var, float *stat,
                         const void *peers, int world, int rank, int64_t seq, int slot, wsis_stream_t stream);
int wsis_bn_bwd_reduce_sync(const float *x, const float *da, int64_t N, int C, const float *stat, int relu, void *ws,
                            double *sums, float *dgamma, float *dbeta, const void *peers, int world, int rank,
                            int64_t seq, int slot, wsis_stream_t stream);
int wsis_bn_stats(const float *x, int64_t N, int C, void *ws, double *sums, wsis_stream_t stream);
int wsis_bn_finalize(const double *sums, int C, const float *gamma, const float *beta, float eps, float momentum,
                     float *running_mean, float *running_var, float *stat, wsis_stream_t stream);
int wsis_bn_bwd_reduce(const float *x, const float *da, int64_t N, int C, const float *stat, int relu, void *ws,
                       double *sums, float *dgamma, float *dbeta, wsis_stream_t stream);
int wsis_bn_bwd_apply(const float *x, const float *da, int64_t N, int C, const float *stat, int relu,
                      const double *sums, const double *count, const float *extra, float *dx, wsis_stream_t stream);
/* Semantic loss of losses_3D_WSIS.py:52-64 (and :72-74 without dice): CrossEntropyLoss(ignore_index) + multi-class dice
 * on the softmax of the labelled rows, forward and backward in two passes over scores float[N,C] (C <= 32).
 *   wsis_ce_dice_fwd: fin float[2 + 2*cp] (cp = C rounded up to 4): fin[0] = loss, fin[1] = #labelled rows, then the
 *                     per-class coefficients the backward needs.  ws: wsis_ce_dice_ws_bytes(N, C).
 *   wsis_ce_dice_bwd: dscores float[N,C] = d loss / d scores * (*grad_out, or 1 when NULL); ignored rows get zeros. */
int64_t wsis_ce_dice_ws_bytes(int64_t N, int C);
int wsis_ce_dice_fwd(const float *scores, const int64_t *labels, int64_t N, int C, int ignore_label, int dice, void *ws,
                     float *fin, wsis_stream_t stream);
int wsis_ce_dice_bwd(const float *scores, const int64_t *labels, int64_t N, int C, int ignore_label, const float *fin,
                     const float *grad_out, float *dscores, wsis_stream_t stream);

/* torch.optim.AdamW (train_scannetv2.py:93-94) over flat buffers, `step` counts from 1.  The gradient is first
 * multiplied by grad_scale (1/world for a sum-all-reduced bucket) and elements [clamp_begin, clamp_end) are
 * clamped to [-1, 1] (the ECC gradient clamp, train_scannetv2.py:246-249). */
int wsis_adamw_step(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int64_t step, float grad_scale, int64_t clamp_begin,
                    int64_t clamp_end, wsis_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WSIS_B200_H_ */
