/* mrgcn_b200.h — C ABI of libmrgcn_b200.so (sm_100a).
 *
 * Drop-in boundary for the relational graph-convolution hot path of wxwilcke/mrgcn
 * (SURVEY.md §8b).  The reference has no FFI: its boundary is the Python nn.Module surface
 * (mrgcn/layers/graph.py:9-11,62 ; mrgcn/models/rgcn.py:63 ; mrgcn/tasks/link_prediction.py:645).
 * The host-side mirror in mrgcn_b200/ keeps those signatures and calls the entry points below
 * through ctypes from a torch.autograd.Function.  Each entry point names the reference lines it
 * replaces.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless its name ends in _host
 *   - all launches go to `stream`; no hidden synchronisation except in *_build (one-time)
 *   - return value: 0 = ok, otherwise a cudaError_t value or MRGCN_E_* (negative);
 *     mrgcn_last_error_string() describes the last failure of the calling thread
 *   - caller owns every buffer (PyTorch's caching allocator on the host side)
 *   - fp32 arithmetic, deterministic (no floating-point atomics anywhere)
 *
 * Edge orders of a relational graph with E stored entries, ND destination rows, NS source
 * columns per relation block and R relation blocks (stacked adjacency A = [A_0|...|A_{R-1}],
 * column c = r*NS + j; mrgcn/encodings/graph_structure.py:33-38):
 *   E1  destination-major  sorted by (dst, rel, src)   rowptr[ND+1]
 *   E2  source-major       sorted by (src, rel, dst)   colptr[NS+1]
 *   E3  relation-major     sorted by (source slab, rel, src, dst)   relptr[n_slabs*R+1]
 *       A slab is `slab_rows` consecutive sources: the relation-major kernels gather feature rows of one slab at a
 *       time, so that the slab (tens of MB) stays L2-resident and every row is fetched from HBM once, not once per edge.
 */
#ifndef MRGCN_B200_H
#define MRGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRGCN_E_BADARG   (-1)
#define MRGCN_E_OVERFLOW (-2)
#define MRGCN_E_NOTSUP   (-3)

typedef void *mrgcn_stream_t; /* cudaStream_t */

int mrgcn_version(void);
const char *mrgcn_last_error_string(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t mrgcn_launch_count(void);
/* Per-kernel timing for bench.py's roofline: when enabled, every kernel launch of the library is bracketed
 * by CUDA events on its own stream.  mrgcn_profile_dump synchronises the device, writes one line per kernel
 * name ("name launches total_ms\n") into buf (NUL terminated, truncated to cap) and clears the records;
 * returns the number of bytes the full text needs. */
void mrgcn_profile_enable(int on);
int64_t mrgcn_profile_dump(char *buf, int64_t cap);

/* Device view of one relational graph; every array is caller-allocated. */
typedef struct mrgcn_graph {
  int64_t E;
  int32_t ND, NS, R, _pad;
  /* E1 */
  int32_t *rowptr;   /* [ND+1] */
  int32_t *e1_src;   /* [E] */
  int32_t *e1_rel;   /* [E] */
  float   *e1_val;   /* [E] */
  int32_t *e1_to_e2; /* [E] position of the same edge in E2 */
  int32_t *e1_to_e3; /* [E] position of the same edge in E3 */
  /* E2 */
  int32_t *colptr;   /* [NS+1] */
  int32_t *e2_src;   /* [E] */
  int32_t *e2_dst;   /* [E] */
  int32_t *e2_rel;   /* [E] */
  float   *e2_val;   /* [E] */
  int32_t *e2_to_e3; /* [E] position of the same edge in E3 */
  /* E3 */
  int32_t *relptr;   /* [n_slabs*R+1] edge range of group g = slab*R + rel */
  int32_t *e3_src;   /* [E] */
  int32_t *e3_dst;   /* [E] */
  float   *e3_val;   /* [E] */
  int32_t *e3_to_e2; /* [E] */
  /* work lists (built by the host side from rowptr/colptr/relptr) */
  int32_t *long_rows;  int32_t n_long_rows;  int32_t long_row_thresh;   /* rows with deg > thresh */
  int32_t *long_cols;  int32_t n_long_cols;  int32_t long_col_thresh;
  /* hubs are cut into segments of long_seg edges, one CTA each; partial sums are combined in segment order */
  int32_t *row_seg_hub;   /* [n_row_segs] index into long_rows */
  int32_t *row_seg_first; /* [n_long_rows+1] first segment of hub h */
  int32_t *col_seg_hub;   /* [n_col_segs] */
  int32_t *col_seg_first; /* [n_long_cols+1] */
  int32_t n_row_segs, n_col_segs, long_seg, _pad3;
  int32_t *chunk_rel;  /* [n_chunks] relation of E3 chunk c */
  int32_t *chunk_ptr;  /* [n_chunks+1] E3 edge range of chunk c (never crosses a relation) */
  int32_t *rel_chunk_ptr; /* [R+1] range in rel_chunk_idx of relation r */
  int32_t *rel_chunk_idx; /* [n_chunks] chunk ids of each relation, in slab order */
  int32_t n_chunks, slab_rows; /* slab_rows: input of mrgcn_graph_build (0 = one slab) */
  /* optional (NULL = natural order): destination rows / source columns sorted by falling entry count, so that the rows a
   * warp of the one-pass narrow-layer kernels (narrow.cu) works on have (nearly) equal lengths */
  int32_t *rows_by_deg; /* [ND] */
  int32_t *cols_by_deg; /* [NS] */
} mrgcn_graph;

/* Work plan of the table-term kernels (tab.cu) over the source-major order E2 of one graph; built by the host side
 * (mrgcn_b200/graph.py: RelGraph.tab_plan) from colptr / e2_rel, once per graph.
 *   task          up to `lt` consecutive E2 edges of one source (a source with more edges has several tasks)
 *   wsrc          the sources with at most long_col_thresh edges (the others are hubs), in node order
 *   tile          consecutive tasks covering at most `tile_slots` E2 edges, one CTA at a time
 *   piece         at most 32 edges of ONE relation inside one tile; tperm lists the tile-local edge slots of every tile
 *                 in (relation, E2 position) order, piece_ptr cuts that list into pieces; the comp gradient of relation
 *                 r is the sum of the records of its pieces, taken in two fixed-order stages: rel_piece_idx lists the pieces
 *                 relation by relation, blk_ptr cuts that list into blocks of at most 128 pieces of one relation, and
 *                 rel_blk_ptr[r .. r+1] is the range of blocks of relation r. */
typedef struct mrgcn_tab_plan {
  int32_t n_tasks, n_wsrc, n_tiles, n_pieces, tile_slots, lt, _pad0, _pad1;
  int32_t *task_src, *task_lo;   /* [n_tasks] source and first E2 edge of task t */
  int32_t *tasks4;               /* [n_tasks][4] {source, first edge, number of edges, 0}: one 16-byte load per task */
  int32_t *wsrc;                 /* [n_wsrc] */
  int32_t *wtasks4;              /* [n_wsrc][4] {source, first edge, number of edges, 0} of the non-hub sources */
  int32_t *tile_task_ptr;        /* [n_tiles+1] */
  int32_t *tile_e0;              /* [n_tiles] first E2 edge of the tile */
  int32_t *tperm;                /* [E] */
  int32_t *piece_ptr;            /* [n_pieces+1] positions in tperm */
  int32_t *tile_piece_ptr;       /* [n_tiles+1] */
  int32_t *rel_piece_ptr;        /* [R+1] */
  int32_t *rel_piece_idx;        /* [n_pieces] */
  int32_t *blk_ptr;              /* [n_blks+1] positions in rel_piece_idx */
  int32_t *rel_blk_ptr;          /* [R+1] */
  int32_t n_blks, _pad2;
} mrgcn_tab_plan;
/* which table-term kernels apply to (B_I identity bases, B_F projected feature bases, out): bit 0 forward messages,
 * bit 1 basis gradient, bit 2 comp gradient without the E x B scratch (cbuf then holds n_pieces x B records). */
int32_t mrgcn_tab_mode(int32_t BI, int32_t BF, int32_t out);
/* Force the choice (same bits); -1 = automatic by shape: messages always, basis gradient for out >= 64 (csrc/tab.cu).
 * MRGCN_TAB=<mask> in the environment forces it at start-up. */
void mrgcn_set_tab_mask(int32_t mask);

/* Identity-term backward (B > 0): 1 (default) = one pass for g_weight_I and the comp-gradient scratch rows
 * (csrc/ident_bwd.cu: even out <= 16, B <= 64), 0 = the separate round-1 kernels.  MRGCN_IDENT_FUSED=0 in the environment
 * does the same at start-up. */
void mrgcn_set_ident_fused(int32_t on);

/* 1 when a feature-only layer of this shape runs in one pass (csrc/narrow.cu: R x in x out weights resident in shared
 * memory, no per-edge message buffer: msg_F / msgx_ws may then be NULL / a dummy); MRGCN_NARROW=0 disables it. */
int32_t mrgcn_narrow_supported(int32_t R, int32_t in_dim, int32_t out_dim);

/* Re-emit the reference's stacked adjacency as E1/E2/E3.
 * Replaces: scipy CSR -> torch COO hand-off (mrgcn/data/utils.py:165-170, mrgcn/data/batch.py:144-149)
 * and the per-call coalesce/sort inside torch.mm(sparse, dense) (mrgcn/layers/graph.py:75,95).
 * coo_row/coo_col: int64[E] = A._indices(); coo_val: float[E] = A._values() cast exactly to fp32.
 * ncols = A.shape[1] = R*NS.  Fills every E1/E2/E3 array of `g` (work lists are not touched).
 * Input need not be sorted or unique (duplicates stay separate edges: a COO sums them). */
int mrgcn_graph_build(const int64_t *coo_row, const int64_t *coo_col, const float *coo_val,
                      int64_t E, int64_t nrows, int64_t ncols, int32_t R,
                      mrgcn_graph *g, mrgcn_stream_t stream);

/* Triples (s,p,o) -> COO of the row-normalised stacked adjacency, bit-exact with
 * mrgcn/encodings/graph_structure.py:70-108,162-169 followed by the float32 cast of
 * mrgcn/data/io/tarball.py:151-157: block 2p holds A[s,o]=1/outdeg_p(s), block 2p+1 the inverse,
 * block R-1 the identity.  triples: int32[T*3] (unique), outputs sized 2T+N (include_inverse=1)
 * or T+N.  Output order is E1 order is NOT guaranteed; feed the result to mrgcn_graph_build. */
int mrgcn_adjacency_from_triples(const int32_t *triples, int64_t T, int32_t N, int32_t P,
                                 int32_t include_inverse,
                                 int64_t *coo_row, int64_t *coo_col, float *coo_val,
                                 mrgcn_stream_t stream);

/* One R-GCN layer, forward.  Replaces GraphConvolution.forward (mrgcn/layers/graph.py:62-102)
 * plus the row-mask multiply and ReLU of RGCN._forward_full_batch (mrgcn/models/rgcn.py:78-87):
 *
 *   out[i,:] = act( mask[i] * ( b + addend[i,:] + sum_{e=(i,r,j)} val_e * ( M_I(r,j) + X[j,:] . W_F(r) ) ) )
 *   M_I(r,j) = weight_I[r*NS+j,:]                       (B == 0)
 *            = sum_b comp_I[r,b] * weight_I[b*NS+j,:]   (B  > 0, graph.py:69-72)
 *   W_F(r)   = weight_F[r] or sum_b comp_F[r,b]*weight_F[b]      (graph.py:83-85)
 *
 * gI: graph of the identity term (may be NULL when weight_I is NULL);
 * gF: graph of the feature term (may be NULL when X is NULL); they differ only in mini-batch mode
 *     (graph.py:88-91).  Both must have the same ND.
 * weight_I [S*NS_I, out], comp_I [R,B] or NULL, X [NS_F, in] or NULL, weight_F [S, in, out],
 * comp_F [R,B] or NULL, bias [out] or NULL, row_mask [ND] or NULL, relu 0/1.
 * addend [ND,out] or NULL: a pre-activation term computed elsewhere (the identity term of a node-partitioned
 *   run, after its reduce-scatter).  Its gradient is `gact` of the backward call.
 * wmix: workspace [R*in*out] (only if B>0 and X given; receives W_F(r), needed by backward)
 * msg_I: workspace [E_I*mrgcn_msg_stride(out)] (only if B>0 and weight_I given); msg_F: workspace
 * [E_F*mrgcn_msg_stride(out)] (per-edge messages, rows padded to 16-byte multiples). */
int32_t mrgcn_msg_stride(int32_t out);
/* Per-basis projection of node features on the tensor cores (tcgen05 + tensor-map TMA, split-TF32 with fp32-grade
 * accuracy): P[j, b*out + o] = sum_k X[j, k] * weight_F[b, k, o].  Replaces the dense half of
 * torch.einsum('ij,bjk->bik', X, W_F) (mrgcn/layers/graph.py:83-94) after re-association over the bases.
 * mrgcn_feat_proj_supported returns ceil32(in) (the row pitch X wants) or 0 when the shape is not handled
 * (in < 32, in > 192, B*out not a multiple of 16 or > 1024). */
int32_t mrgcn_feat_proj_supported(int32_t in_dim, int32_t B, int32_t out_dim);
int mrgcn_feat_proj(const float *X, int64_t N, int32_t in_dim, int32_t x_stride, const float *weight_F, int32_t B,
                    int32_t out_dim, float *vt_ws, float *xpad_ws, float *P, mrgcn_stream_t stream);
/* Pinned host rows [rows][cols] -> device rows of `pitch` floats (cudaMemcpy2DAsync on `stream`); pad columns untouched.
 * The feature upload of MRGCN.forward (mrgcn/models/mrgcn.py:203-204) in the layout mrgcn_feat_proj reads. */
int mrgcn_upload_rows(const float *X_host, int64_t rows, int32_t cols, float *X_dev, int32_t pitch, mrgcn_stream_t stream);
typedef struct mrgcn_layer_args {
  const mrgcn_graph *gI, *gF;
  int32_t in_dim, out_dim, B, relu;
  const float *weight_I, *comp_I, *X, *weight_F, *comp_F, *bias, *row_mask, *addend;
  float *wmix, *msg_I, *msg_F;
  float *hub_ws; /* [max(n_row_segs, n_col_segs) * max(out, in, B*out)] partial sums of hub segments (may be NULL without hubs) */
  float *out; /* [ND, out] */
  /* table-term kernels (tab.cu): plan of gI (NULL = the tile-staging kernels of round 1).
   * proj [NS, B, out] workspace: when given (input layer with identity AND feature term, B > 0, gI == gF, and
   * mrgcn_feat_proj_supported(in, B, out) != 0 and mrgcn_tab_mode(B, B, out) bit 0) the features are projected per basis
   * on the tensor cores, proj[j, b, :] = X[j, :] . weight_F[b] (feat_proj.cu), and mixed with comp_F in the same pass
   * that mixes weight_I with comp_I; msg_F is then not used.  vt_ws [2 * B*out * ceil32(in)]: tf32 pieces of weight_F;
   * xpad_ws [NS * ceil32(in)] or NULL: zero-padded copy of X, needed unless x_stride == ceil32(in) already.
   * x_stride: row pitch of X in floats (0 = in_dim). */
  const mrgcn_tab_plan *plan;
  float *proj, *vt_ws, *xpad_ws;
  int32_t x_stride, _pad;
} mrgcn_layer_args;
int mrgcn_rgcn_layer_fwd(const mrgcn_layer_args *a, mrgcn_stream_t stream);

/* Backward of the same layer (replaces autograd through graph.py:62-102).
 * gout [ND,out] = dL/d out, out = the forward result (for the ReLU mask).
 * Gradients are written (not accumulated); NULL pointer = not wanted.
 *   g_weight_I [S*NS_I,out]  g_comp_I [R,B]  g_weight_F [S,in,out]  g_comp_F [R,B]  g_bias [out]
 *   g_X [NS_F,in]
 * Workspaces: gact [ND*out]; cbuf [E_I*B] (B>0 and identity term; [(n_pieces + n_blks)*B] when f.plan is given and
 *   mrgcn_tab_mode(B, 0, out) has bit 2);
 *   part [n_chunks * max(B, in*out)] ; g_wmix [R*in*out] (B>0 and feature term);
 *   colsum_ws [ceil(ND/1024) * out]; wt_ws, msgx_ws when g_X is wanted. */
typedef struct mrgcn_layer_bwd_args {
  mrgcn_layer_args f;     /* forward arguments (out = forward result, wmix as filled by forward) */
  const float *gout;
  float *g_weight_I, *g_comp_I, *g_weight_F, *g_comp_F, *g_bias, *g_X;
  float *gact, *cbuf, *part, *g_wmix, *colsum_ws;
  float *wt_ws;   /* [R*in*out]  (g_X wanted) transposed weights */
  float *msgx_ws; /* [E_F*mrgcn_msg_stride(in)] (g_X wanted) per-edge input-gradient messages */
  /* phases: 0 = everything in one call; otherwise a mask of MRGCN_BWD_* - the node-partitioned model first asks for
   * ACT | GX (the input gradient feeds a reduce-scatter), records an event, and then for the weight gradients, so that
   * the collective runs under them.  gact must be the same buffer in both calls. */
  int32_t phases, _pad;
} mrgcn_layer_bwd_args;
#define MRGCN_BWD_ACT   1 /* gact = gout * relu' * mask, bias gradient */
#define MRGCN_BWD_IDENT 2 /* g_weight_I, g_comp_I */
#define MRGCN_BWD_FEATW 4 /* g_weight_F, g_comp_F */
#define MRGCN_BWD_GX    8 /* g_X */
int mrgcn_rgcn_layer_bwd(const mrgcn_layer_bwd_args *a, mrgcn_stream_t stream);

/* DistMult scorer.  Replaces score_distmult_bc (mrgcn/tasks/link_prediction.py:645-665), generic path:
 *   score[t] = sum_k E[s_t,k] * Rel[p_t,k] * E[o_t,k]        s,p,o: int64[n] */
int mrgcn_distmult_fwd(const int64_t *s, const int64_t *p, const int64_t *o, int64_t n,
                       const float *E, const float *Rel, int32_t h, float *score,
                       mrgcn_stream_t stream);
/* Backward: gE [N,h] and gRel [NR,h] are fully written (rows without triples = 0).
 * ws: int32 workspace of mrgcn_distmult_bwd_ws_elems(n) elements (incidence lists + the temporary storage of the
 * library radix sort that orders them: nothing is allocated inside the call). */
int64_t mrgcn_distmult_bwd_ws_elems(int64_t n);
int mrgcn_distmult_bwd(const int64_t *s, const int64_t *p, const int64_t *o, int64_t n,
                       const float *gscore, const float *E, const float *Rel,
                       int64_t N, int64_t NR, int32_t h, float *gE, float *gRel, int32_t *ws,
                       mrgcn_stream_t stream);

/* Ranking.  Replaces compute_ranks_fast + filter_scores_ (link_prediction.py:557-643) for one side, in GEMM form
 * (csrc/rank.cu): 32-fact x 128-candidate score tiles in registers, compared with the target's score on the fly - no
 * F x N score matrix.
 *   scores[f,c] = sum_k (E[s,k]*Rel[p_f,k])*E[o,k]  with the candidate c in the subject (head=1) or object (head=0) slot
 *   candidates listed in filt (CSR over facts: filt_ptr[f..f+1] -> candidate ids; the target itself is ignored) do not count
 *   rank[f] = #(scores[f,:] > scores[f,target_f]) + round_half_even((#ties-1)/2) + 1
 * ws: mrgcn_distmult_rank_ws_elems(F) int32 words. */
int64_t mrgcn_distmult_rank_ws_elems(int64_t F);
int mrgcn_distmult_rank(const int64_t *facts /* [F,3] */, int64_t F, int32_t head,
                        const float *E, const float *Rel, int64_t N, int32_t h,
                        const int32_t *filt_ptr, const int32_t *filt_idx, /* NULL = raw */
                        int32_t *ws, int64_t *rank, mrgcn_stream_t stream);

/* ---- either side of the path ------------------------------------------------------------------
 * Fused gradient clipping + Adam.  Replaces, for CUDA parameters, the pair
 *   nn.utils.clip_grad_norm_(model.parameters(), max_norm); optimizer.step()        (torch.optim.Adam, no amsgrad)
 * of mrgcn/tasks/node_classification.py:190-193 / link_prediction.py:324-326 (optimizer built in mrgcn/tasks/utils.py:8-45).
 * mrgcn_grad_sqnorm: total[0] (+)= sum g^2 (double; ws: mrgcn_sqnorm_ws_elems() doubles); call once per gradient tensor
 * with accumulate = 0 for the first.  mrgcn_adam_clip: one pass over p, g, m, v with
 *   coef = min(1, max_norm / (sqrt(total_sq[0]) + 1e-6))   (max_norm <= 0 or total_sq NULL: no clipping)
 * `step` is the 1-based step count of the bias corrections. */
int64_t mrgcn_sqnorm_ws_elems(void);
int mrgcn_grad_sqnorm(const float *g, int64_t n, double *ws, double *total, int32_t accumulate, mrgcn_stream_t stream);
int mrgcn_adam_clip(float *p, const float *g, float *m, float *v, int64_t n, const double *total_sq, float max_norm,
                    float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, mrgcn_stream_t stream);
/* Gated scatter of one modality's encoder output into the node-feature matrix (mrgcn/models/mrgcn.py:295-301):
 *   X[row_idx[i], col0 : col0 + d] = gate[0] * src[i, :]        and its backward (g_src, g_gate; ws as above). */
int mrgcn_scatter_rows(const float *src, const int64_t *row_idx, const float *gate, float *X, int64_t m, int32_t d,
                       int32_t ldx, int32_t col0, mrgcn_stream_t stream);
int mrgcn_scatter_rows_bwd(const float *gX, const int64_t *row_idx, const float *gate, const float *src, float *g_src,
                           float *g_gate, double *ws, int64_t m, int32_t d, int32_t ldx, int32_t col0, mrgcn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
