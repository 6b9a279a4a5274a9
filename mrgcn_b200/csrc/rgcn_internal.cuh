// Internal launch helpers shared by rgcn_fwd.cu and rgcn_bwd.cu.
#pragma once
#include "common.cuh"

namespace mrgcn {

// Row stride (floats) of per-edge message buffers: rows are padded so that a row never straddles a 64-byte DRAM
// access granule more than necessary and can be moved with 16-byte vector accesses (pads are written as zeros).
__host__ __device__ inline int msg_stride(int out) { return out <= 4 ? 4 : out <= 8 ? 8 : ((out + 15) / 16) * 16; }

// store one OC-wide chunk [c0, c0+OC) of a message row (row stride ms), zero beyond `out`, nothing beyond ms
template <int OC>
__device__ __forceinline__ void store_msg_chunk(float *__restrict__ row, int c0, int out, int ms, const float (&val)[OC]) {
  if constexpr (OC % 4 == 0) {
#pragma unroll
    for (int q = 0; q < OC / 4; ++q) {
      const int o = c0 + 4 * q;
      if (o < ms)
        *reinterpret_cast<float4 *>(row + o) = make_float4(o < out ? val[4 * q] : 0.f, o + 1 < out ? val[4 * q + 1] : 0.f,
                                                           o + 2 < out ? val[4 * q + 2] : 0.f, o + 3 < out ? val[4 * q + 3] : 0.f);
    }
  } else {
#pragma unroll
    for (int q = 0; q < OC; ++q)
      if (c0 + q < ms) row[c0 + q] = c0 + q < out ? val[q] : 0.f;
  }
}

// Arguments of the segmented message aggregation (rows = destinations in forward, sources in the input-gradient pass).
struct AggArgs {
  const int32_t *rowptr;
  const int32_t *pI;   // e1_to_e2 of gI (messages of the identity term) or NULL
  const float *msgI;
  const int32_t *pF;   // e1_to_e3 of gF or NULL
  const float *msgF;
  const int32_t *rowptrF;  // rowptr of gF (== rowptr when gI == gF)
  const float *Wd;     // weight_I for the direct gather (B == 0) or NULL
  const int32_t *d_src, *d_rel;
  const float *d_val;
  int64_t NSd;
  const float *bias, *mask, *addend;
  int ms;              // row stride of msgI / msgF (msg_stride(odim))
  float *out;
  int ND, odim, relu, thresh;
};

// epilogue of every aggregation: + addend, + bias, node-dropout mask, ReLU (models/rgcn.py:78-87, layers/graph.py:99-102)
__device__ __forceinline__ void agg_store(const AggArgs &a, int i, int o, float acc) {
  if (a.addend) acc += a.addend[(size_t)i * a.odim + o];
  if (a.bias) acc += a.bias[o];
  if (a.mask) acc *= a.mask[i];
  if (a.relu) acc = fmaxf(acc, 0.f);
  a.out[(size_t)i * a.odim + o] = acc;
}

int launch_basis_mix_fwd(const float *comp, const float *V, float *W, int R, int B, int IO, cudaStream_t st);
// msg[e3,:] = val_e * Xrows[gather[e3], :] . W[r]   (gather = e3_src forward, e3_dst for the input gradient)
int launch_feat_msg(const mrgcn_graph *g, const int32_t *gather, const float *X, int ldx, const float *W, float *msg, int in,
                    int out, cudaStream_t st, const char *prof_name);
// segmented sums of AggArgs over a.ND rows; hub rows (long_rows) are processed one CTA per segment of `seg` edges
// (seg_hub/seg_first describe the segments), partial sums in hub_ws, combined in segment order
struct HubSegs {
  const int32_t *long_ids, *seg_hub, *seg_first;
  int n_long, n_segs, seg;
  float *ws;
};
int launch_agg(const AggArgs &a, const HubSegs &h, cudaStream_t st, const char *prof_name);
// narrow layers (narrow.cu): gather, transform and aggregate in one pass, W resident in shared memory, no message buffer:
//   out[i, :] = epilogue( sum_{e in row i} val_e * X[nbr_e, :] . W[rel_e] )      W: [R][in][out]
// rows/nbr/rel/val are a row-major edge order (E1 for the forward pass, E2 with W^T for the input gradient); `epi` carries
// ND, odim (= out), thresh and the epilogue.  Hub rows go through `h` (one CTA per segment, partials combined in order).
struct NarrowArgs {
  const int32_t *order;   // rows in processing order (by falling length) or NULL
  const int32_t *rowptr, *nbr, *rel;
  const float *val, *X, *W;
  int ldx, R, in, out;
  AggArgs epi;
};
bool narrow_supported(int R, int in, int out);
int launch_narrow(const NarrowArgs &a, const HubSegs &h, cudaStream_t st, const char *prof_name);
// per-basis projection of the node features on the tensor cores (feat_proj.cu): P[j, b*out + o] = X[j, :] . V[b, :, o]
bool feat_proj_supported(int in, int ldx, int B, int out);
int launch_feat_proj(const float *X, int64_t N, int in, int ldx, const float *V, int B, int out, float *vt_ws, float *xpad_ws,
                     float *P, cudaStream_t st);
// table-term kernels (tab.cu)
struct TabGeom { int GS, HS, BPT, NOP, CSP; };
// max_bpt: widest per-lane base count to prefer (more base splits = fewer registers per lane, more resident warps)
constexpr int kBwdWBpt = 20;
bool tab_geometry(int Btot, int out, TabGeom &g, int max_bpt = 40);
bool tab_c_geometry(int BI, int out, int &BC, int &OP);
int launch_tab_msg_fwd(const mrgcn_graph *g, const mrgcn_tab_plan *pl, const float *TI, const float *compI, int BI,
                       const float *TP, const float *compF, int BF, int out, float *msg, cudaStream_t st);
int launch_tab_bwd_w(const mrgcn_graph *g, const mrgcn_tab_plan *pl, const float *compI, int BI, int out,
                     const float *gact, float *gW, cudaStream_t st);
int launch_tab_bwd_c(const mrgcn_graph *g, const mrgcn_tab_plan *pl, const float *TI, int BI, int out, const float *gact,
                     float *rec, cudaStream_t st);
// identity-term backward in one pass (ident_bwd.cu): g_weight_I of the non-hub sources + scratch rows cbuf[e3, 0:B] of every
// edge; returns 1 (nothing launched) when the shape is not handled, 0 when launched, otherwise an error code
int launch_ident_bwd_fused(const mrgcn_graph *g, const float *V, const float *comp, int B, int out, const float *gact,
                           float *gW, float *cbuf, cudaStream_t st);
// feature-term weight gradient for wide outputs (feat_bwd_w.cu): part[c, k, o] per E3 chunk as a register-tiled product;
// returns 1 (nothing launched) when the shape is not handled (out <= 16, out or ldx not multiples of 4, tile > 256 threads)
int launch_feat_bwd_w_tile(const mrgcn_graph *g, const float *X, int ldx, const float *gact, float *part, int in, int out,
                           cudaStream_t st);
int pick_oc(int out);
int ident_tile(int B, int out, int OP);
struct IdentPipe;
int ident_pipe_config(IdentPipe &p, int64_t NS, int B, int out, size_t other_smem);

}  // namespace mrgcn
