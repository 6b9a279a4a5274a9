// Internal launch helpers shared by rgcn_fwd.cu and rgcn_bwd.cu.
#pragma once
#include "common.cuh"

namespace mrgcn {

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
  float *out;
  int ND, odim, relu, thresh;
};

int launch_basis_mix_fwd(const float *comp, const float *V, float *W, int R, int B, int IO, cudaStream_t st);
// msg[e3,:] = val_e * Xrows[gather[e3], :] . W[r]   (gather = e3_src forward, e3_dst for the input gradient)
int launch_feat_msg(const mrgcn_graph *g, const int32_t *gather, const float *X, const float *W, float *msg, int in,
                    int out, cudaStream_t st, const char *prof_name);
// segmented sums of AggArgs over a.ND rows; hub rows (long_rows) are processed one CTA per segment of `seg` edges
// (seg_hub/seg_first describe the segments), partial sums in hub_ws, combined in segment order
struct HubSegs {
  const int32_t *long_ids, *seg_hub, *seg_first;
  int n_long, n_segs, seg;
  float *ws;
};
int launch_agg(const AggArgs &a, const HubSegs &h, cudaStream_t st, const char *prof_name);
// tensor-core (tcgen05, 3xTF32) variant of launch_feat_msg; *launched = 0 when it does not apply
int launch_feat_msg_tc(const mrgcn_graph *g, const int32_t *gather, const float *X, const float *W, float *msg, int in,
                       int out, cudaStream_t st, const char *prof_name, int *launched);
int pick_oc(int out);
int ident_tile(int B, int out, int OP);
struct IdentPipe;
int ident_pipe_config(IdentPipe &p, int64_t NS, int B, int out, size_t other_smem);

}  // namespace mrgcn
