// Narrow layers in one pass (hidden layers of the R-GCN: AM 10 -> 11, AIFB 16 -> 4).
//
// The reference computes  AX = A . X  per relation and then  AX_r . W_r  (layers/graph.py:86-96).  For a layer whose
// R x in x out weights fit in shared memory (AM layer 1: 267 x 10 x 11 floats = 117 KB) the message of an edge is
// cheaper to recompute than to store: the generic path writes one padded message row per edge in relation-major order
// (feat_msg_fwd) and gathers it again in destination-major order (agg_fwd) - 2 x 64 B per edge through HBM plus a
// permutation index - where this kernel reads 12 B of edge structure and one L2-resident 40-byte row of X.
//
//   k_narrow       W staged once per CTA (row stride padded to an odd number of 16-byte chunks: lanes sit on different
//                  relations, their float4 reads then spread over the bank groups); 4 lanes per row split the INPUT
//                  columns, sum val_e * x_e over each run of one relation, multiply the run sum into W[rel] once, a
//                  fixed 2-step butterfly adds the four partial products, every lane stores its quarter of the outputs
//                  through the layer epilogue; rows are taken by falling length so that a warp's rows are alike
//   k_narrow_long  hub rows: one CTA per 512-edge segment, one edge slot per thread (W through L1), fixed-order tree
//   k_narrow_comb  partial sums of multi-segment hubs, in segment order
// The same kernels compute the input gradient dX[j, :] = sum_{e: src = j} val_e * W[rel_e] . gOut[dst_e, :] on the
// source-major order with the transposed weights.  Every sum has a fixed order: results are bit-reproducible.
#include "rgcn_internal.cuh"

namespace mrgcn {
namespace {

constexpr int kNarrowThreads = 1024;
constexpr int kNarrowLongThreads = 256;

__host__ __device__ inline int narrow_row_stride(int in, int OP) {
  const int chunks = in * OP / 4;
  return (chunks % 2 == 0 ? chunks + 1 : chunks) * 4;
}

// acc[0..OP) += sum_k (val * x[k]) * W[k][0..OP)
// SMEM: w = padded rows [in][OP] in shared memory; otherwise w = the layer's own rows [in][out] in global memory
template <int OP, bool SMEM>
__device__ __forceinline__ void narrow_edge(const float *__restrict__ xr, const float *__restrict__ w, int in, int out, float val,
                                            bool vec2, float (&acc)[OP]) {
  float x[16];
  if (vec2) {       // even row stride (AM hidden layer: 10 floats): 8-byte loads
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      float2 v = make_float2(0.f, 0.f);
      if (k + 1 < in) v = __ldg(reinterpret_cast<const float2 *>(xr + k));
      else if (k < in) v.x = __ldg(xr + k);
      x[k] = v.x; x[k + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = k < in ? __ldg(xr + k) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (k < in) {
      const float hk = val * x[k];
      if constexpr (SMEM) {
#pragma unroll
        for (int c = 0; c < OP / 4; ++c) {
          const float4 w4 = *reinterpret_cast<const float4 *>(w + k * OP + 4 * c);
          acc[4 * c + 0] = fmaf(hk, w4.x, acc[4 * c + 0]);
          acc[4 * c + 1] = fmaf(hk, w4.y, acc[4 * c + 1]);
          acc[4 * c + 2] = fmaf(hk, w4.z, acc[4 * c + 2]);
          acc[4 * c + 3] = fmaf(hk, w4.w, acc[4 * c + 3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < OP; ++c)
          if (c < out) acc[c] = fmaf(hk, __ldg(w + k * out + c), acc[c]);
      }
    }
  }
}

// W [R][in][out] is staged into shared memory as [R][in][OP] (zero padded) with row stride RS.
// Four lanes per row.  The row's edges arrive sorted by relation (E1 / E2 orders): the lanes walk them together, lane q
// keeps  xs[t] = sum over the current relation's run of val_e * x_e[q + 4t]  and, when the relation changes, multiplies
// its slice of the run sum into W[rel]  - the reference's own association, (A_r X) W_r (layers/graph.py:86-96) - so the
// shared-memory reads are per (row, relation) run, not per edge.  A butterfly over the four lanes ends the row.
template <int OP>
__global__ void __launch_bounds__(kNarrowThreads, 1) k_narrow(NarrowArgs a, int RS) {
  extern __shared__ __align__(16) float Ws[];
  const int in = a.in, od = a.out, run = in * OP;
  for (int x = threadIdx.x; x < a.R * run; x += kNarrowThreads) {
    const int r = x / run, y = x - r * run;
    const int k = y / OP, c = y - k * OP;
    Ws[(size_t)r * RS + y] = c < od ? __ldg(a.W + ((size_t)r * in + k) * od + c) : 0.f;
  }
  __syncthreads();
  const int q = threadIdx.x & 3;
  const int rows_per_cta = kNarrowThreads / 4;
  const int ngroups = (a.epi.ND + rows_per_cta - 1) / rows_per_cta;
  const bool k1 = q + 4 < in, k2 = q + 8 < in, k3 = q + 12 < in;      // q < in always (in >= 4 is not required: checked below)
  const bool k0 = q < in;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int slot = g * rows_per_cta + (threadIdx.x >> 2);
    const bool valid = slot < a.epi.ND;
    const int i = valid ? (a.order ? a.order[slot] : slot) : 0;
    int lo = 0, hi = 0;
    bool hub = false;
    if (valid) {
      lo = a.rowptr[i];
      hi = a.rowptr[i + 1];
      hub = a.epi.thresh > 0 && hi - lo > a.epi.thresh;            // hub: k_narrow_long
      if (hub) hi = lo;
    }
    float acc[OP];
#pragma unroll
    for (int c = 0; c < OP; ++c) acc[c] = 0.f;
    float xs0 = 0.f, xs1 = 0.f, xs2 = 0.f, xs3 = 0.f;
    int cur = -1;
    auto flush = [&](int r) {
      const float *w = Ws + (size_t)r * RS + q * OP;
      auto mul = [&](float hk, const float *wk) {
#pragma unroll
        for (int c = 0; c < OP / 4; ++c) {
          const float4 w4 = *reinterpret_cast<const float4 *>(wk + 4 * c);
          acc[4 * c + 0] = fmaf(hk, w4.x, acc[4 * c + 0]);
          acc[4 * c + 1] = fmaf(hk, w4.y, acc[4 * c + 1]);
          acc[4 * c + 2] = fmaf(hk, w4.z, acc[4 * c + 2]);
          acc[4 * c + 3] = fmaf(hk, w4.w, acc[4 * c + 3]);
        }
      };
      if (k0) mul(xs0, w);
      if (k1) mul(xs1, w + 4 * OP);
      if (k2) mul(xs2, w + 8 * OP);
      if (k3) mul(xs3, w + 12 * OP);
    };
    // software pipeline: structure words two edges ahead, the gathered row one edge ahead
    int nb1 = 0, rl1 = 0, nb2 = 0, rl2 = 0;
    float vl1 = 0.f, vl2 = 0.f;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
    auto load_x = [&](int nb) {
      const float *xr = a.X + (size_t)nb * a.ldx + q;
      x0 = k0 ? __ldg(xr) : 0.f;
      x1 = k1 ? __ldg(xr + 4) : 0.f;
      x2 = k2 ? __ldg(xr + 8) : 0.f;
      x3 = k3 ? __ldg(xr + 12) : 0.f;
    };
    if (lo < hi) { nb1 = a.nbr[lo]; rl1 = a.rel[lo]; vl1 = a.val[lo]; load_x(nb1); }
    if (lo + 1 < hi) { nb2 = a.nbr[lo + 1]; rl2 = a.rel[lo + 1]; vl2 = a.val[lo + 1]; }
    for (int e = lo; e < hi; ++e) {
      const float c0 = x0, c1 = x1, c2 = x2, c3 = x3, v = vl1;
      const int r = rl1;
      nb1 = nb2; rl1 = rl2; vl1 = vl2;
      if (e + 1 < hi) load_x(nb1);
      if (e + 2 < hi) { nb2 = a.nbr[e + 2]; rl2 = a.rel[e + 2]; vl2 = a.val[e + 2]; }
      if (r != cur) {
        if (cur >= 0) flush(cur);
        cur = r;
        xs0 = xs1 = xs2 = xs3 = 0.f;
      }
      xs0 = fmaf(v, c0, xs0); xs1 = fmaf(v, c1, xs1); xs2 = fmaf(v, c2, xs2); xs3 = fmaf(v, c3, xs3);
    }
    if (cur >= 0) flush(cur);
#pragma unroll
    for (int c = 0; c < OP; ++c) {
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
    }
    if (valid && !hub) {
#pragma unroll
      for (int c = 0; c < OP; ++c)
        if (c / (OP / 4) == q && c < od) agg_store(a.epi, i, c, acc[c]);
    }
  }
}

template <int OP>
__global__ void __launch_bounds__(kNarrowLongThreads) k_narrow_long(NarrowArgs a, HubSegs h) {
  __shared__ float red[kNarrowLongThreads][OP + 1];
  const int sg = blockIdx.x;
  const int hub = h.seg_hub[sg];
  const int i = h.long_ids[hub];
  const int first = h.seg_first[hub], nseg = h.seg_first[hub + 1] - first;
  const int row_lo = a.rowptr[i], row_hi = a.rowptr[i + 1];
  const int lo = min(row_hi, row_lo + (sg - first) * h.seg), hi = min(row_hi, lo + h.seg);
  float acc[OP];
#pragma unroll
  for (int c = 0; c < OP; ++c) acc[c] = 0.f;
  for (int e = lo + threadIdx.x; e < hi; e += kNarrowLongThreads)
    narrow_edge<OP, false>(a.X + (size_t)a.nbr[e] * a.ldx, a.W + (size_t)a.rel[e] * a.in * a.out, a.in, a.out, a.val[e], false, acc);
#pragma unroll
  for (int c = 0; c < OP; ++c) red[threadIdx.x][c] = acc[c];
  __syncthreads();
  for (int s = kNarrowLongThreads / 2; s > 0; s >>= 1) {
    for (int x = threadIdx.x; x < s * OP; x += kNarrowLongThreads) {
      const int t = x / OP, c = x - t * OP;
      red[t][c] += red[t + s][c];
    }
    __syncthreads();
  }
  if (threadIdx.x < a.out) {
    if (nseg == 1) agg_store(a.epi, i, threadIdx.x, red[0][threadIdx.x]);
    else h.ws[(size_t)sg * a.out + threadIdx.x] = red[0][threadIdx.x];
  }
}

__global__ void k_narrow_comb(AggArgs epi, HubSegs h) {
  const int hub = blockIdx.x;
  const int first = h.seg_first[hub], nseg = h.seg_first[hub + 1] - first;
  if (nseg <= 1) return;
  const int i = h.long_ids[hub];
  for (int o = threadIdx.x; o < epi.odim; o += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < nseg; ++s) acc += h.ws[(size_t)(first + s) * epi.odim + o];
    agg_store(epi, i, o, acc);
  }
}

size_t narrow_smem(int R, int in, int OP) { return (size_t)R * narrow_row_stride(in, OP) * 4; }

}  // namespace

static int narrow_op(int out) { return (out + 3) / 4 * 4; }

bool narrow_supported(int R, int in, int out) {
  static int off = -1;
  if (off < 0) { const char *e = getenv("MRGCN_NARROW"); off = (e && e[0] == '0') ? 1 : 0; }
  if (off || in < 1 || in > 16 || out < 1 || out > 16 || R < 1) return false;
  return narrow_smem(R, in, narrow_op(out)) <= 200 * 1024;
}

int launch_narrow(const NarrowArgs &a, const HubSegs &h, cudaStream_t st, const char *prof_name) {
  if (a.epi.ND <= 0) return 0;
  MRGCN_REQUIRE(narrow_supported(a.R, a.in, a.out), MRGCN_E_NOTSUP, "narrow: shape not supported");
  const int OP = narrow_op(a.out);
  const int RS = narrow_row_stride(a.in, OP);
  const size_t smem = narrow_smem(a.R, a.in, OP);
  mrgcn::prof_begin(prof_name, st);
  const int ngroups = (int)cdiv(a.epi.ND, kNarrowThreads / 4);
  const unsigned grid = (unsigned)(ngroups < kNumSMs ? ngroups : kNumSMs);
#define LAUNCH(OPV)                                                                                              \
  do {                                                                                                           \
    if (smem > 48 * 1024)                                                                                        \
      MRGCN_CUDA(cudaFuncSetAttribute(k_narrow<OPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    k_narrow<OPV><<<grid, kNarrowThreads, smem, st>>>(a, RS);                                                    \
    MRGCN_LAUNCH_CHECK();                                                                                        \
    if (h.n_long > 0) {                                                                                          \
      mrgcn::prof_begin("narrow_long", st);                                                                      \
      k_narrow_long<OPV><<<(unsigned)h.n_segs, kNarrowLongThreads, 0, st>>>(a, h);                               \
      MRGCN_LAUNCH_CHECK();                                                                                      \
    }                                                                                                            \
  } while (0)
  MRGCN_REQUIRE(h.n_long == 0 || h.ws || h.n_segs == h.n_long, MRGCN_E_BADARG, "narrow: hub_ws missing");
  switch (OP) {
    case 4: LAUNCH(4); break;
    case 8: LAUNCH(8); break;
    case 12: LAUNCH(12); break;
    default: LAUNCH(16); break;
  }
#undef LAUNCH
  if (h.n_long > 0 && h.n_segs > h.n_long) {
    mrgcn::prof_begin("narrow_combine", st);
    k_narrow_comb<<<(unsigned)h.n_long, 32, 0, st>>>(a.epi, h);
    MRGCN_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace mrgcn

extern "C" int32_t mrgcn_narrow_supported(int32_t R, int32_t in_dim, int32_t out_dim) {
  return mrgcn::narrow_supported(R, in_dim, out_dim) ? 1 : 0;
}
