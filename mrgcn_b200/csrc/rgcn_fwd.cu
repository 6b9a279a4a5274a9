// R-GCN layer forward (replaces GraphConvolution.forward, /root/reference/mrgcn/layers/graph.py:62-102,
// and the mask/ReLU lines of RGCN._forward_full_batch, mrgcn/models/rgcn.py:78-87).
//
// The reference materialises one dense operand row per (relation, node) -- R*N*out floats -- and
// calls torch.mm(sparse COO, dense).  Here nothing of size R*N*out exists:
//   identity term, B>0 : source-major pass (E2).  A CTA stages the basis rows V_I[:, j0:j0+TJ, :] of TJ
//                        consecutive sources ONCE in shared memory (coalesced runs per basis), mixes them
//                        with comp_I[r,:] per edge and writes one `out`-wide message per edge.
//   feature term       : relation-major pass (E3).  A CTA stages W_F(r) once per chunk of edges of one
//                        relation and computes val * X[j,:] . W_F(r) per edge (rows of X staged through
//                        shared memory with coalesced loads).
//   aggregation        : destination-major pass (E1).  Deterministic segmented sum of the messages of a
//                        row (+ the direct weight_I row gather when B == 0) fused with bias, row mask, ReLU.
// All passes are HBM/L2-bound gathers; there is no float atomic anywhere.
#include <stdlib.h>

#include "common.cuh"
#include "pipeline.cuh"
#include "ident_pipe.cuh"
#include "rgcn_internal.cuh"

namespace mrgcn {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// W[r, x] = sum_b comp[r,b] * V[b, x]   (graph.py:83-85; x over in*out).  Tiny.
__global__ void k_basis_mix_fwd(const float *__restrict__ comp, const float *__restrict__ V, float *__restrict__ W,
                                int B, int IO) {
  int r = blockIdx.y;
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= IO) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc = fmaf(__ldg(comp + (size_t)r * B + b), __ldg(V + (size_t)b * IO + x), acc);
  W[(size_t)r * IO + x] = acc;
}

// ------------------------------------------------------------------------------------------------
// Identity term with basis decomposition (graph.py:69-75), source-major.
//   msg[e2, :] = val_e * sum_b comp[r_e, b] * V[b, j_e, :]
// A CTA walks tiles of TJ consecutive sources.  The tile V[:, j0:j0+TJ, :] is B contiguous runs of TJ*out
// floats; it is read from HBM exactly once and mixed per edge out of shared memory.

// acc[0..OC) += c * row[0..OC) for every basis; row stride between bases = bstride floats
template <int OC, int VW>
__device__ __forceinline__ void mix_bases(const float *__restrict__ vp, size_t bstride, const float *__restrict__ cr,
                                          int B, float (&acc)[OC]) {
  float2 a2[OC / 2];
#pragma unroll
  for (int q = 0; q < OC / 2; ++q) a2[q] = make_float2(acc[2 * q], acc[2 * q + 1]);
#pragma unroll 4
  for (int b = 0; b < B; ++b) {
    const float c = cr[b];
    const float *row = vp + (size_t)b * bstride;
    if constexpr (VW == 4) {
#pragma unroll
      for (int q = 0; q < OC / 4; ++q) {
        float4 t = reinterpret_cast<const float4 *>(row)[q];
        fma2(a2[2 * q], c, make_float2(t.x, t.y));
        fma2(a2[2 * q + 1], c, make_float2(t.z, t.w));
      }
    } else if constexpr (VW == 2) {
#pragma unroll
      for (int q = 0; q < OC / 2; ++q) fma2(a2[q], c, reinterpret_cast<const float2 *>(row)[q]);
    } else {
#pragma unroll
      for (int q = 0; q < OC / 2; ++q) fma2(a2[q], c, make_float2(row[2 * q], row[2 * q + 1]));
    }
  }
#pragma unroll
  for (int q = 0; q < OC / 2; ++q) { acc[2 * q] = a2[q].x; acc[2 * q + 1] = a2[q].y; }
}

// per-edge work of one staged tile: RS = row stride (floats) of a (basis, source) row in the tile
template <int OC, int VW>
__device__ __forceinline__ void ident_msg_tile(const float *__restrict__ Vs, size_t bstride, int RS,
                                               const float *__restrict__ comp_s, int CS, const float *__restrict__ comp,
                                               int B, int out, int j0, int e_lo, int e_hi,
                                               const int32_t *__restrict__ e2_src, const int32_t *__restrict__ e2_rel,
                                               const float *__restrict__ e2_val, float *__restrict__ msg) {
  for (int e = e_lo + threadIdx.x; e < e_hi; e += kThreads) {
    const int jl = e2_src[e] - j0, r = e2_rel[e];
    const float v = e2_val[e];
    const float *cr = comp_s ? comp_s + r * CS : comp + (size_t)r * B;
    const int ms = msg_stride(out);
    for (int c0 = 0; c0 < ms; c0 += OC) {
      float acc[OC];
#pragma unroll
      for (int o = 0; o < OC; ++o) acc[o] = 0.f;
      if (c0 < out) mix_bases<OC, VW>(Vs + (size_t)jl * RS + c0, bstride, cr, B, acc);
#pragma unroll
      for (int o = 0; o < OC; ++o) acc[o] *= v;
      store_msg_chunk<OC>(msg + (size_t)e * ms, c0, out, ms, acc);
    }
  }
}

__device__ __forceinline__ void load_comp_smem(float *comp_s, const float *__restrict__ comp, int R, int B, int CS) {
  for (int r = threadIdx.x / 32; r < R; r += kThreads / 32)
    for (int b = threadIdx.x & 31; b < B; b += 32) comp_s[r * CS + b] = __ldg(comp + (size_t)r * B + b);
}

// (a) TMA-engine variant (ident_pipe.cuh): producer warp + ring of stages + consumer warps; needs 16-byte aligned
//     runs: (NS*out) % 4 == 0, (TJ*out) % 4 == 0, NS >= TJ.  The last tile is shifted back to NS-TJ so that every
//     tile is full; the sources it shares with its predecessor produce identical messages twice.
template <int OC, int VW>
__global__ void __launch_bounds__(kPipeThreads)
k_ident_msg_fwd_bulk(const float *__restrict__ V, const float *__restrict__ comp, const int32_t *__restrict__ colptr,
                     const int32_t *__restrict__ e2_src, const int32_t *__restrict__ e2_rel,
                     const float *__restrict__ e2_val, float *__restrict__ msg, IdentPipe p, int R, int CS) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *empty = full + p.S;
  float *comp_s = reinterpret_cast<float *>(smem_raw + 16 * ((2 * p.S * 8 + 15) / 16));
  unsigned char *stages = reinterpret_cast<unsigned char *>(comp_s + ((R * CS + 3) & ~3));
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kPipeConsumerWarps); }
    mbar_fence_init();
  }
  for (int r = threadIdx.x / 32; r < R; r += kPipeThreads / 32)
    for (int b = threadIdx.x & 31; b < p.B; b += 32) comp_s[r * CS + b] = __ldg(comp + (size_t)r * p.B + b);
  __syncthreads();
  const int B = p.B, out = p.out;
  const size_t bstride = (size_t)p.TJ * out;
  ident_pipeline(p, V, colptr, e2_src, e2_rel, e2_val, stages, full, empty,
                 [&](const float *vs, int j0, int e, int src, int rel, float v) {
                   const float *cr = comp_s + rel * CS;
                   const float *vrow = vs + (size_t)(src - j0) * out;
                   const int ms = msg_stride(out);
                   for (int c0 = 0; c0 < ms; c0 += OC) {
                     float acc[OC];
#pragma unroll
                     for (int o = 0; o < OC; ++o) acc[o] = 0.f;
                     if (c0 < out) mix_bases<OC, VW>(vrow + c0, bstride, cr, B, acc);
#pragma unroll
                     for (int o = 0; o < OC; ++o) acc[o] *= v;
                     store_msg_chunk<OC>(msg + (size_t)e * ms, c0, out, ms, acc);
                   }
                 });
}

// (b) generic variant (any alignment): cooperative coalesced loads into a padded tile.
//     smem: comp_s[R][CS] then Vs[B][TJ][OP] (OP = out rounded up to a multiple of OC; rows 16B aligned).
template <int OC>
__global__ void __launch_bounds__(kThreads)
k_ident_msg_fwd(const float *__restrict__ V, const float *__restrict__ comp, const int32_t *__restrict__ colptr,
                const int32_t *__restrict__ e2_src, const int32_t *__restrict__ e2_rel,
                const float *__restrict__ e2_val, float *__restrict__ msg, int NS, int R, int B, int out, int OP,
                int TJ, int CS, int comp_smem) {
  extern __shared__ __align__(16) float smem[];
  float *comp_s = smem;
  float *Vs = smem + (comp_smem ? ((R * CS + 3) & ~3) : 0);
  const int tid = threadIdx.x;
  if (comp_smem) load_comp_smem(comp_s, comp, R, B, CS);
  for (int j0 = blockIdx.x * TJ; j0 < NS; j0 += gridDim.x * TJ) {
    const int tjw = min(TJ, NS - j0);
    const int e_lo = colptr[j0], e_hi = colptr[j0 + tjw];
    if (e_lo == e_hi) continue;
    __syncthreads();  // previous tile fully consumed (and comp_s visible)
    const int run = tjw * out;
    for (int x = tid; x < run; x += kThreads) {
      int jl = x / out, o = x - jl * out;
      const float *src = V + (size_t)j0 * out + x;
      float *dst = Vs + jl * OP + o;
      for (int b0 = 0; b0 < B; b0 += 8) {  // 8 independent loads in flight per thread
        float tmp[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) tmp[q] = (b0 + q < B) ? __ldg(src + (size_t)(b0 + q) * NS * out) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (b0 + q < B) dst[(size_t)(b0 + q) * TJ * OP] = tmp[q];
      }
    }
    __syncthreads();
    ident_msg_tile<OC, 4>(Vs, (size_t)TJ * OP, OP, comp_smem ? comp_s : nullptr, CS, comp, B, out, j0, e_lo, e_hi, e2_src,
                          e2_rel, e2_val, msg);
  }
}

// ------------------------------------------------------------------------------------------------
// Feature term (graph.py:93-95 re-associated: no (R,N,out) projection), relation-major.
//   msg[e3, :] = val_e * X[j_e, :] . W[r, :, :]          one CTA = one chunk of edges of ONE relation
// smem: Ws[INP][OC] for the current output chunk (INP = in rounded up to KC, pad rows zero),
//       Xs[warp][2][32][KC+1]: rows of X for 32 edges, KC columns at a time, double buffered with cp.async
//       (coalesced 128-byte row segments in flight while the previous segment is multiplied).
constexpr int KC = 32;
// KCT = columns staged per round: 32, or 16 for narrow inputs (in <= 16: hidden layers, input-gradient messages), where a
// cp.async instruction then moves 16 columns of TWO rows so that no lane idles on zero padding.
template <int OC, int KCT>
__global__ void __launch_bounds__(kThreads)
k_feat_msg_fwd(const float *__restrict__ X, const float *__restrict__ W, const int32_t *__restrict__ chunk_rel,
               const int32_t *__restrict__ chunk_ptr, const int32_t *__restrict__ e3_src,
               const float *__restrict__ e3_val, float *__restrict__ msg, int in, int out, int INP, int ldx) {  // e3_src: gather index; ldx: row pitch of X
  extern __shared__ __align__(16) float smem[];
  float *Ws = smem;                         // [INP][OC]
  float *Xs_all = smem + (size_t)INP * OC;  // [nwarps][2][32][KCT+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = kThreads / 32;
  constexpr int XB = 32 * (KCT + 1);
  float *Xs = Xs_all + (size_t)warp * 2 * XB;
  const int c = blockIdx.x;
  const int r = chunk_rel[c];
  const int e_lo = chunk_ptr[c], e_hi = chunk_ptr[c + 1];
  const float *Wr = W + (size_t)r * in * out;
  const int nkc = INP / KCT;
  for (int c0 = 0; c0 < out; c0 += OC) {
    __syncthreads();
    for (int x = tid; x < INP * OC; x += kThreads) {
      int k = x / OC, o = x - k * OC;
      Ws[x] = (k < in && c0 + o < out) ? __ldg(Wr + (size_t)k * out + c0 + o) : 0.f;
    }
    __syncthreads();
    for (int eb = e_lo + warp * 32; eb < e_hi; eb += nwarps * 32) {
      const int e = eb + lane;
      const bool live = e < e_hi;
      const int j = live ? e3_src[e] : -1;
      auto prefetch = [&](int kc, int buf) {
        if constexpr (KCT == 32) {
          float *dst = Xs + buf * XB + lane;
          const int k = kc * KCT + lane;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const int ji = __shfl_sync(0xffffffffu, j, i);
            const bool ok = ji >= 0 && k < in;
            cp_async4(dst + i * (KCT + 1), X + (ok ? (size_t)ji * ldx + k : 0), ok);
          }
        } else {
          const int half = lane >> 4, col = lane & 15;
          float *dst = Xs + buf * XB + col;
          const int k = kc * KCT + col;
#pragma unroll 8
          for (int i = 0; i < 32; i += 2) {
            const int ji = __shfl_sync(0xffffffffu, j, i + half);
            const bool ok = ji >= 0 && k < in;
            cp_async4(dst + (i + half) * (KCT + 1), X + (ok ? (size_t)ji * ldx + k : 0), ok);
          }
        }
        cp_async_commit();
      };
      float2 acc[OC / 2];
#pragma unroll
      for (int o = 0; o < OC / 2; ++o) acc[o] = make_float2(0.f, 0.f);
      __syncwarp();
      prefetch(0, 0);
      for (int kc = 0; kc < nkc; ++kc) {
        if (kc + 1 < nkc) {
          prefetch(kc + 1, (kc + 1) & 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
        const float *xrow = Xs + (kc & 1) * XB + lane * (KCT + 1);
        const float *wk = Ws + (size_t)kc * KCT * OC;
#pragma unroll 4
        for (int kk = 0; kk < KCT; ++kk) {
          const float x = xrow[kk];
          const float4 *w4 = reinterpret_cast<const float4 *>(wk + kk * OC);
#pragma unroll
          for (int q = 0; q < OC / 4; ++q) {
            float4 t = w4[q];
            fma2(acc[2 * q], x, make_float2(t.x, t.y));
            fma2(acc[2 * q + 1], x, make_float2(t.z, t.w));
          }
        }
        __syncwarp();  // all lanes done with buffer (kc & 1) before it is refilled two chunks later
      }
      if (live) {
        const float v = e3_val[e];
        const int ms = msg_stride(out);
        float vals[OC];
#pragma unroll
        for (int o = 0; o < OC / 2; ++o) { vals[2 * o] = v * acc[o].x; vals[2 * o + 1] = v * acc[o].y; }
        store_msg_chunk<OC>(msg + (size_t)e * ms, c0, out, ms, vals);
        if (c0 + OC >= out)   // last computed chunk: zero the rest of the padded row
          for (int o = c0 + OC; o < ms; o += 4) *reinterpret_cast<float4 *>(msg + (size_t)e * ms + o) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Aggregation over destination rows (E1), fused with bias, row mask and ReLU (AggArgs: rgcn_internal.cuh).

// sum of msg[perm[e]][o] over e = lo+beg, lo+beg+step, ... < hi; four gathers in flight, fixed order
__device__ __forceinline__ float gather_sum(const float *__restrict__ msg, const int32_t *__restrict__ perm, int lo, int hi,
                                            int beg, int step, int od /* row stride */, int o, float acc) {
  int e = lo + beg;
  for (; e + 3 * step < hi; e += 4 * step) {
    const int p0 = perm[e], p1 = perm[e + step], p2 = perm[e + 2 * step], p3 = perm[e + 3 * step];
    const float v0 = msg[(size_t)p0 * od + o], v1 = msg[(size_t)p1 * od + o];
    const float v2 = msg[(size_t)p2 * od + o], v3 = msg[(size_t)p3 * od + o];
    acc += v0; acc += v1; acc += v2; acc += v3;
  }
  for (; e < hi; e += step) acc += msg[(size_t)perm[e] * od + o];
  return acc;
}

// sum over the sub-range [off, off+len) of row i's edges (len < 0: whole row), slots beg, beg+step, ...
__device__ __forceinline__ float agg_edges(const AggArgs &a, int i, int o, int beg, int step, int off = 0, int len = -1) {
  float acc = 0.f;
  const int od = a.odim;
  if (a.rowptr) {
    int lo = a.rowptr[i], hi = a.rowptr[i + 1];
    if (len >= 0) { lo = min(hi, lo + off); hi = min(hi, lo + len); }
    if (a.msgI) acc = gather_sum(a.msgI, a.pI, lo, hi, beg, step, a.ms, o, acc);
    if (a.Wd) {
      int e = lo + beg;
      for (; e + step < hi; e += 2 * step) {
        const float w0 = a.Wd[((size_t)a.d_rel[e] * a.NSd + a.d_src[e]) * od + o];
        const float w1 = a.Wd[((size_t)a.d_rel[e + step] * a.NSd + a.d_src[e + step]) * od + o];
        acc = fmaf(a.d_val[e], w0, acc);
        acc = fmaf(a.d_val[e + step], w1, acc);
      }
      for (; e < hi; e += step) acc = fmaf(a.d_val[e], a.Wd[((size_t)a.d_rel[e] * a.NSd + a.d_src[e]) * od + o], acc);
    }
  }
  if (a.msgF) {
    int lo = a.rowptrF[i], hi = a.rowptrF[i + 1];
    if (len >= 0) { lo = min(hi, lo + off); hi = min(hi, lo + len); }
    acc = gather_sum(a.msgF, a.pF, lo, hi, beg, step, a.ms, o, acc);
  }
  return acc;
}

__device__ __forceinline__ int row_degree(const AggArgs &a, int i) {
  int d = 0;
  if (a.rowptr) d = a.rowptr[i + 1] - a.rowptr[i];
  if (a.msgF) d = max(d, a.rowptrF[i + 1] - a.rowptrF[i]);
  return d;
}


// short rows: one sub-warp group of `odim` lanes per row (32/odim rows per warp), or a whole warp per row
__global__ void __launch_bounds__(kThreads) k_agg_fwd(AggArgs a) {
  const int od = a.odim;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (od >= 32) {
    int i = gw;
    if (i >= a.ND) return;
    if (a.thresh > 0 && row_degree(a, i) > a.thresh) return;
    for (int o = lane; o < od; o += 32) agg_store(a, i, o, agg_edges(a, i, o, 0, 1));
  } else {
    const int rpw = 32 / od;
    const int slot = lane / od, o = lane - slot * od;
    int i = gw * rpw + slot;
    if (slot >= rpw || i >= a.ND) return;
    if (a.thresh > 0 && row_degree(a, i) > a.thresh) return;
    agg_store(a, i, o, agg_edges(a, i, o, 0, 1));
  }
}

// wide rows (32 <= out <= 256, messages only): a warp per destination row keeps all NC = ceil(out/32) column chunks of the
// row in registers and walks the row's edges ONCE, two message rows in flight (k_agg_fwd re-walks the edge list and its
// permutation once per 32 columns).  Same summation order per element as k_agg_fwd: identical results.
template <int NC>
__global__ void __launch_bounds__(kThreads) k_agg_fwd_wide(AggArgs a) {
  const int lane = threadIdx.x & 31;
  const int i = (int)(((size_t)blockIdx.x * kThreads + threadIdx.x) >> 5);
  if (i >= a.ND) return;
  if (a.thresh > 0 && row_degree(a, i) > a.thresh) return;
  const int od = a.odim, ms = a.ms;
  float acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.f;
  auto walk = [&](const float *__restrict__ msg, const int32_t *__restrict__ perm, int lo, int hi) {
    int e = lo;
    for (; e + 1 < hi; e += 2) {
      const float *r0 = msg + (size_t)perm[e] * ms + lane, *r1 = msg + (size_t)perm[e + 1] * ms + lane;
      float v0[NC], v1[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const bool in = lane + 32 * c < od;
        v0[c] = in ? r0[32 * c] : 0.f;
        v1[c] = in ? r1[32 * c] : 0.f;
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) { acc[c] += v0[c]; acc[c] += v1[c]; }
    }
    if (e < hi) {
      const float *r0 = msg + (size_t)perm[e] * ms + lane;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (lane + 32 * c < od) acc[c] += r0[32 * c];
    }
  };
  if (a.rowptr && a.msgI) walk(a.msgI, a.pI, a.rowptr[i], a.rowptr[i + 1]);
  if (a.msgF) walk(a.msgF, a.pF, a.rowptrF[i], a.rowptrF[i + 1]);
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (lane + 32 * c < od) agg_store(a, i, lane + 32 * c, acc[c]);
}

// short rows, vector variant: MS/4 lanes per row, each lane owns 4 consecutive outputs and gathers 16-byte pieces of the
// (64-byte aligned) message rows; four gathers in flight; fixed order.
__device__ __forceinline__ void gather_sum4(const float *__restrict__ msg, const int32_t *__restrict__ perm, int lo, int hi,
                                            int ms, int o0, float4 &acc) {
  int e = lo;
  for (; e + 3 < hi; e += 4) {
    const int p0 = perm[e], p1 = perm[e + 1], p2 = perm[e + 2], p3 = perm[e + 3];
    const float4 v0 = *reinterpret_cast<const float4 *>(msg + (size_t)p0 * ms + o0);
    const float4 v1 = *reinterpret_cast<const float4 *>(msg + (size_t)p1 * ms + o0);
    const float4 v2 = *reinterpret_cast<const float4 *>(msg + (size_t)p2 * ms + o0);
    const float4 v3 = *reinterpret_cast<const float4 *>(msg + (size_t)p3 * ms + o0);
    acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
    acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
    acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
    acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
  }
  for (; e < hi; ++e) {
    const float4 v = *reinterpret_cast<const float4 *>(msg + (size_t)perm[e] * ms + o0);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
}
__global__ void __launch_bounds__(kThreads) k_agg_fwd_v4(AggArgs a) {
  const int od = a.odim, ms = a.ms;
  const int G = ms >> 2;                       // lanes per row (1, 2, 4, 8, ...)
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int rpw = 32 / G;
  const int slot = lane / G, o0 = (lane - slot * G) * 4;
  const int i = gw * rpw + slot;
  if (slot >= rpw || i >= a.ND || o0 >= od) return;   // G need not divide 32: the spare lanes must not start the next warp's row
  if (a.thresh > 0 && row_degree(a, i) > a.thresh) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.rowptr) {
    const int lo = a.rowptr[i], hi = a.rowptr[i + 1];
    if (a.msgI) gather_sum4(a.msgI, a.pI, lo, hi, ms, o0, acc);
    if (a.Wd) {
      for (int e = lo; e < hi; ++e) {
        const float *wr = a.Wd + ((size_t)a.d_rel[e] * a.NSd + a.d_src[e]) * od + o0;
        const float v = a.d_val[e];
        acc.x = fmaf(v, wr[0], acc.x);
        if (o0 + 1 < od) acc.y = fmaf(v, wr[1], acc.y);
        if (o0 + 2 < od) acc.z = fmaf(v, wr[2], acc.z);
        if (o0 + 3 < od) acc.w = fmaf(v, wr[3], acc.w);
      }
    }
  }
  if (a.msgF) gather_sum4(a.msgF, a.pF, a.rowptrF[i], a.rowptrF[i + 1], ms, o0, acc);
  agg_store(a, i, o0, acc.x);
  if (o0 + 1 < od) agg_store(a, i, o0 + 1, acc.y);
  if (o0 + 2 < od) agg_store(a, i, o0 + 2, acc.z);
  if (o0 + 3 < od) agg_store(a, i, o0 + 3, acc.w);
}

// hubs: one CTA (1024 threads) per SEGMENT of a long row; edge slots strided over the segment, fixed-order tree over
// the slots.  A single-segment hub is finished here; otherwise the partial goes to ws[seg] for k_agg_combine.
constexpr int kLongThreads = 1024;
__global__ void __launch_bounds__(kLongThreads) k_agg_fwd_long(AggArgs a, HubSegs h) {
  extern __shared__ float red[];  // [nslots][oc]
  const int od = a.odim;
  const int sg = blockIdx.x;
  const int hub = h.seg_hub[sg];
  const int i = h.long_ids[hub];
  const int first = h.seg_first[hub], nseg = h.seg_first[hub + 1] - first;
  const int off = (sg - first) * h.seg;
  const int oc = min(od, kLongThreads);
  const int nslots = kLongThreads / oc;
  const int slot = threadIdx.x / oc, ol = threadIdx.x - slot * oc;
  for (int o0 = 0; o0 < od; o0 += oc) {
    const int o = o0 + ol;
    float acc = 0.f;
    if (slot < nslots && o < od) acc = agg_edges(a, i, o, slot, nslots, off, h.seg);
    if (slot < nslots) red[slot * oc + ol] = acc;
    __syncthreads();
    for (int s = 1; s < nslots; s <<= 1) {
      if (slot < nslots && (slot % (2 * s)) == 0 && slot + s < nslots) red[slot * oc + ol] += red[(slot + s) * oc + ol];
      __syncthreads();
    }
    if (slot == 0 && o < od) {
      if (nseg == 1) agg_store(a, i, o, red[ol]);
      else h.ws[(size_t)sg * od + o] = red[ol];
    }
    __syncthreads();
  }
}
__global__ void k_agg_combine(AggArgs a, HubSegs h) {
  const int hub = blockIdx.x;
  const int first = h.seg_first[hub], nseg = h.seg_first[hub + 1] - first;
  if (nseg <= 1) return;
  const int i = h.long_ids[hub];
  for (int o = threadIdx.x; o < a.odim; o += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < nseg; ++s) acc += h.ws[(size_t)(first + s) * a.odim + o];
    agg_store(a, i, o, acc);
  }
}

template <class K>
static unsigned persistent_grid(K kernel, int threads, size_t smem, int64_t max_ctas) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  int64_t g = (int64_t)kNumSMs * per_sm;
  return (unsigned)(g < max_ctas ? g : (max_ctas > 0 ? max_ctas : 1));
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) MRGCN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace

// ---- launchers (also used by rgcn_bwd.cu) ------------------------------------------------------
int launch_basis_mix_fwd(const float *comp, const float *V, float *W, int R, int B, int IO, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(IO, 128), (unsigned)R);
  MRGCN_PROF("basis_mix_fwd");
  k_basis_mix_fwd<<<grid, 128, 0, st>>>(comp, V, W, B, IO);
  MRGCN_LAUNCH_CHECK();
  return 0;
}

int pick_oc(int out) { return out <= 4 ? 4 : out <= 8 ? 8 : out <= 12 ? 12 : 16; }

// tile of sources for the identity-term kernels: tjw*out <= 256 (one coalesced run per basis), smem bounded
int ident_tile(int B, int out, int OP) {
  int tj = 256 / out;
  if (tj < 1) tj = 1;
  while (tj > 1 && (size_t)B * tj * OP * 4 > 56 * 1024) --tj;
  return tj;
}

// tile of the TMA-engine variants: ~48 KB of V per stage, 16-byte runs; 0 = not applicable
int ident_pipe_config(IdentPipe &p, int64_t NS, int B, int out, size_t other_smem) {
  if ((NS * out) % 4 != 0) return 0;
  int tj = (48 * 1024) / (B * out * 4);
  if (tj > 64) tj = 64;
  while (tj > 0 && (tj * out) % 4 != 0) --tj;
  if (tj <= 0 || tj > NS) return 0;
  p.NS = (int)NS; p.B = B; p.out = out; p.TJ = tj;
  p.ntiles = (int)cdiv(NS, tj);
  p.mcap = 1024;
  p.v_floats = ((B * tj * out + 3) & ~3) + 16;
  p.stage_bytes = ident_pipe_stage_bytes(p.v_floats, p.mcap);
  for (p.S = 4; p.S >= 2; --p.S)
    if (16 * ((2 * p.S * 8 + 15) / 16) + other_smem + (size_t)p.S * p.stage_bytes <= 200 * 1024) return 1;
  return 0;
}

static int launch_ident_msg_fwd(const mrgcn_graph *g, const float *V, const float *comp, float *msg, int B, int out,
                                cudaStream_t st) {
  const int OC = pick_oc(out);
  const int CS = B | 1;
  const int comp_smem = ((size_t)g->R * CS * 4 <= 64 * 1024) ? 1 : 0;
  const size_t comp_bytes = (comp_smem ? (((size_t)g->R * CS + 3) & ~(size_t)3) : 0) * 4;
  unsigned grid = 0;
  IdentPipe p;
  if (comp_smem && ident_pipe_config(p, g->NS, B, out, comp_bytes)) {
    const int VW = (out % 4 == 0) ? 4 : (out % 2 == 0) ? 2 : 1;
    const size_t smem = 16 * ((2 * p.S * 8 + 15) / 16) + comp_bytes + (size_t)p.S * p.stage_bytes;
    MRGCN_PROF("ident_msg_fwd");
#define LAUNCH(OCV, VWV)                                                                                            \
  do {                                                                                                              \
    if (int rc = set_smem(k_ident_msg_fwd_bulk<OCV, VWV>, smem)) return rc;                                         \
    grid = persistent_grid(k_ident_msg_fwd_bulk<OCV, VWV>, kPipeThreads, smem, p.ntiles);                           \
    k_ident_msg_fwd_bulk<OCV, VWV><<<grid, kPipeThreads, smem, st>>>(V, comp, g->colptr, g->e2_src, g->e2_rel,     \
                                                                     g->e2_val, msg, p, g->R, CS);                  \
  } while (0)
#define LAUNCH_VW(OCV)                \
  do {                                \
    if (VW == 4) LAUNCH(OCV, 4);      \
    else if (VW == 2) LAUNCH(OCV, 2); \
    else LAUNCH(OCV, 1);              \
  } while (0)
    switch (OC) {
      case 4: LAUNCH_VW(4); break;
      case 8: LAUNCH_VW(8); break;
      case 12: LAUNCH_VW(12); break;
      default: LAUNCH_VW(16); break;
    }
#undef LAUNCH_VW
#undef LAUNCH
    MRGCN_LAUNCH_CHECK();
    return 0;
  }
  const int OP = (int)cdiv(out, OC) * OC;
  const int TJ = ident_tile(B, out, OP);
  size_t smem = comp_bytes + (size_t)B * TJ * OP * 4;
  MRGCN_REQUIRE(smem <= 220 * 1024, MRGCN_E_NOTSUP, "ident_msg_fwd: B*out too large for shared memory (%zu B)", smem);
#define LAUNCH(OCV)                                                                                               \
  do {                                                                                                            \
    if (int rc = set_smem(k_ident_msg_fwd<OCV>, smem)) return rc;                                                 \
    grid = persistent_grid(k_ident_msg_fwd<OCV>, kThreads, smem, cdiv(g->NS, TJ));                                \
    k_ident_msg_fwd<OCV><<<grid, kThreads, smem, st>>>(V, comp, g->colptr, g->e2_src, g->e2_rel, g->e2_val, msg, \
                                                       g->NS, g->R, B, out, OP, TJ, CS, comp_smem);              \
  } while (0)
  MRGCN_PROF("ident_msg_fwd");
  switch (OC) {
    case 4: LAUNCH(4); break;
    case 8: LAUNCH(8); break;
    case 12: LAUNCH(12); break;
    default: LAUNCH(16); break;
  }
#undef LAUNCH
  MRGCN_LAUNCH_CHECK();
  return 0;
}

int launch_feat_msg(const mrgcn_graph *g, const int32_t *gather, const float *X, int ldx, const float *W, float *msg, int in,
                    int out, cudaStream_t st, const char *prof_name) {
  if (g->n_chunks == 0) return 0;
  const int OC = pick_oc(out);
  const int KCT = in <= 16 ? 16 : KC;
  const int INP = (int)cdiv(in, KCT) * KCT;
  size_t smem = ((size_t)INP * OC + (size_t)(kThreads / 32) * 2 * 32 * (KCT + 1)) * 4;
  MRGCN_REQUIRE(smem <= 220 * 1024, MRGCN_E_NOTSUP, "feat_msg: in too large for shared memory (%zu B)", smem);
#define LAUNCH(OCV, KV)                                                                                            \
  do {                                                                                                             \
    if (int rc = set_smem(k_feat_msg_fwd<OCV, KV>, smem)) return rc;                                               \
    k_feat_msg_fwd<OCV, KV><<<(unsigned)g->n_chunks, kThreads, smem, st>>>(X, W, g->chunk_rel, g->chunk_ptr,      \
                                                                          gather, g->e3_val, msg, in, out, INP, ldx); \
  } while (0)
#define LAUNCH_K(OCV)                 \
  do {                                \
    if (KCT == 16) LAUNCH(OCV, 16);   \
    else LAUNCH(OCV, 32);             \
  } while (0)
  mrgcn::prof_begin(prof_name, st);
  switch (OC) {
    case 4: LAUNCH_K(4); break;
    case 8: LAUNCH_K(8); break;
    case 12: LAUNCH_K(12); break;
    default: LAUNCH_K(16); break;
  }
#undef LAUNCH_K
#undef LAUNCH
  MRGCN_LAUNCH_CHECK();
  return 0;
}

int launch_agg(const AggArgs &g, const HubSegs &h, cudaStream_t st, const char *prof_name) {
  if (g.ND <= 0) return 0;
  mrgcn::prof_begin(prof_name, st);
  if (g.ms <= 128) {
    const int rows_per_warp = 32 / (g.ms >> 2);
    unsigned grid = (unsigned)cdiv(cdiv(g.ND, rows_per_warp) * 32, kThreads);
    k_agg_fwd_v4<<<grid, kThreads, 0, st>>>(g);
  } else if (g.odim >= 32 && g.odim <= 256 && !g.Wd && (g.msgI || g.msgF)) {
    unsigned grid = (unsigned)cdiv((int64_t)g.ND * 32, kThreads);
    switch ((g.odim + 31) / 32) {
      case 1: k_agg_fwd_wide<1><<<grid, kThreads, 0, st>>>(g); break;
      case 2: k_agg_fwd_wide<2><<<grid, kThreads, 0, st>>>(g); break;
      case 3: k_agg_fwd_wide<3><<<grid, kThreads, 0, st>>>(g); break;
      case 4: k_agg_fwd_wide<4><<<grid, kThreads, 0, st>>>(g); break;
      case 5: k_agg_fwd_wide<5><<<grid, kThreads, 0, st>>>(g); break;
      case 6: k_agg_fwd_wide<6><<<grid, kThreads, 0, st>>>(g); break;
      case 7: k_agg_fwd_wide<7><<<grid, kThreads, 0, st>>>(g); break;
      default: k_agg_fwd_wide<8><<<grid, kThreads, 0, st>>>(g); break;
    }
  } else {
    const int rows_per_warp = g.odim >= 32 ? 1 : 32 / g.odim;
    unsigned grid = (unsigned)cdiv(cdiv(g.ND, rows_per_warp) * 32, kThreads);
    k_agg_fwd<<<grid, kThreads, 0, st>>>(g);
  }
  MRGCN_LAUNCH_CHECK();
  if (h.n_long > 0) {
    MRGCN_REQUIRE(h.ws || h.n_segs == h.n_long, MRGCN_E_BADARG, "agg: hub_ws missing");
    mrgcn::prof_begin("agg_long", st);
    k_agg_fwd_long<<<(unsigned)h.n_segs, kLongThreads, kLongThreads * sizeof(float), st>>>(g, h);
    MRGCN_LAUNCH_CHECK();
    if (h.n_segs > h.n_long) {
      mrgcn::prof_begin("agg_combine", st);
      k_agg_combine<<<(unsigned)h.n_long, 128, 0, st>>>(g, h);
      MRGCN_LAUNCH_CHECK();
    }
  }
  return 0;
}

}  // namespace mrgcn

using namespace mrgcn;

extern "C" int32_t mrgcn_msg_stride(int32_t out) { return msg_stride(out); }

extern "C" int mrgcn_rgcn_layer_fwd(const mrgcn_layer_args *a, mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(a && a->out && a->out_dim > 0, MRGCN_E_BADARG, "layer_fwd: null/empty output");
  const bool hasI = a->weight_I != nullptr, hasF = a->X != nullptr;
  MRGCN_REQUIRE(hasI || hasF, MRGCN_E_BADARG, "layer_fwd: neither identity nor feature term");
  MRGCN_REQUIRE(!hasI || a->gI, MRGCN_E_BADARG, "layer_fwd: identity term without graph");
  MRGCN_REQUIRE(!hasF || (a->gF && a->weight_F && a->in_dim > 0), MRGCN_E_BADARG, "layer_fwd: feature term incomplete");
  const int B = a->B > 0 ? a->B : 0, out = a->out_dim, in = a->in_dim;
  const int ldx = a->x_stride > 0 ? a->x_stride : in;
  MRGCN_REQUIRE(ldx >= in, MRGCN_E_BADARG, "layer_fwd: x_stride smaller than in_dim");
  const mrgcn_graph *gI = a->gI, *gF = a->gF;
  const int ND = hasI ? gI->ND : gF->ND;
  MRGCN_REQUIRE(!(hasI && hasF) || gI->ND == gF->ND, MRGCN_E_BADARG, "layer_fwd: graphs disagree on rows");

  AggArgs g{};
  g.ND = ND; g.odim = out; g.ms = msg_stride(out); g.relu = a->relu; g.bias = a->bias; g.mask = a->row_mask; g.addend = a->addend; g.out = a->out;
  const mrgcn_graph *gl = hasI ? gI : gF;  // long-row list owner
  bool fused_feat = false;
  if (hasI) {
    g.rowptr = gI->rowptr;
    if (B > 0) {
      MRGCN_REQUIRE(a->comp_I && a->msg_I, MRGCN_E_BADARG, "layer_fwd: comp_I/msg_I missing");
      TabGeom tg;
      // projected features: the feature term rides in the identity term's messages (one table pass, one message per edge)
      fused_feat = hasF && a->proj && a->plan && gI == gF && (mrgcn_tab_mode(B, B, out) & 1) && feat_proj_supported(in, ldx, B, out) && tab_geometry(2 * B, out, tg);
      if (fused_feat) {
        MRGCN_REQUIRE(a->comp_F && a->vt_ws, MRGCN_E_BADARG, "layer_fwd: comp_F/vt_ws missing");
        if (int rc = launch_feat_proj(a->X, gI->NS, in, ldx, a->weight_F, B, out, a->vt_ws, a->xpad_ws, a->proj, st)) return rc;
      }
      if (gI->E > 0) {
        if (fused_feat) {
          if (int rc = launch_tab_msg_fwd(gI, a->plan, a->weight_I, a->comp_I, B, a->proj, a->comp_F, B, out, a->msg_I, st)) return rc;
        } else if (a->plan && (mrgcn_tab_mode(B, 0, out) & 1) && tab_geometry(B, out, tg)) {
          if (int rc = launch_tab_msg_fwd(gI, a->plan, a->weight_I, a->comp_I, B, nullptr, nullptr, 0, out, a->msg_I, st)) return rc;
        } else {
          if (int rc = launch_ident_msg_fwd(gI, a->weight_I, a->comp_I, a->msg_I, B, out, st)) return rc;
        }
      }
      g.pI = gI->e1_to_e2; g.msgI = a->msg_I;
    } else {
      g.Wd = a->weight_I; g.d_src = gI->e1_src; g.d_rel = gI->e1_rel; g.d_val = gI->e1_val; g.NSd = gI->NS;
    }
  }
  if (hasF) {
    const float *W = a->weight_F;
    if (B > 0) {
      MRGCN_REQUIRE(a->comp_F && a->wmix, MRGCN_E_BADARG, "layer_fwd: comp_F/wmix missing");
      if (int rc = launch_basis_mix_fwd(a->comp_F, a->weight_F, a->wmix, gF->R, B, in * out, st)) return rc;
      W = a->wmix;
    }
    if (!hasI && narrow_supported(gF->R, in, out)) {
      // feature-only narrow layer: gather, transform and aggregate in one pass (narrow.cu), no message buffer
      NarrowArgs n{gF->rows_by_deg, gF->rowptr, gF->e1_src, gF->e1_rel, gF->e1_val, a->X, W, ldx, gF->R, in, out, g};
      n.epi.thresh = gF->n_long_rows > 0 ? gF->long_row_thresh : 0;
      HubSegs hs{gF->long_rows, gF->row_seg_hub, gF->row_seg_first, gF->n_long_rows, gF->n_row_segs, gF->long_seg, a->hub_ws};
      return launch_narrow(n, hs, st, "narrow_fwd");
    }
    if (!fused_feat) {
      MRGCN_REQUIRE(a->msg_F, MRGCN_E_BADARG, "layer_fwd: msg_F missing");
      if (gF->E > 0)
        if (int rc = launch_feat_msg(gF, gF->e3_src, a->X, ldx, W, a->msg_F, in, out, st, "feat_msg_fwd")) return rc;
      g.pF = gF->e1_to_e3; g.msgF = a->msg_F; g.rowptrF = gF->rowptr;
    }
  }
  // Long rows are taken from the owner graph (gI when there is an identity term).  The two graphs differ only in
  // mini-batch mode, where gF is a column slice of gI's adjacency, so a row of gF is never longer than the same row of
  // gI; a caller that breaks this (E_F > E_I) is refused instead of leaving hub rows unwritten.
  MRGCN_REQUIRE(!(hasI && hasF) || gF->E <= gI->E, MRGCN_E_BADARG, "layer_fwd: feature graph has more entries than the identity graph");
  g.thresh = gl->n_long_rows > 0 ? gl->long_row_thresh : 0;
  HubSegs hs{gl->long_rows, gl->row_seg_hub, gl->row_seg_first, gl->n_long_rows, gl->n_row_segs, gl->long_seg, a->hub_ws};
  if (int rc = launch_agg(g, hs, st, "agg_fwd")) return rc;
  return 0;
}
