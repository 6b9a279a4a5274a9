// R-GCN layer backward (replaces autograd through GraphConvolution.forward,
// /root/reference/mrgcn/layers/graph.py:62-102).
//
// With t_e = val_e * gact[dst_e, :] (gact = dL/d(pre-activation)):
//   g_bias[o]          = sum_i gact[i,o]                                   two-stage column sum
//   identity, B == 0   : g_weight_I[r*NS+j,:] = sum over the (j,r) run of E2 of t_e
//   identity, B  > 0   : g_weight_I[b*NS+j,:] = sum_{e: src=j} comp_I[r_e,b] * t_e          (E2, per source)
//                        cbuf[e2,b] = <V_I[b,j_e,:], t_e>  ->  g_comp_I[r,b] = sum_{e in r} cbuf   (E3 chunks)
//   feature            : g_W[r,k,o] = sum_{e in r} X[j_e,k] * t_e[o]                          (E3 chunks)
//                        B > 0: g_weight_F[b] = sum_r comp_F[r,b] g_W[r],  g_comp_F[r,b] = <weight_F[b], g_W[r]>
//                        g_X[j,k] = sum_{e: src=j} sum_o t_e[o] * W[r_e,k,o]                   (E2, per source)
// Every reduction is a segmented sum in a fixed order: bit-reproducible, no float atomics.
#include <stdlib.h>

#include "common.cuh"
#include "pipeline.cuh"
#include "ident_pipe.cuh"
#include "rgcn_internal.cuh"

namespace mrgcn {
namespace {

constexpr int kThreads = 256;
constexpr int kColRows = 1024;  // rows per stage-1 block of the bias column sum

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

// gact = gout * relu'(out) * mask ; stage 1 of the bias column sum (rgcn.py:82-87 backwards)
__global__ void __launch_bounds__(kThreads)
k_act_bwd(const float *__restrict__ gout, const float *__restrict__ outv, const float *__restrict__ mask,
          float *__restrict__ gact, float *__restrict__ colsum, int ND, int od, int relu) {
  extern __shared__ double red[];  // [kThreads]; the (small) bias sums are accumulated in double: the result is then as
                                   // close to the exact sum as the fp32 inputs allow, whatever the order
  // block (x, y): rows [1024 x, 1024 x + 1024), columns [oc y, oc y + oc) with oc = min(od, 64): wide layers get several
  // blocks per row range (a 14 541 x 200 layer had 15 blocks in all), every thread keeps four rows' loads in flight
  const int r0 = blockIdx.x * kColRows, r1 = min(ND, r0 + kColRows);
  const int oc = min(od, 64), nslots = kThreads / oc;
  const int slot = threadIdx.x / oc, ol = threadIdx.x - slot * oc;
  const int o = blockIdx.y * oc + ol;
  const bool live = slot < nslots && o < od;
  double acc = 0.0;
  auto one = [&](int i, float g, float ov) {
    if (relu && !(ov > 0.f)) g = 0.f;
    if (mask) g *= mask[i];
    gact[(size_t)i * od + o] = g;
    acc += g;
  };
  if (live) {
    int i = r0 + slot;
    for (; i + 3 * nslots < r1; i += 4 * nslots) {
      const size_t x0 = (size_t)i * od + o, st = (size_t)nslots * od;
      const float g0 = gout[x0], g1 = gout[x0 + st], g2 = gout[x0 + 2 * st], g3 = gout[x0 + 3 * st];
      float v0 = 1.f, v1 = 1.f, v2 = 1.f, v3 = 1.f;
      if (relu) { v0 = outv[x0]; v1 = outv[x0 + st]; v2 = outv[x0 + 2 * st]; v3 = outv[x0 + 3 * st]; }
      one(i, g0, v0); one(i + nslots, g1, v1); one(i + 2 * nslots, g2, v2); one(i + 3 * nslots, g3, v3);
    }
    for (; i < r1; i += nslots) one(i, gout[(size_t)i * od + o], relu ? outv[(size_t)i * od + o] : 1.f);
  }
  if (slot < nslots) red[slot * oc + ol] = acc;
  __syncthreads();
  for (int s = 1; s < nslots; s <<= 1) {
    if (slot < nslots && (slot % (2 * s)) == 0 && slot + s < nslots) red[slot * oc + ol] += red[(slot + s) * oc + ol];
    __syncthreads();
  }
  if (colsum && slot == 0 && o < od) colsum[(size_t)blockIdx.x * od + o] = (float)red[ol];
}

// out[s, x] = sum_c part[idx(c)*width + x] for c in [seg_ptr[s], seg_ptr[s+1]), idx = seg_idx[c] (or c) -- fixed order: 8 chunk slots stride the
// segment (each slot sequential), then a fixed tree over the slots.  seg_ptr == NULL: one segment [0, nall).
__global__ void __launch_bounds__(256)
k_seq_reduce(const float *__restrict__ part, const int32_t *__restrict__ seg_ptr, const int32_t *__restrict__ seg_idx,
             int nall, int width, float *__restrict__ outp) {
  __shared__ double red[8][32];   // partial sums of thousands of fp32 terms: accumulated in double (tiny kernels)
  const int s = blockIdx.y;
  const int xl = threadIdx.x & 31, slot = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + xl;
  const int lo = seg_ptr ? seg_ptr[s] : 0, hi = seg_ptr ? seg_ptr[s + 1] : nall;
  double a0 = 0.0, a1 = 0.0;
  if (x < width) {
    int c = lo + slot;
    for (; c + 8 < hi; c += 16) {
      a0 += part[(size_t)(seg_idx ? seg_idx[c] : c) * width + x];
      a1 += part[(size_t)(seg_idx ? seg_idx[c + 8] : c + 8) * width + x];
    }
    if (c < hi) a0 += part[(size_t)(seg_idx ? seg_idx[c] : c) * width + x];
  }
  red[slot][xl] = a0 + a1;
  __syncthreads();
  if (slot == 0 && x < width) {
    double t = ((red[0][xl] + red[1][xl]) + (red[2][xl] + red[3][xl])) + ((red[4][xl] + red[5][xl]) + (red[6][xl] + red[7][xl]));
    outp[(size_t)s * width + x] = (float)t;
  }
}

// ---- identity term, B == 0: one thread group per E2 edge that starts a (src, rel) run ------------
__global__ void __launch_bounds__(kThreads)
k_ident_bwd_direct(const int32_t *__restrict__ e2_src, const int32_t *__restrict__ e2_rel,
                   const int32_t *__restrict__ e2_dst, const float *__restrict__ e2_val,
                   const float *__restrict__ gact, float *__restrict__ gW, int64_t E, int64_t NS, int od) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  int64_t e; int o0, ostep;
  if (od >= 32) { e = gw; o0 = lane; ostep = 32; }
  else {
    const int epw = 32 / od, slot = lane / od;
    if (slot >= epw) return;
    e = gw * epw + slot; o0 = lane - slot * od; ostep = od;
  }
  if (e >= E) return;
  const int j = e2_src[e], r = e2_rel[e];
  if (e > 0 && e2_src[e - 1] == j && e2_rel[e - 1] == r) return;  // not a run head
  int64_t end = e + 1;
  while (end < E && e2_src[end] == j && e2_rel[end] == r) ++end;
  for (int o = o0; o < od; o += ostep) {
    float acc = 0.f;
    for (int64_t q = e; q < end; ++q) acc = fmaf(e2_val[q], gact[(size_t)e2_dst[q] * od + o], acc);
    gW[((size_t)r * NS + j) * od + o] = acc;
  }
}

// ---- identity term, B > 0, comp gradient contributions: cbuf[e2, b] = <V[b, j_e, :], t_e> ----------
// Same tiling / TMA-engine staging as k_ident_msg_fwd_bulk (rgcn_fwd.cu); one edge per lane, the lane's B
// results are 160 contiguous bytes of cbuf (16-byte stores when B % 4 == 0).
// per-edge body shared by both variants: B dot products <V[b, j, :], t_e> -> cbuf[e, 0:B]
template <int OC, int VW>
__device__ __forceinline__ void ident_bwd_c_edge(const float *__restrict__ vrow, size_t bstride, int B, int out, float v,
                                                 const float *__restrict__ gp, float *__restrict__ cp) {
  for (int c0 = 0; c0 < out; c0 += OC) {
    float t[OC];
#pragma unroll
    for (int o = 0; o < OC; ++o) t[o] = (c0 + o < out) ? v * gp[c0 + o] : 0.f;
    const float *vp = vrow + c0;
    auto dot = [&](int b) {
      const float *row = vp + (size_t)b * bstride;
      float acc = 0.f;
      if constexpr (VW == 4) {
        float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < OC / 4; ++q) {
          float4 w = reinterpret_cast<const float4 *>(row)[q];
          fma2v(a2, make_float2(w.x, w.y), make_float2(t[4 * q + 0], t[4 * q + 1]));
          fma2v(a2, make_float2(w.z, w.w), make_float2(t[4 * q + 2], t[4 * q + 3]));
        }
        acc = a2.x + a2.y;
      } else if constexpr (VW == 2) {
        float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < OC / 2; ++q)
          fma2v(a2, reinterpret_cast<const float2 *>(row)[q], make_float2(t[2 * q + 0], t[2 * q + 1]));
        acc = a2.x + a2.y;
      } else {
#pragma unroll
        for (int q = 0; q < OC; ++q)
          if (c0 + q < out) acc = fmaf(row[q], t[q], acc);
      }
      return acc;
    };
    if ((B & 3) == 0) {
#pragma unroll 2
      for (int b = 0; b < B; b += 4) {
        float4 c = make_float4(dot(b), dot(b + 1), dot(b + 2), dot(b + 3));
        float4 *dst = reinterpret_cast<float4 *>(cp + b);
        if (c0 > 0) { float4 q = *dst; c.x += q.x; c.y += q.y; c.z += q.z; c.w += q.w; }
        *dst = c;
      }
    } else {
      for (int b = 0; b < B; ++b) {
        float c = dot(b);
        if (c0 > 0) c += cp[b];
        cp[b] = c;
      }
    }
  }
}

template <int OC, int VW>
__global__ void __launch_bounds__(kPipeThreads)
k_ident_bwd_c_bulk(const float *__restrict__ V, const int32_t *__restrict__ colptr, const int32_t *__restrict__ e2_src,
                   const int32_t *__restrict__ e2_dst, const float *__restrict__ e2_val, const float *__restrict__ gact,
                   float *__restrict__ cbuf, IdentPipe p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *empty = full + p.S;
  unsigned char *stages = smem_raw + 16 * ((2 * p.S * 8 + 15) / 16);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kPipeConsumerWarps); }
    mbar_fence_init();
  }
  __syncthreads();
  const int B = p.B, out = p.out;
  const size_t bstride = (size_t)p.TJ * out;
  ident_pipeline(p, V, colptr, e2_src, e2_dst, e2_val, stages, full, empty,
                 [&](const float *vs, int j0, int e, int src, int dst, float v) {
                   ident_bwd_c_edge<OC, VW>(vs + (size_t)(src - j0) * out, bstride, B, out, v, gact + (size_t)dst * out,
                                            cbuf + (size_t)e * B);
                 });
}

// generic variant (any alignment): cooperative loads into a padded tile Vs[B][TJ][OP], pad columns zeroed
template <int OC>
__global__ void __launch_bounds__(kThreads)
k_ident_bwd_c(const float *__restrict__ V, const int32_t *__restrict__ colptr, const int32_t *__restrict__ e2_src,
              const int32_t *__restrict__ e2_dst, const float *__restrict__ e2_val, const float *__restrict__ gact,
              float *__restrict__ cbuf, int NS, int B, int out, int OP, int TJ) {
  extern __shared__ __align__(16) float smem[];
  float *Vs = smem;  // [B][TJ][OP]
  const int tid = threadIdx.x;
  for (int j0 = blockIdx.x * TJ; j0 < NS; j0 += gridDim.x * TJ) {
    const int tjw = min(TJ, NS - j0);
    const int e_lo = colptr[j0], e_hi = colptr[j0 + tjw];
    if (e_lo == e_hi) continue;
    __syncthreads();
    const int run = tjw * out;
    for (int x = tid; x < run; x += kThreads) {
      int jl = x / out, o = x - jl * out;
      const float *src = V + (size_t)j0 * out + x;
      float *dst = Vs + jl * OP + o;
      for (int b0 = 0; b0 < B; b0 += 8) {
        float tmp[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) tmp[q] = (b0 + q < B) ? __ldg(src + (size_t)(b0 + q) * NS * out) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (b0 + q < B) dst[(size_t)(b0 + q) * TJ * OP] = tmp[q];
      }
    }
    if (OP > out)  // padding columns meet t = 0 but must not be NaN
      for (int x = tid; x < B * tjw * (OP - out); x += kThreads) {
        int row = x / (OP - out), o = out + x % (OP - out);
        int b = row / tjw, jl = row - b * tjw;
        Vs[((size_t)b * TJ + jl) * OP + o] = 0.f;
      }
    __syncthreads();
    for (int e = e_lo + tid; e < e_hi; e += kThreads)
      ident_bwd_c_edge<OC, 4>(Vs + (size_t)(e2_src[e] - j0) * OP, (size_t)TJ * OP, B, out, e2_val[e],
                              gact + (size_t)e2_dst[e] * out, cbuf + (size_t)e * B);
  }
}

// Wide outputs (out >= 64, few bases: the link-prediction encoders, out 200 / B 2): a WARP per task of the table work plan
// (<= 32 consecutive E2 edges of one source).  The lanes split the columns: the source's rows V[b, j, :] sit in registers
// (NC = ceil(out / 32) floats per lane and basis, two bases at a time), every edge's gact row is read with NC coalesced
// loads and the B dot products are finished with a fixed shuffle tree.  The lane-per-edge kernel above walks an 800-byte
// row per lane (uncoalesced): 2.4 ms on the YAGO3-10+ shape where this one takes the time of the row reads.
template <int NC>
__global__ void __launch_bounds__(256)
k_ident_bwd_c_task(const float *__restrict__ V, const int4 *__restrict__ tasks, int n_tasks, const int32_t *__restrict__ e2_dst,
                   const float *__restrict__ e2_val, const float *__restrict__ gact, float *__restrict__ cbuf, int64_t NS, int B,
                   int out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int ti = blockIdx.x * wpb + (threadIdx.x >> 5); ti < n_tasks; ti += gridDim.x * wpb) {
    const int4 tk = __ldg(tasks + ti);   // {source, first edge, number of edges, 0}
    const int j = tk.x, e0 = tk.y, n = tk.z;
    const int dst_l = lane < n ? e2_dst[e0 + lane] : 0;
    const float val_l = lane < n ? e2_val[e0 + lane] : 0.f;
    for (int b0 = 0; b0 < B; b0 += 2) {
      float v0[NC], v1[NC];
      const bool two = b0 + 1 < B;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int o = lane + 32 * c;
        v0[c] = o < out ? __ldg(V + ((size_t)b0 * NS + j) * out + o) : 0.f;
        v1[c] = (two && o < out) ? __ldg(V + ((size_t)(b0 + 1) * NS + j) * out + o) : 0.f;
      }
      for (int i = 0; i < n; i += 2) {
        // two edges per trip: their row loads are in flight together
        const int d0 = __shfl_sync(0xffffffffu, dst_l, i), d1 = __shfl_sync(0xffffffffu, dst_l, min(i + 1, n - 1));
        const float w0 = __shfl_sync(0xffffffffu, val_l, i), w1 = __shfl_sync(0xffffffffu, val_l, min(i + 1, n - 1));
        float t0[NC], t1[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int o = lane + 32 * c;
          t0[c] = o < out ? __ldg(gact + (size_t)d0 * out + o) : 0.f;
          t1[c] = o < out ? __ldg(gact + (size_t)d1 * out + o) : 0.f;
        }
        float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;   // a[edge][basis]
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          a00 = fmaf(v0[c], t0[c], a00); a01 = fmaf(v1[c], t0[c], a01);
          a10 = fmaf(v0[c], t1[c], a10); a11 = fmaf(v1[c], t1[c], a11);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          a00 += __shfl_xor_sync(0xffffffffu, a00, off); a01 += __shfl_xor_sync(0xffffffffu, a01, off);
          a10 += __shfl_xor_sync(0xffffffffu, a10, off); a11 += __shfl_xor_sync(0xffffffffu, a11, off);
        }
        if (lane == 0) {
          float *c0 = cbuf + (size_t)(e0 + i) * B + b0;
          c0[0] = w0 * a00;
          if (two) c0[1] = w0 * a01;
          if (i + 1 < n) {
            c0[B] = w1 * a10;
            if (two) c0[B + 1] = w1 * a11;
          }
        }
      }
    }
  }
}

// ---- identity term, B > 0, basis gradient: g_weight_I[b, j, o] = sum_{e: src=j} comp[r_e,b] * t_e[o] ----
// A CTA walks tiles of TJ sources (TJ*out <= 256).  Per tile: (1) every thread stages t_e = val_e*gact[dst_e,:]
// of one edge into shared memory (all gathers independent -> deep memory-level parallelism), (2) one thread per
// output column x = (j - j0)*out + o accumulates ALL bases in registers over the source's edges out of shared
// memory, (3) stores are coalesced across the tile (consecutive x for a fixed basis).
// Sources with more than `thresh` edges are left to k_ident_bwd_w_long.
// edges staged per round: as many as fit ~48 KB, at most 512
static int pick_ec(int out) {
  int ec = (48 * 1024) / (4 * (out + 1));
  ec = ec > 512 ? 512 : ec;
  return ec < 32 ? 32 : (ec / 32) * 32;
}
template <int BT>
__global__ void __launch_bounds__(kThreads)
k_ident_bwd_w(const float *__restrict__ comp, const int32_t *__restrict__ colptr, const int32_t *__restrict__ e2_dst,
              const int32_t *__restrict__ e2_rel, const float *__restrict__ e2_val, const float *__restrict__ gact,
              float *__restrict__ gW, int NS, int R, int B, int out, int TJ, int thresh, int EC) {
  extern __shared__ __align__(16) float smem[];
  float *comp_s = smem;                 // [R][BT], zero padded
  float *Ts = smem + (size_t)R * BT;    // [EC][out]
  int *Rs = reinterpret_cast<int *>(Ts + (size_t)EC * out);  // [EC]
  const int tid = threadIdx.x;
  for (int x = tid; x < R * BT; x += kThreads) {
    int r = x / BT, b = x - r * BT;
    comp_s[x] = b < B ? __ldg(comp + (size_t)r * B + b) : 0.f;
  }
  for (int j0 = blockIdx.x * TJ; j0 < NS; j0 += gridDim.x * TJ) {
    const int tjw = min(TJ, NS - j0);
    const int t_lo = colptr[j0], t_hi = colptr[j0 + tjw];
    const bool mine = tid < tjw * out;
    const int jl = mine ? tid / out : 0, o = tid - jl * out;
    int s_lo = 0, s_hi = 0;
    if (mine) {
      s_lo = colptr[j0 + jl];
      s_hi = colptr[j0 + jl + 1];
      if (thresh > 0 && s_hi - s_lo > thresh) s_hi = s_lo;  // hub: other kernel
    }
    float2 acc[BT / 2];
#pragma unroll
    for (int q = 0; q < BT / 2; ++q) acc[q] = make_float2(0.f, 0.f);
    for (int c_lo = t_lo; c_lo < t_hi; c_lo += EC) {
      const int c_hi = min(t_hi, c_lo + EC);
      __syncthreads();
      for (int el = tid; el < c_hi - c_lo; el += kThreads) {
        const int e = c_lo + el;
        const float v = e2_val[e];
        const float *gp = gact + (size_t)e2_dst[e] * out;
        Rs[el] = e2_rel[e];
        for (int q = 0; q < out; ++q) Ts[el * out + q] = v * gp[q];
      }
      __syncthreads();
      const int lo = max(s_lo, c_lo), hi = min(s_hi, c_hi);
      for (int e = lo; e < hi; ++e) {
        const float t = Ts[(e - c_lo) * out + o];
        const float4 *c4 = reinterpret_cast<const float4 *>(comp_s + (size_t)Rs[e - c_lo] * BT);
#pragma unroll
        for (int q = 0; q < BT / 4; ++q) {
          float4 c = c4[q];
          fma2(acc[2 * q], t, make_float2(c.x, c.y));
          fma2(acc[2 * q + 1], t, make_float2(c.z, c.w));
        }
      }
    }
    if (mine && !(thresh > 0 && colptr[j0 + jl + 1] - colptr[j0 + jl] > thresh)) {
#pragma unroll
      for (int q = 0; q < BT / 2; ++q) {
        if (2 * q < B) gW[((size_t)(2 * q) * NS + j0) * out + tid] = acc[q].x;
        if (2 * q + 1 < B) gW[((size_t)(2 * q + 1) * NS + j0) * out + tid] = acc[q].y;
      }
    }
  }
}

// generic fallback (B > 64 or out > 256): one thread per output column, BC bases at a time, no staging
constexpr int BC = 8;
__global__ void __launch_bounds__(kThreads)
k_ident_bwd_w_generic(const float *__restrict__ comp, const int32_t *__restrict__ colptr, const int32_t *__restrict__ e2_dst,
                      const int32_t *__restrict__ e2_rel, const float *__restrict__ e2_val, const float *__restrict__ gact,
                      float *__restrict__ gW, int64_t NS, int B, int out, int thresh) {
  const int64_t x = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (x >= NS * out) return;
  const int64_t j = x / out;
  const int o = (int)(x - j * out);
  const int e_lo = colptr[j], e_hi = colptr[j + 1];
  if (thresh > 0 && e_hi - e_lo > thresh) return;
  for (int b0 = 0; b0 < B; b0 += BC) {
    float acc[BC];
#pragma unroll
    for (int q = 0; q < BC; ++q) acc[q] = 0.f;
    for (int e = e_lo; e < e_hi; ++e) {
      const float t = e2_val[e] * gact[(size_t)e2_dst[e] * out + o];
      const float *cr = comp + (size_t)e2_rel[e] * B + b0;
#pragma unroll
      for (int q = 0; q < BC; ++q)
        if (b0 + q < B) acc[q] = fmaf(__ldg(cr + q), t, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < BC; ++q)
      if (b0 + q < B) gW[((size_t)(b0 + q) * NS + j) * out + o] = acc[q];
  }
}

// hubs: one CTA (1024 threads) per SEGMENT of a long source (HubSegs).  The segment's edges are staged like above; comp
// is resident in shared memory; thread p owns the (basis, o) pairs p, p + 1024, ... and walks the staged edges in order.
// A single-segment hub is finished here; otherwise the partial goes to ws[seg] for k_ident_bwd_w_combine.
constexpr int kHubThreads = 1024;
__global__ void __launch_bounds__(kHubThreads)
k_ident_bwd_w_long(const float *__restrict__ comp, HubSegs h, const int32_t *__restrict__ colptr,
                   const int32_t *__restrict__ e2_dst, const int32_t *__restrict__ e2_rel,
                   const float *__restrict__ e2_val, const float *__restrict__ gact, float *__restrict__ gW, int64_t NS,
                   int R, int B, int out, int comp_smem, int EL) {
  extern __shared__ __align__(16) float smem[];
  float *Ts = smem;                                          // [EL][out]
  int *Rs = reinterpret_cast<int *>(Ts + (size_t)EL * out);  // [EL]  (pre-multiplied by B)
  float *comp_s = reinterpret_cast<float *>(Rs + EL);        // [R][B]
  const int sg = blockIdx.x;
  const int hub = h.seg_hub[sg];
  const int j = h.long_ids[hub];
  const int first = h.seg_first[hub], nseg = h.seg_first[hub + 1] - first;
  const int s_lo = min(colptr[j + 1], colptr[j] + (sg - first) * h.seg);
  const int s_hi = min(colptr[j + 1], s_lo + h.seg);
  const int tid = threadIdx.x;
  if (comp_smem)
    for (int x = tid; x < R * B; x += kHubThreads) comp_s[x] = __ldg(comp + x);
  const float *cbase = comp_smem ? comp_s : comp;
  const int npairs = B * out;
  constexpr int PP = 4;  // pairs per thread per pass
  for (int p0 = 0; p0 < npairs; p0 += PP * kHubThreads) {
    float acc[PP];
    int pb[PP], po[PP];
#pragma unroll
    for (int q = 0; q < PP; ++q) {
      const int p = p0 + q * kHubThreads + tid;
      acc[q] = 0.f;
      pb[q] = p < npairs ? p / out : -1;
      po[q] = p < npairs ? p - pb[q] * out : 0;
    }
    for (int c_lo = s_lo; c_lo < s_hi; c_lo += EL) {
      const int n = min(EL, s_hi - c_lo);
      __syncthreads();
      // all threads stage the chunk's t_e rows, coalesced along the row (a thread per edge walked a whole row alone: 200
      // dependent loads per thread on the link-prediction encoders)
      for (int x = tid; x < n * out; x += kHubThreads) {
        const int el = x / out, q = x - el * out;
        const int e = c_lo + el;
        Ts[x] = e2_val[e] * gact[(size_t)e2_dst[e] * out + q];
      }
      for (int el = tid; el < n; el += kHubThreads) Rs[el] = e2_rel[c_lo + el] * B;
      __syncthreads();
#pragma unroll
      for (int q = 0; q < PP; ++q) {
        if (pb[q] < 0) continue;
        float a0 = 0.f, a1 = 0.f;
        int el = 0;
        for (; el + 1 < n; el += 2) {
          a0 = fmaf(cbase[Rs[el] + pb[q]], Ts[el * out + po[q]], a0);
          a1 = fmaf(cbase[Rs[el + 1] + pb[q]], Ts[(el + 1) * out + po[q]], a1);
        }
        if (el < n) a0 = fmaf(cbase[Rs[el] + pb[q]], Ts[el * out + po[q]], a0);
        acc[q] += a0 + a1;
      }
    }
#pragma unroll
    for (int q = 0; q < PP; ++q) {
      if (pb[q] < 0) continue;
      if (nseg == 1) gW[((size_t)pb[q] * NS + j) * out + po[q]] = acc[q];
      else h.ws[(size_t)sg * npairs + (size_t)pb[q] * out + po[q]] = acc[q];
    }
  }
}
__global__ void k_ident_bwd_w_combine(HubSegs h, float *__restrict__ gW, int64_t NS, int B, int out) {
  const int hub = blockIdx.x;
  const int first = h.seg_first[hub], nseg = h.seg_first[hub + 1] - first;
  if (nseg <= 1) return;
  const int j = h.long_ids[hub];
  const int npairs = B * out;
  for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < nseg; ++s) acc += h.ws[(size_t)(first + s) * npairs + p];
    const int b = p / out, o = p - b * out;
    gW[((size_t)b * NS + j) * out + o] = acc;
  }
}

// ---- relation-chunk reductions -----------------------------------------------------------------------
// part[c, b] = sum over the E3 edges of chunk c of cbuf[e3_to_e2[e], b]   (rows of cbuf are B contiguous floats);
// e3_to_e2 == NULL: the rows are already in E3 order (k_ident_bwd_fused) and are streamed
__global__ void __launch_bounds__(kThreads)
k_comp_chunk_reduce(const float *__restrict__ cbuf, const int32_t *__restrict__ chunk_ptr,
                    const int32_t *__restrict__ e3_to_e2, float *__restrict__ part, int B) {
  __shared__ double red[kThreads];
  const int c = blockIdx.x;
  const int e_lo = chunk_ptr[c], e_hi = chunk_ptr[c + 1];
  const int bc = min(B, kThreads), nslots = kThreads / bc;
  const int slot = threadIdx.x / bc, bl = threadIdx.x - slot * bc;
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int b = b0 + bl;
    double acc = 0.0;
    if (slot < nslots && b < B)
      for (int e = e_lo + slot; e < e_hi; e += nslots) acc += cbuf[(size_t)(e3_to_e2 ? e3_to_e2[e] : e) * B + b];
    if (slot < nslots) red[slot * bc + bl] = acc;
    __syncthreads();
    for (int s = 1; s < nslots; s <<= 1) {
      if (slot < nslots && (slot % (2 * s)) == 0 && slot + s < nslots) red[slot * bc + bl] += red[(slot + s) * bc + bl];
      __syncthreads();
    }
    if (slot == 0 && b < B) part[(size_t)c * B + b] = (float)red[bl];
    __syncthreads();
  }
}

// The same sums when the scratch rows already lie in E3 order (k_ident_bwd_fused) and B % 4 == 0: one WARP per chunk streams
// the chunk's rows with 16-byte loads - lane = (row slot, column group of 4), a slot walks every RS-th row in order, the slots
// are then added in slot order.  A CTA per chunk (above) is launch- and latency-bound on the many short chunks.
__global__ void __launch_bounds__(256)
k_comp_chunk_reduce_rows(const float *__restrict__ cbuf, const int32_t *__restrict__ chunk_ptr, float *__restrict__ part,
                         int n_chunks, int B) {
  const int c = (int)(((size_t)blockIdx.x * 256 + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (c >= n_chunks) return;
  const int G4 = B >> 2, RS = 32 / G4;
  const int slot = lane / G4, q = lane - slot * G4;
  const int e_lo = chunk_ptr[c], e_hi = chunk_ptr[c + 1];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (slot < RS) {
    const float4 *base = reinterpret_cast<const float4 *>(cbuf) + q;
    int e = e_lo + slot;
    for (; e + 3 * RS < e_hi; e += 4 * RS) {
      const float4 x0 = __ldcs(base + (size_t)e * G4), x1 = __ldcs(base + (size_t)(e + RS) * G4);
      const float4 x2 = __ldcs(base + (size_t)(e + 2 * RS) * G4), x3 = __ldcs(base + (size_t)(e + 3 * RS) * G4);
      a0 += x0.x; a1 += x0.y; a2 += x0.z; a3 += x0.w;
      a0 += x1.x; a1 += x1.y; a2 += x1.z; a3 += x1.w;
      a0 += x2.x; a1 += x2.y; a2 += x2.z; a3 += x2.w;
      a0 += x3.x; a1 += x3.y; a2 += x3.z; a3 += x3.w;
    }
    for (; e < e_hi; e += RS) {
      const float4 x0 = __ldcs(base + (size_t)e * G4);
      a0 += x0.x; a1 += x0.y; a2 += x0.z; a3 += x0.w;
    }
  }
  for (int s = 1; s < RS; ++s) {
    const int srcl = s * G4 + q;   // < 32 for slot-0 lanes; other lanes read something valid and ignore it
    const double b0 = __shfl_sync(0xffffffffu, a0, srcl & 31), b1 = __shfl_sync(0xffffffffu, a1, srcl & 31);
    const double b2 = __shfl_sync(0xffffffffu, a2, srcl & 31), b3 = __shfl_sync(0xffffffffu, a3, srcl & 31);
    if (slot == 0) { a0 += b0; a1 += b1; a2 += b2; a3 += b3; }
  }
  if (slot == 0) reinterpret_cast<float4 *>(part + (size_t)c * B)[q] = make_float4((float)a0, (float)a1, (float)a2, (float)a3);
}

// part[c, k, o] = sum over the E3 edges of chunk c (one relation) of X[j_e, k] * t_e[o]
// thread = one k (row of g_W), OC outputs in registers; t_e staged per batch of EB edges in shared memory.
constexpr int EB = 64;
template <int OC>
__global__ void k_feat_bwd_w(const float *__restrict__ X, const float *__restrict__ gact,
                             const int32_t *__restrict__ chunk_ptr, const int32_t *__restrict__ e3_src,
                             const int32_t *__restrict__ e3_dst, const float *__restrict__ e3_val,
                             float *__restrict__ part, int in, int out, int ldx) {
  __shared__ __align__(16) float Ts[EB * OC];
  __shared__ int Js[EB];
  const int c = blockIdx.x;
  const int k = blockIdx.y * blockDim.x + threadIdx.x;
  const int e_lo = chunk_ptr[c], e_hi = chunk_ptr[c + 1];
  const bool kin = k < in;
  for (int c0 = 0; c0 < out; c0 += OC) {
    float2 acc[OC / 2];
#pragma unroll
    for (int o = 0; o < OC / 2; ++o) acc[o] = make_float2(0.f, 0.f);
    for (int eb = e_lo; eb < e_hi; eb += EB) {
      const int nb = min(EB, e_hi - eb);
      __syncthreads();
      for (int x = threadIdx.x; x < EB * OC; x += blockDim.x) {
        int el = x / OC, o = x - el * OC;
        float t = 0.f;
        if (el < nb && c0 + o < out) t = e3_val[eb + el] * gact[(size_t)e3_dst[eb + el] * out + c0 + o];
        Ts[x] = t;
      }
      for (int x = threadIdx.x; x < EB; x += blockDim.x) Js[x] = x < nb ? e3_src[eb + x] : 0;
      __syncthreads();
      if (kin) {
#pragma unroll 8
        for (int el = 0; el < EB; ++el) {  // padded entries have t = 0 and read row Js = 0 (valid memory)
          const float x = X[(size_t)Js[el] * ldx + k];
          const float4 *t4 = reinterpret_cast<const float4 *>(Ts + el * OC);
#pragma unroll
          for (int q = 0; q < OC / 4; ++q) {
            float4 t = t4[q];
            fma2(acc[2 * q], x, make_float2(t.x, t.y));
            fma2(acc[2 * q + 1], x, make_float2(t.z, t.w));
          }
        }
      }
    }
    if (kin) {
      float *pp = part + ((size_t)c * in + k) * out + c0;
#pragma unroll
      for (int o = 0; o < OC / 2; ++o) {
        if (c0 + 2 * o < out) pp[2 * o] = acc[o].x;
        if (c0 + 2 * o + 1 < out) pp[2 * o + 1] = acc[o].y;
      }
    }
  }
}

// Row-per-warp variant of the feature weight gradient (in <= 160, out <= 16): lane l owns rows k = c*32 + l of g_W for
// every 32-wide chunk c (NK*OC accumulators in registers), reads each edge's feature row with NK coalesced 128-byte
// loads and t_e from a per-warp shared-memory batch (broadcast).  4 warps split the chunk's edge batches; their partial
// sums are added in warp order.  ~3x fewer shared-pipe wavefronts than k_feat_bwd_w (ncu r01_feat: that one is bound by
// the L1/shared data pipe).
template <int NK, int OC>
__global__ void __launch_bounds__(128, 6)
k_feat_bwd_w_rw(const float *__restrict__ X, const float *__restrict__ gact, const int32_t *__restrict__ chunk_ptr,
                const int32_t *__restrict__ e3_src, const int32_t *__restrict__ e3_dst,
                const float *__restrict__ e3_val, float *__restrict__ part, int in, int out, int ldx) {
  constexpr int P = OC / 2, NW = 4;
  __shared__ __align__(16) float Ts[NW][32][OC];
  __shared__ int Js[NW][32];
  __shared__ float red[NW - 1][NK * 32 * OC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x;
  const int e_lo = chunk_ptr[c], e_hi = chunk_ptr[c + 1];
  float2 acc[NK][P];
#pragma unroll
  for (int q = 0; q < NK; ++q)
#pragma unroll
    for (int p = 0; p < P; ++p) acc[q][p] = make_float2(0.f, 0.f);
  for (int eb = e_lo + warp * 32; eb < e_hi; eb += NW * 32) {
    const int nb = min(32, e_hi - eb);
    __syncwarp();
    {
      const int e = eb + lane;
      const bool live = lane < nb;
      const float v = live ? e3_val[e] : 0.f;
      const float *gp = gact + (size_t)(live ? e3_dst[e] : 0) * out;
      Js[warp][lane] = live ? e3_src[e] : 0;
#pragma unroll
      for (int o = 0; o < OC; ++o) Ts[warp][lane][o] = (live && o < out) ? v * gp[o] : 0.f;
    }
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {  // padded entries have t = 0 and read row 0 (valid memory)
      const int j = Js[warp][i];
      float x[NK];
#pragma unroll
      for (int q = 0; q < NK; ++q) {
        const int k = q * 32 + lane;
        x[q] = (k < in) ? __ldg(X + (size_t)j * ldx + k) : 0.f;
      }
      float2 t[P];
      if constexpr (OC % 4 == 0) {
#pragma unroll
        for (int p = 0; p < P; p += 2) {
          const float4 t4 = *reinterpret_cast<const float4 *>(&Ts[warp][i][2 * p]);
          t[p] = make_float2(t4.x, t4.y);
          t[p + 1] = make_float2(t4.z, t4.w);
        }
      } else {
#pragma unroll
        for (int p = 0; p < P; ++p) t[p] = *reinterpret_cast<const float2 *>(&Ts[warp][i][2 * p]);
      }
#pragma unroll
      for (int q = 0; q < NK; ++q)
#pragma unroll
        for (int p = 0; p < P; ++p) fma2(acc[q][p], x[q], t[p]);
    }
  }
  // partial sums of warps 1..3 -> shared memory, warp 0 adds them in warp order and stores
  if (warp > 0) {
#pragma unroll
    for (int q = 0; q < NK; ++q)
#pragma unroll
      for (int p = 0; p < P; ++p) {
        red[warp - 1][(q * 32 + lane) * OC + 2 * p] = acc[q][p].x;
        red[warp - 1][(q * 32 + lane) * OC + 2 * p + 1] = acc[q][p].y;
      }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      const int k = q * 32 + lane;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        float a = acc[q][p].x, b = acc[q][p].y;
#pragma unroll
        for (int w2 = 0; w2 < NW - 1; ++w2) {
          a += red[w2][(q * 32 + lane) * OC + 2 * p];
          b += red[w2][(q * 32 + lane) * OC + 2 * p + 1];
        }
        if (k < in) {
          float *pp = part + ((size_t)c * in + k) * out;
          if (2 * p < out) pp[2 * p] = a;
          if (2 * p + 1 < out) pp[2 * p + 1] = b;
        }
      }
    }
  }
}

// Tiny layers (in * ceil(out/4) <= 32: AM hidden layer 10 -> 11, AIFB 16 -> 4): lane = (k, group of 4 outputs), so that
// (nearly) all lanes multiply - in k_feat_bwd_w_rw only the first `in` lanes of a warp own rows of g_W (10 of 32 on AM) and
// every one of them reads the whole t_e row.  Same edge order per accumulator and same warp order: identical results.
template <int OC>
__global__ void __launch_bounds__(128)
k_feat_bwd_w_small(const float *__restrict__ X, const float *__restrict__ gact, const int32_t *__restrict__ chunk_ptr,
                   const int32_t *__restrict__ e3_src, const int32_t *__restrict__ e3_dst, const float *__restrict__ e3_val,
                   float *__restrict__ part, int in, int out, int ldx) {
  constexpr int NW = 4, NOC = OC / 4;
  __shared__ __align__(16) float Ts[NW][32][OC];
  __shared__ int Js[NW][32];
  __shared__ __align__(16) float red[NW - 1][32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = lane / NOC, oc = lane - k * NOC;
  const bool act = k < in;
  const int c = blockIdx.x;
  const int e_lo = chunk_ptr[c], e_hi = chunk_ptr[c + 1];
  float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
  for (int eb = e_lo + warp * 32; eb < e_hi; eb += NW * 32) {
    const int nb = min(32, e_hi - eb);
    __syncwarp();
    {
      const int e = eb + lane;
      const bool live = lane < nb;
      const float v = live ? e3_val[e] : 0.f;
      const float *gp = gact + (size_t)(live ? e3_dst[e] : 0) * out;
      Js[warp][lane] = live ? e3_src[e] : 0;
#pragma unroll
      for (int o = 0; o < OC; ++o) Ts[warp][lane][o] = (live && o < out) ? v * gp[o] : 0.f;
    }
    __syncwarp();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {  // padded entries have t = 0 and read row 0 (valid memory)
      const float x = act ? __ldg(X + (size_t)Js[warp][i] * ldx + k) : 0.f;
      const float4 t4 = *reinterpret_cast<const float4 *>(&Ts[warp][i][4 * oc]);
      fma2(a0, x, make_float2(t4.x, t4.y));
      fma2(a1, x, make_float2(t4.z, t4.w));
    }
  }
  if (warp > 0) *reinterpret_cast<float4 *>(red[warp - 1][lane]) = make_float4(a0.x, a0.y, a1.x, a1.y);
  __syncthreads();
  if (warp == 0 && act) {
    float r[4] = {a0.x, a0.y, a1.x, a1.y};
#pragma unroll
    for (int w2 = 0; w2 < NW - 1; ++w2) {
      const float4 q = *reinterpret_cast<const float4 *>(red[w2][lane]);
      r[0] += q.x; r[1] += q.y; r[2] += q.z; r[3] += q.w;
    }
    float *pp = part + ((size_t)c * in + k) * out + 4 * oc;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (4 * oc + u < out) pp[u] = r[u];
  }
}

// ---- basis gradients of the feature weights (graph.py:83-85 backwards).  Tiny. ---------------------
__global__ void k_basis_mix_bwd_v(const float *__restrict__ comp, const float *__restrict__ gW, float *__restrict__ gV,
                                  int R, int B, int IO) {
  const int b = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= IO) return;
  float acc = 0.f;
  for (int r = 0; r < R; ++r) acc = fmaf(__ldg(comp + (size_t)r * B + b), gW[(size_t)r * IO + x], acc);
  gV[(size_t)b * IO + x] = acc;
}
__global__ void __launch_bounds__(kThreads)
k_basis_mix_bwd_c(const float *__restrict__ V, const float *__restrict__ gW, float *__restrict__ gcomp, int B, int IO) {
  __shared__ float red[kThreads];
  const int r = blockIdx.x, b = blockIdx.y;
  float acc = 0.f;
  for (int x = threadIdx.x; x < IO; x += kThreads) acc = fmaf(V[(size_t)b * IO + x], gW[(size_t)r * IO + x], acc);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) gcomp[(size_t)r * B + b] = red[0];
}

// ---- input gradient --------------------------------------------------------------------------------
// g_X[j,:] = sum_{e: src=j} val_e * gact[dst_e,:] . W[r_e]^T  is the forward pass on the transposed graph: the
// relation-major message kernel gathers gact rows by e3_dst and multiplies with W^T, the segmented sum runs over
// sources (E2) through e2_to_e3.  Wt[r][o][k] = W[r][k][o]:
__global__ void k_transpose_w(const float *__restrict__ W, float *__restrict__ Wt, int in, int out) {
  const int r = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= in * out) return;
  const int k = x / out, o = x - k * out;
  Wt[((size_t)r * out + o) * in + k] = W[(size_t)r * in * out + x];
}

}  // namespace
}  // namespace mrgcn

using namespace mrgcn;

extern "C" int mrgcn_rgcn_layer_bwd(const mrgcn_layer_bwd_args *a, mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(a && a->gout && a->gact, MRGCN_E_BADARG, "layer_bwd: gout/gact missing");
  const mrgcn_layer_args &f = a->f;
  const bool hasI = f.weight_I != nullptr, hasF = f.X != nullptr;
  const int B = f.B > 0 ? f.B : 0, out = f.out_dim, in = f.in_dim;
  const int ldx = f.x_stride > 0 ? f.x_stride : in;
  const mrgcn_graph *gI = f.gI, *gF = f.gF;
  const int ND = hasI ? gI->ND : gF->ND;
  MRGCN_REQUIRE(!f.relu || f.out, MRGCN_E_BADARG, "layer_bwd: relu needs the forward output");
  const int ph = a->phases ? a->phases : (MRGCN_BWD_ACT | MRGCN_BWD_IDENT | MRGCN_BWD_FEATW | MRGCN_BWD_GX);

  // 1. gact and bias gradient
  const int nblk = (int)cdiv(ND > 0 ? ND : 1, kColRows);
  MRGCN_REQUIRE(!a->g_bias || a->colsum_ws, MRGCN_E_BADARG, "layer_bwd: colsum_ws missing");
  if (ND > 0 && (ph & MRGCN_BWD_ACT)) {
    MRGCN_PROF("act_bwd");
  k_act_bwd<<<dim3((unsigned)nblk, (unsigned)cdiv(out, out < 64 ? out : 64)), kThreads, kThreads * sizeof(double), st>>>(a->gout, f.out, f.row_mask, a->gact,
                                                                 a->g_bias ? a->colsum_ws : nullptr, ND, out, f.relu);
    MRGCN_LAUNCH_CHECK();
  }
  if (a->g_bias && (ph & MRGCN_BWD_ACT)) {
    if (ND > 0) {
      MRGCN_PROF("bias_reduce");
  k_seq_reduce<<<dim3((unsigned)cdiv(out, 32), 1), 256, 0, st>>>(a->colsum_ws, nullptr, nullptr, nblk, out, a->g_bias);
      MRGCN_LAUNCH_CHECK();
    } else {
      MRGCN_CUDA(cudaMemsetAsync(a->g_bias, 0, sizeof(float) * out, st));
    }
  }

  // 2. identity term
  if (hasI && a->g_weight_I && (ph & MRGCN_BWD_IDENT)) {
    const int64_t NS = gI->NS;
    if (B == 0) {
      MRGCN_CUDA(cudaMemsetAsync(a->g_weight_I, 0, sizeof(float) * (size_t)gI->R * NS * out, st));
      if (gI->E > 0) {
        const int epw = out >= 32 ? 1 : 32 / out;
        MRGCN_PROF("ident_bwd_direct");
  k_ident_bwd_direct<<<(unsigned)cdiv(cdiv(gI->E, epw) * 32, kThreads), kThreads, 0, st>>>(
            gI->e2_src, gI->e2_rel, gI->e2_dst, gI->e2_val, a->gact, a->g_weight_I, gI->E, NS, out);
        MRGCN_LAUNCH_CHECK();
      }
    } else {
      const int OC = pick_oc(out);
      const int thresh = gI->n_long_cols > 0 ? gI->long_col_thresh : 0;
      TabGeom tg;
      const bool tab_w = f.plan && (mrgcn_tab_mode(B, 0, out) & 2) && tab_geometry(B, out, tg, kBwdWBpt);
      int tBC0 = 0, tOP0 = 0;
      const bool tab_c0 = f.plan && (mrgcn_tab_mode(B, 0, out) & 4) && tab_c_geometry(B, out, tBC0, tOP0);
      // one pass for both gradients (ident_bwd.cu) when neither table kernel is asked for and the shape fits
      bool fused = false;
      if (!tab_w && !tab_c0 && a->g_comp_I && a->cbuf && a->part && gI->E > 0) {
        const int rc = launch_ident_bwd_fused(gI, f.weight_I, f.comp_I, B, out, a->gact, a->g_weight_I, a->cbuf, st);
        if (rc < 0 || rc > 1) return rc;
        fused = rc == 0;
      }
      {  // basis gradient
        if (fused) {
        } else if (tab_w) {
          if (int rc = launch_tab_bwd_w(gI, f.plan, f.comp_I, B, out, a->gact, a->g_weight_I, st)) return rc;
        } else if (B <= 64 && out <= 256) {
          const int BT = (int)cdiv(B, 8) * 8;
          const int TJ = 256 / out;
          const int EC = pick_ec(out);
          size_t smem = ((size_t)gI->R * BT + (size_t)EC * out + EC) * 4;
          MRGCN_REQUIRE(smem <= 200 * 1024, MRGCN_E_NOTSUP, "ident_bwd_w: R*B too large for shared memory");
          unsigned grid = 0;
          MRGCN_PROF("ident_bwd_w");
#define LAUNCH(BTV)                                                                                              \
  do {                                                                                                           \
    if (int rc = set_smem(k_ident_bwd_w<BTV>, smem)) return rc;                                                  \
    grid = persistent_grid(k_ident_bwd_w<BTV>, kThreads, smem, cdiv(NS, TJ));                                    \
    k_ident_bwd_w<BTV><<<grid, kThreads, smem, st>>>(f.comp_I, gI->colptr, gI->e2_dst, gI->e2_rel, gI->e2_val,   \
                                                     a->gact, a->g_weight_I, (int)NS, gI->R, B, out, TJ, thresh, EC); \
  } while (0)
          switch (BT) {
            case 8: LAUNCH(8); break;
            case 16: LAUNCH(16); break;
            case 24: LAUNCH(24); break;
            case 32: LAUNCH(32); break;
            case 40: LAUNCH(40); break;
            case 48: LAUNCH(48); break;
            case 56: LAUNCH(56); break;
            default: LAUNCH(64); break;
          }
#undef LAUNCH
          MRGCN_LAUNCH_CHECK();
        } else {
          MRGCN_PROF("ident_bwd_w");
          k_ident_bwd_w_generic<<<(unsigned)cdiv(NS * out, kThreads), kThreads, 0, st>>>(
              f.comp_I, gI->colptr, gI->e2_dst, gI->e2_rel, gI->e2_val, a->gact, a->g_weight_I, NS, B, out, thresh);
          MRGCN_LAUNCH_CHECK();
        }
        if (gI->n_long_cols > 0) {
          const int comp_smem = (size_t)gI->R * B * 4 <= 64 * 1024 ? 1 : 0;
          HubSegs hs{gI->long_cols, gI->col_seg_hub, gI->col_seg_first, gI->n_long_cols, gI->n_col_segs, gI->long_seg, f.hub_ws};
          MRGCN_REQUIRE(hs.ws || hs.n_segs == hs.n_long, MRGCN_E_BADARG, "ident_bwd_w: hub_ws missing");
          int EL = hs.seg;
          while (EL > 32 && ((size_t)EL * out + EL) * 4 > 64 * 1024) EL >>= 1;
          size_t smem = ((size_t)EL * out + EL + (comp_smem ? (size_t)gI->R * B : 0)) * 4;
          MRGCN_REQUIRE(smem <= 200 * 1024, MRGCN_E_NOTSUP, "ident_bwd_w_long: out too large for shared memory");
          if (int rc = set_smem(k_ident_bwd_w_long, smem)) return rc;
          MRGCN_PROF("ident_bwd_w_long");
          k_ident_bwd_w_long<<<(unsigned)hs.n_segs, kHubThreads, smem, st>>>(f.comp_I, hs, gI->colptr, gI->e2_dst, gI->e2_rel,
                                                                            gI->e2_val, a->gact, a->g_weight_I, NS, gI->R, B,
                                                                            out, comp_smem, EL);
          MRGCN_LAUNCH_CHECK();
          if (hs.n_segs > hs.n_long) {
            MRGCN_PROF("ident_bwd_w_combine");
            k_ident_bwd_w_combine<<<(unsigned)hs.n_long, 256, 0, st>>>(hs, a->g_weight_I, NS, B, out);
            MRGCN_LAUNCH_CHECK();
          }
        }
      }
      if (a->g_comp_I) {
        MRGCN_REQUIRE(a->cbuf && a->part, MRGCN_E_BADARG, "layer_bwd: cbuf/part missing");
        int tBC = 0, tOP = 0;
        const bool tab_c = f.plan && (mrgcn_tab_mode(B, 0, out) & 4) && tab_c_geometry(B, out, tBC, tOP);
        if (tab_c) {
          // records of the (tile, relation) pieces, then the fixed-order sum per relation
          if (gI->E > 0)
            if (int rc = launch_tab_bwd_c(gI, f.plan, f.weight_I, B, out, a->gact, a->cbuf, st)) return rc;
          // stage 1: blocks of <= 128 records of one relation; stage 2: the blocks of every relation
          float *blk = a->cbuf + (size_t)f.plan->n_pieces * B;
          if (f.plan->n_blks > 0) {
            MRGCN_PROF("comp_block_reduce");
            k_seq_reduce<<<dim3((unsigned)cdiv(B, 32), (unsigned)f.plan->n_blks), 256, 0, st>>>(a->cbuf, f.plan->blk_ptr,
                                                                                                 f.plan->rel_piece_idx,
                                                                                                 f.plan->n_pieces, B, blk);
            MRGCN_LAUNCH_CHECK();
          }
          MRGCN_PROF("comp_reduce");
          k_seq_reduce<<<dim3((unsigned)cdiv(B, 32), (unsigned)gI->R), 256, 0, st>>>(blk, f.plan->rel_blk_ptr, nullptr,
                                                                                      f.plan->n_blks, B, a->g_comp_I);
          MRGCN_LAUNCH_CHECK();
        } else {
        if (gI->E > 0) {
          unsigned grid = 0;
          size_t smem = 0;
          IdentPipe p;
          if (fused) {
            // scratch rows already written, in E3 order
          } else if (f.plan && f.plan->n_tasks > 0 && f.plan->lt <= 32 && out >= 64 && out <= 256 && B <= 8) {
            const int NC = (int)cdiv(out, 32);
            const int64_t blocks = cdiv(f.plan->n_tasks, 8);
            const unsigned gridt = (unsigned)(blocks < 8 * kNumSMs ? blocks : 8 * kNumSMs);
            const int4 *tk = reinterpret_cast<const int4 *>(f.plan->tasks4);
            MRGCN_PROF("ident_bwd_c");
#define LAUNCH_T(NCV) \
  k_ident_bwd_c_task<NCV><<<gridt, 256, 0, st>>>(f.weight_I, tk, f.plan->n_tasks, gI->e2_dst, gI->e2_val, a->gact, a->cbuf, NS, B, out)
            switch (NC) {
              case 2: LAUNCH_T(2); break;
              case 3: LAUNCH_T(3); break;
              case 4: LAUNCH_T(4); break;
              case 5: LAUNCH_T(5); break;
              case 6: LAUNCH_T(6); break;
              case 7: LAUNCH_T(7); break;
              default: LAUNCH_T(8); break;
            }
#undef LAUNCH_T
            MRGCN_LAUNCH_CHECK();
          } else if (ident_pipe_config(p, NS, B, out, 0)) {
            const int VW = (out % 4 == 0) ? 4 : (out % 2 == 0) ? 2 : 1;
            smem = 16 * ((2 * p.S * 8 + 15) / 16) + (size_t)p.S * p.stage_bytes;
            MRGCN_PROF("ident_bwd_c");
#define LAUNCH(OCV, VWV)                                                                                          \
  do {                                                                                                            \
    if (int rc = set_smem(k_ident_bwd_c_bulk<OCV, VWV>, smem)) return rc;                                         \
    grid = persistent_grid(k_ident_bwd_c_bulk<OCV, VWV>, kPipeThreads, smem, p.ntiles);                           \
    k_ident_bwd_c_bulk<OCV, VWV><<<grid, kPipeThreads, smem, st>>>(f.weight_I, gI->colptr, gI->e2_src, gI->e2_dst, \
                                                                   gI->e2_val, a->gact, a->cbuf, p);               \
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
          } else {
            const int OP = (int)cdiv(out, OC) * OC;
            const int TJ = ident_tile(B, out, OP);
            smem = (size_t)B * TJ * OP * 4;
            MRGCN_REQUIRE(smem <= 220 * 1024, MRGCN_E_NOTSUP, "ident_bwd_c: B*out too large for shared memory");
            MRGCN_PROF("ident_bwd_c");
#define LAUNCH(OCV)                                                                                              \
  do {                                                                                                           \
    if (int rc = set_smem(k_ident_bwd_c<OCV>, smem)) return rc;                                                  \
    grid = persistent_grid(k_ident_bwd_c<OCV>, kThreads, smem, cdiv(NS, TJ));                                    \
    k_ident_bwd_c<OCV><<<grid, kThreads, smem, st>>>(f.weight_I, gI->colptr, gI->e2_src, gI->e2_dst, gI->e2_val, \
                                                     a->gact, a->cbuf, (int)NS, B, out, OP, TJ);                 \
  } while (0)
            switch (OC) {
              case 4: LAUNCH(4); break;
              case 8: LAUNCH(8); break;
              case 12: LAUNCH(12); break;
              default: LAUNCH(16); break;
            }
#undef LAUNCH
            MRGCN_LAUNCH_CHECK();
          }
          MRGCN_PROF("comp_chunk_reduce");
          if (fused && (B & 3) == 0 && B <= 128 && (((uintptr_t)a->cbuf | (uintptr_t)a->part) & 15) == 0)
            k_comp_chunk_reduce_rows<<<(unsigned)cdiv(gI->n_chunks, 8), 256, 0, st>>>(a->cbuf, gI->chunk_ptr, a->part,
                                                                                      gI->n_chunks, B);
          else
            k_comp_chunk_reduce<<<(unsigned)gI->n_chunks, kThreads, 0, st>>>(a->cbuf, gI->chunk_ptr,
                                                                             fused ? nullptr : gI->e3_to_e2, a->part, B);
          MRGCN_LAUNCH_CHECK();
        }
        MRGCN_PROF("comp_reduce");
        k_seq_reduce<<<dim3((unsigned)cdiv(B, 32), (unsigned)gI->R), 256, 0, st>>>(a->part, gI->rel_chunk_ptr,
                                                                                    gI->rel_chunk_idx, gI->n_chunks, B,
                                                                                    a->g_comp_I);
        MRGCN_LAUNCH_CHECK();
        }
      }
    }
  }

  // 3. feature term
  if (hasF) {
    const int IO = in * out;
    const float *W = B > 0 ? f.wmix : f.weight_F;
    if ((a->g_weight_F || a->g_comp_F) && (ph & MRGCN_BWD_FEATW)) {
      float *gW = B > 0 ? a->g_wmix : a->g_weight_F;
      MRGCN_REQUIRE(gW && a->part, MRGCN_E_BADARG, "layer_bwd: g_wmix/part missing");
      static int rw_mode = -1;   // MRGCN_FEAT_RW=0 selects the thread-per-row kernel everywhere
      if (rw_mode < 0) { const char *e = getenv("MRGCN_FEAT_RW"); rw_mode = (e && e[0] == '0') ? 0 : 1; }
      int tile_rc = 1;
      if (gF->E > 0 && out <= 12 && in * (((out + 3) / 4)) <= 32 && rw_mode == 1) {
        const int OCS = ((out + 3) / 4) * 4;
        MRGCN_PROF("feat_bwd_w");
#define LAUNCH_S(OCV) \
  k_feat_bwd_w_small<OCV><<<(unsigned)gF->n_chunks, 128, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out, ldx)
        switch (OCS) {
          case 4: LAUNCH_S(4); break;
          case 8: LAUNCH_S(8); break;
          default: LAUNCH_S(12); break;
        }
#undef LAUNCH_S
        MRGCN_LAUNCH_CHECK();
      } else if (gF->E > 0 && in <= 160 && out <= 16 && rw_mode == 1) {
        const int NK = (int)cdiv(in, 32);
        const int OCR = out <= 4 ? 4 : out <= 8 ? 8 : out <= 10 ? 10 : out <= 12 ? 12 : 16;
        MRGCN_PROF("feat_bwd_w");
#define LAUNCH_RW(NKV, OCV) \
  k_feat_bwd_w_rw<NKV, OCV><<<(unsigned)gF->n_chunks, 128, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out, ldx)
#define LAUNCH_NK(OCV)                  \
  switch (NK) {                         \
    case 1: LAUNCH_RW(1, OCV); break;   \
    case 2: LAUNCH_RW(2, OCV); break;   \
    case 3: LAUNCH_RW(3, OCV); break;   \
    case 4: LAUNCH_RW(4, OCV); break;   \
    default: LAUNCH_RW(5, OCV); break;  \
  }
        switch (OCR) {
          case 4: LAUNCH_NK(4); break;
          case 8: LAUNCH_NK(8); break;
          case 10: LAUNCH_NK(10); break;
          case 12: LAUNCH_NK(12); break;
          default: LAUNCH_NK(16); break;
        }
#undef LAUNCH_NK
#undef LAUNCH_RW
        MRGCN_LAUNCH_CHECK();
      } else if (gF->E > 0 && (tile_rc = launch_feat_bwd_w_tile(gF, f.X, ldx, a->gact, a->part, in, out, st)) != 1) {
        // wide outputs: register-tiled product per chunk (feat_bwd_w.cu)
        if (tile_rc != 0) return tile_rc;
      } else if (gF->E > 0) {
        const int OC = pick_oc(out);
        int bt = (int)cdiv(in < 256 ? in : 256, 32) * 32;
        dim3 grid((unsigned)gF->n_chunks, (unsigned)cdiv(in, bt));
        MRGCN_PROF("feat_bwd_w");
  switch (OC) {
          case 4: k_feat_bwd_w<4><<<grid, bt, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out, ldx); break;
          case 8: k_feat_bwd_w<8><<<grid, bt, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out, ldx); break;
          case 12: k_feat_bwd_w<12><<<grid, bt, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out, ldx); break;
          default: k_feat_bwd_w<16><<<grid, bt, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out, ldx); break;
        }
        MRGCN_LAUNCH_CHECK();
      }
      MRGCN_PROF("feat_w_reduce");
  k_seq_reduce<<<dim3((unsigned)cdiv(IO, 32), (unsigned)gF->R), 256, 0, st>>>(a->part, gF->rel_chunk_ptr,
                                                                                   gF->rel_chunk_idx, gF->n_chunks, IO, gW);
      MRGCN_LAUNCH_CHECK();
      if (B > 0) {
        if (a->g_weight_F) {
          MRGCN_PROF("basis_mix_bwd_v");
  k_basis_mix_bwd_v<<<dim3((unsigned)cdiv(IO, 128), (unsigned)B), 128, 0, st>>>(f.comp_F, gW, a->g_weight_F, gF->R, B, IO);
          MRGCN_LAUNCH_CHECK();
        }
        if (a->g_comp_F) {
          MRGCN_PROF("basis_mix_bwd_c");
  k_basis_mix_bwd_c<<<dim3((unsigned)gF->R, (unsigned)B), kThreads, 0, st>>>(f.weight_F, gW, a->g_comp_F, B, IO);
          MRGCN_LAUNCH_CHECK();
        }
      }
    }
    if (a->g_X && (ph & MRGCN_BWD_GX)) {
      MRGCN_REQUIRE(a->wt_ws && a->msgx_ws, MRGCN_E_BADARG, "layer_bwd: wt_ws/msgx_ws missing");
      const int64_t NS = gF->NS;
      if (NS > 0) {
        MRGCN_PROF("transpose_w");
        k_transpose_w<<<dim3((unsigned)cdiv(IO, 128), (unsigned)gF->R), 128, 0, st>>>(W, a->wt_ws, in, out);
        MRGCN_LAUNCH_CHECK();
        AggArgs g{};
        g.ND = (int)NS; g.odim = in; g.ms = msg_stride(in); g.out = a->g_X;
        g.thresh = gF->n_long_cols > 0 ? gF->long_col_thresh : 0;
        HubSegs hs{gF->long_cols, gF->col_seg_hub, gF->col_seg_first, gF->n_long_cols, gF->n_col_segs, gF->long_seg, f.hub_ws};
        if (narrow_supported(gF->R, out, in)) {
          // dX[j, :] = sum_{e: src = j} val_e * W[rel_e] . gact[dst_e, :] in one pass over the source-major order (narrow.cu)
          NarrowArgs n{gF->cols_by_deg, gF->colptr, gF->e2_dst, gF->e2_rel, gF->e2_val, a->gact, a->wt_ws, out, gF->R, out, in, g};
          if (int rc = launch_narrow(n, hs, st, "narrow_bwd_x")) return rc;
        } else {
          if (gF->E > 0)
            if (int rc = launch_feat_msg(gF, gF->e3_dst, a->gact, out, a->wt_ws, a->msgx_ws, out, in, st, "feat_bwd_x_msg")) return rc;
          g.msgF = a->msgx_ws; g.pF = gF->e2_to_e3; g.rowptrF = gF->colptr;
          if (int rc = launch_agg(g, hs, st, "feat_bwd_x_agg")) return rc;
        }
      }
    }
  }
  return 0;
}
