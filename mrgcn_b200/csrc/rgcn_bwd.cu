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
#include "common.cuh"

namespace mrgcn {
int launch_basis_mix_fwd(const float *comp, const float *V, float *W, int R, int B, int IO, cudaStream_t st);
int pick_oc(int out);
int ident_tile(int B, int out, int OP);

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
  extern __shared__ float red[];  // [kThreads]
  const int r0 = blockIdx.x * kColRows, r1 = min(ND, r0 + kColRows);
  const int oc = min(od, kThreads), nslots = kThreads / oc;
  const int slot = threadIdx.x / oc, ol = threadIdx.x - slot * oc;
  for (int o0 = 0; o0 < od; o0 += oc) {
    const int o = o0 + ol;
    float acc = 0.f;
    if (slot < nslots && o < od) {
      for (int i = r0 + slot; i < r1; i += nslots) {
        size_t x = (size_t)i * od + o;
        float g = gout[x];
        if (relu && !(outv[x] > 0.f)) g = 0.f;
        if (mask) g *= mask[i];
        gact[x] = g;
        acc += g;
      }
    }
    if (slot < nslots) red[slot * oc + ol] = acc;
    __syncthreads();
    for (int s = 1; s < nslots; s <<= 1) {
      if (slot < nslots && (slot % (2 * s)) == 0 && slot + s < nslots) red[slot * oc + ol] += red[(slot + s) * oc + ol];
      __syncthreads();
    }
    if (colsum && slot == 0 && o < od) colsum[(size_t)blockIdx.x * od + o] = red[ol];
    __syncthreads();
  }
}

// out[x] = sum_c part[c*stride + x] for c in [lo, hi) -- sequential, fixed order.
// seg_ptr == NULL: one segment [0, nseg_total).
__global__ void k_seq_reduce(const float *__restrict__ part, const int32_t *__restrict__ seg_ptr, int nall, int width,
                             float *__restrict__ outp) {
  const int s = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= width) return;
  const int lo = seg_ptr ? seg_ptr[s] : 0, hi = seg_ptr ? seg_ptr[s + 1] : nall;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int c = lo;
  for (; c + 4 <= hi; c += 4) {  // 4 interleaved chains: fixed association, more loads in flight
    a0 += part[(size_t)c * width + x];
    a1 += part[(size_t)(c + 1) * width + x];
    a2 += part[(size_t)(c + 2) * width + x];
    a3 += part[(size_t)(c + 3) * width + x];
  }
  for (; c < hi; ++c) a0 += part[(size_t)c * width + x];
  outp[(size_t)s * width + x] = (a0 + a1) + (a2 + a3);
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
// Same tiling as k_ident_msg_fwd (V tile staged once per TJ sources); rows of cbuf are staged per warp
// in shared memory so that the global stores are coalesced.
template <int OC>
__global__ void __launch_bounds__(kThreads)
k_ident_bwd_c(const float *__restrict__ V, const int32_t *__restrict__ colptr, const int32_t *__restrict__ e2_src,
              const int32_t *__restrict__ e2_dst, const float *__restrict__ e2_val, const float *__restrict__ gact,
              float *__restrict__ cbuf, int NS, int B, int out, int OP, int TJ, int BS) {
  extern __shared__ __align__(16) float smem[];
  float *Vs = smem;                               // [B][TJ][OP]
  float *Cs_all = smem + (size_t)B * TJ * OP;     // [nwarps][32][BS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float *Cs = Cs_all + (size_t)warp * 32 * BS;
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
#pragma unroll 8
      for (int b = 0; b < B; ++b) dst[(size_t)b * TJ * OP] = ldg_stream(src + (size_t)b * NS * out);
    }
    // zero the padding columns (they are multiplied with t = 0 below, but must not be NaN)
    if (OP > out)
      for (int x = tid; x < B * tjw * (OP - out); x += kThreads) {
        int row = x / (OP - out), o = out + x % (OP - out);
        int b = row / tjw, jl = row - b * tjw;
        Vs[((size_t)b * TJ + jl) * OP + o] = 0.f;
      }
    __syncthreads();
    for (int eb = e_lo + warp * 32; eb < e_hi; eb += (kThreads / 32) * 32) {
      const int e = eb + lane;
      const bool live = e < e_hi;
      const int jl = live ? e2_src[e] - j0 : 0;
      const float v = live ? e2_val[e] : 0.f;
      const float *gp = gact + (size_t)(live ? e2_dst[e] : 0) * out;
      for (int b = 0; b < B; ++b) Cs[lane * BS + b] = 0.f;
      for (int c0 = 0; c0 < OP; c0 += OC) {
        float t[OC];
#pragma unroll
        for (int o = 0; o < OC; ++o) t[o] = (live && c0 + o < out) ? v * gp[c0 + o] : 0.f;
        const float *vp = Vs + jl * OP + c0;
        for (int b = 0; b < B; ++b) {
          const float4 *v4 = reinterpret_cast<const float4 *>(vp + (size_t)b * TJ * OP);
          float acc = 0.f;
#pragma unroll
          for (int q = 0; q < OC / 4; ++q) {
            float4 w = v4[q];
            acc = fmaf(w.x, t[4 * q + 0], acc);
            acc = fmaf(w.y, t[4 * q + 1], acc);
            acc = fmaf(w.z, t[4 * q + 2], acc);
            acc = fmaf(w.w, t[4 * q + 3], acc);
          }
          Cs[lane * BS + b] += acc;
        }
      }
      __syncwarp();
      // coalesced write-out of the 32 x B block (rows are consecutive in cbuf)
      const int nlive = min(32, e_hi - eb);
      float *cp = cbuf + (size_t)eb * B;
      for (int x = lane; x < nlive * B; x += 32) {
        int row = x / B, b = x - row * B;
        cp[x] = Cs[row * BS + b];
      }
      __syncwarp();
    }
  }
}

// ---- identity term, B > 0, basis gradient: g_weight_I[b, j, o] = sum_{e: src=j} comp[r_e,b] * t_e[o] ----
// One thread per output column x = (j - j0)*out + o of a tile of TJ sources; BC bases at a time in registers.
// Stores are coalesced across the tile (consecutive x for a fixed basis).  Sources with degree > thresh are
// left to k_ident_bwd_w_long.
constexpr int BC = 8;
__global__ void __launch_bounds__(kThreads)
k_ident_bwd_w(const float *__restrict__ comp, const int32_t *__restrict__ colptr, const int32_t *__restrict__ e2_dst,
              const int32_t *__restrict__ e2_rel, const float *__restrict__ e2_val, const float *__restrict__ gact,
              float *__restrict__ gW, int NS, int R, int B, int out, int TJ, int CS, int thresh) {
  extern __shared__ __align__(16) float comp_s[];  // [R][CS], CS = B rounded up to BC, zero padded
  const int tid = threadIdx.x;
  for (int x = tid; x < R * CS; x += kThreads) {
    int r = x / CS, b = x - r * CS;
    comp_s[x] = b < B ? __ldg(comp + (size_t)r * B + b) : 0.f;
  }
  __syncthreads();
  for (int j0 = blockIdx.x * TJ; j0 < NS; j0 += gridDim.x * TJ) {
    const int tjw = min(TJ, NS - j0);
    for (int x = tid; x < tjw * out; x += kThreads) {
      const int jl = x / out, o = x - jl * out;
      const int e_lo = colptr[j0 + jl], e_hi = colptr[j0 + jl + 1];
      const bool skip = thresh > 0 && e_hi - e_lo > thresh;
      for (int b0 = 0; b0 < B; b0 += BC) {
        float acc[BC];
#pragma unroll
        for (int q = 0; q < BC; ++q) acc[q] = 0.f;
        if (!skip)
          for (int e = e_lo; e < e_hi; ++e) {
            const float t = e2_val[e] * gact[(size_t)e2_dst[e] * out + o];
            const float4 *c4 = reinterpret_cast<const float4 *>(comp_s + (size_t)e2_rel[e] * CS + b0);
#pragma unroll
            for (int q = 0; q < BC / 4; ++q) {
              float4 c = c4[q];
              acc[4 * q + 0] = fmaf(c.x, t, acc[4 * q + 0]);
              acc[4 * q + 1] = fmaf(c.y, t, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(c.z, t, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(c.w, t, acc[4 * q + 3]);
            }
          }
        if (!skip) {
#pragma unroll
          for (int q = 0; q < BC; ++q)
            if (b0 + q < B) gW[((size_t)(b0 + q) * NS + j0) * out + x] = acc[q];
        }
      }
    }
  }
}

// hubs: one CTA per long source; edge slots strided over the source's edges, fixed-order tree over slots
__global__ void __launch_bounds__(kThreads)
k_ident_bwd_w_long(const float *__restrict__ comp, const int32_t *__restrict__ long_cols,
                   const int32_t *__restrict__ colptr, const int32_t *__restrict__ e2_dst,
                   const int32_t *__restrict__ e2_rel, const float *__restrict__ e2_val,
                   const float *__restrict__ gact, float *__restrict__ gW, int NS, int B, int out) {
  __shared__ float red[kThreads];
  const int j = long_cols[blockIdx.x];
  const int e_lo = colptr[j], e_hi = colptr[j + 1];
  const int oc = min(out, kThreads), nslots = kThreads / oc;
  const int slot = threadIdx.x / oc, ol = threadIdx.x - slot * oc;
  for (int o0 = 0; o0 < out; o0 += oc) {
    const int o = o0 + ol;
    for (int b = 0; b < B; ++b) {
      float acc = 0.f;
      if (slot < nslots && o < out)
        for (int e = e_lo + slot; e < e_hi; e += nslots)
          acc = fmaf(__ldg(comp + (size_t)e2_rel[e] * B + b), e2_val[e] * gact[(size_t)e2_dst[e] * out + o], acc);
      if (slot < nslots) red[slot * oc + ol] = acc;
      __syncthreads();
      for (int s = 1; s < nslots; s <<= 1) {
        if (slot < nslots && (slot % (2 * s)) == 0 && slot + s < nslots) red[slot * oc + ol] += red[(slot + s) * oc + ol];
        __syncthreads();
      }
      if (slot == 0 && o < out) gW[((size_t)b * NS + j) * out + o] = red[ol];
      __syncthreads();
    }
  }
}

// ---- relation-chunk reductions -----------------------------------------------------------------------
// part[c, b] = sum over the E3 edges of chunk c of cbuf[e3_to_e2[e], b]   (rows of cbuf are B contiguous floats)
__global__ void __launch_bounds__(kThreads)
k_comp_chunk_reduce(const float *__restrict__ cbuf, const int32_t *__restrict__ chunk_ptr,
                    const int32_t *__restrict__ e3_to_e2, float *__restrict__ part, int B) {
  __shared__ float red[kThreads];
  const int c = blockIdx.x;
  const int e_lo = chunk_ptr[c], e_hi = chunk_ptr[c + 1];
  const int bc = min(B, kThreads), nslots = kThreads / bc;
  const int slot = threadIdx.x / bc, bl = threadIdx.x - slot * bc;
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int b = b0 + bl;
    float acc = 0.f;
    if (slot < nslots && b < B)
      for (int e = e_lo + slot; e < e_hi; e += nslots) acc += cbuf[(size_t)e3_to_e2[e] * B + b];
    if (slot < nslots) red[slot * bc + bl] = acc;
    __syncthreads();
    for (int s = 1; s < nslots; s <<= 1) {
      if (slot < nslots && (slot % (2 * s)) == 0 && slot + s < nslots) red[slot * bc + bl] += red[(slot + s) * bc + bl];
      __syncthreads();
    }
    if (slot == 0 && b < B) part[(size_t)c * B + b] = red[bl];
    __syncthreads();
  }
}

// part[c, k, o] = sum over the E3 edges of chunk c (one relation) of X[j_e, k] * t_e[o]
// thread = one k (row of g_W), OC outputs in registers; t_e staged per batch of EB edges in shared memory.
constexpr int EB = 64;
template <int OC>
__global__ void k_feat_bwd_w(const float *__restrict__ X, const float *__restrict__ gact,
                             const int32_t *__restrict__ chunk_ptr, const int32_t *__restrict__ e3_src,
                             const int32_t *__restrict__ e3_dst, const float *__restrict__ e3_val,
                             float *__restrict__ part, int in, int out) {
  __shared__ __align__(16) float Ts[EB * OC];
  __shared__ int Js[EB];
  const int c = blockIdx.x;
  const int k = blockIdx.y * blockDim.x + threadIdx.x;
  const int e_lo = chunk_ptr[c], e_hi = chunk_ptr[c + 1];
  const bool kin = k < in;
  for (int c0 = 0; c0 < out; c0 += OC) {
    float acc[OC];
#pragma unroll
    for (int o = 0; o < OC; ++o) acc[o] = 0.f;
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
          const float x = X[(size_t)Js[el] * in + k];
          const float4 *t4 = reinterpret_cast<const float4 *>(Ts + el * OC);
#pragma unroll
          for (int q = 0; q < OC / 4; ++q) {
            float4 t = t4[q];
            acc[4 * q + 0] = fmaf(x, t.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(x, t.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(x, t.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(x, t.w, acc[4 * q + 3]);
          }
        }
      }
    }
    if (kin) {
      float *pp = part + ((size_t)c * in + k) * out + c0;
#pragma unroll
      for (int o = 0; o < OC; ++o)
        if (c0 + o < out) pp[o] = acc[o];
    }
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

// ---- input gradient: g_X[j,k] = sum_{e: src=j} val_e * sum_o gact[dst_e,o] * W[r_e,k,o] -------------
// one thread per (j,k); hubs handled by the _long variant.
__global__ void __launch_bounds__(kThreads)
k_feat_bwd_x(const float *__restrict__ W, const int32_t *__restrict__ colptr, const int32_t *__restrict__ e2_dst,
             const int32_t *__restrict__ e2_rel, const float *__restrict__ e2_val, const float *__restrict__ gact,
             float *__restrict__ gX, int64_t NS, int in, int out, int thresh) {
  const int64_t x = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (x >= NS * in) return;
  const int64_t j = x / in;
  const int k = (int)(x - j * in);
  const int e_lo = colptr[j], e_hi = colptr[j + 1];
  if (thresh > 0 && e_hi - e_lo > thresh) return;
  float acc = 0.f;
  for (int e = e_lo; e < e_hi; ++e) {
    const float *gp = gact + (size_t)e2_dst[e] * out;
    const float *wp = W + ((size_t)e2_rel[e] * in + k) * out;
    float d = 0.f;
    for (int o = 0; o < out; ++o) d = fmaf(gp[o], __ldg(wp + o), d);
    acc = fmaf(e2_val[e], d, acc);
  }
  gX[x] = acc;
}
__global__ void __launch_bounds__(kThreads)
k_feat_bwd_x_long(const float *__restrict__ W, const int32_t *__restrict__ long_cols, const int32_t *__restrict__ colptr,
                  const int32_t *__restrict__ e2_dst, const int32_t *__restrict__ e2_rel,
                  const float *__restrict__ e2_val, const float *__restrict__ gact, float *__restrict__ gX, int in,
                  int out) {
  __shared__ float red[kThreads];
  const int j = long_cols[blockIdx.x];
  const int e_lo = colptr[j], e_hi = colptr[j + 1];
  const int kc = min(in, kThreads), nslots = kThreads / kc;
  const int slot = threadIdx.x / kc, kl = threadIdx.x - slot * kc;
  for (int k0 = 0; k0 < in; k0 += kc) {
    const int k = k0 + kl;
    float acc = 0.f;
    if (slot < nslots && k < in)
      for (int e = e_lo + slot; e < e_hi; e += nslots) {
        const float *gp = gact + (size_t)e2_dst[e] * out;
        const float *wp = W + ((size_t)e2_rel[e] * in + k) * out;
        float d = 0.f;
        for (int o = 0; o < out; ++o) d = fmaf(gp[o], __ldg(wp + o), d);
        acc = fmaf(e2_val[e], d, acc);
      }
    if (slot < nslots) red[slot * kc + kl] = acc;
    __syncthreads();
    for (int s = 1; s < nslots; s <<= 1) {
      if (slot < nslots && (slot % (2 * s)) == 0 && slot + s < nslots) red[slot * kc + kl] += red[(slot + s) * kc + kl];
      __syncthreads();
    }
    if (slot == 0 && k < in) gX[(size_t)j * in + k] = red[kl];
    __syncthreads();
  }
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
  const mrgcn_graph *gI = f.gI, *gF = f.gF;
  const int ND = hasI ? gI->ND : gF->ND;
  MRGCN_REQUIRE(!f.relu || f.out, MRGCN_E_BADARG, "layer_bwd: relu needs the forward output");

  // 1. gact and bias gradient
  const int nblk = (int)cdiv(ND > 0 ? ND : 1, kColRows);
  MRGCN_REQUIRE(!a->g_bias || a->colsum_ws, MRGCN_E_BADARG, "layer_bwd: colsum_ws missing");
  if (ND > 0) {
    MRGCN_PROF("act_bwd");
  k_act_bwd<<<nblk, kThreads, kThreads * sizeof(float), st>>>(a->gout, f.out, f.row_mask, a->gact,
                                                                 a->g_bias ? a->colsum_ws : nullptr, ND, out, f.relu);
    MRGCN_LAUNCH_CHECK();
  }
  if (a->g_bias) {
    if (ND > 0) {
      MRGCN_PROF("bias_reduce");
  k_seq_reduce<<<dim3((unsigned)cdiv(out, 128), 1), 128, 0, st>>>(a->colsum_ws, nullptr, nblk, out, a->g_bias);
      MRGCN_LAUNCH_CHECK();
    } else {
      MRGCN_CUDA(cudaMemsetAsync(a->g_bias, 0, sizeof(float) * out, st));
    }
  }

  // 2. identity term
  if (hasI && a->g_weight_I) {
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
      const int OP = (int)cdiv(out, OC) * OC;
      const int TJ = ident_tile(B, out, OP);
      {  // basis gradient
        const int CS = (int)cdiv(B, BC) * BC;
        size_t smem = (size_t)gI->R * CS * 4;
        MRGCN_REQUIRE(smem <= 200 * 1024, MRGCN_E_NOTSUP, "ident_bwd_w: R*B too large for shared memory");
        if (int rc = set_smem(k_ident_bwd_w, smem)) return rc;
        const int thresh = gI->n_long_cols > 0 ? gI->long_col_thresh : 0;
        unsigned grid = persistent_grid(k_ident_bwd_w, kThreads, smem, cdiv(NS, TJ));
        MRGCN_PROF("ident_bwd_w");
  k_ident_bwd_w<<<grid, kThreads, smem, st>>>(f.comp_I, gI->colptr, gI->e2_dst, gI->e2_rel, gI->e2_val, a->gact,
                                                    a->g_weight_I, (int)NS, gI->R, B, out, TJ, CS, thresh);
        MRGCN_LAUNCH_CHECK();
        if (gI->n_long_cols > 0) {
          MRGCN_PROF("ident_bwd_w_long");
  k_ident_bwd_w_long<<<(unsigned)gI->n_long_cols, kThreads, 0, st>>>(f.comp_I, gI->long_cols, gI->colptr,
                                                                             gI->e2_dst, gI->e2_rel, gI->e2_val, a->gact,
                                                                             a->g_weight_I, (int)NS, B, out);
          MRGCN_LAUNCH_CHECK();
        }
      }
      if (a->g_comp_I) {
        MRGCN_REQUIRE(a->cbuf && a->part, MRGCN_E_BADARG, "layer_bwd: cbuf/part missing");
        if (gI->E > 0) {
          const int BS = B | 1;
          size_t smem = ((size_t)B * TJ * OP + (size_t)(kThreads / 32) * 32 * BS) * 4;
          MRGCN_REQUIRE(smem <= 220 * 1024, MRGCN_E_NOTSUP, "ident_bwd_c: B*out too large for shared memory");
          unsigned grid = 0;
#define LAUNCH(OCV)                                                                                              \
  do {                                                                                                           \
    if (int rc = set_smem(k_ident_bwd_c<OCV>, smem)) return rc;                                                  \
    grid = persistent_grid(k_ident_bwd_c<OCV>, kThreads, smem, cdiv(NS, TJ));                                    \
    k_ident_bwd_c<OCV><<<grid, kThreads, smem, st>>>(f.weight_I, gI->colptr, gI->e2_src, gI->e2_dst, gI->e2_val, \
                                                     a->gact, a->cbuf, (int)NS, B, out, OP, TJ, BS);             \
  } while (0)
          MRGCN_PROF("ident_bwd_c");
  switch (OC) {
            case 4: LAUNCH(4); break;
            case 8: LAUNCH(8); break;
            case 12: LAUNCH(12); break;
            default: LAUNCH(16); break;
          }
#undef LAUNCH
          MRGCN_LAUNCH_CHECK();
          MRGCN_PROF("comp_chunk_reduce");
  k_comp_chunk_reduce<<<(unsigned)gI->n_chunks, kThreads, 0, st>>>(a->cbuf, gI->chunk_ptr, gI->e3_to_e2, a->part, B);
          MRGCN_LAUNCH_CHECK();
        }
        MRGCN_PROF("comp_reduce");
  k_seq_reduce<<<dim3((unsigned)cdiv(B, 128), (unsigned)gI->R), 128, 0, st>>>(a->part, gI->rel_chunk_ptr,
                                                                                    gI->n_chunks, B, a->g_comp_I);
        MRGCN_LAUNCH_CHECK();
      }
    }
  }

  // 3. feature term
  if (hasF) {
    const int IO = in * out;
    const float *W = B > 0 ? f.wmix : f.weight_F;
    if (a->g_weight_F || a->g_comp_F) {
      float *gW = B > 0 ? a->g_wmix : a->g_weight_F;
      MRGCN_REQUIRE(gW && a->part, MRGCN_E_BADARG, "layer_bwd: g_wmix/part missing");
      if (gF->E > 0) {
        const int OC = pick_oc(out);
        int bt = (int)cdiv(in < 256 ? in : 256, 32) * 32;
        dim3 grid((unsigned)gF->n_chunks, (unsigned)cdiv(in, bt));
        MRGCN_PROF("feat_bwd_w");
  switch (OC) {
          case 4: k_feat_bwd_w<4><<<grid, bt, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out); break;
          case 8: k_feat_bwd_w<8><<<grid, bt, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out); break;
          case 12: k_feat_bwd_w<12><<<grid, bt, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out); break;
          default: k_feat_bwd_w<16><<<grid, bt, 0, st>>>(f.X, a->gact, gF->chunk_ptr, gF->e3_src, gF->e3_dst, gF->e3_val, a->part, in, out); break;
        }
        MRGCN_LAUNCH_CHECK();
      }
      MRGCN_PROF("feat_w_reduce");
  k_seq_reduce<<<dim3((unsigned)cdiv(IO, 128), (unsigned)gF->R), 128, 0, st>>>(a->part, gF->rel_chunk_ptr,
                                                                                   gF->n_chunks, IO, gW);
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
    if (a->g_X) {
      const int64_t NS = gF->NS;
      const int thresh = gF->n_long_cols > 0 ? gF->long_col_thresh : 0;
      if (NS > 0) {
        MRGCN_PROF("feat_bwd_x");
  k_feat_bwd_x<<<(unsigned)cdiv(NS * in, kThreads), kThreads, 0, st>>>(W, gF->colptr, gF->e2_dst, gF->e2_rel, gF->e2_val,
                                                                             a->gact, a->g_X, NS, in, out, thresh);
        MRGCN_LAUNCH_CHECK();
        if (gF->n_long_cols > 0) {
          MRGCN_PROF("feat_bwd_x_long");
  k_feat_bwd_x_long<<<(unsigned)gF->n_long_cols, kThreads, 0, st>>>(W, gF->long_cols, gF->colptr, gF->e2_dst, gF->e2_rel,
                                                                            gF->e2_val, a->gact, a->g_X, in, out);
          MRGCN_LAUNCH_CHECK();
        }
      }
    }
  }
  return 0;
}
