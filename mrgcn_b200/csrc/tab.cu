// Table term of an input layer with basis decomposition (identity term of /root/reference/mrgcn/layers/graph.py:67-75
// and, when the features were projected per basis first, the feature term of graph.py:83-95 re-associated):
//
//   msg[e, :] = val_e * ( sum_b comp_I[r_e, b] * T_I[b, j_e, :]  +  sum_b comp_F[r_e, b] * T_P[j_e, b, :] )
//
// T_I = weight_I in the reference layout [B_I][NS][out] (basis-major); T_P[j, b, :] = X[j, :] . V_F[b] is the per-basis
// projection of the features (feat_proj.cu, node-major [NS][B_F][out]).  Both tables are indexed by the SOURCE node, so all
// three passes walk the source-major edge order E2 and keep the table rows of a source in REGISTERS:
//
//   tab_msg_fwd : a group of lanes owns (task, output pair[, half of the bases]); a task is up to 32 consecutive E2 edges of
//                 one source.  The group loads the source's table rows once (coalesced: the lanes of a group cover the
//                 contiguous `out` floats of a row, consecutive sources are consecutive groups) and then produces one
//                 message per edge from comp rows read out of shared memory (16-byte loads, broadcast inside the group).
//   tab_bwd_w   : same ownership, one task per (non-hub) source; accumulates g_T_I[b, j, :] = sum_e comp_I[r_e, b] t_e in
//                 registers, t_e = val_e * gact[dst_e, :] gathered straight from global memory, four edges in flight.
//   tab_bwd_c   : a lane owns (task, chunk of bases) and all outputs: c_e[b] = <T_I[b, j_e, :], t_e> goes to a shared-memory
//                 tile (<= 543 edges x B), which the CTA then sums per relation in a precomputed tile-local relation order
//                 (pieces of <= 32 edges) -> one record per piece; g_comp_I[r, :] = fixed-order sum of the records of r.
//                 Nothing of size E x B touches HBM.
// The reference layout of weight_I makes a per-edge gather of its B rows hostile (B separate 4*out-byte segments N*out*4
// bytes apart); here every table row is read from HBM exactly once per pass, straight into registers: no shared-memory
// traffic for the table at all (the round-1 kernels were bound by the shared-memory pipe: one 8-byte read per 2 FMAs).
// All reductions have a fixed order: bit-reproducible, no float atomics.
#include <stdlib.h>

#include "common.cuh"
#include "pipeline.cuh"
#include "rgcn_internal.cuh"

namespace mrgcn {
namespace {

constexpr int kTabThreads = 256;
constexpr int kTabWarps = kTabThreads / 32;
constexpr int LT = 32;  // edges per task (mrgcn_tab_plan.lt)

struct TabTables {
  const float *TI;  // [BI][NS][out] or NULL
  const float *TP;  // [NS][BF][out] or NULL
  int BI, BF, out;
  int64_t NS;
};

// lane geometry shared by the kernels that own (task, output pair, base split)
struct Geo {
  int LPT, TPW, tslot, within, hs, op0;
  bool lane_on;
  __device__ Geo(int GS, int HS, int lane) {
    LPT = GS > 32 ? 32 : GS * HS;
    TPW = 32 / LPT;
    tslot = lane / LPT;
    within = lane - tslot * LPT;
    hs = GS > 32 ? 0 : within / GS;
    op0 = GS > 32 ? within : within - hs * GS;
    lane_on = tslot < TPW;
  }
};

__device__ __forceinline__ void fill_comp(float *comp_s, const float *__restrict__ compI, const float *__restrict__ compF,
                                          int R, int BI, int BF, int CSP) {
  for (int x = threadIdx.x; x < R * CSP; x += kTabThreads) {
    const int r = x / CSP, bb = x - r * CSP;
    float c = 0.f;
    if (bb < BI) c = __ldg(compI + (size_t)r * BI + bb);
    else if (bb < BI + BF) c = __ldg(compF + (size_t)r * BF + (bb - BI));
    comp_s[x] = c;
  }
}

// 4-byte asynchronous global -> shared copy (LDGSTS): metadata of the NEXT item lands while the current one is computed
__device__ __forceinline__ void cp4(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// {source, first E2 edge, number of edges, -} of task / source t, or an empty task
__device__ __forceinline__ int4 ld_task(const int4 *__restrict__ tasks, int t, int n, bool lane_on) {
  return (lane_on && t < n) ? __ldg(tasks + t) : make_int4(0, 0, 0, 0);
}

template <bool EVEN>
__device__ __forceinline__ float2 ld2(const float *p, bool second) {
  if constexpr (EVEN) return *reinterpret_cast<const float2 *>(p);
  else return make_float2(p[0], second ? p[1] : 0.f);
}

// The BPT table rows (x NOP output pairs) of one lane: bases [b0, b0 + BPT) of source j, the first n1 of them from the
// basis-major identity table, the rest from the node-major projection; everything beyond Btot or `out` is zero.
// The two common cases - all of the lane's bases in ONE table - walk a pointer (two instructions per load).
// Lanes without a task pass j = 0: they load (valid) rows they never use, so that the whole warp takes the same path.
template <int BPT, int NOP, bool EVEN>
__device__ __forceinline__ void load_rows(float2 (&T)[BPT][NOP], const TabTables &tb, int b0, int64_t j, int op0) {
  const int out = tb.out, Btot = tb.BI + tb.BF;
  const int n1 = min(max(tb.BI - b0, 0), BPT);
  const int nvalid = min(max(Btot - b0, 0), BPT);
  const size_t s1 = (size_t)tb.NS * out;
  const bool o_ok = NOP > 1 || 2 * op0 < out;
  if (nvalid == BPT && o_ok && (n1 == BPT || n1 == 0)) {
    const float *p = n1 ? tb.TI + ((size_t)b0 * tb.NS + j) * out + 2 * op0 : tb.TP + ((size_t)j * tb.BF + (b0 - tb.BI)) * out + 2 * op0;
    const size_t step = n1 ? s1 : (size_t)out;
#pragma unroll
    for (int b = 0; b < BPT; ++b, p += step) {
#pragma unroll
      for (int q = 0; q < NOP; ++q) {
        const int o = 2 * (op0 + 32 * q);
        T[b][q] = (NOP == 1 || o < out) ? ld2<EVEN>(p + 64 * q, o + 1 < out) : make_float2(0.f, 0.f);
      }
    }
    return;
  }
  const float *p1 = tb.TI + ((size_t)min(b0, tb.BI - 1) * tb.NS + j) * out + 2 * op0;
  const float *p2 = tb.TP ? tb.TP + ((size_t)j * tb.BF + max(b0 - tb.BI, 0)) * out + 2 * op0 : p1;
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    const float *p = b < n1 ? p1 + (size_t)b * s1 : p2 + (ptrdiff_t)(b - n1) * out;
#pragma unroll
    for (int q = 0; q < NOP; ++q) {
      const int o = 2 * (op0 + 32 * q);
      T[b][q] = (b < nvalid && o < out) ? ld2<EVEN>(p + 64 * q, o + 1 < out) : make_float2(0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
template <int BPT, int NOP, bool EVEN>
__global__ void __launch_bounds__(kTabThreads, 2)
k_tab_msg_fwd(TabTables tb, const float *__restrict__ compI, const float *__restrict__ compF, int R,
              const int4 *__restrict__ tasks, int n_tasks, const int32_t *__restrict__ e2_rel,
              const float *__restrict__ e2_val, float *__restrict__ msg, int ms, int GS, int HS, int CSP) {
  extern __shared__ __align__(16) float smem[];
  float *comp_s = smem;                                             // [R][CSP]
  int2 *meta = reinterpret_cast<int2 *>(smem + (size_t)R * CSP);    // [warps][TPW*LT] (rel, val)
  fill_comp(comp_s, compI, compF, R, tb.BI, tb.BF, CSP);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Geo g(GS, HS, lane);
  const int out = tb.out;
  int2 *tmeta = meta + ((size_t)warp * g.TPW + (g.lane_on ? g.tslot : 0)) * LT;   // this lane's task
  const float *cbase = comp_s + g.hs * BPT;
  const int n_items = (n_tasks + g.TPW - 1) / g.TPW;
  const bool writer = g.hs == 0 && g.lane_on;
  const int stride = gridDim.x * kTabWarps;
  int wi = blockIdx.x * kTabWarps + warp;
  // 16-byte task descriptors {source, first edge, edges}, fetched one item ahead
  int4 cur = ld_task(tasks, wi * g.TPW + g.tslot, n_tasks, g.lane_on && wi < n_items);
  for (; wi < n_items; wi += stride) {
    const int4 nxt = ld_task(tasks, (wi + stride) * g.TPW + g.tslot, n_tasks, g.lane_on && wi + stride < n_items);
    const int len = cur.z;
    __syncwarp();                              // previous item done with the metadata
    for (int s = g.within; s < len; s += g.LPT)
      tmeta[s] = make_int2(ldg_stream(e2_rel + cur.y + s), __float_as_int(ldg_stream(e2_val + cur.y + s)));
    float2 T[BPT][NOP];
    load_rows<BPT, NOP, EVEN>(T, tb, g.hs * BPT, cur.x, g.op0);
    __syncwarp();
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    float *mrow = msg + (size_t)cur.y * ms + 2 * g.op0;
    // one edge of the task per step: comp row of its relation (16-byte shared loads) against the register-resident rows
    auto edge = [&](int s, bool live) {
      int2 m = make_int2(0, 0);
      if (live) m = tmeta[s];
      const float *cr = cbase + m.x * CSP;
      float2 acc[NOP], acc1[NOP];   // two chains per output pair (even / odd groups of four bases), added at the end
#pragma unroll
      for (int q = 0; q < NOP; ++q) { acc[q] = make_float2(0.f, 0.f); acc1[q] = make_float2(0.f, 0.f); }
#pragma unroll
      for (int b = 0; b < BPT; b += 4) {
        const float4 c = *reinterpret_cast<const float4 *>(cr + b);
#pragma unroll
        for (int q = 0; q < NOP; ++q) {
          float2 &a = ((b >> 2) & 1) ? acc1[q] : acc[q];
          fma2(a, c.x, T[b][q]);
          fma2(a, c.y, T[b + 1][q]);
          fma2(a, c.z, T[b + 2][q]);
          fma2(a, c.w, T[b + 3][q]);
        }
      }
#pragma unroll
      for (int q = 0; q < NOP; ++q) { acc[q].x += acc1[q].x; acc[q].y += acc1[q].y; }
      if (HS == 2) {  // partial sums of the two base splits
#pragma unroll
        for (int q = 0; q < NOP; ++q) {
          acc[q].x += __shfl_down_sync(0xffffffffu, acc[q].x, GS);
          acc[q].y += __shfl_down_sync(0xffffffffu, acc[q].y, GS);
        }
      } else if (HS > 2) {   // ... of more splits, added in split order
#pragma unroll
        for (int q = 0; q < NOP; ++q) {
          const float2 own = acc[q];
          for (int h = 1; h < HS; ++h) {
            acc[q].x += __shfl_down_sync(0xffffffffu, own.x, h * GS);
            acc[q].y += __shfl_down_sync(0xffffffffu, own.y, h * GS);
          }
        }
      }
      if (live && writer) {
        // rows are padded to `ms` floats; the pad columns are never read into a stored sum (agg kernels) and stay unwritten
        const float v = __int_as_float(m.y);
        float *row = mrow + (size_t)s * ms;
#pragma unroll
        for (int q = 0; q < NOP; ++q)
          if (NOP == 1 || 2 * (g.op0 + 32 * q) < out) *reinterpret_cast<float2 *>(row + 64 * q) = make_float2(v * acc[q].x, v * acc[q].y);
      }
    };
    int s = 0;
    for (; s + 1 < maxlen; s += 2) {      // two edges per trip: their shared-memory loads and FMA chains interleave
      edge(s, s < len);
      edge(s + 1, s + 1 < len);
    }
    if (s < maxlen) edge(s, s < len);
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct __align__(16) EdgeMeta { int d, r; float v; int pad; };

template <int BPT, int NOP, bool EVEN>
__global__ void __launch_bounds__(kTabThreads, (BPT * NOP > 24) ? 1 : 2)
k_tab_bwd_w(const float *__restrict__ compI, int R, int BI, int64_t NS, int out, const int4 *__restrict__ wtasks, int n_src,
            const int32_t *__restrict__ e2_dst, const int32_t *__restrict__ e2_rel, const float *__restrict__ e2_val,
            const float *__restrict__ gact, float *__restrict__ gW, int GS, int HS, int CSP) {
  extern __shared__ __align__(16) float smem[];
  float *comp_s = smem;                                                     // [R][CSP]
  EdgeMeta *meta = reinterpret_cast<EdgeMeta *>(smem + (size_t)R * CSP);    // [warps][2][TPW*32], double buffered
  fill_comp(comp_s, compI, nullptr, R, BI, 0, CSP);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Geo g(GS, HS, lane);
  EdgeMeta *tmeta = meta + ((size_t)warp * 2 * g.TPW + (g.lane_on ? g.tslot : 0)) * 32;
  const int bufstride = g.TPW * 32;
  const float *cbase = comp_s + g.hs * BPT;
  const int n_items = (n_src + g.TPW - 1) / g.TPW;
  const int stride = gridDim.x * kTabWarps;
  auto stage = [&](int buf, const int4 &ti, int c0) {      // edges [c0, c0 + 32) of this lane's source
    EdgeMeta *dst = tmeta + buf * bufstride;
    const int n = min(32, ti.z - c0);
    for (int s = g.within; s < n; s += g.LPT) {
      const int e = ti.y + c0 + s;
      cp4(&dst[s].d, e2_dst + e);
      cp4(&dst[s].r, e2_rel + e);
      cp4(&dst[s].v, e2_val + e);
    }
    cp_async_commit();
  };
  int wi = blockIdx.x * kTabWarps + warp;
  int4 cur = ld_task(wtasks, wi * g.TPW + g.tslot, n_src, g.lane_on && wi < n_items);
  int4 nxt = ld_task(wtasks, (wi + stride) * g.TPW + g.tslot, n_src, g.lane_on && wi + stride < n_items);
  stage(0, cur, 0);
  int buf = 0;
  for (; wi < n_items; wi += stride) {
    const int4 nn = ld_task(wtasks, (wi + 2 * stride) * g.TPW + g.tslot, n_src, g.lane_on && wi + 2 * stride < n_items);
    const bool on = g.lane_on && wi * g.TPW + g.tslot < n_src;
    const int j = cur.x, len = cur.z;
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    float2 acc[BPT][NOP];
#pragma unroll
    for (int b = 0; b < BPT; ++b)
#pragma unroll
      for (int q = 0; q < NOP; ++q) acc[b][q] = make_float2(0.f, 0.f);
    int c0 = 0;
    do {
      // the next 32 edges of these sources, or the first 32 of the next item's, are fetched under this chunk's arithmetic
      if (c0 + 32 < maxlen) stage(buf ^ 1, cur, c0 + 32);
      else stage(buf ^ 1, nxt, 0);
      cp_async_wait<1>();
      __syncwarp();
      const EdgeMeta *mbuf = tmeta + buf * bufstride;
      const int mylen = min(32, max(0, len - c0));
      const int nch = min(32, maxlen - c0);
      // groups of U edges; the gathers of group k+1 are in flight while group k is accumulated
      constexpr int U = (BPT * NOP >= 16) ? 2 : 4;
      EdgeMeta m[U], mn[U];
      float2 t[U][NOP], tn[U][NOP];
      auto fetch = [&](int s0, EdgeMeta (&mm)[U], float2 (&tt)[U][NOP]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool live = s0 + u < mylen;
          mm[u].d = 0; mm[u].r = 0; mm[u].v = 0.f;
          if (live) mm[u] = mbuf[s0 + u];
          const float *gp = gact + (size_t)mm[u].d * out + 2 * g.op0;
#pragma unroll
          for (int q = 0; q < NOP; ++q) {
            const int o = 2 * (g.op0 + 32 * q);
            tt[u][q] = (live && o < out) ? ld2<EVEN>(gp + 64 * q, o + 1 < out) : make_float2(0.f, 0.f);
          }
        }
      };
      fetch(0, mn, tn);
      for (int s0 = 0; s0 < nch; s0 += U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          m[u] = mn[u];
#pragma unroll
          for (int q = 0; q < NOP; ++q) t[u][q] = tn[u][q];
        }
        if (s0 + U < nch) fetch(s0 + U, mn, tn);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float *cr = cbase + m[u].r * CSP;
#pragma unroll
          for (int q = 0; q < NOP; ++q) { t[u][q].x *= m[u].v; t[u][q].y *= m[u].v; }
#pragma unroll
          for (int b = 0; b < BPT; b += 4) {
            const float4 c = *reinterpret_cast<const float4 *>(cr + b);
#pragma unroll
            for (int q = 0; q < NOP; ++q) {
              fma2(acc[b][q], c.x, t[u][q]);
              fma2(acc[b + 1][q], c.y, t[u][q]);
              fma2(acc[b + 2][q], c.z, t[u][q]);
              fma2(acc[b + 3][q], c.w, t[u][q]);
            }
          }
        }
      }
      __syncwarp();      // everybody is done with this buffer before it is staged again
      buf ^= 1;
      c0 += 32;
    } while (c0 < maxlen);
    if (on) {
      const int b0 = g.hs * BPT;
      const size_t s1 = (size_t)NS * out;
      float *row = gW + ((size_t)b0 * NS + j) * out + 2 * g.op0;
#pragma unroll
      for (int b = 0; b < BPT; ++b) {
        if (b0 + b < BI) {
#pragma unroll
          for (int q = 0; q < NOP; ++q) {
            const int o = 2 * (g.op0 + 32 * q);
            if (o < out) {
              float *p = row + (size_t)b * s1 + 64 * q;
              if (EVEN || o + 1 < out) {
                if constexpr (EVEN) *reinterpret_cast<float2 *>(p) = acc[b][q];
                else { p[0] = acc[b][q].x; p[1] = acc[b][q].y; }
              } else {
                p[0] = acc[b][q].x;
              }
            }
          }
        }
      }
    }
    cur = nxt;
    nxt = nn;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------------------------
constexpr int kCThreads = 128;   // comp-gradient kernel: small CTAs, four per SM, so that a CTA waiting at its tile barrier
constexpr int kCWarps = kCThreads / 32;   // leaves the SM to the others

template <int BC, int OP, bool EVEN>
__global__ void __launch_bounds__(kCThreads, 4)
k_tab_bwd_c(const float *__restrict__ TI, int BI, int64_t NS, int out, const int32_t *__restrict__ colptr,
            const int32_t *__restrict__ task_src, const int32_t *__restrict__ task_lo,
            const int32_t *__restrict__ tile_task_ptr, const int32_t *__restrict__ tile_e0, int n_tiles,
            const int32_t *__restrict__ e2_dst, const float *__restrict__ e2_val, const float *__restrict__ gact,
            const int32_t *__restrict__ tperm, const int32_t *__restrict__ piece_ptr,
            const int32_t *__restrict__ tile_piece_ptr, float *__restrict__ rec, int LPT, int BSP, int tile_slots) {
  extern __shared__ __align__(16) float smem[];
  float *Cs = smem;                                                       // [tile_slots][BSP]
  int2 *meta = reinterpret_cast<int2 *>(smem + (size_t)tile_slots * BSP); // [warps][TPW*LT] (dst, val)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int TPW = 32 / LPT;
  const int tslot = lane / LPT, gch = lane - tslot * LPT;   // base chunk of this lane
  const bool lane_on = tslot < TPW;
  int2 *tmeta = meta + ((size_t)warp * TPW + (lane_on ? tslot : 0)) * LT;
  const size_t s1 = (size_t)NS * out;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int e0 = tile_e0[tile], t_lo = tile_task_ptr[tile], t_hi = tile_task_ptr[tile + 1];
    const int n_items = (t_hi - t_lo + TPW - 1) / TPW;
    for (int wi = warp; wi < n_items; wi += kCWarps) {
      const int t = t_lo + wi * TPW + tslot;
      const bool on = lane_on && t < t_hi;
      int j = 0, lo = 0, len = 0;
      if (on) { j = task_src[t]; lo = task_lo[t]; len = min(LT, colptr[j + 1] - lo); }
      __syncwarp();
      for (int s = gch; s < len; s += LPT)
        tmeta[s] = make_int2(ldg_stream(e2_dst + lo + s), __float_as_int(ldg_stream(e2_val + lo + s)));
      float2 T[BC][OP];
      {
        const int nvalid = on ? min(max(BI - gch * BC, 0), BC) : 0;
        const float *p1 = TI + ((size_t)min(gch * BC, BI - 1) * NS + j) * out;
#pragma unroll
        for (int b = 0; b < BC; ++b)
#pragma unroll
          for (int q = 0; q < OP; ++q)
            T[b][q] = (b < nvalid && 2 * q < out) ? ld2<EVEN>(p1 + (size_t)b * s1 + 2 * q, 2 * q + 1 < out) : make_float2(0.f, 0.f);
      }
      __syncwarp();
      const int maxlen = __reduce_max_sync(0xffffffffu, len);
      // t_e of the next edge is fetched while the current one is multiplied
      float2 tn[OP];
      float vn = 0.f;
      {
        int2 m = make_int2(0, 0);
        if (0 < len) m = tmeta[0];
        vn = __int_as_float(m.y);
        const float *gp = gact + (size_t)m.x * out;
#pragma unroll
        for (int q = 0; q < OP; ++q) tn[q] = (0 < len && 2 * q < out) ? ld2<EVEN>(gp + 2 * q, 2 * q + 1 < out) : make_float2(0.f, 0.f);
      }
      float *crow = Cs + (size_t)(lo - e0) * BSP + gch * BC;
      for (int s = 0; s < maxlen; ++s, crow += BSP) {
        const bool live = s < len;
        float2 tv[OP];
        const float v = vn;
#pragma unroll
        for (int q = 0; q < OP; ++q) tv[q] = make_float2(tn[q].x * v, tn[q].y * v);
        {
          const bool nlive = s + 1 < len;
          int2 m = make_int2(0, 0);
          if (nlive) m = tmeta[s + 1];
          vn = __int_as_float(m.y);
          const float *gp = gact + (size_t)m.x * out;
#pragma unroll
          for (int q = 0; q < OP; ++q) tn[q] = (nlive && 2 * q < out) ? ld2<EVEN>(gp + 2 * q, 2 * q + 1 < out) : make_float2(0.f, 0.f);
        }
        if (live) {
#pragma unroll
          for (int b = 0; b < BC; b += 4) {
            float c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
              for (int q = 0; q < OP; ++q) fma2v(a2, T[b + u][q], tv[q]);
              c[u] = a2.x + a2.y;
            }
            *reinterpret_cast<float4 *>(crow + b) = make_float4(c[0], c[1], c[2], c[3]);
          }
        }
      }
    }
    __syncthreads();
    // per-relation pieces of the tile, summed in the precomputed order
    const int p_lo = tile_piece_ptr[tile], p_hi = tile_piece_ptr[tile + 1];
    for (int x = threadIdx.x; x < (p_hi - p_lo) * BI; x += kCThreads) {
      const int pc = p_lo + x / BI, b = x - (pc - p_lo) * BI;
      const int q_lo = piece_ptr[pc], q_hi = piece_ptr[pc + 1];
      float acc = 0.f;
      for (int q = q_lo; q < q_hi; ++q) acc += Cs[(size_t)tperm[q] * BSP + b];
      rec[(size_t)pc * BI + b] = acc;
    }
    __syncthreads();
  }
}

template <class K>
static int tab_set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) MRGCN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
template <class K>
static unsigned tab_grid(K kernel, size_t smem, int64_t max_ctas, int threads = kTabThreads) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t g = (int64_t)kNumSMs * per_sm;
  return (unsigned)(g < max_ctas ? g : (max_ctas > 0 ? max_ctas : 1));
}

}  // namespace

// ---- geometry --------------------------------------------------------------------------------------------------
// GS = output pairs, HS = base splits, BPT = bases per lane (multiple of 4), NOP = output pairs per lane.
bool tab_geometry(int Btot, int out, TabGeom &g, int max_bpt) {
  if (Btot <= 0 || out <= 0 || (out & 1)) return false;   // odd widths stay on the tile-staging kernels (8-byte row accesses)
  g.GS = out / 2;
  if (g.GS > 32) {
    g.HS = 1;
    g.NOP = (g.GS + 31) / 32;
    g.BPT = ((Btot + 3) / 4) * 4;
    if (g.NOP == 3) g.NOP = 4;
    if (g.BPT > 4) g.BPT = 8;
    if (!((g.NOP == 2 || g.NOP == 4) && Btot <= 8)) return false;
  } else {
    g.NOP = 1;
    const int max_hs = 32 / g.GS;
    g.HS = 0;
    for (int hs = 1; hs <= max_hs; ++hs) {
      const int bpt = (((Btot + hs - 1) / hs + 3) / 4) * 4;
      if (bpt <= max_bpt) { g.HS = hs; g.BPT = bpt; break; }
    }
    if (!g.HS) {   // cannot honour the preferred width: take the widest split
      g.HS = max_hs;
      g.BPT = (((Btot + max_hs - 1) / max_hs + 3) / 4) * 4;
      if (g.BPT > 40) return false;
    }
    // instantiated widths: 4, 8, 12, 16, 20, 24, 32, 40
    if (g.BPT > 24 && g.BPT <= 32) g.BPT = 32;
    else if (g.BPT > 32) g.BPT = 40;
  }
  const int cs = g.HS * g.BPT;
  g.CSP = ((cs / 4) & 1) ? cs : cs + 4;   // row pitch with an odd number of 16-byte groups: rows spread over the banks
  return true;
}

bool tab_c_geometry(int BI, int out, int &BC, int &OP) {
  if (BI < 8 || (out & 1)) return false;   // tiny B: the E x B scratch of the generic path is tiny as well
  if (out <= 4) { BC = 8; OP = 2; }
  else if (out <= 10) { BC = 8; OP = 5; }
  else if (out <= 16) { BC = 4; OP = 8; }
  else return false;
  return (BI + BC - 1) / BC <= 32;
}

#define TAB_DISPATCH(CALL)                                                                               \
  do {                                                                                                   \
    if (geo.NOP == 1) {                                                                                  \
      switch (geo.BPT) {                                                                                 \
        case 4: CALL(4, 1); break;                                                                       \
        case 8: CALL(8, 1); break;                                                                       \
        case 12: CALL(12, 1); break;                                                                     \
        case 16: CALL(16, 1); break;                                                                     \
        case 20: CALL(20, 1); break;                                                                     \
        case 24: CALL(24, 1); break;                                                                     \
        case 32: CALL(32, 1); break;                                                                     \
        default: CALL(40, 1); break;                                                                     \
      }                                                                                                  \
    } else if (geo.NOP == 2) {                                                                           \
      if (geo.BPT == 4) CALL(4, 2); else CALL(8, 2);                                                     \
    } else {                                                                                             \
      if (geo.BPT == 4) CALL(4, 4); else CALL(8, 4);                                                     \
    }                                                                                                    \
  } while (0)

int launch_tab_msg_fwd(const mrgcn_graph *g, const mrgcn_tab_plan *pl, const float *TI, const float *compI, int BI,
                       const float *TP, const float *compF, int BF, int out, float *msg, cudaStream_t st) {
  TabGeom geo;
  MRGCN_REQUIRE(tab_geometry(BI + BF, out, geo, 40), MRGCN_E_NOTSUP, "tab_msg_fwd: unsupported shape B=%d out=%d", BI + BF, out);
  if (pl->n_tasks == 0) return 0;
  TabTables tb{TI, TP, BI, BF, out, (int64_t)g->NS};
  const int LPT = geo.GS > 32 ? 32 : geo.GS * geo.HS, TPW = 32 / LPT;
  const size_t smem = ((size_t)g->R * geo.CSP) * 4 + (size_t)kTabWarps * TPW * LT * sizeof(int2);
  MRGCN_REQUIRE(smem <= 110 * 1024, MRGCN_E_NOTSUP, "tab_msg_fwd: R*B too large for shared memory (%zu B)", smem);
  const int ms = msg_stride(out);
  const int64_t items = cdiv(pl->n_tasks, TPW);
  MRGCN_PROF("tab_msg_fwd");
#define CALL(BPTV, NOPV)                                                                                              \
  do {                                                                                                                \
    if (int rc = tab_set_smem(k_tab_msg_fwd<BPTV, NOPV, true>, smem)) return rc;                                            \
    const unsigned grid = tab_grid(k_tab_msg_fwd<BPTV, NOPV, true>, smem, cdiv(items, kTabWarps));                          \
    k_tab_msg_fwd<BPTV, NOPV, true><<<grid, kTabThreads, smem, st>>>(tb, compI, compF, g->R,                             \
                                                               reinterpret_cast<const int4 *>(pl->tasks4), pl->n_tasks, \
                                                               g->e2_rel, g->e2_val, msg, ms, geo.GS, geo.HS, geo.CSP); \
  } while (0)
  TAB_DISPATCH(CALL);
#undef CALL
  MRGCN_LAUNCH_CHECK();
  return 0;
}

int launch_tab_bwd_w(const mrgcn_graph *g, const mrgcn_tab_plan *pl, const float *compI, int BI, int out,
                     const float *gact, float *gW, cudaStream_t st) {
  TabGeom geo;
  MRGCN_REQUIRE(tab_geometry(BI, out, geo, kBwdWBpt), MRGCN_E_NOTSUP, "tab_bwd_w: unsupported shape B=%d out=%d", BI, out);
  if (pl->n_wsrc == 0) return 0;
  const int LPT = geo.GS > 32 ? 32 : geo.GS * geo.HS, TPW = 32 / LPT;
  const size_t smem = ((size_t)g->R * geo.CSP) * 4 + (size_t)kTabWarps * 2 * TPW * 32 * sizeof(EdgeMeta);
  MRGCN_REQUIRE(smem <= 110 * 1024, MRGCN_E_NOTSUP, "tab_bwd_w: R*B too large for shared memory (%zu B)", smem);
  const int64_t items = cdiv(pl->n_wsrc, TPW);
  MRGCN_PROF("tab_bwd_w");
#define CALL(BPTV, NOPV)                                                                                              \
  do {                                                                                                                \
    if (int rc = tab_set_smem(k_tab_bwd_w<BPTV, NOPV, true>, smem)) return rc;                                              \
    const unsigned grid = tab_grid(k_tab_bwd_w<BPTV, NOPV, true>, smem, cdiv(items, kTabWarps));                            \
    k_tab_bwd_w<BPTV, NOPV, true><<<grid, kTabThreads, smem, st>>>(compI, g->R, BI, (int64_t)g->NS, out,                   \
                                                             reinterpret_cast<const int4 *>(pl->wtasks4), pl->n_wsrc, \
                                                             g->e2_dst, g->e2_rel, g->e2_val, gact, gW, geo.GS,       \
                                                             geo.HS, geo.CSP);                                        \
  } while (0)
  TAB_DISPATCH(CALL);
#undef CALL
  MRGCN_LAUNCH_CHECK();
  return 0;
}

int launch_tab_bwd_c(const mrgcn_graph *g, const mrgcn_tab_plan *pl, const float *TI, int BI, int out, const float *gact,
                     float *rec, cudaStream_t st) {
  int BC = 0, OP = 0;
  MRGCN_REQUIRE(tab_c_geometry(BI, out, BC, OP), MRGCN_E_NOTSUP, "tab_bwd_c: unsupported shape B=%d out=%d", BI, out);
  if (pl->n_tiles == 0) return 0;
  const int LPT = (BI + BC - 1) / BC, TPW = 32 / LPT;
  const int BSP = LPT * BC;
  const size_t smem = ((size_t)pl->tile_slots * BSP) * 4 + (size_t)kCWarps * TPW * LT * sizeof(int2);
  MRGCN_REQUIRE(smem <= 110 * 1024, MRGCN_E_NOTSUP, "tab_bwd_c: tile too large for shared memory (%zu B)", smem);
  MRGCN_PROF("tab_bwd_c");
#define CALL(BCV, OPV)                                                                                                 \
  do {                                                                                                                 \
    if (int rc = tab_set_smem(k_tab_bwd_c<BCV, OPV, true>, smem)) return rc;                                                 \
    const unsigned grid = tab_grid(k_tab_bwd_c<BCV, OPV, true>, smem, pl->n_tiles, kCThreads);                               \
    k_tab_bwd_c<BCV, OPV, true><<<grid, kCThreads, smem, st>>>(TI, BI, (int64_t)g->NS, out, g->colptr, pl->task_src,      \
                                                           pl->task_lo, pl->tile_task_ptr, pl->tile_e0, pl->n_tiles,   \
                                                           g->e2_dst, g->e2_val, gact, pl->tperm, pl->piece_ptr,       \
                                                           pl->tile_piece_ptr, rec, LPT, BSP, pl->tile_slots);         \
  } while (0)
  if (OP == 2) CALL(8, 2);
  else if (OP == 5) CALL(8, 5);
  else CALL(4, 8);
#undef CALL
  MRGCN_LAUNCH_CHECK();
  return 0;
}

}  // namespace mrgcn

// bit 0: tab_msg_fwd, bit 1: tab_bwd_w (both need B_I + B_F resp. B_I bases to fit the lane geometry),
// bit 2: tab_bwd_c (no E x B scratch: `cbuf` then holds the n_pieces x B records)
// MRGCN_TAB=<mask> in the environment (or mrgcn_set_tab_mask) forces the table-term kernels; without it the choice is by
// shape, from measurements (profiles/r02_experiments.md):
//   bit 0 always - the message kernel + projection beat the round-1 forward on every shape;
//   bit 1 for wide outputs (out >= 64) - on the FB15k-237 layer (out 200, 2 bases) tab_bwd_w takes 0.09 ms where the
//         tile-staging ident_bwd_w takes 1.9 ms; on AM (out 10, 40 bases) it loses 2.5 ms to 1.8 ms;
//   bit 2 never by default - tab_bwd_c ties on FB15k-237 (0.76 ms both) and loses on AM (3.8 ms vs 2.2 + 0.5 ms).
static int g_tab_mask = -1;      // -1: not read yet, -2: automatic
static int tab_env_mask() {
  if (g_tab_mask == -1) { const char *e = getenv("MRGCN_TAB"); g_tab_mask = e ? atoi(e) : -2; }
  return g_tab_mask;
}
extern "C" void mrgcn_set_tab_mask(int32_t mask) { g_tab_mask = mask < 0 ? -2 : mask; }

extern "C" int32_t mrgcn_tab_mode(int32_t BI, int32_t BF, int32_t out) {
  mrgcn::TabGeom geo;
  int m = 0;
  if (BI > 0 && mrgcn::tab_geometry(BI + BF, out, geo, 40)) m |= 1;
  if (BI > 0 && mrgcn::tab_geometry(BI, out, geo, mrgcn::kBwdWBpt)) m |= 2;
  int BC, OP;
  if (BI > 0 && mrgcn::tab_c_geometry(BI, out, BC, OP)) m |= 4;
  const int forced = tab_env_mask();
  return m & (forced >= 0 ? forced : (1 | (out >= 64 ? 2 : 0)));
}
