// Identity-term backward with basis decomposition in ONE pass over the source-major edge order E2
// (autograd through /root/reference/mrgcn/layers/graph.py:69-75):
//
//   t_e                = val_e * gact[dst_e, :]
//   g_weight_I[b,j,:]  = sum_{e: src = j} comp_I[rel_e, b] * t_e                    (written once, 4*B*NS*out bytes)
//   cbuf[e3(e), b]     = <weight_I[b, j, :], t_e>     ->  g_comp_I[r, b] = fixed-order sum over the E3 range of r
//
// Round 1 did this in three kernels (ident_bwd_w, ident_bwd_c, comp_chunk_reduce) that read the edges, gact and the
// scratch rows several times.  Here a lane owns one basis b (two when B > 32): the rows weight_I[b, j, :] and the
// accumulators g_weight_I[b, j, :] of the warp's current source live in registers, so per edge the warp reads only the
// edge's t_e (broadcast) and comp_I[rel_e, :] (one coalesced row) from shared memory.  The scratch rows are written at
// the edge's position in the relation-major order E3, so that the per-relation reduction streams them contiguously.
//
// One persistent CTA per SM walks tiles of TJ consecutive sources.  Five roles, joined by mbarriers only (no CTA-wide
// barrier after start-up), over TWO rings of shared memory so that the edge side runs ahead of the table side:
//   edge ring (Sm slots: header, the tile's slices of e2_dst / e2_rel / e2_val / e2_to_e3, t_e rows)
//     edge loader    cp.async.bulk (TMA engine) of the four slices; colptr of the tile into the header
//     gather warps   t_e of every staged edge - all gathers of a tile in flight at once, a few tiles before they are used
//   table ring (Sv slots: the B runs weight_I[b, j0 : j0+TJ, :])
//     table loader   cp.async.bulk, one run per basis
//     compute warps  sources of the tile, claimed from a counter in the slot header; g_weight_I goes back IN PLACE
//     storer         a finished slot leaves for HBM with cp.async.bulk shared -> global (B runs: full-line stores)
// weight_I, g_weight_I and the scratch rows are touched once: they carry the L2 evict-first policy, so that gact (the only
// re-used operand, ND*out floats) stays L2-resident under 10 GB of streaming traffic.
// Hub sources (more than `thresh` edges) keep their g_weight_I in k_ident_bwd_w_long (launched afterwards: it overwrites
// the rows this kernel leaves untouched); their scratch rows are produced here, chunk by chunk across the compute warps.
// Every sum is taken in edge order by one thread: bit-reproducible, no atomics on floats.
#include <stdlib.h>

#include "common.cuh"
#include "pipeline.cuh"
#include "ident_pipe.cuh"
#include "rgcn_internal.cuh"

namespace mrgcn {
namespace {

constexpr int kCW = 12;  // compute warps (19 warps in all: at most 5 per scheduler, so 96 registers per thread)
constexpr int kGW = 4;   // gather warps: warp g takes the tiles k = g (mod kGW) whole, so kGW tiles' gathers are in flight
constexpr int kFusedThreads = (kCW + kGW + 3) * 32;   // + edge loader + table loader + storer
constexpr int kHdrInts = 64;   // slot header: 0 e_lo, 1 e_hi, 3 a_lo, 4 source counter, 5 hub bits, 16.. colptr[j0 .. j0+TJ]

struct FusedCfg {
  int NS, R, B, out, TJ, Sm, Sv, ntiles, mcap, thresh;
  int vstride;       // floats between the runs of consecutive bases inside a table slot (TJ*out, padded against bank conflicts)
  int off_meta, off_ts, mslot_bytes, vslot_bytes;
  int comp_bytes;
  int dbg;   // MRGCN_IDF_DBG (timing experiments only, results are wrong): (1 unused) 2 no gact gathers, 4 no bulk stores, 8 compute warps idle, 16 one bulk load per tile
};

// position in a ring of S slots: slot index and phase parity of its current use, advanced without divisions
struct RingPos {
  int s, S;
  uint32_t ph;
  __device__ __forceinline__ RingPos(int S_, int k0 = 0) : s(0), S(S_), ph(0) { advance(k0); }
  __device__ __forceinline__ void advance(int n = 1) {
    s += n;
    while (s >= S) { s -= S; ph ^= 1; }
  }
};

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void *dst, const void *src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(src)),
               "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void stg_hint(float *p, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ float2 ldg2_hint(const float2 *p, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}

// one edge: the warp's lanes hold v (rows of weight_I) and g (accumulators) of NB bases each
template <int OUT, int NB>
__device__ __forceinline__ void edge_body(const float2 (&t)[OUT / 2], const float (&c)[NB], const float2 (&v)[NB][OUT / 2],
                                          float2 (&g)[NB][OUT / 2], float (&d)[NB], bool do_g) {
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < OUT / 2; ++q) fma2v(a2, v[nb][q], t[q]);
    d[nb] = a2.x + a2.y;
    if (do_g) {
#pragma unroll
      for (int q = 0; q < OUT / 2; ++q) fma2(g[nb][q], c[nb], t[q]);
    }
  }
}

template <int OUT, int NB, int NG>
__global__ void __launch_bounds__(kFusedThreads, 1)
k_ident_bwd_fused(const float *__restrict__ V, const float *__restrict__ comp, const int32_t *__restrict__ colptr,
                  const int32_t *__restrict__ e2_dst, const int32_t *__restrict__ e2_rel, const float *__restrict__ e2_val,
                  const int32_t *__restrict__ e2_to_e3, const float *__restrict__ gact, float *__restrict__ gW,
                  float *__restrict__ cbuf, FusedCfg p) {
  constexpr int TP = (OUT + 3) & ~3;   // floats per staged t_e row
  constexpr int Q = OUT / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Sm = p.Sm, Sv = p.Sv;
  uint64_t *mfull = reinterpret_cast<uint64_t *>(smem_raw);   // edge slices landed (tx count)
  uint64_t *tready = mfull + Sm;                               // t_e rows written (by the tile's gather warp)
  uint64_t *mdone = tready + Sm;                               // edge slot drained (one arrival per compute warp)
  uint64_t *vfull = mdone + Sm;                                // table runs landed (tx count)
  uint64_t *vdone = vfull + Sv;                                // g_weight_I of the tile complete (one arrival per compute warp)
  uint64_t *vfree = vdone + Sv;                                // bulk store has read the slot
  float *comp_s = reinterpret_cast<float *>(smem_raw + 256);
  unsigned char *mslots = smem_raw + 256 + p.comp_bytes;
  unsigned char *vslots = mslots + (size_t)Sm * p.mslot_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = p.B, TJ = p.TJ;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Sm; ++s) { mbar_init(&mfull[s], 1); mbar_init(&tready[s], 1); mbar_init(&mdone[s], kCW); }
    for (int s = 0; s < Sv; ++s) { mbar_init(&vfull[s], 1); mbar_init(&vdone[s], kCW); mbar_init(&vfree[s], 1); }
    mbar_fence_init();
  }
  for (int x = threadIdx.x; x < p.R * B; x += kFusedThreads) comp_s[x] = __ldg(comp + x);
  __syncthreads();
  const int ntiles_mine = (int)blockIdx.x < p.ntiles ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto tile_j0 = [&](int k) {
    int j0 = ((int)blockIdx.x + k * (int)gridDim.x) * TJ;
    return j0 + TJ > p.NS ? p.NS - TJ : j0;   // the last tile overlaps its neighbour: the same values are written twice
  };
  const uint32_t run_bytes = (uint32_t)TJ * OUT * 4;

  if (warp == kCW + kGW) {
    // ============================== edge loader ==============================
    int cpv = 0;   // colptr[j0 + lane] of the next tile: loaded one tile ahead, off the critical path
    if (ntiles_mine > 0 && lane <= TJ) cpv = colptr[tile_j0(0) + lane];
    RingPos r(Sm);
    for (int k = 0; k < ntiles_mine; ++k, r.advance()) {
      const int s = r.s;
      unsigned char *st = mslots + (size_t)s * p.mslot_bytes;
      int *hdr = reinterpret_cast<int *>(st);
      const int cur = cpv;
      if (k + 1 < ntiles_mine && lane <= TJ) cpv = colptr[tile_j0(k + 1) + lane];
      if (k >= Sm) {
        mbar_wait(&mdone[s], r.ph ^ 1, 7, 64);
        fence_proxy_async();   // the slot's generic reads (acquired above) before the async-proxy writes below
      }
      const int e_lo = __shfl_sync(0xffffffffu, cur, 0), e_hi = __shfl_sync(0xffffffffu, cur, TJ);
      const int nxt = __shfl_down_sync(0xffffffffu, cur, 1);
      const unsigned hubs = __ballot_sync(0xffffffffu, lane < TJ && p.thresh > 0 && nxt - cur > p.thresh);
      if (lane <= TJ) hdr[16 + lane] = cur;
      const int a_lo = e_lo & ~3;
      int cnt = min(e_hi - a_lo, p.mcap);
      cnt = e_hi > e_lo ? ((cnt + 3) & ~3) : 0;
      int *mD = reinterpret_cast<int *>(st + p.off_meta);
      __syncwarp();
      if (lane == 0) {
        hdr[0] = e_lo; hdr[1] = e_hi; hdr[3] = a_lo; hdr[4] = 0; hdr[5] = (int)hubs;
        fence_proxy_async();
        mbar_expect_tx(&mfull[s], 4u * cnt * 4u);
      }
      __syncwarp();
      if (cnt > 0) {
        const int ms = p.mcap + 4;
        if (lane == 0) bulk_g2s(mD, e2_dst + a_lo, cnt * 4u, &mfull[s]);
        if (lane == 1) bulk_g2s(mD + ms, e2_rel + a_lo, cnt * 4u, &mfull[s]);
        if (lane == 2) bulk_g2s(mD + 2 * ms, e2_val + a_lo, cnt * 4u, &mfull[s]);
        if (lane == 3) bulk_g2s(mD + 3 * ms, e2_to_e3 + a_lo, cnt * 4u, &mfull[s]);
      }
    }
    return;
  }
  if (warp == kCW + kGW + 1) {
    // ============================== table loader ==============================
    const uint64_t pol = policy_evict_first();
    RingPos r(Sv);
    for (int k = 0; k < ntiles_mine; ++k, r.advance()) {
      const int s = r.s;
      float *vs = reinterpret_cast<float *>(vslots + (size_t)s * p.vslot_bytes);
      const int j0 = tile_j0(k);
      if (k >= Sv) {
        mbar_wait(&vfree[s], r.ph ^ 1, 8, 64);   // the slot's previous tile has left for HBM
        fence_proxy_async();
      }
      if (lane == 0) mbar_expect_tx(&vfull[s], run_bytes * B);
      __syncwarp();
      if (p.dbg & 16) {
        if (lane == 0) bulk_g2s_hint(vs, V + (size_t)j0 * OUT * B, run_bytes * B, &vfull[s], pol);
      } else
      for (int b = lane; b < B; b += 32)
        bulk_g2s_hint(vs + (size_t)b * p.vstride, V + ((size_t)b * p.NS + j0) * OUT, run_bytes, &vfull[s], pol);
    }
    return;
  }
  if (warp == kCW + kGW + 2) {
    // ============================== storer ==============================
    const uint64_t pol = policy_evict_first();
    RingPos r(Sv);
    for (int k = 0; k < ntiles_mine; ++k, r.advance()) {
      const int s = r.s;
      const float *vs = reinterpret_cast<const float *>(vslots + (size_t)s * p.vslot_bytes);
      const int j0 = tile_j0(k);
      mbar_wait(&vdone[s], r.ph, 9, 64);
      fence_proxy_async();   // the compute warps' generic writes of g (acquired above) before the async-proxy reads below
      if (!(p.dbg & 4))
        for (int b = lane; b < B; b += 32)
          bulk_s2g_hint(gW + ((size_t)b * p.NS + j0) * OUT, vs + (size_t)b * p.vstride, run_bytes, pol);
      bulk_commit();
      bulk_wait_read0();
      __syncwarp();
      if (lane == 0) mbar_arrive(&vfree[s]);
    }
    bulk_wait0();
    return;
  }

  if (warp >= kCW) {
    // ============================== gather warps: t_e of the staged edges ==============================
    const int gw = warp - kCW;
    const uint64_t keep = policy_evict_last();
    RingPos r(Sm);
    int mine = gw;   // tiles until this warp's next one
    for (int k = 0; k < ntiles_mine; ++k, r.advance()) {
      const int s = r.s;
      // Every gather warp waits for every tile, in order, although it gathers only its own: a parity wait cannot tell phase n
      // from phase n +- 2, so a waiter must never be more than one phase away from the barrier - a warp that skipped the
      // slot's previous use could find the barrier still in that phase (wait passes at once, on stale data).
      mbar_wait(&mfull[s], r.ph, 1, 32);
      if (mine > 0) { --mine; continue; }
      mine = kGW - 1;
      unsigned char *st = mslots + (size_t)s * p.mslot_bytes;
      const int *hdr = reinterpret_cast<const int *>(st);
      const int e_lo = hdr[0], e_hi = hdr[1], a_lo = hdr[3];
      const int ms = p.mcap + 4;
      const int *mD = reinterpret_cast<const int *>(st + p.off_meta);
      const float *mV = reinterpret_cast<const float *>(mD + 2 * ms);
      float *Ts = reinterpret_cast<float *>(st + p.off_ts);
      const int n_st = min(e_hi - a_lo, p.mcap);
      // four edges per lane and trip (128 per warp): their gathers are in flight together
      constexpr int U = 4;
      for (int i0 = (e_lo - a_lo) + lane; i0 < n_st; i0 += U * 32) {
        float r_[U][TP];
        float val[U];
        const float2 *gp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = i0 + 32 * u < n_st ? i0 + 32 * u : i0;
          val[u] = mV[idx];
          gp[u] = reinterpret_cast<const float2 *>(gact + (size_t)mD[idx] * OUT);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            float2 x = (p.dbg & 2) ? make_float2(1.f, 2.f) : ldg2_hint(gp[u] + q, keep);
            r_[u][2 * q] = x.x; r_[u][2 * q + 1] = x.y;
          }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (i0 + 32 * u < n_st) {
#pragma unroll
            for (int q = 0; q < OUT; ++q) r_[u][q] *= val[u];
#pragma unroll
            for (int q = OUT; q < TP; ++q) r_[u][q] = 0.f;
            float4 *dst = reinterpret_cast<float4 *>(Ts + (size_t)(i0 + 32 * u) * TP);
#pragma unroll
            for (int q = 0; q < TP / 4; ++q) dst[q] = make_float4(r_[u][4 * q], r_[u][4 * q + 1], r_[u][4 * q + 2], r_[u][4 * q + 3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tready[s]);
    }
    return;
  }

  // ============================== compute warps ==============================
  const int b0 = lane, b1 = lane + 32;
  // B > 32: the bases 32 .. B-1 keep only B - 32 lanes busy.  The staged loop therefore takes them for `ng` edges at a time:
  // lane group qg (gs lanes) works on edge i + qg, lane li of the group on basis 32 + li.  Every group then holds a partial
  // g_weight_I of those bases (edges i = qg mod ng); the groups are added at the end of the source in a fixed butterfly.
  constexpr int ng = NG, gs = 32 / NG;   // NG = 4 for B <= 40, 2 for B <= 48, else 1 (the host picks the instance)
  const int qg = lane / gs, li = lane & (gs - 1);
  const int bb = 32 + li;
  const uint64_t pol = policy_evict_first();
  RingPos rm(Sm), rv(Sv);
  for (int k = 0; k < ntiles_mine; ++k, rm.advance(), rv.advance()) {
    const int sm = rm.s, sv = rv.s;
    mbar_wait(&tready[sm], rm.ph, 3);   // implies mfull (the tile's gather warp waited for it)
    mbar_wait(&vfull[sv], rv.ph, 2);
    unsigned char *st = mslots + (size_t)sm * p.mslot_bytes;
    const int *hdr = reinterpret_cast<const int *>(st);
    const int a_lo = hdr[3];
    const unsigned hubs = (unsigned)hdr[5];
    const int *cp = hdr + 16;
    float *vs = reinterpret_cast<float *>(vslots + (size_t)sv * p.vslot_bytes);
    const int ms = p.mcap + 4;
    const int *mD = reinterpret_cast<const int *>(st + p.off_meta);
    const int *mR = mD + ms;
    const int *mP = mD + 3 * ms;
    const float *Ts = reinterpret_cast<const float *>(st + p.off_ts);

    // v/g of one source: v from the table slot, g back into the same place
    auto load_v = [&](int jl, float2 (&v)[NB][Q]) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int b = nb == 0 ? lane : bb;   // second set: every lane group holds the same rows 32 .. B-1
        const float2 *vp = reinterpret_cast<const float2 *>(vs + (size_t)b * p.vstride + jl * OUT);
#pragma unroll
        for (int q = 0; q < Q; ++q) v[nb][q] = b < B ? vp[q] : make_float2(0.f, 0.f);
      }
    };
    // edges [e0, e0+n) (n <= 32) from global memory: lane i gathers t of edge i, the warp then walks the edges
    auto chunk_global = [&](int e0, int n, const float2 (&v)[NB][Q], float2 (&g)[NB][Q], bool do_g) {
      float2 tl[Q];
      int rel_l = 0, pos_l = 0;
      if (lane < n) {
        const int e = e0 + lane;
        const float val = e2_val[e];
        rel_l = e2_rel[e]; pos_l = e2_to_e3[e];
        const float2 *gp = reinterpret_cast<const float2 *>(gact + (size_t)e2_dst[e] * OUT);
#pragma unroll
        for (int q = 0; q < Q; ++q) { float2 x = __ldg(gp + q); tl[q] = make_float2(val * x.x, val * x.y); }
      } else {
#pragma unroll
        for (int q = 0; q < Q; ++q) tl[q] = make_float2(0.f, 0.f);
      }
      for (int i = 0; i < n; ++i) {
        float2 t[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) { t[q].x = __shfl_sync(0xffffffffu, tl[q].x, i); t[q].y = __shfl_sync(0xffffffffu, tl[q].y, i); }
        const int rel = __shfl_sync(0xffffffffu, rel_l, i), pos = __shfl_sync(0xffffffffu, pos_l, i);
        float c[NB], d[NB];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) c[nb] = (lane + 32 * nb < B) ? comp_s[rel * B + lane + 32 * nb] : 0.f;
        edge_body<OUT, NB>(t, c, v, g, d, do_g);
        float *crow = cbuf + (size_t)pos * B;
        if (b0 < B) stg_hint(crow + b0, d[0], pol);
        if (NB > 1 && b1 < B) stg_hint(crow + b1, d[NB - 1], pol);
      }
    };

    // Sources are claimed from a counter in the slot header, not dealt out by warp index: source lengths are heavy-tailed,
    // and with a fixed deal the slowest of the kCW warps held every slot of the (short) ring - the others waited for
    // slots about half of the time.  A warp that finds the counter exhausted moves on to the next tile.
    int *ctr = const_cast<int *>(hdr) + 4;
    for (;;) {
      if (p.dbg & 8) break;
      int jl = 0;
      if (lane == 0) jl = atomicAdd(ctr, 1);
      jl = __shfl_sync(0xffffffffu, jl, 0);
      if (jl >= TJ) break;
      if ((hubs >> jl) & 1u) continue;   // hub: see below
      const int s_lo = cp[jl], s_hi = cp[jl + 1];
      float2 v[NB][Q], g[NB][Q];
      load_v(jl, v);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int q = 0; q < Q; ++q) g[nb][q] = make_float2(0.f, 0.f);
      if (s_hi - a_lo <= p.mcap) {
        // every edge of the source is staged
        // (the metadata arrays have 4 words of slack: the look-ahead of the last edge stays inside them)
        const int *rp = mR + (s_lo - a_lo), *pp = mP + (s_lo - a_lo);
        const float *tptr = Ts + (size_t)(s_lo - a_lo) * TP;
        const float *cl = comp_s + lane;
        float *cb = cbuf + lane;
        auto load_t = [&](const float *tq, float2 (&t)[Q]) {
          const float4 *tp = reinterpret_cast<const float4 *>(tq);
#pragma unroll
          for (int q = 0; q < Q / 2; ++q) { float4 x = tp[q]; t[2 * q] = make_float2(x.x, x.y); t[2 * q + 1] = make_float2(x.z, x.w); }
          if constexpr (OUT % 4 != 0) t[Q - 1] = *reinterpret_cast<const float2 *>(tq + OUT - 2);
        };
        if constexpr (NB == 1) {
          int rel = rp[0], pos = pp[0];
          for (int n = s_hi - s_lo; n > 0; --n) {
            ++rp; ++pp;
            const int rel_n = rp[0], pos_n = pp[0];
            float2 t[Q];
            load_t(tptr, t);
            tptr += TP;
            float c[NB], d[NB];
            c[0] = (lane < B) ? cl[rel * B] : 0.f;
            edge_body<OUT, NB>(t, c, v, g, d, true);
            if (b0 < B) stg_hint(cb + (size_t)pos * B, d[0], pol);
            rel = rel_n; pos = pos_n;
          }
        } else {
          const int n = s_hi - s_lo;
          for (int i = 0; i < n; i += ng) {
            const int m = min(ng, n - i);
            // bases 0 .. 31: one edge at a time, t_e broadcast
#pragma unroll
            for (int u = 0; u < NG; ++u) {
              if (u >= m) break;
              const int rel = rp[u], pos = pp[u];
              float2 t[Q];
              load_t(tptr + u * TP, t);
              const float c0 = cl[rel * B];
              float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
              for (int q = 0; q < Q; ++q) fma2v(a2, v[0][q], t[q]);
#pragma unroll
              for (int q = 0; q < Q; ++q) fma2(g[0][q], c0, t[q]);
              stg_hint(cb + (size_t)pos * B, a2.x + a2.y, pol);
            }
            // bases 32 .. B-1: lane group qg takes edge i + qg
            {
              const bool on = qg < m && bb < B;
              const int u = qg < m ? qg : 0;
              const int rel = rp[u], pos = pp[u];
              float2 t[Q];
              load_t(tptr + u * TP, t);
              const float c1 = on ? comp_s[rel * B + bb] : 0.f;
              float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
              for (int q = 0; q < Q; ++q) fma2v(a2, v[NB - 1][q], t[q]);
#pragma unroll
              for (int q = 0; q < Q; ++q) fma2(g[NB - 1][q], c1, t[q]);
              if (on) stg_hint(cbuf + (size_t)pos * B + bb, a2.x + a2.y, pol);
            }
            rp += ng; pp += ng; tptr += ng * TP;
          }
        }
      } else {
        for (int e0 = s_lo; e0 < s_hi; e0 += 32) chunk_global(e0, min(32, s_hi - e0), v, g, true);
      }
      if constexpr (NB > 1) {
        for (int off = gs; off < 32; off <<= 1) {
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            g[NB - 1][q].x += __shfl_xor_sync(0xffffffffu, g[NB - 1][q].x, off);
            g[NB - 1][q].y += __shfl_xor_sync(0xffffffffu, g[NB - 1][q].y, off);
          }
        }
      }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int b = nb == 0 ? lane : bb;
        if (b < B && (nb == 0 || qg == 0)) {
          float2 *vp = reinterpret_cast<float2 *>(vs + (size_t)b * p.vstride + jl * OUT);
#pragma unroll
          for (int q = 0; q < Q; ++q) vp[q] = g[nb][q];
        }
      }
    }
    // hubs of the tile: scratch rows only, 32-edge chunks dealt round-robin to the compute warps (their rows of the slot
    // still hold weight_I: nobody writes g for them)
    for (unsigned hm = (p.dbg & 8) ? 0u : hubs; hm; hm &= hm - 1) {
      const int jl = __ffs(hm) - 1;
      const int s_lo = cp[jl], s_hi = cp[jl + 1];
      float2 v[NB][Q], g[NB][Q];
      load_v(jl, v);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int q = 0; q < Q; ++q) g[nb][q] = make_float2(0.f, 0.f);
      for (int e0 = s_lo + warp * 32; e0 < s_hi; e0 += kCW * 32) chunk_global(e0, min(32, s_hi - e0), v, g, false);
    }
    // No proxy fence here: fence.proxy.async is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in SASS and would wait for the warp's
    // scratch-row stores to be acknowledged by L2 (about a microsecond per tile).  The fence is executed by the storer / the
    // edge loader instead, after they have acquired these arrivals and before they issue their bulk copies.
    __syncwarp();
    if (lane == 0) { mbar_arrive(&vdone[sv]); mbar_arrive(&mdone[sm]); }
  }
}

static int g_fused = -1;   // -1: not read yet
static bool fused_enabled() {
  if (g_fused < 0) { const char *e = getenv("MRGCN_IDENT_FUSED"); g_fused = (e && e[0] == '0') ? 0 : 1; }
  return g_fused == 1;
}
static int env_int(const char *name, int dflt) { const char *e = getenv(name); return e && e[0] ? atoi(e) : dflt; }

static bool fused_config(FusedCfg &p, int64_t NS, int R, int B, int out, int thresh) {
  if (B <= 0 || B > 64) return false;
  if (!(out == 4 || out == 8 || out == 10 || out == 12 || out == 16)) return false;
  if ((NS * out) % 4 != 0 || NS > INT32_MAX) return false;
  p.comp_bytes = (int)((((size_t)R * B * 4) + 15) & ~(size_t)15);
  if (p.comp_bytes > 64 * 1024) return false;
  int tj = env_int("MRGCN_IDF_TJ", (26 * 1024) / (B * out * 4));
  if (tj > 28) tj = 28;
  while (tj > 0 && (tj * out) % 4 != 0) --tj;
  if (tj >= kCW) {   // whole rounds of the compute warps (source jl belongs to warp jl mod kCW)
    int m = (tj / kCW) * kCW;
    while (m > 0 && (m * out) % 4 != 0) m -= kCW;
    if (m > 0) tj = m;
  }
  if (tj <= 0 || tj > NS) return false;
  p.NS = (int)NS; p.R = R; p.B = B; p.out = out; p.TJ = tj; p.thresh = thresh;
  p.ntiles = (int)cdiv(NS, tj);
  int mcap = env_int("MRGCN_IDF_MCAP", tj * 16);
  mcap = mcap < 64 ? 64 : mcap > 1024 ? 1024 : mcap;
  p.mcap = (mcap + 3) & ~3;
  const int run = tj * out;
  p.vstride = ((run / 4) & 1) ? run : run + 4;   // 4 x odd floats: lane-strided 8-byte reads meet 2 banks at most
  const int tp = (out + 3) & ~3;
  p.off_meta = kHdrInts * 4;
  p.off_ts = p.off_meta + 4 * (p.mcap + 4) * 4;
  p.mslot_bytes = p.off_ts + p.mcap * tp * 4;
  p.vslot_bytes = B * p.vstride * 4;
  const int budget = 227 * 1024 - 256 - p.comp_bytes;
  // table ring first (it carries the HBM traffic), the edge ring gets the rest: it should be the deeper one
  int sv = env_int("MRGCN_IDF_SV", 4), sm = env_int("MRGCN_IDF_SM", 0);
  if (sv > 5) sv = 5;
  while (sv >= 2 && budget - sv * p.vslot_bytes < 2 * p.mslot_bytes) --sv;
  if (sv < 2) return false;
  const int sm_max = (budget - sv * p.vslot_bytes) / p.mslot_bytes;
  if (sm <= 0 || sm > sm_max) sm = sm_max;
  if (sm > 5) sm = 5;   // 3*Sm + 3*Sv barriers in the first 256 bytes
  if (sm < 2) return false;
  p.dbg = env_int("MRGCN_IDF_DBG", 0);
  p.Sv = sv; p.Sm = sm;   // (fewer edge slots than gather warps only serialises the gather warps)
  return true;
}

}  // namespace

// g_weight_I (non-hub sources) and the scratch rows cbuf[e3, :] of every edge; returns 1 when the shape is not handled
int launch_ident_bwd_fused(const mrgcn_graph *g, const float *V, const float *comp, int B, int out, const float *gact,
                           float *gW, float *cbuf, cudaStream_t st) {
  if (!fused_enabled()) return 1;
  FusedCfg p;
  const int thresh = g->n_long_cols > 0 ? g->long_col_thresh : 0;
  if (!fused_config(p, g->NS, g->R, B, out, thresh)) return 1;
  if (((uintptr_t)V | (uintptr_t)gW) & 15) return 1;
  const size_t smem = 256 + (size_t)p.comp_bytes + (size_t)p.Sm * p.mslot_bytes + (size_t)p.Sv * p.vslot_bytes;
  // MRGCN_IDF_GRID: fewer CTAs than SMs, so that small test graphs also walk many tiles per CTA (ring wrap-around)
  const int gmax = env_int("MRGCN_IDF_GRID", kNumSMs);
  const unsigned grid = (unsigned)(p.ntiles < gmax ? p.ntiles : gmax);
  const int NB = B > 32 ? 2 : 1;
  MRGCN_PROF("ident_bwd_fused");
#define LAUNCH(OUTV, NBV, NGV)                                                                                              \
  do {                                                                                                                      \
    MRGCN_CUDA(cudaFuncSetAttribute(k_ident_bwd_fused<OUTV, NBV, NGV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_ident_bwd_fused<OUTV, NBV, NGV><<<grid, kFusedThreads, smem, st>>>(V, comp, g->colptr, g->e2_dst, g->e2_rel, g->e2_val, \
                                                                         g->e2_to_e3, gact, gW, cbuf, p);                   \
  } while (0)
#define LAUNCH_NB(OUTV)                    \
  do {                                     \
    if (NB == 1) LAUNCH(OUTV, 1, 1);       \
    else if (B <= 40) LAUNCH(OUTV, 2, 4);  \
    else if (B <= 48) LAUNCH(OUTV, 2, 2);  \
    else LAUNCH(OUTV, 2, 1);               \
  } while (0)
  switch (out) {
    case 4: LAUNCH_NB(4); break;
    case 8: LAUNCH_NB(8); break;
    case 10: LAUNCH_NB(10); break;
    case 12: LAUNCH_NB(12); break;
    default: LAUNCH_NB(16); break;
  }
#undef LAUNCH_NB
#undef LAUNCH
  MRGCN_LAUNCH_CHECK();
  return 0;
}

}  // namespace mrgcn

extern "C" void mrgcn_set_ident_fused(int32_t on) { mrgcn::g_fused = on ? 1 : 0; }
