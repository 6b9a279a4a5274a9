// Producer/consumer pipeline shared by the identity-term kernels that stream the basis table V_I
// (ident_msg_fwd_bulk, ident_bwd_c_bulk).
//
// One persistent CTA per SM walks tiles of TJ consecutive sources.  A dedicated producer warp feeds a ring of S
// shared-memory stages with the TMA engine (cp.async.bulk, SASS UBLKCP): per tile B runs of TJ*out floats of
// V_I[b, j0:j0+TJ, :] plus the tile's slice of the three E2 edge arrays, so the consumer warps read everything
// they need from shared memory.  full[s] (tx-count) / empty[s] (one arrival per consumer warp) mbarriers; consumer
// warps never meet at a CTA barrier, so a slow warp does not stall the others.
#pragma once
#include "common.cuh"
#include "pipeline.cuh"

namespace mrgcn {

constexpr int kPipeConsumerWarps = 8;
constexpr int kPipeThreads = (kPipeConsumerWarps + 1) * 32;

struct IdentPipe {
  int NS, B, out, TJ, S, ntiles;
  int mcap;          // edges of metadata staged per tile (multiple of 4); the rest is read from global memory
  int v_floats;      // B*TJ*out rounded up to 4, + 16 floats of slack for chunked row reads
  int stage_bytes;   // 16 (header) + v_floats*4 + 3*(mcap+4)*4, multiple of 16
};

__host__ __device__ inline int ident_pipe_stage_bytes(int v_floats, int mcap) { return 16 + v_floats * 4 + 3 * (mcap + 4) * 4; }

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// fn(Vs, j0, e, a, b, v): one edge; a/b/v are the three metadata words of edge e (E2 order)
template <class EdgeFn>
__device__ __forceinline__ void ident_pipeline(const IdentPipe &p, const float *__restrict__ V,
                                               const int32_t *__restrict__ colptr, const int32_t *__restrict__ gA,
                                               const int32_t *__restrict__ gB, const float *__restrict__ gC,
                                               unsigned char *stage_base, uint64_t *full, uint64_t *empty, EdgeFn fn) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  if (warp == kPipeConsumerWarps) {
    // ===== producer warp: lane 0 owns the barriers, all 32 lanes issue bulk copies (one thread alone cannot issue
    //       B + 3 small copies per tile fast enough to keep HBM busy) =====
    const uint32_t run_bytes = (uint32_t)p.TJ * p.out * 4;
    for (int k = 0;; ++k) {
      const int t = blockIdx.x + k * gridDim.x;
      if (t >= p.ntiles) break;
      const int s = k % S;
      int j0 = t * p.TJ;
      if (j0 + p.TJ > p.NS) j0 = p.NS - p.TJ;
      const int e_lo = colptr[j0], e_hi = colptr[j0 + p.TJ];
      const int a_lo = e_lo & ~3;
      int cnt = min(e_hi - a_lo, p.mcap);
      cnt = e_hi > e_lo ? ((cnt + 3) & ~3) : 0;
      unsigned char *st = stage_base + (size_t)s * p.stage_bytes;
      float *vs = reinterpret_cast<float *>(st + 16);
      int *mA = reinterpret_cast<int *>(st + 16 + (size_t)p.v_floats * 4);
      int *mB = mA + p.mcap + 4;
      float *mC = reinterpret_cast<float *>(mB + p.mcap + 4);
      if (lane == 0) {
        mbar_wait(&empty[s], ((k / S) & 1) ^ 1, 7);   // stage drained by every consumer warp
        int *hdr = reinterpret_cast<int *>(st);
        hdr[0] = e_lo; hdr[1] = e_hi; hdr[2] = j0; hdr[3] = a_lo;
        fence_proxy_async();
        mbar_expect_tx(&full[s], run_bytes * p.B + 3u * cnt * 4u);
      }
      __syncwarp();
      for (int b = lane; b < p.B; b += 32)
        bulk_g2s(vs + (size_t)b * p.TJ * p.out, V + ((size_t)b * p.NS + j0) * p.out, run_bytes, &full[s]);
      if (cnt > 0) {
        if (lane == 29) bulk_g2s(mA, gA + a_lo, cnt * 4u, &full[s]);
        if (lane == 30) bulk_g2s(mB, gB + a_lo, cnt * 4u, &full[s]);
        if (lane == 31) bulk_g2s(mC, gC + a_lo, cnt * 4u, &full[s]);
      }
    }
    return;
  }
  // ===== consumers =====
  for (int k = 0;; ++k) {
    const int t = blockIdx.x + k * gridDim.x;
    if (t >= p.ntiles) break;
    const int s = k % S;
    mbar_wait(&full[s], (k / S) & 1);
    const unsigned char *st = stage_base + (size_t)s * p.stage_bytes;
    const int *hdr = reinterpret_cast<const int *>(st);
    const int e_lo = hdr[0], e_hi = hdr[1], j0 = hdr[2], a_lo = hdr[3];
    const float *vs = reinterpret_cast<const float *>(st + 16);
    const int *mA = reinterpret_cast<const int *>(st + 16 + (size_t)p.v_floats * 4);
    const int *mB = mA + p.mcap + 4;
    const float *mC = reinterpret_cast<const float *>(mB + p.mcap + 4);
    for (int e = e_lo + warp * 32 + lane; e < e_hi; e += kPipeConsumerWarps * 32) {
      const int idx = e - a_lo;
      int a, b;
      float v;
      if (idx < p.mcap) { a = mA[idx]; b = mB[idx]; v = mC[idx]; }
      else { a = gA[e]; b = gB[e]; v = gC[e]; }
      fn(vs, j0, e, a, b, v);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
}

}  // namespace mrgcn
