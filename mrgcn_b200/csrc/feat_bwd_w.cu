// Feature-term weight gradient for WIDE outputs (autograd through torch.einsum('ij,bjk->bik') + torch.mm(A, .) of
// /root/reference/mrgcn/layers/graph.py:83-97; the YAGO3-10+ encoder 145 -> 200 is the shape):
//
//   part[c, k, o] = sum over the E3 edges e of chunk c (one relation) of X[src_e, k] * val_e * gact[dst_e, o]
//
// Per chunk this is a small GEMM  X_c^T [in x n] . T_c [n x out]  over gathered rows.  The round-1 kernel
// (k_feat_bwd_w: a thread per k, 16 outputs in registers, t_e restaged once per 16-column pass) ran at 24 % of the
// FP32 pipe on this shape; here it is a register-tiled SGEMM: a thread owns 8 rows x 16 columns of the chunk's
// product (128 accumulators, FFMA2), the gathered rows are staged by four producer warps through a ring of
// shared-memory stages (16-byte loads, all gathers of a 32-edge batch in flight), eight consumer warps never wait
// for global memory.  mbarriers only; every sum is taken in edge order by one thread: bit-reproducible.
#include "common.cuh"
#include "pipeline.cuh"
#include "ident_pipe.cuh"
#include "rgcn_internal.cuh"

namespace mrgcn {
namespace {

constexpr int kKT = 8, kOT = 16;   // rows (k) x columns (o) of the product per thread
constexpr int kEB = 32;            // edges per stage
constexpr int kCWt = 8, kPWt = 4;  // consumer / producer warps
constexpr int kTileThreads = (kCWt + kPWt) * 32;

struct TileCfg {
  int in, out, ldx, NTK, NTO, KP, OP, S, stage_floats;
};

__global__ void __launch_bounds__(kTileThreads, 1)
k_feat_bwd_w_tile(const float *__restrict__ X, const float *__restrict__ gact, const int32_t *__restrict__ chunk_ptr,
                  const int32_t *__restrict__ e3_src, const int32_t *__restrict__ e3_dst, const float *__restrict__ e3_val,
                  float *__restrict__ part, TileCfg p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *empty = full + p.S;
  float *stages = reinterpret_cast<float *>(smem_raw + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) { mbar_init(&full[s], kPWt); mbar_init(&empty[s], kCWt); }
    mbar_fence_init();
  }
  __syncthreads();
  const int c = blockIdx.x;
  const int e_lo = chunk_ptr[c], e_hi = chunk_ptr[c + 1];
  const int nbatch = (e_hi - e_lo + kEB - 1) / kEB;
  const int KP = p.KP, OP = p.OP;

  if (warp >= kCWt) {
    // ===== producers: warp pw stages rows [8 pw, 8 pw + 8) of every batch =====
    // Lane r first fetches the metadata of the batch's edge r (one coalesced load per array); the rows are then gathered four
    // at a time, all 16-byte loads of the four rows issued before the first store: three global round trips per batch
    // and warp (the first version took one per row and piece and starved the consumers: 26 % issue slots used).
    const int pw = warp - kCWt;
    const int kv = KP / 4, ov = OP / 4;   // 16-byte pieces per staged row (<= 64 each)
    constexpr int RPW = kEB / kPWt, G = 4;
    int s = 0;
    uint32_t ph = 0;
    for (int b = 0; b < nbatch; ++b) {
      const int eb = e_lo + b * kEB;
      const bool lv = eb + lane < e_hi;
      const int src_l = lv ? e3_src[eb + lane] : 0, dst_l = lv ? e3_dst[eb + lane] : 0;
      const float val_l = lv ? e3_val[eb + lane] : 0.f;   // padded edges: t = 0, row 0 (valid memory)
      if (b >= p.S) mbar_wait(&empty[s], ph ^ 1, 11, 32);
      float *Xs = stages + (size_t)s * p.stage_floats;
      float *Ts = Xs + kEB * KP;
      for (int r0 = pw * RPW; r0 < (pw + 1) * RPW; r0 += G) {
        float4 xa[G][2], ta[G][2];
        float val[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const int src = __shfl_sync(0xffffffffu, src_l, r0 + u), dst = __shfl_sync(0xffffffffu, dst_l, r0 + u);
          val[u] = __shfl_sync(0xffffffffu, val_l, r0 + u);
          const float4 *xrow = reinterpret_cast<const float4 *>(X + (size_t)src * p.ldx);
          const float4 *grow = reinterpret_cast<const float4 *>(gact + (size_t)dst * p.out);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q = lane + 32 * h;
            xa[u][h] = (q < kv && 4 * q < p.in) ? __ldg(xrow + q) : make_float4(0.f, 0.f, 0.f, 0.f);   // 4q + 3 < ldx
            ta[u][h] = (q < ov && 4 * q < p.out) ? __ldg(grow + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) {
          float4 *xd = reinterpret_cast<float4 *>(Xs + (r0 + u) * KP);
          float4 *td = reinterpret_cast<float4 *>(Ts + (r0 + u) * OP);
          const bool live = eb + r0 + u < e_hi;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q = lane + 32 * h;
            float4 x = xa[u][h], t = ta[u][h];
            const int k0 = 4 * q;
            if (!live) x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + 1 >= p.in) x.y = 0.f;
            if (k0 + 2 >= p.in) x.z = 0.f;
            if (k0 + 3 >= p.in) x.w = 0.f;
            t.x *= val[u]; t.y *= val[u]; t.z *= val[u]; t.w *= val[u];
            if (q < kv) xd[q] = x;
            if (q < ov) td[q] = t;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      if (++s == p.S) { s = 0; ph ^= 1; }
    }
    return;
  }

  // ===== consumers: thread (tk, to) owns rows k = 8 tk + i and columns o = 4 to + 4 NTO q + j (q, j < 4) =====
  const int t = threadIdx.x;
  const bool act = t < p.NTK * p.NTO;
  const int to = act ? t % p.NTO : 0, tk = act ? t / p.NTO : 0;
  float2 acc[kKT][kOT / 2];
#pragma unroll
  for (int i = 0; i < kKT; ++i)
#pragma unroll
    for (int j = 0; j < kOT / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);
  int s = 0;
  uint32_t ph = 0;
  const int oseg = 4 * p.NTO;
  for (int b = 0; b < nbatch; ++b) {
    mbar_wait(&full[s], ph, 12);
    const float *Xs = stages + (size_t)s * p.stage_floats + 8 * tk;
    const float *Ts = stages + (size_t)s * p.stage_floats + kEB * KP + 4 * to;
    if (act) {
#pragma unroll 2
      for (int el = 0; el < kEB; ++el) {
        const float4 x0 = *reinterpret_cast<const float4 *>(Xs + el * KP);
        const float4 x1 = *reinterpret_cast<const float4 *>(Xs + el * KP + 4);
        const float x[kKT] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        float2 tv[kOT / 2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 tq = *reinterpret_cast<const float4 *>(Ts + el * OP + q * oseg);
          tv[2 * q] = make_float2(tq.x, tq.y);
          tv[2 * q + 1] = make_float2(tq.z, tq.w);
        }
#pragma unroll
        for (int i = 0; i < kKT; ++i)
#pragma unroll
          for (int j = 0; j < kOT / 2; ++j) fma2(acc[i][j], x[i], tv[j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == p.S) { s = 0; ph ^= 1; }
  }
  if (act) {
#pragma unroll
    for (int i = 0; i < kKT; ++i) {
      const int k = 8 * tk + i;
      if (k >= p.in) continue;
      float *row = part + ((size_t)c * p.in + k) * p.out;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int o = 4 * to + q * oseg;
        if (o < p.out)
          *reinterpret_cast<float4 *>(row + o) = make_float4(acc[i][2 * q].x, acc[i][2 * q].y, acc[i][2 * q + 1].x, acc[i][2 * q + 1].y);
      }
    }
  }
}

}  // namespace

// returns 1 (nothing launched) when the shape is not handled
int launch_feat_bwd_w_tile(const mrgcn_graph *g, const float *X, int ldx, const float *gact, float *part, int in, int out,
                           cudaStream_t st) {
  if (out <= 16 || (out & 3) || (ldx & 3) || ldx < in || in < 8) return 1;
  if ((((uintptr_t)X | (uintptr_t)gact | (uintptr_t)part) & 15) != 0) return 1;
  TileCfg p;
  p.in = in; p.out = out; p.ldx = ldx;
  p.NTK = (in + kKT - 1) / kKT;
  p.NTO = (out + kOT - 1) / kOT;
  if (p.NTK * p.NTO > kCWt * 32) return 1;
  p.KP = p.NTK * kKT;
  p.OP = p.NTO * kOT;
  if (p.KP > 256 || p.OP > 256) return 1;   // two 16-byte pieces per lane and staged row
  p.stage_floats = kEB * (p.KP + p.OP);
  const size_t stage_bytes = (size_t)p.stage_floats * 4;
  p.S = (int)((200 * 1024 - 128) / stage_bytes);
  if (p.S > 4) p.S = 4;
  if (p.S < 2) return 1;
  const size_t smem = 128 + (size_t)p.S * stage_bytes;
  MRGCN_CUDA(cudaFuncSetAttribute(k_feat_bwd_w_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MRGCN_PROF("feat_bwd_w");
  k_feat_bwd_w_tile<<<(unsigned)g->n_chunks, kTileThreads, smem, st>>>(X, gact, g->chunk_ptr, g->e3_src, g->e3_dst, g->e3_val, part,
                                                                       p);
  MRGCN_LAUNCH_CHECK();
  return 0;
}

}  // namespace mrgcn
