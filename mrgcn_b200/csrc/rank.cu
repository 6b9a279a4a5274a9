// Filtered ranking in GEMM form (replaces compute_ranks_fast + filter_scores_,
// /root/reference/mrgcn/tasks/link_prediction.py:557-643).
//
// For every fact f and side, the reference scores all N candidates, overwrites the other true triples with -inf, and
// counts how many scores beat / tie the target's.  Here no F x N score matrix exists:
//   rank_target   t_f = score of the fact's own (s, p, o), one thread per fact
//   rank_tile     a 32-fact x 128-candidate tile of scores per CTA, K staged through shared memory in chunks of 32
//                 (candidate rows are read once per 32 facts instead of once per fact), each thread 4 x 4 scores in
//                 registers; the epilogue compares with t_f and adds #(score > t_f), #(score == t_f) to two integer
//                 counters per fact (integer atomics: order-independent, deterministic)
//   rank_filter   the filter list (CSR over facts, built on the device by sort + segment in
//                 mrgcn_b200/tasks/link_prediction.py) is applied as a correction: every listed candidate other than the
//                 target is scored again - same arithmetic, same bits - and taken out of the counters
//   rank_final    rank = #greater + round_half_even((#ties - 1) / 2) + 1
// Every score is the same sequential chain  acc = fma(x_k * y_k, z_k, acc), k = 0 .. h-1  with the operand order of
// score_distmult_bc's generic path ((s * p) * o, link_prediction.py:665), so the target's score inside the tile is
// bit-identical to t_f (it ties with itself, as in the reference) and a filtered candidate is removed exactly.
#include "common.cuh"

namespace mrgcn {
namespace {

constexpr int TF = 32, TC = 128, TK = 32;      // facts x candidates per CTA, K chunk
constexpr int kRankThreads = 256;

// operands of one (fact, candidate): head side: (e_c * rel) * e_fixed ; tail side: (e_fixed * rel) * e_c
__device__ __forceinline__ float chain(const float *x, const float *y, const float *z, int h) {
  float acc = 0.f;
  for (int k = 0; k < h; ++k) acc = fmaf(x[k] * y[k], z[k], acc);
  return acc;
}

__global__ void k_rank_target(const int64_t *__restrict__ facts, int64_t F, int head, const float *__restrict__ E,
                              const float *__restrict__ Rel, int h, float *__restrict__ tscore, int32_t *__restrict__ cnt) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float *es = E + (size_t)facts[3 * f] * h, *rp = Rel + (size_t)facts[3 * f + 1] * h, *eo = E + (size_t)facts[3 * f + 2] * h;
  tscore[f] = chain(es, rp, eo, h);       // (s * p) * o on either side
  cnt[2 * f] = 0;
  cnt[2 * f + 1] = 0;
  (void)head;
}

__global__ void __launch_bounds__(kRankThreads)
k_rank_tile(const int64_t *__restrict__ facts, int64_t F, int head, const float *__restrict__ E, const float *__restrict__ Rel,
            int64_t N, int h, const float *__restrict__ tscore, int32_t *__restrict__ cnt) {
  __shared__ float Fx[TF][TK + 1];      // per fact: the fixed entity row
  __shared__ float Fr[TF][TK + 1];      // per fact: the relation row
  __shared__ float Cs[TC][TK + 1];      // candidate rows
  const int64_t f0 = (int64_t)blockIdx.y * TF, c0 = (int64_t)blockIdx.x * TC;
  const int tid = threadIdx.x;
  const int tf = tid / 32, tc = tid % 32;        // this thread: facts tf*4 .. +4, candidates tc, tc+32, tc+64, tc+96
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < h; k0 += TK) {
    __syncthreads();
    for (int x = tid; x < TF * TK; x += kRankThreads) {
      const int i = x / TK, k = x - i * TK;
      const int64_t f = f0 + i;
      float fx = 0.f, fr = 0.f;
      if (f < F && k0 + k < h) {
        const int64_t fixed = head ? facts[3 * f + 2] : facts[3 * f];
        fx = E[(size_t)fixed * h + k0 + k];
        fr = Rel[(size_t)facts[3 * f + 1] * h + k0 + k];
      }
      Fx[i][k] = fx;
      Fr[i][k] = fr;
    }
    for (int x = tid; x < TC * TK; x += kRankThreads) {
      const int i = x / TK, k = x - i * TK;
      const int64_t c = c0 + i;
      Cs[i][k] = (c < N && k0 + k < h) ? E[(size_t)c * h + k0 + k] : 0.f;
    }
    __syncthreads();
    const int kn = min(TK, h - k0);
    for (int k = 0; k < kn; ++k) {
      float cv[4], xv[4], rv[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) cv[b] = Cs[tc + 32 * b][k];
#pragma unroll
      for (int a = 0; a < 4; ++a) { xv[a] = Fx[tf * 4 + a][k]; rv[a] = Fr[tf * 4 + a][k]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          acc[a][b] = head ? fmaf(cv[b] * rv[a], xv[a], acc[a][b]) : fmaf(xv[a] * rv[a], cv[b], acc[a][b]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t f = f0 + tf * 4 + a;
    if (f >= F) continue;
    const float t = tscore[f];
    int gt = 0, eq = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t c = c0 + tc + 32 * b;
      if (c < N) { gt += acc[a][b] > t; eq += acc[a][b] == t; }
    }
    // the 32 lanes of a warp hold the same fact: one pair of atomics per warp and fact
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { gt += __shfl_xor_sync(0xffffffffu, gt, s); eq += __shfl_xor_sync(0xffffffffu, eq, s); }
    if (tc == 0) {
      if (gt) atomicAdd(cnt + 2 * f, gt);
      if (eq) atomicAdd(cnt + 2 * f + 1, eq);
    }
  }
}

__global__ void k_rank_filter(const int64_t *__restrict__ facts, int64_t F, int head, const float *__restrict__ E,
                              const float *__restrict__ Rel, int h, const int32_t *__restrict__ fptr,
                              const int32_t *__restrict__ fidx, const float *__restrict__ tscore, int32_t *__restrict__ cnt) {
  const int64_t f = blockIdx.x;
  const int64_t target = head ? facts[3 * f] : facts[3 * f + 2];
  const int64_t fixed = head ? facts[3 * f + 2] : facts[3 * f];
  const float *fx = E + (size_t)fixed * h, *fr = Rel + (size_t)facts[3 * f + 1] * h;
  const float t = tscore[f];
  int gt = 0, eq = 0;
  for (int x = fptr[f] + threadIdx.x; x < fptr[f + 1]; x += blockDim.x) {
    const int64_t c = fidx[x];
    if (c == target) continue;                     // link_prediction.py:566,572: the target itself is not filtered
    const float *ec = E + (size_t)c * h;
    const float sc = head ? chain(ec, fr, fx, h) : chain(fx, fr, ec, h);
    gt += sc > t;
    eq += sc == t;
  }
  if (gt) atomicSub(cnt + 2 * f, gt);
  if (eq) atomicSub(cnt + 2 * f + 1, eq);
}

__global__ void k_rank_final(const int32_t *__restrict__ cnt, int64_t F, int64_t *__restrict__ rank) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  // rank = #greater + round_half_even((#ties - 1) / 2) + 1      (link_prediction.py:632-643)
  const int64_t m = (int64_t)cnt[2 * f + 1] - 1;
  int64_t half = m / 2;
  if (m > 0 && (m & 1) && (half & 1)) half += 1;
  if (m < 0) half = 0;  // target score NaN: it does not tie with itself; torch.round(-0.5) = -0 -> 0
  rank[f] = cnt[2 * f] + half + 1;
}

}  // namespace
}  // namespace mrgcn

using namespace mrgcn;

extern "C" int64_t mrgcn_distmult_rank_ws_elems(int64_t F) { return 3 * (F > 0 ? F : 1); }

extern "C" int mrgcn_distmult_rank(const int64_t *facts, int64_t F, int32_t head, const float *E, const float *Rel,
                                   int64_t N, int32_t h, const int32_t *filt_ptr, const int32_t *filt_idx,
                                   int32_t *ws, int64_t *rank, mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(F >= 0 && N > 0 && h > 0 && ws && rank, MRGCN_E_BADARG, "distmult_rank: bad arguments");
  if (F == 0) return 0;
  MRGCN_REQUIRE(cdiv(F, TF) < 65536, MRGCN_E_BADARG, "distmult_rank: at most %d facts per call", 65535 * TF);
  int32_t *cnt = ws;
  float *tscore = reinterpret_cast<float *>(ws + 2 * F);
  MRGCN_PROF("rank_target");
  k_rank_target<<<(unsigned)cdiv(F, 128), 128, 0, st>>>(facts, F, head, E, Rel, h, tscore, cnt);
  MRGCN_LAUNCH_CHECK();
  MRGCN_PROF("rank_tile");
  k_rank_tile<<<dim3((unsigned)cdiv(N, TC), (unsigned)cdiv(F, TF)), kRankThreads, 0, st>>>(facts, F, head, E, Rel, N, h, tscore, cnt);
  MRGCN_LAUNCH_CHECK();
  if (filt_ptr && filt_idx) {
    MRGCN_PROF("rank_filter");
    k_rank_filter<<<(unsigned)F, 128, 0, st>>>(facts, F, head, E, Rel, h, filt_ptr, filt_idx, tscore, cnt);
    MRGCN_LAUNCH_CHECK();
  }
  MRGCN_PROF("rank_final");
  k_rank_final<<<(unsigned)cdiv(F, 128), 128, 0, st>>>(cnt, F, rank);
  MRGCN_LAUNCH_CHECK();
  return 0;
}
