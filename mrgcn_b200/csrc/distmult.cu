// DistMult scorer and its backward
// (replaces /root/reference/mrgcn/tasks/link_prediction.py:645-665 score_distmult_bc; the ranking pass is rank.cu).
// Fused gather-multiply-reduce: nothing of shape (n, h) or (b, N, h) is materialised.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace mrgcn {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// score[t] = sum_k (E[s,k] * Rel[p,k]) * E[o,k]      one warp per triple
__global__ void __launch_bounds__(kThreads)
k_distmult_fwd(const int64_t *__restrict__ s, const int64_t *__restrict__ p, const int64_t *__restrict__ o, int64_t n,
               const float *__restrict__ E, const float *__restrict__ Rel, int h, float *__restrict__ score) {
  const int lane = threadIdx.x & 31;
  const int64_t t = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (t >= n) return;
  const float *es = E + (size_t)s[t] * h, *rp = Rel + (size_t)p[t] * h, *eo = E + (size_t)o[t] * h;
  float acc = 0.f;
  for (int k = lane; k < h; k += 32) acc = fmaf(es[k] * rp[k], eo[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) score[t] = acc;
}

// incidence keys: entry q < n is the subject role of triple q, entry q >= n the object role of triple q-n
__global__ void k_incidence(const int64_t *__restrict__ s, const int64_t *__restrict__ p, const int64_t *__restrict__ o,
                            int64_t n, int32_t *__restrict__ nk, int32_t *__restrict__ nv, int32_t *__restrict__ pk,
                            int32_t *__restrict__ pv) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  nk[t] = (int32_t)s[t]; nv[t] = (int32_t)t;
  nk[n + t] = (int32_t)o[t]; nv[n + t] = (int32_t)(n + t);
  pk[t] = (int32_t)p[t]; pv[t] = (int32_t)t;
}

// gE[node,:] = sum over the node's incidence run (sorted, stable => fixed order); one warp per run head
__global__ void __launch_bounds__(kThreads)
k_distmult_bwd_nodes(const int32_t *__restrict__ keys, const int32_t *__restrict__ vals, int64_t m, int64_t n,
                     const int64_t *__restrict__ s, const int64_t *__restrict__ p, const int64_t *__restrict__ o,
                     const float *__restrict__ g, const float *__restrict__ E, const float *__restrict__ Rel, int h,
                     float *__restrict__ gE) {
  const int lane = threadIdx.x & 31;
  const int64_t q = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (q >= m) return;
  const int node = keys[q];
  if (q > 0 && keys[q - 1] == node) return;
  int64_t end = q + 1;
  while (end < m && keys[end] == node) ++end;
  // the run's entries in order (fixed order => reproducible); per entry the index chain vals -> (p, other, g) is followed once
  // for all columns (KB x 32 of them in registers), two entries per trip so that their row loads are in flight together
  constexpr int KB = 8;
  for (int k0 = 0; k0 < h; k0 += 32 * KB) {
    float acc[KB];
#pragma unroll
    for (int c = 0; c < KB; ++c) acc[c] = 0.f;
    for (int64_t x = q; x < end; x += 2) {
      const bool two = x + 1 < end;
      const int v0 = vals[x], v1 = two ? vals[x + 1] : vals[x];
      const int64_t t0 = v0 < n ? v0 : v0 - n, t1 = v1 < n ? v1 : v1 - n;
      const float *r0 = Rel + (size_t)p[t0] * h, *r1 = Rel + (size_t)p[t1] * h;
      const float *e0 = E + (size_t)(v0 < n ? o[t0] : s[t0]) * h, *e1 = E + (size_t)(v1 < n ? o[t1] : s[t1]) * h;
      const float g0 = g[t0], g1 = g[t1];
      float a[KB], b[KB];
#pragma unroll
      for (int c = 0; c < KB; ++c) {
        const int k = k0 + lane + 32 * c;
        a[c] = k < h ? r0[k] * e0[k] : 0.f;
        b[c] = (two && k < h) ? r1[k] * e1[k] : 0.f;
      }
#pragma unroll
      for (int c = 0; c < KB; ++c) {
        acc[c] = fmaf(g0, a[c], acc[c]);
        if (two) acc[c] = fmaf(g1, b[c], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < KB; ++c) {
      const int k = k0 + lane + 32 * c;
      if (k < h) gE[(size_t)node * h + k] = acc[c];
    }
  }
}

__global__ void __launch_bounds__(kThreads)
k_distmult_bwd_rels(const int32_t *__restrict__ keys, const int32_t *__restrict__ vals, int64_t n,
                    const int64_t *__restrict__ s, const int64_t *__restrict__ o, const float *__restrict__ g,
                    const float *__restrict__ E, int h, float *__restrict__ gRel) {
  const int lane = threadIdx.x & 31;
  const int64_t q = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (q >= n) return;
  const int rel = keys[q];
  if (q > 0 && keys[q - 1] == rel) return;
  int64_t end = q + 1;
  while (end < n && keys[end] == rel) ++end;
  constexpr int KB = 8;   // as above: a popular relation holds a fifth of the batch, its run was one long dependent chain
  for (int k0 = 0; k0 < h; k0 += 32 * KB) {
    float acc[KB];
#pragma unroll
    for (int c = 0; c < KB; ++c) acc[c] = 0.f;
    for (int64_t x = q; x < end; x += 2) {
      const bool two = x + 1 < end;
      const int64_t t0 = vals[x], t1 = two ? vals[x + 1] : vals[x];
      const float *s0 = E + (size_t)s[t0] * h, *o0 = E + (size_t)o[t0] * h;
      const float *s1 = E + (size_t)s[t1] * h, *o1 = E + (size_t)o[t1] * h;
      const float g0 = g[t0], g1 = g[t1];
      float a[KB], b[KB];
#pragma unroll
      for (int c = 0; c < KB; ++c) {
        const int k = k0 + lane + 32 * c;
        a[c] = k < h ? s0[k] * o0[k] : 0.f;
        b[c] = (two && k < h) ? s1[k] * o1[k] : 0.f;
      }
#pragma unroll
      for (int c = 0; c < KB; ++c) {
        acc[c] = fmaf(g0, a[c], acc[c]);
        if (two) acc[c] = fmaf(g1, b[c], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < KB; ++c) {
      const int k = k0 + lane + 32 * c;
      if (k < h) gRel[(size_t)rel * h + k] = acc[c];
    }
  }
}

// temporary storage of the radix sort of m (key, value) pairs: part of the caller's workspace (no allocation in here)
static size_t sort_temp_bytes(int64_t m) {
  size_t bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                                  (const int32_t *)nullptr, (int32_t *)nullptr, (int)m, 0, 32, (cudaStream_t)0);
  if (e != cudaSuccess) {   // no device (host-only tests): a bound that only sizes a buffer nobody will use
    (void)cudaGetLastError();
    bytes = (size_t)(1 << 20) + 64 * (size_t)m;
  }
  return (bytes + 255) & ~(size_t)255;
}
static int sort_i32(int32_t *kin, int32_t *kout, int32_t *vin, int32_t *vout, int64_t m, void *tmp, size_t tmp_bytes,
                    cudaStream_t st) {
  size_t bytes = 0;
  MRGCN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)m, 0, 32, st));
  MRGCN_REQUIRE(bytes <= tmp_bytes, MRGCN_E_BADARG, "distmult_bwd: workspace too small for the sort (%zu > %zu bytes)", bytes,
                tmp_bytes);
  MRGCN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, (int)m, 0, 32, st));
  count_launch(3);
  return 0;
}

}  // namespace
}  // namespace mrgcn

using namespace mrgcn;

extern "C" int mrgcn_distmult_fwd(const int64_t *s, const int64_t *p, const int64_t *o, int64_t n, const float *E,
                                  const float *Rel, int32_t h, float *score, mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(h > 0 && n >= 0, MRGCN_E_BADARG, "distmult_fwd: bad sizes");
  if (n == 0) return 0;
  MRGCN_PROF("distmult_fwd");
  k_distmult_fwd<<<(unsigned)cdiv(n * 32, kThreads), kThreads, 0, st>>>(s, p, o, n, E, Rel, h, score);
  MRGCN_LAUNCH_CHECK();
  return 0;
}

// 12 n index words (incidence lists and their sorted copies), padded to 256 bytes, + the sort's temporary storage
static int64_t ws_index_words(int64_t n) { return ((12 * n + 63) / 64) * 64; }
extern "C" int64_t mrgcn_distmult_bwd_ws_elems(int64_t n) {
  n = n > 0 ? n : 1;
  return ws_index_words(n) + (int64_t)(sort_temp_bytes(2 * n) / 4);
}

extern "C" int mrgcn_distmult_bwd(const int64_t *s, const int64_t *p, const int64_t *o, int64_t n, const float *gscore,
                                  const float *E, const float *Rel, int64_t N, int64_t NR, int32_t h, float *gE,
                                  float *gRel, int32_t *ws, mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(h > 0 && n >= 0 && N < (1ll << 31) && NR < (1ll << 31) && n < (1ll << 29), MRGCN_E_BADARG,
                "distmult_bwd: bad sizes");
  if (gE) MRGCN_CUDA(cudaMemsetAsync(gE, 0, sizeof(float) * (size_t)N * h, st));
  if (gRel) MRGCN_CUDA(cudaMemsetAsync(gRel, 0, sizeof(float) * (size_t)NR * h, st));
  if (n == 0) return 0;
  int32_t *nk = ws, *nko = ws + 2 * n, *nv = ws + 4 * n, *nvo = ws + 6 * n;
  int32_t *pk = ws + 8 * n, *pko = ws + 9 * n, *pv = ws + 10 * n, *pvo = ws + 11 * n;
  void *tmp = ws + ws_index_words(n);
  const size_t tmp_bytes = sort_temp_bytes(2 * n);
  k_incidence<<<(unsigned)cdiv(n, kThreads), kThreads, 0, st>>>(s, p, o, n, nk, nv, pk, pv);
  MRGCN_LAUNCH_CHECK();
  if (gE) {
    if (int rc = sort_i32(nk, nko, nv, nvo, 2 * n, tmp, tmp_bytes, st)) return rc;
    MRGCN_PROF("distmult_bwd_nodes");
  k_distmult_bwd_nodes<<<(unsigned)cdiv(2 * n * 32, kThreads), kThreads, 0, st>>>(nko, nvo, 2 * n, n, s, p, o, gscore, E,
                                                                                    Rel, h, gE);
    MRGCN_LAUNCH_CHECK();
  }
  if (gRel) {
    if (int rc = sort_i32(pk, pko, pv, pvo, n, tmp, tmp_bytes, st)) return rc;
    MRGCN_PROF("distmult_bwd_rels");
  k_distmult_bwd_rels<<<(unsigned)cdiv(n * 32, kThreads), kThreads, 0, st>>>(pko, pvo, n, s, o, gscore, E, h, gRel);
    MRGCN_LAUNCH_CHECK();
  }
  return 0;
}
