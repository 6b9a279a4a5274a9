// Fused gradient clipping + Adam for the (sharded) parameters of the relational model.
//
// The reference's task loops run, per step, `nn.utils.clip_grad_norm_(model.parameters(), 1.0)` followed by
// `optimizer.step()` of a stock torch.optim.Adam (/root/reference/mrgcn/tasks/node_classification.py:190-193,
// mrgcn/tasks/utils.py:8-45).  On the AM identity table (66.7 M x 10 floats = 2.67 GB) that is ~14 passes over the
// table: norm (1 read), scale in place (1 read + 1 write), then Adam's foreach kernels over p, g, m, v.  Here:
//   mrgcn_grad_sqnorm   one read of every gradient: block partial sums of squares in double, fixed-order final sum
//                       -> one double per call, accumulated by the caller over tensors (and all-reduced over ranks when
//                       weight_I is sharded: every rank adds the sum of its shard, replicated tensors count once)
//   mrgcn_adam_clip     one pass: g' = g * min(1, max_norm / (sqrt(total) + 1e-6)) (clip_grad_norm_'s coefficient),
//                       m = b1 m + (1-b1) g', v = b2 v + (1-b2) g'^2, p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
//                       - the arithmetic of torch.optim.Adam (no amsgrad; weight decay added to g' as L2)
// 1 + 4 reads and 3 writes of the table instead of ~14 passes.
#include "common.cuh"

namespace mrgcn {
namespace {

constexpr int kOptThreads = 256;
constexpr int kOptBlocks = 148 * 8;

__global__ void __launch_bounds__(kOptThreads)
k_sqnorm_partial(const float *__restrict__ g, int64_t n, double *__restrict__ part) {
  __shared__ double red[kOptThreads];
  double acc = 0.0;
  const int64_t n4 = n / 4;
  const float4 *g4 = reinterpret_cast<const float4 *>(g);
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOptThreads) {
    const float4 v = g4[i];
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (int64_t i = 4 * n4 + threadIdx.x; i < n; i += kOptThreads) acc += (double)g[i] * g[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kOptThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// total[0] (+)= sum of part[0..nb) in a fixed order
__global__ void __launch_bounds__(kOptThreads)
k_sqnorm_final(const double *__restrict__ part, int nb, double *__restrict__ total, int accumulate) {
  __shared__ double red[kOptThreads];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += kOptThreads) acc += part[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kOptThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = (accumulate ? total[0] : 0.0) + red[0];
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, max_norm, bc1, bc2_sqrt;   // bc1 = 1 - beta1^t, bc2_sqrt = sqrt(1 - beta2^t)
};

__device__ __forceinline__ void adam1(float &p, float g, float &m, float &v, const AdamArgs &a, float coef) {
  g *= coef;
  if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
  m = fmaf(a.beta1, m, (1.f - a.beta1) * g);          // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.beta2, v, (1.f - a.beta2) * g * g);      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= (a.lr / a.bc1) * (m / denom);
}

__global__ void __launch_bounds__(kOptThreads)
k_adam_clip(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int64_t n,
            const double *__restrict__ total_sq, AdamArgs a) {
  float coef = 1.f;
  if (a.max_norm > 0.f && total_sq) {
    const float norm = (float)sqrt(total_sq[0]);
    coef = fminf(a.max_norm / (norm + 1e-6f), 1.f);   // torch.nn.utils.clip_grad_norm_: clamp(max_norm / (total + 1e-6), max = 1)
  }
  const int64_t n4 = n / 4;
  float4 *p4 = reinterpret_cast<float4 *>(p), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
  const float4 *g4 = reinterpret_cast<const float4 *>(g);
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOptThreads) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    adam1(pp.x, gg.x, mm.x, vv.x, a, coef);
    adam1(pp.y, gg.y, mm.y, vv.y, a, coef);
    adam1(pp.z, gg.z, mm.z, vv.z, a, coef);
    adam1(pp.w, gg.w, mm.w, vv.w, a, coef);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0)
    for (int64_t i = 4 * n4 + threadIdx.x; i < n; i += kOptThreads) adam1(p[i], g[i], m[i], v[i], a, coef);
}

// X[row_idx[i], col0 + c] = gate * src[i, c]   (gated scatter of one modality's encoder output, mrgcn.py:295-301)
__global__ void k_scatter_rows(const float *__restrict__ src, const int64_t *__restrict__ row_idx, const float *__restrict__ gate,
                               float *__restrict__ X, int64_t m, int d, int ldx, int col0) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= m * d) return;
  const int64_t i = x / d;
  const int c = (int)(x - i * d);
  X[row_idx[i] * ldx + col0 + c] = gate[0] * src[x];
}
// backward: g_src[i, c] = gate * gX[row_idx[i], col0 + c];  g_gate partial per block = sum src * gX
__global__ void __launch_bounds__(kOptThreads)
k_scatter_rows_bwd(const float *__restrict__ gX, const int64_t *__restrict__ row_idx, const float *__restrict__ gate,
                   const float *__restrict__ src, float *__restrict__ g_src, double *__restrict__ part, int64_t m, int d, int ldx,
                   int col0) {
  __shared__ double red[kOptThreads];
  double acc = 0.0;
  for (int64_t x = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; x < m * d; x += (int64_t)gridDim.x * kOptThreads) {
    const int64_t i = x / d;
    const int c = (int)(x - i * d);
    const float gx = gX[row_idx[i] * ldx + col0 + c];
    if (g_src) g_src[x] = gate[0] * gx;
    acc += (double)src[x] * gx;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kOptThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void k_sum_to_float(const double *__restrict__ part, int nb, float *__restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double acc = 0.0;
    for (int i = 0; i < nb; ++i) acc += part[i];
    out[0] = (float)acc;
  }
}

}  // namespace
}  // namespace mrgcn

using namespace mrgcn;

extern "C" int64_t mrgcn_sqnorm_ws_elems(void) { return kOptBlocks; }

extern "C" int mrgcn_grad_sqnorm(const float *g, int64_t n, double *ws, double *total, int32_t accumulate, mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(g && ws && total && n >= 0, MRGCN_E_BADARG, "grad_sqnorm: bad arguments");
  MRGCN_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, MRGCN_E_BADARG, "grad_sqnorm: gradient must be 16-byte aligned");
  int nb = (int)(cdiv(n > 0 ? n : 1, (int64_t)kOptThreads * 4) < kOptBlocks ? cdiv(n > 0 ? n : 1, (int64_t)kOptThreads * 4) : kOptBlocks);
  MRGCN_PROF("grad_sqnorm");
  k_sqnorm_partial<<<nb, kOptThreads, 0, st>>>(g, n, ws);
  MRGCN_LAUNCH_CHECK();
  k_sqnorm_final<<<1, kOptThreads, 0, st>>>(ws, nb, total, accumulate);
  MRGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mrgcn_adam_clip(float *p, const float *g, float *m, float *v, int64_t n, const double *total_sq, float max_norm,
                               float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                               mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(p && g && m && v && n >= 0 && step >= 1, MRGCN_E_BADARG, "adam_clip: bad arguments");
  MRGCN_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 15) == 0, MRGCN_E_BADARG, "adam_clip: tensors must be 16-byte aligned");
  if (n == 0) return 0;
  AdamArgs a{lr, beta1, beta2, eps, weight_decay, max_norm, (float)(1.0 - pow((double)beta1, (double)step)),
             (float)sqrt(1.0 - pow((double)beta2, (double)step))};
  const int64_t want = cdiv(n, (int64_t)kOptThreads * 4);
  const int nb = (int)(want < kOptBlocks ? want : kOptBlocks);
  MRGCN_PROF("adam_clip");
  k_adam_clip<<<nb, kOptThreads, 0, st>>>(p, g, m, v, n, total_sq, a);
  MRGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mrgcn_scatter_rows(const float *src, const int64_t *row_idx, const float *gate, float *X, int64_t m, int32_t d,
                                  int32_t ldx, int32_t col0, mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(src && row_idx && gate && X && d > 0 && ldx >= col0 + d, MRGCN_E_BADARG, "scatter_rows: bad arguments");
  if (m == 0) return 0;
  MRGCN_PROF("scatter_rows");
  k_scatter_rows<<<(unsigned)cdiv(m * d, 256), 256, 0, st>>>(src, row_idx, gate, X, m, d, ldx, col0);
  MRGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mrgcn_scatter_rows_bwd(const float *gX, const int64_t *row_idx, const float *gate, const float *src, float *g_src,
                                      float *g_gate, double *ws, int64_t m, int32_t d, int32_t ldx, int32_t col0,
                                      mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(gX && row_idx && gate && src && g_gate && ws && d > 0, MRGCN_E_BADARG, "scatter_rows_bwd: bad arguments");
  const int64_t want = cdiv(m * d > 0 ? m * d : 1, (int64_t)kOptThreads);
  const int nb = (int)(want < kOptBlocks ? want : kOptBlocks);
  MRGCN_PROF("scatter_rows_bwd");
  k_scatter_rows_bwd<<<nb, kOptThreads, 0, st>>>(gX, row_idx, gate, src, g_src, ws, m, d, ldx, col0);
  MRGCN_LAUNCH_CHECK();
  k_sum_to_float<<<1, 32, 0, st>>>(ws, nb, g_gate);
  MRGCN_LAUNCH_CHECK();
  return 0;
}
