// Feature-term kernels on the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
// The feature term of one relation is a gathered GEMM:  Z[e, :] = X[j_e, :] . W_r  over the relation's edges
// (forward, /root/reference/mrgcn/layers/graph.py:93-95 re-associated) and  dW_r = sum_e X[j_e, :]^T t_e  (weight
// gradient).  Both read the same shared-memory image of the gathered rows of X: tiles of 128 edges x 32 features
// (128-byte rows, 128B swizzle).  Read with M = edges it is the canonical K-major A operand of the forward MMA, read
// with M = features it is the canonical MN-major A operand of the gradient MMA (CUTLASS cute/atom/mma_traits_sm100.hpp,
// make_umma_desc).
//
// fp32 parity (1e-5 relative / 1e-6 absolute, element-wise) rules out plain TF32 (11-bit significand) and also the
// usual 3xTF32 (two pieces keep 22 of the 24 bits: measured 5e-6 absolute error on O(1) outputs).  Every operand is
// split on the fly into THREE tf32 pieces x = h + m + l (exact) and six MMAs are issued per K step
// (h*h + h*m + m*h + m*m + h*l + l*h; everything dropped is below 2^-33), fp32 accumulation in TMEM.  The split
// happens in the producer warps while the rows are on their way from HBM to shared memory; the tensor pipe has
// >10x headroom here, so the kernels stay HBM-bound.  MRGCN_FEAT_TC_PIECES=2 selects plain 3xTF32.
//
// Warp roles (one persistent CTA per SM): warps 0-3 epilogue (tcgen05.ld -> sum partial accumulators -> scale by val ->
// global), warp 4 MMA issuer (one elected thread) and TMEM owner, warps 5-28 producers.  A stage (128 rows x 128 B, all
// pieces) is filled by four producer warps, one quarter (32 rows) each; the 24 producer warps take quarters round-robin,
// so every warp has its 32 coalesced 128-byte loads in flight (~100 KB per SM) while the stages ahead of it are being
// consumed, and the edge indices of its next quarter are prefetched one round earlier.
// mbarrier rings: a_full/a_empty (A stages), w_full/w_empty (W_r double buffer), d_full/d_empty (2 accumulator sets).
// Accumulation in TMEM rounds toward zero; to keep the bias at fp32-FMA level every k-chunk's h*h products get their own
// accumulator (4 accumulations each) and all correction products share one; the epilogue adds the NKC+1 partials.
#include <stdlib.h>

#include "common.cuh"
#include "pipeline.cuh"
#include "rgcn_internal.cuh"

namespace mrgcn {
namespace {

constexpr int TM = 128;             // edges per tile (UMMA M)
constexpr int KCB = 128;            // bytes per row of a k-chunk (32 fp32)
constexpr int A_TILE = TM * KCB;    // 16 KB: one tf32 piece of a stage
constexpr int kEpiWarps = 4, kProdWarps = 24;
constexpr int kTcThreads = (kEpiWarps + 1 + kProdWarps) * 32;   // 928
constexpr int kMaxMyChunks = 1024;  // chunk table of one CTA in shared memory
// "stage consumed" barriers are indexed by the stage counter mod NB (not by ring slot): a producer group revisits a
// slot only every 6 stages, i.e. up to two ring laps later, and a parity wait is only valid within one phase.
constexpr int NB = 12;

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// byte offset of (row, byte column c) inside a 128B-swizzled tile with 128-byte rows (Swizzle<3,4,3>)
__device__ __forceinline__ uint32_t swz128(int row, int c) { return row * 128 + ((((c >> 4) ^ (row & 7)) << 4) | (c & 15)); }

__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO(=1)<<16 | SBO(1024B>>4)<<32 | version 1<<46 | layout 2<<61
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive1(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NP>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&d)[NP]);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&d)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float (&d)[32]) {
  float lo[16], hi[16];
  tmem_ld<16>(taddr, lo);
  tmem_ld<16>(taddr + 16, hi);
#pragma unroll
  for (int i = 0; i < 16; ++i) { d[i] = lo[i]; d[16 + i] = hi[i]; }
}

// ---------------------------------------------------------------------------------------------------------------
// forward:  msg[e3, 0:out] = val_e * X[gather[e3], :] . W[r]
template <int NP, int NS>
__global__ void __launch_bounds__(kTcThreads, 1)
k_feat_msg_tc(const float *__restrict__ X, const float *__restrict__ W, const int32_t *__restrict__ chunk_rel,
              const int32_t *__restrict__ chunk_ptr, const int32_t *__restrict__ gather,
              const float *__restrict__ val, float *__restrict__ msg, int in, int out, int n_chunks, int NKC, int SA,
              int w_bytes, int tcols, int dbg) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte aligned base (swizzle atoms)
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char *a_st = base;                              // [SA][NS pieces][128 rows][128 B]
  unsigned char *w_st = a_st + (size_t)SA * NS * A_TILE;   // [2][NKC][NS pieces][NP rows][128 B]
  uint64_t *bars = reinterpret_cast<uint64_t *>(w_st + 2 * (size_t)w_bytes);
  uint64_t *a_full = bars, *a_empty = bars + SA, *w_full = a_empty + NB, *w_empty = w_full + 2, *d_full = w_empty + 2,
           *d_empty = d_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_empty + 2);
  int *ctab = reinterpret_cast<int *>(tmem_slot + 4);      // [n_my][3]: e_lo, e_hi, rel of this CTA's chunks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (n_chunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) mbar_init(&a_full[s], 4);
    for (int s = 0; s < NB; ++s) mbar_init(&a_empty[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&w_full[s], kProdWarps); mbar_init(&w_empty[s], 1);
      mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], kEpiWarps);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < n_my; i += kTcThreads) {
    const int c = blockIdx.x + i * gridDim.x;
    ctab[3 * i] = chunk_ptr[c]; ctab[3 * i + 1] = chunk_ptr[c + 1]; ctab[3 * i + 2] = chunk_rel[c];
  }
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tcols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ACC = (NKC + 1) * NP;   // TMEM columns of one accumulator set: NKC main partials + 1 correction

  if (warp < kEpiWarps) {
    // ===================== epilogue =====================
    int it = 0;
    for (int ci = 0; ci < n_my; ++ci) {
      const int e_lo = ctab[3 * ci], e_hi = ctab[3 * ci + 1];
      for (int e0 = e_lo; e0 < e_hi; e0 += TM, ++it) {
        const int acc = it & 1;
        mbar_wait(&d_full[acc], (it >> 1) & 1, 1, 100);
        tc_fence_after();
        float d[NP];
        const uint32_t t0 = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * ACC;
        tmem_ld<NP>(t0 + NKC * NP, d);             // correction terms first (smallest)
        for (int kc = 0; kc < NKC; ++kc) {
          float p[NP];
          tmem_ld<NP>(t0 + kc * NP, p);
#pragma unroll
          for (int o = 0; o < NP; ++o) d[o] += p[o];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive1(&d_empty[acc]);
        const int e = e0 + warp * 32 + lane;
        if (e < e_hi && !(dbg & 4)) {
          const float v = val[e];
          const int ms = msg_stride(out);
#pragma unroll
          for (int o = 0; o < NP; ++o) d[o] *= v;
          store_msg_chunk<NP>(msg + (size_t)e * ms, 0, out, ms, d);
        }
      }
    }
  } else if (warp == kEpiWarps) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      int it = 0, ks = 0;
      for (int ci = 0; ci < n_my; ++ci) {
        const int e_lo = ctab[3 * ci], e_hi = ctab[3 * ci + 1];
        const int wb = ci & 1;
        mbar_wait(&w_full[wb], (ci >> 1) & 1, 2);
        const uint32_t wbase = smem_u32(w_st + (size_t)wb * w_bytes);
        for (int e0 = e_lo; e0 < e_hi; e0 += TM, ++it) {
          const int acc = it & 1;
          mbar_wait(&d_empty[acc], ((it >> 1) & 1) ^ 1, 3);
          tc_fence_after();
          const uint32_t d_set = tmem_base + acc * ACC;
          const uint32_t d_corr = d_set + NKC * NP;
          for (int kc = 0; kc < NKC; ++kc, ++ks) {
            const int s = ks % SA;
            mbar_wait(&a_full[s], (ks / SA) & 1, 4);
            tc_fence_after();
            const uint32_t a0 = smem_u32(a_st + (size_t)s * NS * A_TILE);
            const uint32_t b0 = wbase + (uint32_t)(kc * NS) * (NP * KCB);
            const uint32_t d_main = d_set + kc * NP;
            if (!(dbg & 2)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {  // 4 K-steps of 8 tf32 (32 bytes) inside the 128-byte row
                uint64_t da[NS], db[NS];
#pragma unroll
                for (int q = 0; q < NS; ++q) {
                  da[q] = make_desc_k_sw128(a0 + q * A_TILE + j * 32);
                  db[q] = make_desc_k_sw128(b0 + q * (NP * KCB) + j * 32);
                }
                mma_tf32(d_main, da[0], db[0], idesc, j != 0);
                const uint32_t first = (kc | j) != 0;
                if constexpr (NS == 3) {   // smallest products first
                  mma_tf32(d_corr, da[2], db[0], idesc, first);
                  mma_tf32(d_corr, da[0], db[2], idesc, 1);
                  mma_tf32(d_corr, da[1], db[1], idesc, 1);
                  mma_tf32(d_corr, da[1], db[0], idesc, 1);
                  mma_tf32(d_corr, da[0], db[1], idesc, 1);
                } else {
                  mma_tf32(d_corr, da[1], db[0], idesc, first);
                  mma_tf32(d_corr, da[0], db[1], idesc, 1);
                }
              }
            }
            tc_commit(&a_empty[ks % NB]);   // stage ks consumed once these MMAs have read it
          }
          tc_commit(&d_full[acc]);    // accumulator set complete
        }
        tc_commit(&w_empty[wb]);      // W buffer free once every MMA of the chunk is done
      }
    }
  } else {
    // ===================== producers =====================
    const int pw = warp - kEpiWarps - 1;          // 0..23
    const int ptid = pw * 32 + lane;
    const int q = pw & 3;                         // my quarter (32 rows) of every stage I fill
    const int sphase = pw >> 2;                   // I fill stages ks with ks % 6 == sphase
    constexpr int SPER = kProdWarps / 4;          // 6
    // iterator over (chunk, tile) pairs, flattened with the k-chunk index into the stage counter ks
    int ks = 0;
    // prefetch state: edge indices of my next quarter
    // walk: for each chunk, W staging by everybody, then stages
    for (int ci = 0; ci < n_my; ++ci) {
      const int e_lo = ctab[3 * ci], e_hi = ctab[3 * ci + 1], r = ctab[3 * ci + 2];
      const int wb = ci & 1;
      mbar_wait(&w_empty[wb], ((ci >> 1) & 1) ^ 1, 5, 200);
      {
        unsigned char *wbuf = w_st + (size_t)wb * w_bytes;
        const float *Wr = W + (size_t)r * in * out;
        const int total = NKC * 32 * NP;
        for (int idx = ptid; idx < total; idx += kProdWarps * 32) {
          const int k = idx / NP, o = idx - k * NP;
          const float w = (k < in && o < out) ? __ldg(Wr + (size_t)k * out + o) : 0.f;
          unsigned char *t = wbuf + (size_t)((k >> 5) * NS) * (NP * KCB) + swz128(o, (k & 31) * 4);
          float rem = w;
#pragma unroll
          for (int p = 0; p < NS; ++p) {
            const uint32_t piece = tf32_rna(rem);
            rem -= __uint_as_float(piece);
            *reinterpret_cast<uint32_t *>(t + p * (NP * KCB)) = piece;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive1(&w_full[wb]);
      }
      for (int e0 = e_lo; e0 < e_hi; e0 += TM) {
        int j = -2;   // not loaded yet for this tile
        for (int kc = 0; kc < NKC; ++kc, ++ks) {
          if (ks % SPER != sphase) continue;
          if (j == -2) {
            const int er = e0 + q * 32 + lane;
            j = (er < e_hi) ? gather[er] : -1;
          }
          const int k = kc * 32 + lane;
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int ji = __shfl_sync(0xffffffffu, j, i);
            x[i] = (ji >= 0 && k < in && !(dbg & 1)) ? __ldg(X + (size_t)ji * in + k) : 0.f;
          }
          const int s = ks % SA;
          if (ks >= SA) mbar_wait(&a_empty[(ks - SA) % NB], ((ks - SA) / NB) & 1, 6, 100);   // stage ks-SA consumed: slot free
          unsigned char *ah = a_st + (size_t)s * NS * A_TILE;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            unsigned char *t = ah + swz128(q * 32 + i, lane * 4);
            float rem = x[i];
#pragma unroll
            for (int p = 0; p < NS; ++p) {
              const uint32_t piece = tf32_rna(rem);
              rem -= __uint_as_float(piece);
              *reinterpret_cast<uint32_t *>(t + p * A_TILE) = piece;
            }
          }
          if (!(dbg & 8)) fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive1(&a_full[s]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tcols));
  }
}

}  // namespace

static int g_feat_tc_mode = -1;   // -1: read MRGCN_FEAT_TC on first use
void set_feat_tc_mode(int mode) { g_feat_tc_mode = mode; }

// returns 1 when the tensor-core path applies (and was launched), 0 when the caller must use the CUDA-core kernel
int launch_feat_msg_tc(const mrgcn_graph *g, const int32_t *gather, const float *X, const float *W, float *msg, int in,
                       int out, cudaStream_t st, const char *prof_name, int *launched) {
  *launched = 0;
  if (out > 32 || g->n_chunks == 0) return 0;
  int &enabled = g_feat_tc_mode;
  if (enabled < 0) {
    // MRGCN_FEAT_TC: 0 = never, 1 = whenever it applies, unset = only where it pays.  The gather + tf32 split costs
    // about as many issue slots per element as `out` <= 16 FMAs do on the CUDA cores, so skinny layers stay there.
    const char *e = getenv("MRGCN_FEAT_TC");
    enabled = !e ? 2 : (e[0] == '0' ? 0 : 1);
  }
  if (!enabled || (enabled == 2 && out <= 16)) return 0;
  static int dbg = -1;
  if (dbg < 0) { const char *e = getenv("MRGCN_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  static int pieces = 0;
  if (!pieces) {
    const char *e = getenv("MRGCN_FEAT_TC_PIECES");
    pieces = (e && e[0] == '2') ? 2 : 3;   // three pieces unless MRGCN_FEAT_TC_PIECES=2 (plain 3xTF32) is asked for
  }
  const int NP = out <= 16 ? 16 : 32;
  const int NKC = (int)cdiv(in, 32);
  const int w_bytes = NKC * pieces * NP * KCB;
  int grid = g->n_chunks < kNumSMs ? g->n_chunks : kNumSMs;
  const int n_my = (int)cdiv(g->n_chunks, grid);
  int tcols = 32;
  while (tcols < 2 * (NKC + 1) * NP) tcols <<= 1;
  if (tcols > 512 || n_my > kMaxMyChunks) return 0;
  int SA = 4;
  size_t smem = 0;
  for (; SA >= 2; --SA) {
    smem = 1024 + (size_t)SA * pieces * A_TILE + 2 * (size_t)w_bytes + (SA + NB + 8) * 8 + 32 + (size_t)n_my * 12;
    if (smem <= 220 * 1024) break;
  }
  if (SA < 2) return 0;
  static thread_local char tc_name[64];
  snprintf(tc_name, sizeof(tc_name), "%s_tc", prof_name);
  mrgcn::prof_begin(tc_name, st);
#define LAUNCH(NPV, NSV)                                                                                            \
  do {                                                                                                              \
    MRGCN_CUDA(cudaFuncSetAttribute(k_feat_msg_tc<NPV, NSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_feat_msg_tc<NPV, NSV><<<grid, kTcThreads, smem, st>>>(X, W, g->chunk_rel, g->chunk_ptr, gather, g->e3_val, msg, in, \
                                                            out, g->n_chunks, NKC, SA, w_bytes, tcols, dbg);        \
  } while (0)
  if (NP == 16) { if (pieces == 3) LAUNCH(16, 3); else LAUNCH(16, 2); }
  else { if (pieces == 3) LAUNCH(32, 3); else LAUNCH(32, 2); }
#undef LAUNCH
  MRGCN_LAUNCH_CHECK();
  *launched = 1;
  return 0;
}

}  // namespace mrgcn

extern "C" void mrgcn_set_feat_tc(int mode) { mrgcn::set_feat_tc_mode(mode < 0 ? -1 : (mode > 2 ? 2 : mode)); }
