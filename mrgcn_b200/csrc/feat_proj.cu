// Per-basis projection of the node features on the 5th-generation tensor cores (tcgen05, accumulators in TMEM),
// operands moved by the TMA engine through tensor maps (cp.async.bulk.tensor, SASS UTMALDG).
//
// With basis decomposition the feature term of the input layer (/root/reference/mrgcn/layers/graph.py:83-95) is
//     X[j,:] . W_F(r) = sum_b comp_F[r,b] * ( X[j,:] . V_F[b] )
// so the only dense contraction that survives is ONE genuine GEMM over the nodes, not over the edges:
//     P[j, b*out + o] = sum_k X[j,k] * V_F[b,k,o]            (AM: 1 666 764 x 151 x 400)
// after which the per-edge work is the same basis mixing the identity term needs (tab.cu mixes both tables in one pass).
//
// fp32 parity (1e-5 relative / 1e-6 absolute against the reference's fp32 einsum) rules out plain TF32.  Both operands
// are split into two tf32 pieces, x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi) (|x - hi - lo| <= 2^-24 |x|), and
// three MMAs are issued per K step (hi*hi + lo*hi + hi*lo; the dropped lo*lo is below 2^-22 of the product).  TMEM
// accumulation rounds toward zero (measured in round 1), so the hi*hi products of every PAIR of 32-wide K chunks get
// their own accumulator (8 accumulations each), all correction products share one, and the epilogue adds the partial
// accumulators in registers (round to nearest): the result is as close to the exact product as an fp32 FMA loop is.
//
// One persistent CTA per SM; a CTA keeps one NB-column slice of V (both pieces, all K chunks) resident in shared memory
// and walks the 128-row tiles of X.  The A operand is fed from TENSOR MEMORY: shared memory then only holds the raw fp32
// tiles on their way in (6 x 16 KB in flight per SM - the loads are latency-bound, profiles/r02d_ncu_full_summary.txt) and
// the MMA reads nothing but the V slice from shared memory.
//   warp 4      TMA producer: one 128 x 32 fp32 box of X per stage (tensor map, 128B swizzle)
//   warps 6-9   converters: lane = row; read the row's 32 values (un-swizzling the 16-byte chunks), split them into the
//               tf32 pieces hi / lo in registers and tcgen05.st both into a ring of 3 TMEM stages (64 columns each)
//   warp 5      MMA issuer (one thread) and TMEM owner: D[tmem] += A[tmem] . B[smem]
//   warps 0-3   epilogue: tcgen05.ld the 4 partial accumulators, add, release TMEM, store P rows (node-major)
// mbarrier rings: raw_full (TMA tx) / raw_empty (4 converter warps), a_full (4 converter warps) / a_empty (tcgen05.commit),
// tmem_full / acc_free[3] (the epilogue hands every accumulator back as soon as it has read it).  TMEM columns: 4 x NB
// accumulators + 3 x 64 operand stages <= 512.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "pipeline.cuh"
#include "rgcn_internal.cuh"

namespace mrgcn {
namespace {

constexpr int BM = 128;               // rows of X per tile (UMMA M)
constexpr int KCB = 128;              // bytes per row of a K chunk (32 fp32)
constexpr int A_TILE = BM * KCB;      // 16 KB
constexpr int RS = 6;                 // raw X stages in shared memory (TMA destinations)
constexpr int AS = 3;                 // operand stages in tensor memory (hi | lo, 64 columns each)
constexpr int kProjThreads = 320;     // 10 warps
constexpr int kMaxNKC = 6;

__device__ __forceinline__ uint32_t rna_tf32_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO(=1)<<16 | SBO(1024B>>4)<<32 | version 1<<46 | layout 2<<61
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// D[tmem] (+)= A[tmem] . B[smem]: A = 128 lanes x 8 columns (one tf32 per column) at a_tmem
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_one(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 2-D tiled TMA load: box at element coordinates (c0 = innermost, c1) of the tensor map -> shared memory
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32-byte global store (sm_100: 256-bit accesses): one whole sector per lane and instruction
__device__ __forceinline__ void st_global_v8(float *p, const float *v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
// one lane of a converged warp (warp-uniform control flow around it keeps the operands in uniform registers: a
// `lane == 0` branch makes the compiler wrap every tcgen05 / TMA instruction in an elect-and-broadcast loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Vt_hi / Vt_lo [NCpad][KP]: row c = b*out + o holds V[b, 0:in, o] (K-major), split into two tf32 pieces; zero padding
__global__ void k_vcat_split(const float *__restrict__ V, float *__restrict__ vt_hi, float *__restrict__ vt_lo, int B, int in,
                             int out, int KP, int NCpad) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= NCpad * KP) return;
  const int c = x / KP, k = x - c * KP;
  float w = 0.f;
  if (c < B * out && k < in) {
    const int b = c / out, o = c - b * out;
    w = __ldg(V + ((size_t)b * in + k) * out + o);
  }
  const uint32_t h = rna_tf32_bits(w);
  vt_hi[x] = __uint_as_float(h);
  vt_lo[x] = __uint_as_float(rna_tf32_bits(w - __uint_as_float(h)));
}

// X [N][in] -> Xp [N][KP] (zero padded rows of KP floats: 16-byte multiples, what a tensor map needs)
__global__ void k_pad_rows(const float *__restrict__ X, float *__restrict__ Xp, int64_t N, int in, int ldx, int KP) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte piece of a padded row
  const int q = KP / 4;
  if (x >= N * q) return;
  const int64_t j = x / q;
  const int k = 4 * (int)(x - j * q);
  const float *src = X + j * ldx + k;
  float4 v;
  v.x = k < in ? ldg_stream(src) : 0.f;
  v.y = k + 1 < in ? ldg_stream(src + 1) : 0.f;
  v.z = k + 2 < in ? ldg_stream(src + 2) : 0.f;
  v.w = k + 3 < in ? ldg_stream(src + 3) : 0.f;
  reinterpret_cast<float4 *>(Xp)[x] = v;
}

template <int NB>
__global__ void __launch_bounds__(kProjThreads, 1)
k_feat_proj(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmVhi,
            const __grid_constant__ CUtensorMap tmVlo, float *__restrict__ P, int64_t N, int NC, int NKC, int n_row_tiles,
            int NCH) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  constexpr int B_TILE = NB * KCB;                         // one K chunk of one piece of the V slice
  unsigned char *b_hi = base;                              // [NKC][NB rows][128 B]
  unsigned char *b_lo = b_hi + (size_t)NKC * B_TILE;
  unsigned char *a_raw = b_lo + (size_t)NKC * B_TILE;      // [RS][128 rows][128 B] raw fp32, 128B-swizzled by the TMA
  uint64_t *bars = reinterpret_cast<uint64_t *>(a_raw + (size_t)RS * A_TILE);
  uint64_t *raw_full = bars, *raw_empty = bars + RS, *a_full = bars + 2 * RS, *a_empty = a_full + AS, *b_full = a_empty + AS,
           *tmem_full = b_full + 1, *acc_free = b_full + 2;      // acc_free[0]: correction + main 0, [1]: main 1, [2]: main 2
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(b_full + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % NCH, rg = blockIdx.x / NCH, GR = gridDim.x / NCH;
  const int n_my = rg < n_row_tiles ? (n_row_tiles - rg + GR - 1) / GR : 0;
  const int NG = (NKC + 1) / 2;                            // main accumulators (one per pair of K chunks)
  constexpr int A_COL0 = 4 * NB;                           // operand stages behind the (at most) 4 accumulators

  if (threadIdx.x == 0) {
    for (int s = 0; s < RS; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 4); }
    for (int s = 0; s < AS; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    mbar_init(b_full, 1); mbar_init(tmem_full, 1);
    for (int s = 0; s < 3; ++s) mbar_init(&acc_free[s], 4);
    mbar_fence_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== epilogue =====================
    for (int it = 0; it < n_my; ++it) {
      const int t = rg + it * GR;
      mbar_wait(tmem_full, it & 1, 1, 64);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(warp * 32) << 16);
      // accumulator-major: the correction terms and main 0 first - the next tile's MMAs start on exactly those - and
      // each accumulator is handed back as soon as it has been read, so that the MMAs restart under the rest of the read
      float d[NB];
      {
        uint32_t r[NB / 16][16];
#pragma unroll
        for (int cb = 0; cb < NB / 16; ++cb) tmem_ld16(t0 + NG * NB + cb * 16, r[cb]);
        tmem_wait_ld();
#pragma unroll
        for (int cb = 0; cb < NB / 16; ++cb)
#pragma unroll
          for (int i = 0; i < 16; ++i) d[cb * 16 + i] = __uint_as_float(r[cb][i]);
      }
#pragma unroll
      for (int gq = 0; gq < 3; ++gq) {
        if (gq < NG) {
          uint32_t r[NB / 16][16];
#pragma unroll
          for (int cb = 0; cb < NB / 16; ++cb) tmem_ld16(t0 + gq * NB + cb * 16, r[cb]);
          tmem_wait_ld();
#pragma unroll
          for (int cb = 0; cb < NB / 16; ++cb)
#pragma unroll
            for (int i = 0; i < 16; ++i) d[cb * 16 + i] += __uint_as_float(r[cb][i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_one(&acc_free[gq]);
      }
      const int64_t j = (int64_t)t * BM + warp * 32 + lane;
      if (j < N) {
        float *row = P + j * NC + (size_t)chunk * NB;
        // a lane owns 4*NB contiguous bytes of its row: whole 32-byte sectors per store when the row pieces are 32-byte
        // aligned (half-sector 16-byte stores from 32 different rows were the limiter: stall lg_throttle, r02m profile)
        if ((NC & 7) == 0) {
#pragma unroll
          for (int c = 0; c < NB; c += 8)
            if (chunk * NB + c < NC) st_global_v8(row + c, d + c);
        } else {
#pragma unroll
          for (int c = 0; c < NB; c += 4)
            if (chunk * NB + c < NC) *reinterpret_cast<float4 *>(row + c) = make_float4(d[c], d[c + 1], d[c + 2], d[c + 3]);
        }
      }
    }
  } else if (warp == 4) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(b_full, (uint32_t)(2 * NKC * B_TILE));
      for (int kc = 0; kc < NKC; ++kc) {
        tma_load_2d(b_hi + (size_t)kc * B_TILE, &tmVhi, kc * 32, chunk * NB, b_full);
        tma_load_2d(b_lo + (size_t)kc * B_TILE, &tmVlo, kc * 32, chunk * NB, b_full);
      }
    }
    __syncwarp();
    int ks = 0;
    for (int it = 0; it < n_my; ++it) {
      const int t = rg + it * GR;
      for (int kc = 0; kc < NKC; ++kc, ++ks) {
        const int s = ks % RS;
        mbar_wait(&raw_empty[s], ((ks / RS) & 1) ^ 1, 2, 32);
        if (elect_one()) {
          mbar_expect_tx(&raw_full[s], A_TILE);
          tma_load_2d(a_raw + (size_t)s * A_TILE, &tmX, kc * 32, t * BM, &raw_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);      // warp-uniform for the compiler
      mbar_wait(b_full, 0, 3);
      // ONE lane issues every MMA of the CTA and its instruction stream is the critical path: the first versions branched
      // on `lane == 0` (every MMA then sits in an elect-and-broadcast loop) and rebuilt both 64-bit descriptors per MMA,
      // ~1100 issue cycles for a 480-cycle stage (profiles/r02m_*).  Here the whole warp runs the loop, one elected lane
      // issues, and the B descriptors - which differ only in their 14-bit start address - are advanced by constants.
      const uint64_t dbh_base = desc_k_sw128(smem_u32(b_hi)), dbl_base = desc_k_sw128(smem_u32(b_lo));
      const uint32_t d_corr = tb + NG * NB;
      int ks = 0;
      for (int it = 0; it < n_my; ++it) {
        for (int kc = 0; kc < NKC; ++kc, ++ks) {
          if ((kc & 1) == 0) {      // first use of main accumulator kc/2 (and, at kc = 0, of the correction accumulator)
            mbar_wait(&acc_free[kc >> 1], (it & 1) ^ 1, 4);
          }
          const int s = ks % AS;
          mbar_wait(&a_full[s], (ks / AS) & 1, 5);
          tc_fence_after();
          const uint32_t a_hi = tb + A_COL0 + s * 64, a_lo = a_hi + 32;
          const uint64_t dbh = dbh_base + (uint64_t)((kc * B_TILE) >> 4), dbl = dbl_base + (uint64_t)((kc * B_TILE) >> 4);
          const uint32_t d_main = tb + (kc >> 1) * NB;
          if (elect_one()) {
            // 4 K steps of 8 tf32: 8 TMEM columns of A, 32 bytes (2 descriptor units) of every B row
            mma_tf32_ts(d_main, a_hi, dbh, idesc, kc & 1);
            mma_tf32_ts(d_corr, a_lo, dbh, idesc, kc);
            mma_tf32_ts(d_corr, a_hi, dbl, idesc, 1);
#pragma unroll
            for (int j = 1; j < 4; ++j) {
              mma_tf32_ts(d_main, a_hi + j * 8, dbh + 2 * j, idesc, 1);
              mma_tf32_ts(d_corr, a_lo + j * 8, dbh + 2 * j, idesc, 1);
              mma_tf32_ts(d_corr, a_hi + j * 8, dbl + 2 * j, idesc, 1);
            }
            tc_commit(&a_empty[s]);        // operand stage free once these MMAs have read it
            if (kc == NKC - 1) tc_commit(tmem_full);      // all accumulators of the tile complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== converters =====================
    const int q4 = warp & 3;             // TMEM lane quarter this warp may access = rows [32 q4, 32 q4 + 32) of the tile
    const int row = q4 * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16) + A_COL0;
    int ks = 0;
    for (int it = 0; it < n_my; ++it) {
      for (int kc = 0; kc < NKC; ++kc, ++ks) {
        const int rs = ks % RS, as = ks % AS;
        mbar_wait(&raw_full[rs], (ks / RS) & 1, 6);
        const unsigned char *rp = a_raw + (size_t)rs * A_TILE + row * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {      // logical 16-byte chunk c of the row sits at physical chunk c ^ (row & 7)
          const float4 x = *reinterpret_cast<const float4 *>(rp + ((c ^ (row & 7)) << 4));
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t h = rna_tf32_bits(xs[u]);
            hi[4 * c + u] = h;
            lo[4 * c + u] = rna_tf32_bits(xs[u] - __uint_as_float(h));
          }
        }
        // The raw tile is in registers: its slot may be refilled.  The refill is an async-proxy (TMA) write after
        // generic-proxy reads of the same bytes: without the proxy fence the new tile was seen landing under the loads.
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_one(&raw_empty[rs]);
        mbar_wait(&a_empty[as], ((ks / AS) & 1) ^ 1, 7);
        tc_fence_after();
        tmem_st32(t_lane + as * 64, hi);
        tmem_st32(t_lane + as * 64 + 32, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_one(&a_full[as]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp32 matrix [rows][cols] with row pitch `pitch` floats, boxes of box_rows x 32 floats, 128B swizzle, zero fill out of bounds
static int make_map(CUtensorMap *m, const float *ptr, int64_t rows, int cols, int pitch, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  MRGCN_REQUIRE(enc, MRGCN_E_NOTSUP, "feat_proj: cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MRGCN_REQUIRE(r == CUDA_SUCCESS, MRGCN_E_BADARG, "feat_proj: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

static int pick_nb(int NC) {
  // widest column slice whose 4 accumulators fit the 512 TMEM columns and whose V slice fits shared memory next to the
  // X stages; NC must be a multiple of it so that every CTA does the same work
  const int cand[] = {80, 64, 48, 32, 16};
  for (int nb : cand)
    if (NC % nb == 0) return nb;
  return 0;
}

}  // namespace

bool feat_proj_supported(int in, int ldx, int B, int out) {
  if (B <= 0 || in < 32 || out <= 0) return false;
  const int NC = B * out;
  if (NC % 4 != 0 || NC > 1024 || pick_nb(NC) == 0) return false;
  const int KP = ((in + 31) / 32) * 32;
  if (KP / 32 > kMaxNKC) return false;
  (void)ldx;
  return true;
}

// vt_ws: 2 * NC * KP floats.  X rows must be KP = ceil32(in) floats apart with zeros beyond `in` (ldx == KP), or a padded
// copy is made into xpad_ws [N * KP].
int launch_feat_proj(const float *X, int64_t N, int in, int ldx, const float *V, int B, int out, float *vt_ws, float *xpad_ws,
                     float *P, cudaStream_t st) {
  MRGCN_REQUIRE(feat_proj_supported(in, ldx, B, out), MRGCN_E_NOTSUP, "feat_proj: unsupported shape in=%d B=%d out=%d", in, B, out);
  if (N == 0) return 0;
  const int NC = B * out, KP = ((in + 31) / 32) * 32, NKC = KP / 32, NB = pick_nb(NC), NCH = NC / NB;
  float *vt_hi = vt_ws, *vt_lo = vt_ws + (size_t)NC * KP;
  MRGCN_PROF("vcat_split");
  k_vcat_split<<<(unsigned)cdiv((int64_t)NC * KP, 256), 256, 0, st>>>(V, vt_hi, vt_lo, B, in, out, KP, NC);
  MRGCN_LAUNCH_CHECK();
  const float *Xp = X;
  if (ldx != KP) {
    MRGCN_REQUIRE(xpad_ws, MRGCN_E_BADARG, "feat_proj: X rows are %d floats apart, need %d (or a padding workspace)", ldx, KP);
    MRGCN_PROF("pad_rows");
    k_pad_rows<<<(unsigned)cdiv(N * (KP / 4), 256), 256, 0, st>>>(X, xpad_ws, N, in, ldx, KP);
    MRGCN_LAUNCH_CHECK();
    Xp = xpad_ws;
  }
  MRGCN_REQUIRE((reinterpret_cast<uintptr_t>(Xp) & 15) == 0 && (reinterpret_cast<uintptr_t>(vt_ws) & 15) == 0, MRGCN_E_BADARG,
                "feat_proj: operands must be 16-byte aligned");
  CUtensorMap tmX, tmVhi, tmVlo;
  if (int rc = make_map(&tmX, Xp, N, KP, KP, BM)) return rc;
  if (int rc = make_map(&tmVhi, vt_hi, NC, KP, KP, NB)) return rc;
  if (int rc = make_map(&tmVlo, vt_lo, NC, KP, KP, NB)) return rc;
  const int n_row_tiles = (int)cdiv(N, BM);
  int GR = kNumSMs / NCH;
  if (GR > n_row_tiles) GR = n_row_tiles;
  if (GR < 1) GR = 1;
  const unsigned grid = (unsigned)(GR * NCH);
  const size_t smem = 1024 + (size_t)2 * NKC * NB * KCB + (size_t)RS * A_TILE + 256;
  MRGCN_REQUIRE(smem <= 227 * 1024, MRGCN_E_NOTSUP, "feat_proj: shared memory (%zu B)", smem);
  MRGCN_PROF("feat_proj");
#define LAUNCH(NBV)                                                                                                   \
  do {                                                                                                                \
    MRGCN_CUDA(cudaFuncSetAttribute(k_feat_proj<NBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    k_feat_proj<NBV><<<grid, kProjThreads, smem, st>>>(tmX, tmVhi, tmVlo, P, N, NC, NKC, n_row_tiles, NCH);           \
  } while (0)
  switch (NB) {
    case 80: LAUNCH(80); break;
    case 64: LAUNCH(64); break;
    case 48: LAUNCH(48); break;
    case 32: LAUNCH(32); break;
    default: LAUNCH(16); break;
  }
#undef LAUNCH
  MRGCN_LAUNCH_CHECK();
  return 0;
}

}  // namespace mrgcn

using namespace mrgcn;

extern "C" int32_t mrgcn_feat_proj_supported(int32_t in_dim, int32_t B, int32_t out_dim) {
  return feat_proj_supported(in_dim, 0, B, out_dim) ? ((in_dim + 31) / 32) * 32 : 0;
}

extern "C" int mrgcn_feat_proj(const float *X, int64_t N, int32_t in_dim, int32_t x_stride, const float *weight_F, int32_t B,
                               int32_t out_dim, float *vt_ws, float *xpad_ws, float *P, mrgcn_stream_t stream) {
  MRGCN_REQUIRE(X && weight_F && vt_ws && P, MRGCN_E_BADARG, "feat_proj: null argument");
  return launch_feat_proj(X, N, in_dim, x_stride > 0 ? x_stride : in_dim, weight_F, B, out_dim, vt_ws, xpad_ws, P,
                          (cudaStream_t)stream);
}

// Host rows (pinned, contiguous [rows][cols] fp32) -> device rows of `pitch` floats (pitch >= cols; the pad columns are not
// touched: the caller zeroes the buffer once).  Replaces the `.to(device)` of the feature matrix in MRGCN's forward
// (/root/reference/mrgcn/models/mrgcn.py:203-204) when the features feed the tensor-map loads of mrgcn_feat_proj.
extern "C" int mrgcn_upload_rows(const float *X_host, int64_t rows, int32_t cols, float *X_dev, int32_t pitch,
                                 mrgcn_stream_t stream) {
  MRGCN_REQUIRE(X_host && X_dev && cols > 0 && pitch >= cols && rows >= 0, MRGCN_E_BADARG, "upload_rows: bad arguments");
  if (rows == 0) return 0;
  MRGCN_CUDA(cudaMemcpy2DAsync(X_dev, (size_t)pitch * 4, X_host, (size_t)cols * 4, (size_t)cols * 4, (size_t)rows,
                               cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
