// Graph construction: the reference's stacked adjacency (scipy CSR -> torch sparse COO,
// /root/reference/mrgcn/data/utils.py:165-170, mrgcn/data/batch.py:144-149) re-emitted as three
// sorted edge orders (E1 dst-major, E2 src-major, E3 rel-major) that the layer kernels stream.
// One-time work per adjacency object; integer/structure results are bit-exact by construction
// (stable radix sorts, no atomics).
#include <cub/device/device_radix_sort.cuh>
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace mrgcn {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- per-kernel profile: events bracket each launch on its own stream; read back after a sync --------
struct ProfRec { std::string name; cudaEvent_t a, b; bool open; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
static thread_local ProfRec g_pending = {std::string(), nullptr, nullptr, false};
static thread_local cudaStream_t g_pending_stream = nullptr;

void prof_begin(const char *name, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfRec r{std::string(name), nullptr, nullptr, true};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_pending = r;
  g_pending_stream = st;
}
void prof_end() {
  if (!g_pending.open) return;
  cudaEventRecord(g_pending.b, g_pending_stream);
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(g_pending);
  }
  g_pending.open = false;
}

namespace {

struct TempBuf {  // stream-ordered scratch, freed on scope exit
  void *p = nullptr;
  cudaStream_t s;
  explicit TempBuf(cudaStream_t st) : s(st) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 16, s); }
  ~TempBuf() {
    if (p) cudaFreeAsync(p, s);
  }
  template <class T>
  T *as() { return reinterpret_cast<T *>(p); }
};

__global__ void k_key1(const int64_t *__restrict__ row, const int64_t *__restrict__ col, int64_t ncols,
                       uint64_t *__restrict__ key, int32_t *__restrict__ idx, int64_t E) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  key[e] = (uint64_t)row[e] * (uint64_t)ncols + (uint64_t)col[e];
  idx[e] = (int32_t)e;
}

// E1 arrays from the (row,col)-sorted permutation; also emits the E2 sort key (src*R + rel).
__global__ void k_fill_e1(const int32_t *__restrict__ perm, const int64_t *__restrict__ row,
                          const int64_t *__restrict__ col, const float *__restrict__ val, int32_t NS,
                          int32_t R, int32_t *__restrict__ e1_dst, int32_t *__restrict__ e1_src,
                          int32_t *__restrict__ e1_rel, float *__restrict__ e1_val,
                          uint64_t *__restrict__ key2, int32_t *__restrict__ ident, int64_t E) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= E) return;
  int32_t e = perm[p];
  int64_t c = col[e];
  int32_t r = (int32_t)(c / NS), j = (int32_t)(c - (int64_t)r * NS);
  e1_dst[p] = (int32_t)row[e];
  e1_src[p] = j;
  e1_rel[p] = r;
  e1_val[p] = val[e];
  key2[p] = (uint64_t)j * (uint64_t)R + (uint64_t)r;
  ident[p] = (int32_t)p;
}

// ptr[i] = first position whose (non-decreasing) key is >= i, i in [0, n]
__global__ void k_ptr(const int32_t *__restrict__ keys, int64_t E, int32_t n, int32_t *__restrict__ ptr) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  int64_t lo = 0, hi = E;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < (int32_t)i) lo = mid + 1; else hi = mid;
  }
  ptr[i] = (int32_t)lo;
}

__global__ void k_fill_e2(const int32_t *__restrict__ e2_to_e1, const int32_t *__restrict__ e1_dst,
                          const int32_t *__restrict__ e1_src, const int32_t *__restrict__ e1_rel,
                          const float *__restrict__ e1_val, int32_t *__restrict__ e2_src,
                          int32_t *__restrict__ e2_dst, int32_t *__restrict__ e2_rel,
                          float *__restrict__ e2_val, int32_t *__restrict__ e1_to_e2,
                          uint32_t *__restrict__ key3, int32_t *__restrict__ ident, int64_t E, int32_t slab_rows,
                          int32_t R) {
  int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= E) return;
  int32_t p = e2_to_e1[q];
  e2_src[q] = e1_src[p];
  e2_dst[q] = e1_dst[p];
  int32_t r = e1_rel[p];
  e2_rel[q] = r;
  e2_val[q] = e1_val[p];
  e1_to_e2[p] = (int32_t)q;
  key3[q] = (uint32_t)(e1_src[p] / slab_rows) * (uint32_t)R + (uint32_t)r;   // E3 group = (source slab, relation)
  ident[q] = (int32_t)q;
}

__global__ void k_fill_e3(const int32_t *__restrict__ e3_to_e2_in, const int32_t *__restrict__ e2_to_e1,
                          const int32_t *__restrict__ e2_src, const int32_t *__restrict__ e2_dst,
                          const int32_t *__restrict__ e2_rel, const float *__restrict__ e2_val,
                          int32_t *__restrict__ e3_src, int32_t *__restrict__ e3_dst,
                          int32_t *__restrict__ e3_rel, float *__restrict__ e3_val,
                          int32_t *__restrict__ e3_to_e2, int32_t *__restrict__ e1_to_e3, int32_t *__restrict__ e2_to_e3,
                          int64_t E, int32_t slab_rows, int32_t R) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= E) return;
  int32_t q = e3_to_e2_in[t];
  e3_src[t] = e2_src[q];
  e3_dst[t] = e2_dst[q];
  e3_rel[t] = (e2_src[q] / slab_rows) * R + e2_rel[q];   // group id (slab, rel): non-decreasing in E3 order
  e3_val[t] = e2_val[q];
  e3_to_e2[t] = q;
  e2_to_e3[q] = (int32_t)t;
  e1_to_e3[e2_to_e1[q]] = (int32_t)t;
}

static int bits_for(uint64_t maxval) {
  int b = 1;
  while (b < 64 && (maxval >> b)) ++b;
  return b;
}

template <class K>
static int sort_pairs(K *keys_in, K *keys_out, int32_t *vals_in, int32_t *vals_out, int64_t E, int end_bit,
                      cudaStream_t st) {
  size_t bytes = 0;
  MRGCN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (int)E, 0,
                                             end_bit, st));
  TempBuf tmp(st);
  MRGCN_CUDA(tmp.alloc(bytes));
  MRGCN_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys_in, keys_out, vals_in, vals_out, (int)E, 0,
                                             end_bit, st));
  count_launch(4);
  return 0;
}

// ---- triples -> normalised stacked adjacency (graph_structure.py:70-108,162-169) ----------------
__global__ void k_triple_keys(const int32_t *__restrict__ tr, int64_t T, int32_t P, uint64_t *__restrict__ ksp,
                              uint64_t *__restrict__ kop) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint64_t s = tr[3 * t], p = tr[3 * t + 1], o = tr[3 * t + 2];
  ksp[t] = s * P + p;
  kop[t] = o * P + p;
}

__device__ __forceinline__ int64_t run_length(const uint64_t *__restrict__ sorted, int64_t T, uint64_t key) {
  int64_t lo = 0, hi = T;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] < key) lo = mid + 1; else hi = mid;
  }
  int64_t first = lo;
  hi = T;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] <= key) lo = mid + 1; else hi = mid;
  }
  return lo - first;
}

__global__ void k_adjacency(const int32_t *__restrict__ tr, int64_t T, int32_t N, int32_t P, int inv,
                            const uint64_t *__restrict__ sp_sorted, const uint64_t *__restrict__ op_sorted,
                            int64_t *__restrict__ row, int64_t *__restrict__ col, float *__restrict__ val) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t nt = inv ? 2 * T : T;
  if (t < T) {
    int64_t s = tr[3 * t], p = tr[3 * t + 1], o = tr[3 * t + 2];
    int64_t blk = inv ? 2 * p : p;
    // 1/rowsum in float64 (graph_structure.py:164-166) then the float32 cast of tarball.py:153-157
    int64_t d = run_length(sp_sorted, T, (uint64_t)(s * P + p));
    int64_t k = inv ? 2 * t : t;
    row[k] = s;
    col[k] = blk * N + o;
    val[k] = (float)(1.0 / (double)d);
    if (inv) {
      int64_t di = run_length(op_sorted, T, (uint64_t)(o * P + p));
      row[k + 1] = o;
      col[k + 1] = (blk + 1) * N + s;
      val[k + 1] = (float)(1.0 / (double)di);
    }
  } else if (t < T + N) {
    int64_t i = t - T;
    int64_t R = (inv ? 2 * (int64_t)P : (int64_t)P) + 1;
    row[nt + i] = i;
    col[nt + i] = (R - 1) * N + i;  // identity block appended last (graph_structure.py:33-35)
    val[nt + i] = 1.0f;
  }
}

}  // namespace
}  // namespace mrgcn

using namespace mrgcn;

extern "C" int mrgcn_version(void) { return 100; }
extern "C" const char *mrgcn_last_error_string(void) { return g_err; }
extern "C" int64_t mrgcn_launch_count(void) { return g_launches.load(); }

extern "C" void mrgcn_profile_enable(int on) { g_prof_on = on != 0; }

extern "C" int64_t mrgcn_profile_dump(char *buf, int64_t cap) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<int64_t, double>> acc;
  std::vector<std::string> order;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &r : g_prof) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
        auto it = acc.find(r.name);
        if (it == acc.end()) { order.push_back(r.name); it = acc.emplace(r.name, std::make_pair(0, 0.0)).first; }
        it->second.first += 1;
        it->second.second += ms;
      }
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    g_prof.clear();
  }
  std::string out;
  char line[256];
  for (auto &n : order) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", n.c_str(), (long long)acc[n].first, acc[n].second);
    out += line;
  }
  if (buf && cap > 0) {
    int64_t n = (int64_t)out.size() < cap - 1 ? (int64_t)out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

extern "C" int mrgcn_graph_build(const int64_t *coo_row, const int64_t *coo_col, const float *coo_val,
                                 int64_t E, int64_t nrows, int64_t ncols, int32_t R, mrgcn_graph *g,
                                 mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(g && R > 0 && ncols % R == 0 && E >= 0, MRGCN_E_BADARG, "graph_build: bad shape (ncols=%lld R=%d)",
                (long long)ncols, R);
  MRGCN_REQUIRE(E < (1ll << 31) - 1024 && nrows < (1ll << 31) && ncols / R < (1ll << 31), MRGCN_E_OVERFLOW,
                "graph_build: E/ND/NS exceed int32");
  MRGCN_REQUIRE((double)nrows * (double)ncols < 9.0e18, MRGCN_E_OVERFLOW, "graph_build: nrows*ncols exceeds 63 bits");
  int32_t NS = (int32_t)(ncols / R), ND = (int32_t)nrows;
  g->E = E; g->ND = ND; g->NS = NS; g->R = R;
  const int32_t slab_rows = g->slab_rows > 0 ? g->slab_rows : (NS > 0 ? NS : 1);
  const int64_t n_slabs = cdiv(NS > 0 ? NS : 1, slab_rows);
  const int64_t n_groups = n_slabs * R;
  MRGCN_REQUIRE(n_groups < (1ll << 31), MRGCN_E_OVERFLOW, "graph_build: too many (slab, relation) groups");
  const int T = 256;
  unsigned gridE = (unsigned)cdiv(E > 0 ? E : 1, T);

  TempBuf kA(st), kB(st), vA(st), vB(st), e1dst(st), e2to1(st), e3rel(st);
  MRGCN_CUDA(kA.alloc(E * 8)); MRGCN_CUDA(kB.alloc(E * 8));
  MRGCN_CUDA(vA.alloc(E * 4)); MRGCN_CUDA(vB.alloc(E * 4));
  MRGCN_CUDA(e1dst.alloc(E * 4)); MRGCN_CUDA(e2to1.alloc(E * 4)); MRGCN_CUDA(e3rel.alloc(E * 4));

  if (E > 0) {
    // E1: sort by (row, col) = (dst, rel, src)
    k_key1<<<gridE, T, 0, st>>>(coo_row, coo_col, ncols, kA.as<uint64_t>(), vA.as<int32_t>(), E);
    MRGCN_LAUNCH_CHECK();
    int b1 = bits_for((uint64_t)nrows * (uint64_t)ncols);
    if (int rc = sort_pairs(kA.as<uint64_t>(), kB.as<uint64_t>(), vA.as<int32_t>(), vB.as<int32_t>(), E, b1, st)) return rc;
    k_fill_e1<<<gridE, T, 0, st>>>(vB.as<int32_t>(), coo_row, coo_col, coo_val, NS, R, e1dst.as<int32_t>(), g->e1_src,
                                   g->e1_rel, g->e1_val, kA.as<uint64_t>(), vA.as<int32_t>(), E);
    MRGCN_LAUNCH_CHECK();
    // E2: stable sort of E1 by (src, rel) -> (src, rel, dst)
    int b2 = bits_for((uint64_t)NS * (uint64_t)R);
    if (int rc = sort_pairs(kA.as<uint64_t>(), kB.as<uint64_t>(), vA.as<int32_t>(), e2to1.as<int32_t>(), E, b2, st)) return rc;
    k_fill_e2<<<gridE, T, 0, st>>>(e2to1.as<int32_t>(), e1dst.as<int32_t>(), g->e1_src, g->e1_rel, g->e1_val, g->e2_src,
                                   g->e2_dst, g->e2_rel, g->e2_val, g->e1_to_e2, kA.as<uint32_t>(), vA.as<int32_t>(), E, slab_rows, R);
    MRGCN_LAUNCH_CHECK();
    // E3: stable sort of E2 by rel -> (rel, src, dst)
    int b3 = bits_for((uint64_t)n_groups);
    if (int rc = sort_pairs(kA.as<uint32_t>(), kB.as<uint32_t>(), vA.as<int32_t>(), vB.as<int32_t>(), E, b3, st)) return rc;
    k_fill_e3<<<gridE, T, 0, st>>>(vB.as<int32_t>(), e2to1.as<int32_t>(), g->e2_src, g->e2_dst, g->e2_rel, g->e2_val,
                                   g->e3_src, g->e3_dst, e3rel.as<int32_t>(), g->e3_val, g->e3_to_e2, g->e1_to_e3, g->e2_to_e3, E, slab_rows, R);
    MRGCN_LAUNCH_CHECK();
  }
  k_ptr<<<(unsigned)cdiv((int64_t)ND + 1, T), T, 0, st>>>(e1dst.as<int32_t>(), E, ND, g->rowptr);
  MRGCN_LAUNCH_CHECK();
  k_ptr<<<(unsigned)cdiv((int64_t)NS + 1, T), T, 0, st>>>(g->e2_src, E, NS, g->colptr);
  MRGCN_LAUNCH_CHECK();
  k_ptr<<<(unsigned)cdiv((int64_t)n_groups + 1, T), T, 0, st>>>(e3rel.as<int32_t>(), E, (int32_t)n_groups, g->relptr);
  MRGCN_LAUNCH_CHECK();
  return 0;
}

extern "C" int mrgcn_adjacency_from_triples(const int32_t *triples, int64_t T, int32_t N, int32_t P,
                                            int32_t include_inverse, int64_t *coo_row, int64_t *coo_col,
                                            float *coo_val, mrgcn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MRGCN_REQUIRE(T >= 0 && N > 0 && P > 0, MRGCN_E_BADARG, "adjacency_from_triples: bad sizes");
  MRGCN_REQUIRE(T < (1ll << 30), MRGCN_E_OVERFLOW, "adjacency_from_triples: too many triples");
  TempBuf ksp(st), kop(st), ssp(st), sop(st), tmp(st);
  MRGCN_CUDA(ksp.alloc(T * 8)); MRGCN_CUDA(kop.alloc(T * 8));
  MRGCN_CUDA(ssp.alloc(T * 8)); MRGCN_CUDA(sop.alloc(T * 8));
  const int B = 256;
  if (T > 0) {
    k_triple_keys<<<(unsigned)cdiv(T, B), B, 0, st>>>(triples, T, P, ksp.as<uint64_t>(), kop.as<uint64_t>());
    MRGCN_LAUNCH_CHECK();
    int bits = bits_for((uint64_t)N * (uint64_t)P);
    size_t bytes = 0;
    MRGCN_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, ksp.as<uint64_t>(), ssp.as<uint64_t>(), (int)T, 0, bits, st));
    MRGCN_CUDA(tmp.alloc(bytes));
    MRGCN_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, ksp.as<uint64_t>(), ssp.as<uint64_t>(), (int)T, 0, bits, st));
    MRGCN_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, kop.as<uint64_t>(), sop.as<uint64_t>(), (int)T, 0, bits, st));
    count_launch(8);
  }
  k_adjacency<<<(unsigned)cdiv(T + N, B), B, 0, st>>>(triples, T, N, P, include_inverse ? 1 : 0, ssp.as<uint64_t>(),
                                                       sop.as<uint64_t>(), coo_row, coo_col, coo_val);
  MRGCN_LAUNCH_CHECK();
  return 0;
}
