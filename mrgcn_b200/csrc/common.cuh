// Shared helpers for libmrgcn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mrgcn_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmrgcn_b200 is written for sm_100a (B200) only"
#endif

namespace mrgcn {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// optional per-kernel timing (CUDA events on the launching stream), see mrgcn_profile_enable
void prof_begin(const char *name, cudaStream_t st);
void prof_end();
#define MRGCN_PROF(name) mrgcn::prof_begin(name, st)

#define MRGCN_CUDA(expr)                                                                 \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      mrgcn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

#define MRGCN_LAUNCH_CHECK()                                                             \
  do {                                                                                   \
    mrgcn::prof_end();                                                                   \
    mrgcn::count_launch();                                                               \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      mrgcn::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

#define MRGCN_REQUIRE(cond, code, ...)                                                   \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      mrgcn::set_error(__VA_ARGS__);                                                     \
      return (code);                                                                     \
    }                                                                                    \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// streaming (read-once) loads: keep them out of L1 so gathered operands stay resident
__device__ __forceinline__ float ldg_stream(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_stream(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// Packed fp32 FMA (Blackwell FFMA2): two independent round-to-nearest FMAs per instruction -> half the issue slots
// of the scalar form, bit-identical results.  acc.{x,y} += s * w.{x,y}
__device__ __forceinline__ void fma2(float2 &acc, float s, float2 w) { acc = __ffma2_rn(make_float2(s, s), w, acc); }
__device__ __forceinline__ void fma2v(float2 &acc, float2 a, float2 b) { acc = __ffma2_rn(a, b, acc); }

}  // namespace mrgcn
