// common.cuh -- shared host/device helpers of libetgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/etgpu.h"

#define ET_ABI_VERSION 2

// ---- error plumbing -------------------------------------------------------------------------
void et_set_error(const char *fmt, ...);

struct EtError {
  int code;
};

#define ET_FAIL(code, ...)        \
  do {                            \
    et_set_error(__VA_ARGS__);    \
    throw EtError{code};          \
  } while (0)

#define CUDA_CHECK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      int _c = (_e == cudaErrorMemoryAllocation) ? ET_ENOMEM : ET_ECUDA;                         \
      ET_FAIL(_c, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,        \
              cudaGetErrorString(_e));                                                           \
    }                                                                                            \
  } while (0)

// ---- device buffer that only grows ----------------------------------------------------------
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  // contents are NOT preserved
  void ensure(size_t n, double slack = 1.25) {
    if (n <= cap) return;
    release();
    size_t want = (size_t)((double)n * slack) + 16;
    cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = n;
      e = cudaMalloc((void **)&p, want * sizeof(T));
    }
    if (e != cudaSuccess) {
      p = nullptr;
      cudaGetLastError();
      ET_FAIL(ET_ENOMEM, "device allocation of %zu bytes failed", want * sizeof(T));
    }
    cap = want;
  }
  // contents preserved (device-to-device copy on `st`)
  void grow_keep(size_t n, size_t used, cudaStream_t st) {
    if (n <= cap) return;
    size_t want = (size_t)((double)n * 1.5) + 16;
    T *q = nullptr;
    cudaError_t e = cudaMalloc((void **)&q, want * sizeof(T));
    if (e != cudaSuccess) {
      cudaGetLastError();
      ET_FAIL(ET_ENOMEM, "device allocation of %zu bytes failed", want * sizeof(T));
    }
    if (p && used) {
      cudaMemcpyAsync(q, p, used * sizeof(T), cudaMemcpyDeviceToDevice, st);
      cudaStreamSynchronize(st);
    }
    if (p) cudaFree(p);
    p = q;
    cap = want;
  }
};

// ---- exact FP64 arithmetic shared by host and device ----------------------------------------
// The reference runs on the JVM: every operation is an individually rounded IEEE double op and
// a*b+c is never fused.  Device code uses the _rn intrinsics (never contracted); the host
// translation unit is compiled with -ffp-contract=off semantics (nvcc host pass: no contraction
// on x86-64 SSE2).
#ifdef __CUDA_ARCH__
#define ET_ADD(a, b) __dadd_rn((a), (b))
#define ET_SUB(a, b) __dsub_rn((a), (b))
#define ET_MUL(a, b) __dmul_rn((a), (b))
#define ET_DIV(a, b) __ddiv_rn((a), (b))
#else
static inline double et_host_add(double a, double b) {
  volatile double r = a + b;
  return r;
}
static inline double et_host_mul(double a, double b) {
  volatile double r = a * b;
  return r;
}
#define ET_ADD(a, b) et_host_add((a), (b))
#define ET_SUB(a, b) et_host_add((a), -(b))
#define ET_MUL(a, b) et_host_mul((a), (b))
#define ET_DIV(a, b) ((a) / (b))
#endif

#define ET_HD __host__ __device__ __forceinline__

ET_HD uint64_t et_d2u(double x) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
ET_HD double et_u2d(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}

// Result of `acc = 0.0; repeat h times: acc += c` in round-to-nearest-even FP64, for a positive
// normal c -- the reference's unweighted class distribution adds 1/s once per sample
// (pkg:905-911) instead of computing h/s, and the two differ in the last bits.  Inside one binade
// the rounded increment is a constant number of ulps (ties-to-even settles to the even choice
// after one step), so whole binades are jumped in O(1); the step that crosses a binade boundary
// is done with a real addition.
__host__ __device__ inline double et_repeat_add(double c, int64_t h) {
  if (h <= 0) return 0.0;
  double acc = c;  // 0.0 + c is exact
  h -= 1;
  if (h < 24) {
    while (h-- > 0) acc = ET_ADD(acc, c);
    return acc;
  }
  const uint64_t MANT = (1ULL << 52) - 1;
  const uint64_t cb = et_d2u(c);
  const int ec = (int)((cb >> 52) & 0x7ff);
  const uint64_t mc = (cb & MANT) | (1ULL << 52);
  while (h > 0) {
    acc = ET_ADD(acc, c);
    h -= 1;
    if (h == 0) break;
    uint64_t ab = et_d2u(acc);
    int ea = (int)((ab >> 52) & 0x7ff);
    uint64_t A = (ab & MANT) | (1ULL << 52);
    int s = ea - ec;  // >= 0 because acc >= c
    if (s >= 54) return acc;  // c < ulp/2: every further add is a no-op
    uint64_t inc;
    if (s == 0) {
      inc = mc;
    } else {
      uint64_t q = mc >> s;
      uint64_t rem = mc & ((1ULL << s) - 1);
      uint64_t half = 1ULL << (s - 1);
      if (rem < half)
        inc = q;
      else if (rem > half)
        inc = q + 1;
      else {
        if (A & 1ULL) continue;  // one more real add makes A even
        inc = (q & 1ULL) ? q + 1 : q;
      }
    }
    if (inc == 0) return acc;
    uint64_t room = ((1ULL << 53) - 1 - A) / inc;  // steps that stay strictly inside the binade
    uint64_t j = room < (uint64_t)h ? room : (uint64_t)h;
    if (j > 0) {
      A += j * inc;
      h -= (int64_t)j;
      acc = et_u2d(((uint64_t)ea << 52) | (A & MANT));
    }
  }
  return acc;
}

// ---- counter-based RNG of the free-running mode ---------------------------------------------
ET_HD uint64_t et_mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
ET_HD uint64_t et_tree_key(uint64_t seed, uint64_t tree_id) {
  return et_mix64(et_mix64(seed) ^ (tree_id * 0xd1342543de82ef95ULL + 0x2545f4914f6cdd1dULL));
}
ET_HD uint64_t et_child_key(uint64_t parent, uint32_t side) {
  return et_mix64(parent * 0x9fb21c651e98df25ULL + side + 1);
}
ET_HD uint64_t et_draw(uint64_t node_key, uint32_t counter) {
  return et_mix64(node_key ^ ((uint64_t)counter * 0xa0761d6478bd642fULL));
}
// same mapping as spire's nextDouble(): 53 random bits * 2^-53
ET_HD double et_u01(uint64_t bits) { return (double)(bits >> 11) * 1.1102230246251565e-16; }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
