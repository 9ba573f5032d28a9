// Shared helpers for the b200tok kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b200tok.h"

// ---- host-side error plumbing -----------------------------------------------------------------
void b2t_set_error(const char* fmt, ...);
void b2t_count_launch(int n = 1);

#define B2T_REQUIRE(cond, code, ...)                 \
  do {                                               \
    if (!(cond)) {                                   \
      b2t_set_error(__VA_ARGS__);                    \
      return (code);                                 \
    }                                                \
  } while (0)

#define B2T_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      b2t_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return B2T_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define B2T_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      b2t_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return B2T_ERR_CUDA;                                                          \
    }                                                                               \
    b2t_count_launch();                                                             \
  } while (0)

int b2t_num_sms();
int b2t_arch_ok();   // B2T_OK or B2T_ERR_ARCH for the current device

// ---- device helpers ----------------------------------------------------------------------------
#define B2T_DEVICE __device__ __forceinline__

B2T_DEVICE float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <bool kRound>
B2T_DEVICE float r16(float x) {
  if constexpr (kRound) return bf16_round(x); else return x;
}

// fast-intrinsic forms (MUFU.EX2 + MUFU.RCP, ~2 ulp): the IEEE division path costs ~10x more in the
// GEMM epilogues, and every consumer rounds to bf16 or tolerates 1e-6
B2T_DEVICE float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
B2T_DEVICE float swishf_(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

B2T_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
B2T_DEVICE double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
B2T_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// last index i in [0, n) with off[i] <= x   (off is a non-decreasing prefix array of n+1 entries)
B2T_DEVICE int find_segment(const int32_t* __restrict__ off, int n, int x) {
  int lo = 0, hi = n;  // invariant: off[lo] <= x < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// typed activation load/store (bf16 or fp32 storage, fp32 maths)
template <typename T> B2T_DEVICE float ld_act(const T* p);
template <> B2T_DEVICE float ld_act<float>(const float* p) { return *p; }
template <> B2T_DEVICE float ld_act<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> B2T_DEVICE void st_act(T* p, float v);
template <> B2T_DEVICE void st_act<float>(float* p, float v) { *p = v; }
template <> B2T_DEVICE void st_act<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
