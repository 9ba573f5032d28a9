// Shared helpers for the b200tok kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <assert.h>
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
    if (b2t_debug_sync() && b2t_debug_sync_check(__FILE__, __LINE__) != B2T_OK) return B2T_ERR_CUDA; \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device attribute: opt in once per (call site, device).
#define B2T_SMEM_OPT_IN(bytes, ...)                                                                     \
  do {                                                                                                  \
    static bool done__[B2T_MAX_DEVICES] = {};                                                           \
    const int dev__ = b2t_device_index();                                                               \
    if (!done__[dev__]) {                                                                               \
      B2T_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      done__[dev__] = true;                                                                             \
    }                                                                                                   \
  } while (0)

// B2T_DEBUG_SYNC=1 (environment) or b2t_set_option("debug_sync", 1): synchronise the device after every launch and
// name the launch site whose kernel faulted (an asynchronous fault otherwise surfaces at some later, unrelated call).
bool b2t_debug_sync();
int b2t_debug_sync_check(const char* file, int line);
void b2t_set_debug_sync(int on);

int b2t_num_sms();          // of the current device
unsigned* b2t_trap_rec();   // mapped host record for b2t_trap_record (same pointer on host and device), or null
int b2t_device_index();     // current device, clamped to [0, B2T_MAX_DEVICES)
constexpr int B2T_MAX_DEVICES = 64;
int b2t_arch_ok();   // B2T_OK or B2T_ERR_ARCH for the current device

// ---- device helpers ----------------------------------------------------------------------------
#define B2T_DEVICE __device__ __forceinline__

// A protocol wait that ran out of its bound: say where (device printf + device assert, both reported by the host's
// next synchronisation as cudaErrorAssert) instead of a bare trap that only leaves "unspecified launch failure".
__device__ __noinline__ inline void b2t_trap_report(const char* what, unsigned a, unsigned b) {
  printf("b200tok device trap: %s a=0x%x b=%u block=(%d,%d,%d) thread=%d\n", what, a, b, (int)blockIdx.x, (int)blockIdx.y,
         (int)blockIdx.z, (int)threadIdx.x);
  __assert_fail(what, __FILE__, __LINE__, "b2t_trap_report");
}

// The same for register-critical code (the single-pass attention kernels re-partition registers with setmaxnreg; an
// ABI call such as the printf above makes ptxas spill the whole region around it): the site is recorded with plain
// stores into a 32-byte record in MAPPED HOST memory (b2t_trap_rec(), passed as a kernel argument), which survives the
// trap; the host reads it back through b2t_last_device_trap.  rec = {code, a, b, blockIdx.x, blockIdx.y, threadIdx.x}.
B2T_DEVICE void b2t_trap_record(unsigned* rec, unsigned code, unsigned a, unsigned b) {
  if (rec != nullptr) {
    volatile unsigned* r = rec;
    r[1] = a; r[2] = b; r[3] = blockIdx.x; r[4] = blockIdx.y; r[5] = threadIdx.x;
    __threadfence_system();
    r[0] = code;
    __threadfence_system();
  }
  __trap();
}

B2T_DEVICE float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <bool kRound>
B2T_DEVICE float r16(float x) {
  if constexpr (kRound) return bf16_round(x); else return x;
}

// fast-intrinsic forms (MUFU.EX2 + MUFU.RCP, ~2 ulp): the IEEE division path costs ~10x more in the
// GEMM epilogues, and every consumer rounds to bf16 or tolerates 1e-6
B2T_DEVICE float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
B2T_DEVICE float swishf_(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// erf-form GELU, 0.5 x (1 + erf(x / sqrt 2)), through Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, i.e. an absolute
// error below 1e-7 |x| on the result): two MUFU ops + 8 FMA-pipe instructions instead of the ~30 of erff(), which made the
// conv0 / conv-GEMM epilogues of the mHuBERT path instruction-bound.  The negative branch uses the tail directly
// (1 + erf(-z) = poly * exp(-z^2)), so there is no cancellation.
B2T_DEVICE float geluf_(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float y = p * t * __expf(-z * z);          // 1 - erf(z)
  return 0.5f * x * (x < 0.f ? y : 2.0f - y);
}

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
