// FP32 FMA issue rate on sm_100a: packed fma.rn.f32x2 against scalar fma.rn.f32, register operands, 16 warps per SM
// (the shape of the depthwise-conv tap loop: 16 independent accumulator pairs, weights in registers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float seed) {
  float2 acc[16], w[8], v[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, seed);
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = make_float2(1.0f + 1e-6f * (i + seed), 1.0f - 1e-6f * i);
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = make_float2(1e-7f * (i + 1), seed * 1e-7f);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) {
          asm volatile("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\tmov.b64 c, {%4, %5};\n\t"
                       "fma.rn.f32x2 a, b, c, a;\n\tmov.b64 {%0, %1}, a;\n\t}"
                       : "+f"(acc[i].x), "+f"(acc[i].y) : "f"(w[(r + i) & 7].x), "f"(w[(r + i) & 7].y), "f"(v[r & 3].x), "f"(v[r & 3].y));
        } else {
          asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(w[(r + i) & 7].x), "f"(v[r & 3].x));
          asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(w[(r + i) & 7].y), "f"(v[r & 3].y));
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[gridDim.x * blockDim.x / 2 + 8] = t1 - t0;
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 512 * 4 + 1024);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      if (mode == 0) k<0><<<148, 512>>>(d, iters, 1.0f); else k<1><<<148, 512>>>(d, iters, 1.0f);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      long long clk; cudaMemcpy(&clk, (char*)d + (148 * 512 / 2 + 8) * 8, 8, cudaMemcpyDeviceToHost);
      const double fma = 2.0 * 16 * 8 * (double)iters * 512;          // per SM
      printf("%s: %.3f ms, %lld clk, %.1f FMA/clk/SM, %.1f TFLOP/s (148 SMs), clock %.2f GHz\n", mode == 0 ? "fma.rn.f32x2" : "fma.rn.f32  ",
             ms, clk, fma / clk, 2 * fma * 148 / (ms * 1e-3) / 1e12, clk / (ms * 1e-3) / 1e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
