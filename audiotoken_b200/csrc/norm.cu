// Row LayerNorm over 1024 channels (HF Wav2Vec2Bert LayerNorms, eps 1e-5; affine-free variant of
// reference audiotoken/encoder.py:138-144).  One warp per row, the row lives in registers
// (8 x float4 per lane), two-pass mean / variance, 128-bit loads and stores.  HBM-bound:
// 4 KB read + 2 KB (bf16) or 4 KB (fp32) written per row.
#include "common.cuh"

namespace {

template <typename OutT, bool kAffine>
__global__ void __launch_bounds__(256)
layernorm1024_kernel(const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ b, const uint8_t* __restrict__ row_valid,
                     OutT* __restrict__ out, int rows) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)r * 1024);
  float4 v[8];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = xr[lane + 32 * j];
    sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  sum = warp_sum(sum);
  const float mu = sum * (1.0f / 1024.f);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j].x -= mu; v[j].y -= mu; v[j].z -= mu; v[j].w -= mu;
    sq += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  sq = warp_sum(sq);
  float rstd = rsqrtf(sq * (1.0f / 1024.f) + 1e-5f);
  const bool zero = row_valid != nullptr && row_valid[r] == 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c4 = lane + 32 * j;
    float4 y;
    if constexpr (kAffine) {
      float4 g = reinterpret_cast<const float4*>(w)[c4];
      float4 be = reinterpret_cast<const float4*>(b)[c4];
      y.x = v[j].x * rstd * g.x + be.x; y.y = v[j].y * rstd * g.y + be.y;
      y.z = v[j].z * rstd * g.z + be.z; y.w = v[j].w * rstd * g.w + be.w;
    } else {
      y.x = v[j].x * rstd; y.y = v[j].y * rstd; y.z = v[j].z * rstd; y.w = v[j].w * rstd;
    }
    if (zero) y = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (sizeof(OutT) == 4) {
      reinterpret_cast<float4*>(out + (size_t)r * 1024)[c4] = y;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(out + (size_t)r * 1024)[c4] = pk;
    }
  }
}

// Fused residual update + LayerNorm (one warp per row, the row stays in registers):
//   t = x + alpha * d            (d = the bf16/fp32 output of the preceding Linear; optional bf16 rounding of t:
//                                 the autocast path keeps a bf16 stream inside layer 0)
//   kChain == false:  x <- t;               out <- LN(t; w1, b1)  [rows with row_valid == 0 -> 0]
//   kChain == true :  x <- LN(t; w1, b1);   out <- LN(x; w2, b2)  (final_layer_norm followed by the next
//                                                                   layer's ffn1_layer_norm; out may be NULL)
// This replaces a read-modify-write GEMM epilogue plus a separate LayerNorm pass: 6 KB read + 6 KB written
// per row instead of 8 + 6.
template <typename ActT, bool kChain>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(float* __restrict__ x, const ActT* __restrict__ d, float alpha, int round_bf16,
                     const float* __restrict__ w1, const float* __restrict__ b1,
                     const float* __restrict__ w2, const float* __restrict__ b2,
                     const uint8_t* __restrict__ row_valid, ActT* __restrict__ out, int rows) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float4* xr = reinterpret_cast<float4*>(x + (size_t)r * 1024);
  float4 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = xr[lane + 32 * j];
  if (d != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 dd;
      if constexpr (sizeof(ActT) == 4) {
        dd = reinterpret_cast<const float4*>(d + (size_t)r * 1024)[lane + 32 * j];
      } else {
        const uint2 raw = reinterpret_cast<const uint2*>(d + (size_t)r * 1024)[lane + 32 * j];
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
        dd = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
      }
      v[j].x += alpha * dd.x; v[j].y += alpha * dd.y; v[j].z += alpha * dd.z; v[j].w += alpha * dd.w;
      if (round_bf16) { v[j].x = bf16_round(v[j].x); v[j].y = bf16_round(v[j].y); v[j].z = bf16_round(v[j].z); v[j].w = bf16_round(v[j].w); }
    }
  }
  if constexpr (!kChain) {
    if (d != nullptr) {                      // d == NULL: the producing GEMM's epilogue has already updated x
#pragma unroll
      for (int j = 0; j < 8; ++j) xr[lane + 32 * j] = v[j];
    }
  }
  auto normalise = [&](const float* w, const float* b) {
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    sum = warp_sum(sum);
    const float mu = sum * (1.0f / 1024.f);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j].x -= mu; v[j].y -= mu; v[j].z -= mu; v[j].w -= mu;
      sq += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    sq = warp_sum(sq);
    const float rstd = rsqrtf(sq * (1.0f / 1024.f) + 1e-5f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 g = reinterpret_cast<const float4*>(w)[lane + 32 * j];
      const float4 be = reinterpret_cast<const float4*>(b)[lane + 32 * j];
      v[j].x = v[j].x * rstd * g.x + be.x; v[j].y = v[j].y * rstd * g.y + be.y;
      v[j].z = v[j].z * rstd * g.z + be.z; v[j].w = v[j].w * rstd * g.w + be.w;
    }
  };
  normalise(w1, b1);
  if constexpr (kChain) {
#pragma unroll
    for (int j = 0; j < 8; ++j) xr[lane + 32 * j] = v[j];
    if (out == nullptr) return;
    normalise(w2, b2);
  }
  const bool zero = row_valid != nullptr && row_valid[r] == 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 y = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : v[j];
    if constexpr (sizeof(ActT) == 4) {
      reinterpret_cast<float4*>(out + (size_t)r * 1024)[lane + 32 * j] = y;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(out + (size_t)r * 1024)[lane + 32 * j] = pk;
    }
  }
}

}  // namespace

extern "C" int b2t_add_layernorm(float* x, const void* delta, float alpha, int round_x_bf16, const float* w1,
                                 const float* b1, const float* w2, const float* b2, const uint8_t* row_valid,
                                 void* out, int rows, int precision, void* stream) {
  B2T_REQUIRE(x && w1 && b1, B2T_ERR_ARG, "b2t_add_layernorm: null argument");
  B2T_REQUIRE((w2 == nullptr) == (b2 == nullptr), B2T_ERR_ARG, "b2t_add_layernorm: w2 and b2 go together");
  B2T_REQUIRE(w2 == nullptr || out != nullptr, B2T_ERR_ARG, "b2t_add_layernorm: the chained LayerNorm needs out");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (rows <= 0) return B2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (rows + 7) / 8;
  const bool chain = w2 != nullptr || out == nullptr;
  if (precision == B2T_PREC_BF16) {
    auto* dd = (const __nv_bfloat16*)delta; auto* oo = (__nv_bfloat16*)out;
    if (chain) add_layernorm_kernel<__nv_bfloat16, true><<<blocks, 256, 0, st>>>(x, dd, alpha, round_x_bf16, w1, b1, w2, b2, row_valid, oo, rows);
    else add_layernorm_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>(x, dd, alpha, round_x_bf16, w1, b1, w2, b2, row_valid, oo, rows);
  } else {
    auto* dd = (const float*)delta; auto* oo = (float*)out;
    if (chain) add_layernorm_kernel<float, true><<<blocks, 256, 0, st>>>(x, dd, alpha, 0, w1, b1, w2, b2, row_valid, oo, rows);
    else add_layernorm_kernel<float, false><<<blocks, 256, 0, st>>>(x, dd, alpha, 0, w1, b1, w2, b2, row_valid, oo, rows);
  }
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}

extern "C" int b2t_layernorm(const float* x, const float* weight, const float* bias,
                             const uint8_t* row_valid, void* out, int rows, int cols,
                             int out_precision, void* stream) {
  B2T_REQUIRE(x && out, B2T_ERR_ARG, "b2t_layernorm: null argument");
  B2T_REQUIRE(cols == 1024, B2T_ERR_ARG, "b2t_layernorm: cols must be 1024 (got %d)", cols);
  B2T_REQUIRE((weight == nullptr) == (bias == nullptr), B2T_ERR_ARG,
              "b2t_layernorm: weight and bias must both be given or both be NULL");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (rows <= 0) return B2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (rows + 7) / 8;
  const bool bf = out_precision == B2T_PREC_BF16;
  if (weight) {
    if (bf) layernorm1024_kernel<__nv_bfloat16, true><<<blocks, 256, 0, st>>>(x, weight, bias, row_valid, (__nv_bfloat16*)out, rows);
    else layernorm1024_kernel<float, true><<<blocks, 256, 0, st>>>(x, weight, bias, row_valid, (float*)out, rows);
  } else {
    if (bf) layernorm1024_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>(x, nullptr, nullptr, row_valid, (__nv_bfloat16*)out, rows);
    else layernorm1024_kernel<float, false><<<blocks, 256, 0, st>>>(x, nullptr, nullptr, row_valid, (float*)out, rows);
  }
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
