// Fused GEMM epilogues shared by the SIMT and the tcgen05 kernels (see b2t_epilogue in b200tok.h).
#pragma once
#include <type_traits>
#include "common.cuh"

B2T_DEVICE uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

struct EpiParams {
  const float* bias;
  void* out;
  int ldo;
  float* resid;
  const uint8_t* row_valid;
  int M, N;
  float alpha;
  int round_resid;
};

// Store NV consecutive accumulator columns [col0, col0+NV) of one output row.
// kBF16: autocast rounding points active and `out` (where it is an activation) is bf16.
template <int EPI, bool kBF16, int NV>
B2T_DEVICE void epilogue_store(const EpiParams& p, int row, int col0, const float* acc) {
  static_assert(NV % 4 == 0, "NV must be a multiple of 4");
  using OutT = typename std::conditional<kBF16, __nv_bfloat16, float>::type;
  if (row >= p.M) return;
  float v[NV];
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
      v[i] = r16<kBF16>(acc[i] + b.x); v[i + 1] = r16<kBF16>(acc[i + 1] + b.y);
      v[i + 2] = r16<kBF16>(acc[i + 2] + b.z); v[i + 3] = r16<kBF16>(acc[i + 3] + b.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = r16<kBF16>(acc[i]);
  }
  if constexpr (EPI == B2T_EPI_BIAS || EPI == B2T_EPI_BIAS_SWISH || EPI == B2T_EPI_BIAS_GELU) {
    OutT* o = reinterpret_cast<OutT*>(p.out) + (size_t)row * p.ldo + col0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if constexpr (EPI == B2T_EPI_BIAS_SWISH) v[i] = swishf_(v[i]);
      if constexpr (EPI == B2T_EPI_BIAS_GELU) v[i] = geluf_(v[i]);
    }
    if constexpr (kBF16 && NV % 8 == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 8) {
        uint4 pk;
        pk.x = pack2_bf16(v[i], v[i + 1]); pk.y = pack2_bf16(v[i + 2], v[i + 3]);
        pk.z = pack2_bf16(v[i + 4], v[i + 5]); pk.w = pack2_bf16(v[i + 6], v[i + 7]);
        *reinterpret_cast<uint4*>(o + i) = pk;
      }
    } else if constexpr (kBF16) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        uint2 pk;
        pk.x = pack2_bf16(v[i], v[i + 1]); pk.y = pack2_bf16(v[i + 2], v[i + 3]);
        *reinterpret_cast<uint2*>(o + i) = pk;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; i += 4)
        *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  } else if constexpr (EPI == B2T_EPI_RESID) {
    float* o = p.resid + (size_t)row * p.N + col0;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 x = *reinterpret_cast<float4*>(o + i);
      x.x += p.alpha * v[i]; x.y += p.alpha * v[i + 1]; x.z += p.alpha * v[i + 2]; x.w += p.alpha * v[i + 3];
      if (kBF16 && p.round_resid) {
        x.x = bf16_round(x.x); x.y = bf16_round(x.y); x.z = bf16_round(x.z); x.w = bf16_round(x.w);
      }
      *reinterpret_cast<float4*>(o + i) = x;
    }
  } else if constexpr (EPI == B2T_EPI_GLU) {
    // columns interleaved (a_j, g_j): NV accumulators -> NV/2 outputs at column col0/2
    OutT* o = reinterpret_cast<OutT*>(p.out) + (size_t)row * p.ldo + (col0 >> 1);
    float g[NV / 2];
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) g[i] = v[2 * i] * sigmoidf_(v[2 * i + 1]);
    if constexpr (kBF16 && NV % 16 == 0) {
#pragma unroll
      for (int i = 0; i < NV / 2; i += 8) {
        uint4 pk;
        pk.x = pack2_bf16(g[i], g[i + 1]); pk.y = pack2_bf16(g[i + 2], g[i + 3]);
        pk.z = pack2_bf16(g[i + 4], g[i + 5]); pk.w = pack2_bf16(g[i + 6], g[i + 7]);
        *reinterpret_cast<uint4*>(o + i) = pk;
      }
    } else if constexpr (kBF16) {
#pragma unroll
      for (int i = 0; i < NV / 2; i += 2) *reinterpret_cast<uint32_t*>(o + i) = pack2_bf16(g[i], g[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < NV / 2; i += 2) *reinterpret_cast<float2*>(o + i) = make_float2(g[i], g[i + 1]);
    }
  } else if constexpr (EPI == B2T_EPI_BIAS_MASK) {
    float* o = p.resid + (size_t)row * p.N + col0;
    const bool valid = p.row_valid[row] != 0;
#pragma unroll
    for (int i = 0; i < NV; i += 4)
      *reinterpret_cast<float4*>(o + i) = valid ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3])
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
