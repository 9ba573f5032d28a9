// Internal interface between acoustic.cu (C ABI, model registry) and seanet_tc.cu (tensor-core encoder).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200tok.h"

struct SeanetTcWeights {
  const float* conv0_w; const float* conv0_b;                                   // fp32 [32][16], [32]
  const __nv_bfloat16* k3_w[4]; const float* k3_b[4]; int k3_kpad[4];           // [C/2][pad64(3C)]
  const __nv_bfloat16* res_w[4]; const float* res_b[4]; int res_kpad[4];        // [C][pad64(1.5C)] = [shortcut | k1]
  const __nv_bfloat16* down_w[4]; const float* down_b[4]; int down_kpad[4];     // [2C][2sC]
  const __nv_bfloat16* lstm_w[2]; const float* lstm_b[2];                       // [2048][1024] gate-interleaved rows
  const __nv_bfloat16* final_w; const float* final_b;                           // [128][3584]
};

size_t b2t_seanet_tc_workspace_bytes(const b2t_acoustic_batch* b);
int b2t_seanet_tc_encode(const SeanetTcWeights& wt, const float* wave, const b2t_acoustic_batch* b, void* workspace,
                         size_t workspace_bytes, float* emb, const int32_t* active_host, cudaStream_t st);
float* b2t_seanet_tc_emb(void* workspace, const b2t_acoustic_batch* b);
