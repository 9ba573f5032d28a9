// CUDA-core GEMM with the fused epilogues: out = A[M,K] . W[N,K]^T.
// Used (a) for B2T_PREC_FP32 — the high-precision mode that demonstrates <=1e-4 embedding error
// against the fp32 oracle — and (b) as the on-device cross-check of the tcgen05 kernel.
// 64x64x16 tiles, 256 threads, 4x4 outputs per thread, fp32 FMA accumulation.
#include "gemm_epilogue.cuh"

namespace {

template <typename T> B2T_DEVICE void load4(const T* p, float (&v)[4]);
template <> B2T_DEVICE void load4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> B2T_DEVICE void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}

template <typename T, int EPI, bool kBF16>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W, int K, EpiParams p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int lr = tid >> 2, lk = (tid & 3) * 4;   // loader: row 0..63, k offset 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, w[4];
    if (m0 + lr < p.M) load4<T>(A + (size_t)(m0 + lr) * lda + k0 + lk, a);
    load4<T>(W + (size_t)(n0 + lr) * K + k0 + lk, w);
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[lk + i][lr] = a[i]; Ws[lk + i][lr] = w[i]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) epilogue_store<EPI, kBF16, 4>(p, m0 + ty * 4 + i, n0 + tx * 4, acc[i]);
}

template <typename T, bool kBF16>
int launch_simt(const b2t_gemm_args* a, const EpiParams& p, cudaStream_t st) {
  dim3 grid(a->N / 64, (a->M + 63) / 64);
  const T* A = (const T*)a->A;
  const T* W = (const T*)a->W;
  switch (a->epilogue) {
    case B2T_EPI_BIAS: gemm_simt_kernel<T, B2T_EPI_BIAS, kBF16><<<grid, 256, 0, st>>>(A, a->lda, W, a->K, p); break;
    case B2T_EPI_BIAS_SWISH: gemm_simt_kernel<T, B2T_EPI_BIAS_SWISH, kBF16><<<grid, 256, 0, st>>>(A, a->lda, W, a->K, p); break;
    case B2T_EPI_RESID: gemm_simt_kernel<T, B2T_EPI_RESID, kBF16><<<grid, 256, 0, st>>>(A, a->lda, W, a->K, p); break;
    case B2T_EPI_GLU: gemm_simt_kernel<T, B2T_EPI_GLU, kBF16><<<grid, 256, 0, st>>>(A, a->lda, W, a->K, p); break;
    case B2T_EPI_BIAS_MASK: gemm_simt_kernel<T, B2T_EPI_BIAS_MASK, kBF16><<<grid, 256, 0, st>>>(A, a->lda, W, a->K, p); break;
    case B2T_EPI_BIAS_GELU: gemm_simt_kernel<T, B2T_EPI_BIAS_GELU, kBF16><<<grid, 256, 0, st>>>(A, a->lda, W, a->K, p); break;
    default: b2t_set_error("b2t_gemm: unknown epilogue %d", a->epilogue); return B2T_ERR_ARG;
  }
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}

}  // namespace

int b2t_gemm_tensor(const b2t_gemm_args* a, const EpiParams& p, cudaStream_t st);  // gemm_tc.cu

extern "C" int b2t_gemm(const b2t_gemm_args* a, void* stream) {
  B2T_REQUIRE(a && a->A && a->W, B2T_ERR_ARG, "b2t_gemm: null argument");
  B2T_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, B2T_ERR_ARG, "b2t_gemm: bad shape");
  B2T_REQUIRE(a->N % 64 == 0 && a->K % 16 == 0 && a->lda % 8 == 0, B2T_ERR_ARG,
              "b2t_gemm: N%%64, K%%16, lda%%8 must be 0 (N=%d K=%d lda=%d)", a->N, a->K, a->lda);
  const int e = a->epilogue;
  B2T_REQUIRE(e >= B2T_EPI_BIAS && e <= B2T_EPI_BIAS_GELU, B2T_ERR_ARG, "b2t_gemm: unknown epilogue %d", e);
  if (e == B2T_EPI_RESID || e == B2T_EPI_BIAS_MASK)
    B2T_REQUIRE(a->resid, B2T_ERR_ARG, "b2t_gemm: epilogue %d needs resid", e);
  else
    B2T_REQUIRE(a->out && a->ldo % 8 == 0, B2T_ERR_ARG, "b2t_gemm: epilogue %d needs out with ldo%%8==0", e);
  if (e == B2T_EPI_BIAS_MASK) B2T_REQUIRE(a->row_valid, B2T_ERR_ARG, "b2t_gemm: BIAS_MASK needs row_valid");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (a->M == 0) return B2T_OK;
  EpiParams p{a->bias, a->out, a->ldo, a->resid, a->row_valid, a->M, a->N, a->alpha, a->round_resid_bf16};
  cudaStream_t st = (cudaStream_t)stream;
  if (a->precision == B2T_PREC_FP32) {
    B2T_REQUIRE(a->impl != B2T_IMPL_TENSOR, B2T_ERR_ARG, "b2t_gemm: the tensor path is bf16 only");
    return launch_simt<float, false>(a, p, st);
  }
  if (a->impl == B2T_IMPL_SIMT) return launch_simt<__nv_bfloat16, true>(a, p, st);
  return b2t_gemm_tensor(a, p, st);
}
