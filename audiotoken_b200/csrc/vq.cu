// Nearest-centroid assignment (reference audiotoken/encoder.py:100-101 k-means `cdist`+`argmin`,
// :180 VectorQuantize eval forward) with an exactness guarantee:
//   1. fast pass   : scores A_k = x.c_k - 0.5|c_k|^2 for all k, never written to memory; per row only
//                    the best two candidates and the third-best score survive (fused epilogue).
//   2. finalize    : the two candidates are re-scored in fp64 (exact squared distance); if the
//                    third-best fast score is within 2*delta of the best, the row is not certified and
//                    is re-scanned: every centroid whose fast score is within 2*delta of the running
//                    maximum is re-scored in fp64.  delta bounds |A_k - exact|.
// Result == exact fp64 argmin, first index on ties.  Optional fused affine-free LayerNorm(1024)
// (reference encoder.py:175-176).
#include "common.cuh"

namespace {

struct Cand {
  float v1, v2, v3;   // best, second, third fast score
  int i1, i2;         // their indices
};

B2T_DEVICE void top3_insert(float v, int idx, float& v1, float& v2, float& v3, int& i1, int& i2) {
  if (v > v1) { v3 = v2; v2 = v1; i2 = i1; v1 = v; i1 = idx; }
  else if (v > v2) { v3 = v2; v2 = v; i2 = idx; }
  else if (v > v3) { v3 = v; }
}

__global__ void half_norm_kernel(const float* __restrict__ cb, int K, int D, float* __restrict__ hn) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  double acc = 0.0;
  for (int d = lane; d < D; d += 32) { double c = cb[(size_t)k * D + d]; acc += c * c; }
  acc = warp_sum(acc);
  if (lane == 0) hn[k] = (float)(0.5 * acc);
}

// affine-free LayerNorm over 1024, fp32 -> fp32 (warp per row)
__global__ void __launch_bounds__(256)
vq_ln_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int rows) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)r * ldx);
  float4 v[8];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = xr[lane + 32 * j]; sum += (v[j].x + v[j].y) + (v[j].z + v[j].w); }
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
  for (int j = 0; j < 8; ++j)
    reinterpret_cast<float4*>(y + (size_t)r * 1024)[lane + 32 * j] =
        make_float4(v[j].x * rstd, v[j].y * rstd, v[j].z * rstd, v[j].w * rstd);
}

// CUDA-core fast pass: block = 64 rows, loops over all centroids in tiles of 64.
__global__ void __launch_bounds__(256)
vq_scan_simt_kernel(const float* __restrict__ x, int ldx, int M, int D, const float* __restrict__ cb,
                    const float* __restrict__ hn, int K, Cand* __restrict__ cand) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  __shared__ Cand s_c[64][16];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * 64;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float v1[4], v2[4], v3[4];
  int i1[4], i2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { v1[i] = v2[i] = v3[i] = -INFINITY; i1[i] = i2[i] = 0; }

  for (int n0 = 0; n0 < K; n0 += 64) {
    float acc[4][4] = {};
    for (int k0 = 0; k0 < D; k0 += 16) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
      if (m0 + lr < M) a = *reinterpret_cast<const float4*>(x + (size_t)(m0 + lr) * ldx + k0 + lk);
      if (n0 + lr < K) w = *reinterpret_cast<const float4*>(cb + (size_t)(n0 + lr) * D + k0 + lk);
      As[lk][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
      Ws[lk][lr] = w.x; Ws[lk + 1][lr] = w.y; Ws[lk + 2][lr] = w.z; Ws[lk + 3][lr] = w.w;
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
    for (int j = 0; j < 4; ++j) {
      const int k = n0 + tx * 4 + j;
      if (k < K) {
        const float h = __ldg(hn + k);
#pragma unroll
        for (int i = 0; i < 4; ++i) top3_insert(acc[i][j] - h, k, v1[i], v2[i], v3[i], i1[i], i2[i]);
      }
    }
  }
  // merge the 16 per-thread lists of each row (ascending tx keeps the lowest index on ties)
#pragma unroll
  for (int i = 0; i < 4; ++i) s_c[ty * 4 + i][tx] = Cand{v1[i], v2[i], v3[i], i1[i], i2[i]};
  __syncthreads();
  if (tid < 64 && m0 + tid < M) {
    float b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
    int j1 = 0, j2 = 0;
    for (int t = 0; t < 16; ++t) {
      const Cand c = s_c[tid][t];
      // candidates arrive with strictly larger indices only within a thread; across threads compare
      // (value, -index) so that equal values keep the smaller index first
      auto ins = [&](float v, int idx) {
        if (v > b1 || (v == b1 && idx < j1)) { b3 = b2; b2 = b1; j2 = j1; b1 = v; j1 = idx; }
        else if (v > b2 || (v == b2 && idx < j2)) { b3 = b2; b2 = v; j2 = idx; }
        else if (v > b3) { b3 = v; }
      };
      if (c.v1 > -INFINITY) ins(c.v1, c.i1);
      if (c.v2 > -INFINITY) ins(c.v2, c.i2);
      if (c.v3 > b3) b3 = c.v3;
    }
    cand[m0 + tid] = Cand{b1, b2, b3, j1, j2};
  }
}

B2T_DEVICE double exact_dist(const float* __restrict__ xr, const float* __restrict__ c, int D, int lane) {
  double acc = 0.0;
  for (int d = lane; d < D; d += 32) { double t = (double)xr[d] - (double)c[d]; acc += t * t; }
  return warp_sum(acc);
}

// one warp per row
__global__ void __launch_bounds__(256)
vq_finalize_kernel(const float* __restrict__ x, int ldx, int M, int D, const float* __restrict__ cb,
                   const float* __restrict__ hn, int K, const Cand* __restrict__ cand, float rel_eps,
                   const float* __restrict__ cmax_half_ptr /* max_k 0.5|c_k|^2 */,
                   int16_t* __restrict__ out, int32_t* __restrict__ out32,
                   unsigned int* __restrict__ n_fallback) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= M) return;
  const float* xr = x + (size_t)r * ldx;
  const Cand c = cand[r];
  // |x|: bound on the fast-pass error  delta = rel_eps * (|x| |c|max + 0.5 |c|max^2)
  double xx = 0.0;
  for (int d = lane; d < D; d += 32) xx += (double)xr[d] * (double)xr[d];
  xx = warp_sum(xx);
  const float cmax_half = *cmax_half_ptr;
  const float cmax = sqrtf(2.0f * cmax_half);
  const float delta = rel_eps * ((float)sqrt(xx) * cmax + cmax_half);
  int best;
  if (K == 1) {
    best = 0;
  } else if (!(c.v3 >= c.v1 - 2.0f * delta)) {
    const double d1 = exact_dist(xr, cb + (size_t)c.i1 * D, D, lane);
    const double d2 = exact_dist(xr, cb + (size_t)c.i2 * D, D, lane);
    best = (d2 < d1 || (d2 == d1 && c.i2 < c.i1)) ? c.i2 : c.i1;
  } else {
    if (lane == 0 && n_fallback) atomicAdd(n_fallback, 1u);
    // re-scan: lane-strided centroids, fp32 filter, fp64 re-score of everything near the running max
    float run = -INFINITY;
    double bd = INFINITY;
    int bi = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
      const float* ck = cb + (size_t)k * D;
      float a = 0.f;
      for (int d = 0; d < D; d += 4) {
        float4 xv = *reinterpret_cast<const float4*>(xr + d), cv = *reinterpret_cast<const float4*>(ck + d);
        a = fmaf(xv.x, cv.x, a); a = fmaf(xv.y, cv.y, a); a = fmaf(xv.z, cv.z, a); a = fmaf(xv.w, cv.w, a);
      }
      a -= __ldg(hn + k);
      if (a >= run - 2.0f * delta) {
        double dd = 0.0;
        for (int d = 0; d < D; ++d) { double t = (double)xr[d] - (double)ck[d]; dd += t * t; }
        if (dd < bd || (dd == bd && k < bi)) { bd = dd; bi = k; }
      }
      run = fmaxf(run, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double od = __shfl_xor_sync(0xffffffffu, bd, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    best = bi;
  }
  if (lane == 0) {
    if (out) out[r] = (int16_t)best;
    if (out32) out32[r] = best;
  }
}

__global__ void max_reduce_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  __shared__ float s[32];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, v[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : -INFINITY;
    m = warp_max(m);
    if (threadIdx.x == 0) *out = m;
  }
}

}  // namespace

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t b2t_vq_workspace_bytes(int rows, int dim, int codebook_size) {
  size_t b = 0;
  b += align_up((size_t)rows * dim * 4, 256);          // LayerNormed rows
  b += align_up((size_t)rows * sizeof(Cand), 256);     // candidates
  b += align_up((size_t)codebook_size * 4, 256);       // half norms
  b += 256;                                            // cmax_half, fallback counter
  return b;
}

extern "C" int b2t_vq_argmin(const float* x, int ldx, int rows, int dim, const float* codebook,
                             const float* half_norm, int codebook_size, int apply_ln, int16_t* out,
                             int32_t* out_i32, void* workspace, size_t workspace_bytes, void* stream) {
  B2T_REQUIRE(x && codebook && (out || out_i32) && workspace, B2T_ERR_ARG, "b2t_vq_argmin: null argument");
  B2T_REQUIRE(dim % 32 == 0 && dim >= 32 && ldx % 4 == 0 && ldx >= dim, B2T_ERR_ARG,
              "b2t_vq_argmin: dim must be a multiple of 32 and ldx a multiple of 4 (dim=%d ldx=%d)", dim, ldx);
  B2T_REQUIRE(codebook_size >= 1 && codebook_size <= 32768, B2T_ERR_ARG,
              "b2t_vq_argmin: codebook_size must be in [1, 32768] for int16 tokens (got %d)", codebook_size);
  B2T_REQUIRE(!apply_ln || dim == 1024, B2T_ERR_ARG, "b2t_vq_argmin: fused LayerNorm needs dim == 1024");
  B2T_REQUIRE(workspace_bytes >= b2t_vq_workspace_bytes(rows, dim, codebook_size), B2T_ERR_WORKSPACE,
              "b2t_vq_argmin: workspace too small");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (rows <= 0) return B2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  float* xn = (float*)ws;                 ws += align_up((size_t)rows * dim * 4, 256);
  Cand* cand = (Cand*)ws;                 ws += align_up((size_t)rows * sizeof(Cand), 256);
  float* hn = (float*)ws;                 ws += align_up((size_t)codebook_size * 4, 256);
  float* cmax = (float*)ws;
  unsigned int* nfb = (unsigned int*)(ws + 16);

  const float* xs = x;
  int lds = ldx;
  if (apply_ln) {
    vq_ln_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, ldx, xn, rows);
    B2T_LAUNCH_CHECK();
    xs = xn; lds = 1024;
  }
  if (half_norm == nullptr) {
    half_norm_kernel<<<(codebook_size + 7) / 8, 256, 0, st>>>(codebook, codebook_size, dim, hn);
    B2T_LAUNCH_CHECK();
  } else {
    B2T_CUDA(cudaMemcpyAsync(hn, half_norm, (size_t)codebook_size * 4, cudaMemcpyDeviceToDevice, st));
  }
  max_reduce_kernel<<<1, 1024, 0, st>>>(hn, codebook_size, cmax);
  B2T_LAUNCH_CHECK();
  B2T_CUDA(cudaMemsetAsync(nfb, 0, 4, st));
  vq_scan_simt_kernel<<<(rows + 63) / 64, 256, 0, st>>>(xs, lds, rows, dim, codebook, hn, codebook_size, cand);
  B2T_LAUNCH_CHECK();
  // fp32 FMA dot of length D: |err| <= ~D * 2^-24 * sum|x_d c_d| <= D*2^-24 * |x||c|; 8x margin
  const float rel_eps = 8.0f * (float)dim * 5.9604645e-8f;
  vq_finalize_kernel<<<(rows + 7) / 8, 256, 0, st>>>(xs, lds, rows, dim, codebook, hn, codebook_size, cand,
                                                        rel_eps, cmax, out, out_i32, nfb);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
