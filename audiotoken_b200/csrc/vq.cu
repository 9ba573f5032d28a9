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
#include <string.h>
#include "vq_cand.cuh"

namespace {

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
  Cand loc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) loc[i] = cand_empty();

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
        for (int i = 0; i < 4; ++i) cand_insert_ordered(loc[i], acc[i][j] - h, k);
      }
    }
  }
  // merge the 16 per-thread lists of each row
#pragma unroll
  for (int i = 0; i < 4; ++i) s_c[ty * 4 + i][tx] = loc[i];
  __syncthreads();
  if (tid < 64 && m0 + tid < M) {
    Cand best = cand_empty();
    for (int t = 0; t < 16; ++t) cand_merge(best, s_c[tid][t]);
    cand[m0 + tid] = best;
  }
}

// x (optionally LayerNormed) -> error-compensated bf16 pair: hi = bf16(x), lo = bf16(x - hi); [rows, 2*D]
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ x, int ldx, int rows, int D, __nv_bfloat16* __restrict__ out) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = x[(size_t)r * ldx + d];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[(size_t)r * 2 * D + d] = hi;
    out[(size_t)r * 2 * D + D + d] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

__global__ void fill_inf_kernel(float* __restrict__ p, int from, int to) {
  const int i = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < to) p[i] = INFINITY;
}

B2T_DEVICE double exact_dist(const float* __restrict__ xr, const float* __restrict__ c, int D, int lane) {
  double acc = 0.0;
  for (int d = lane; d < D; d += 32) { double t = (double)xr[d] - (double)c[d]; acc += t * t; }
  return warp_sum(acc);
}

// one warp per row
__global__ void __launch_bounds__(256)
vq_finalize_kernel(const float* __restrict__ x, int ldx, int M, int D, const float* __restrict__ cb,
                   const float* __restrict__ hn, int K, const Cand* __restrict__ cand, int slices, float rel_eps,
                   const float* __restrict__ cmax_half_ptr /* max_k 0.5|c_k|^2 */,
                   int16_t* __restrict__ out, int32_t* __restrict__ out32,
                   unsigned int* __restrict__ n_fallback, unsigned int* __restrict__ max_err,
                   int* __restrict__ rescan_rows, float* __restrict__ rescan_thr, int rescan_cap) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= M) return;
  const float* xr = x + (size_t)r * ldx;
  // merge the per-slice records of this row (lane-strided, then a shuffle tree)
  Cand c = cand_empty();
  for (int sidx = lane; sidx < slices; sidx += 32) cand_merge(c, cand[(size_t)r * slices + sidx]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Cand other;
    other.v1 = __shfl_xor_sync(0xffffffffu, c.v1, o); other.v2 = __shfl_xor_sync(0xffffffffu, c.v2, o);
    other.v3 = __shfl_xor_sync(0xffffffffu, c.v3, o); other.v4 = __shfl_xor_sync(0xffffffffu, c.v4, o);
    other.i1 = __shfl_xor_sync(0xffffffffu, c.i1, o); other.i2 = __shfl_xor_sync(0xffffffffu, c.i2, o);
    other.i3 = __shfl_xor_sync(0xffffffffu, c.i3, o);
    cand_merge(c, other);
  }
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
  } else if (!(c.v4 >= c.v1 - 2.0f * delta) && c.i1 < K) {
    // certified: only the three candidates can be the exact winner; skip those the bound already excludes
    const double d1 = exact_dist(xr, cb + (size_t)c.i1 * D, D, lane);
    best = c.i1;
    double bd = d1;
    if (c.i2 < K && c.v2 >= c.v1 - 2.0f * delta) {
      const double d2 = exact_dist(xr, cb + (size_t)c.i2 * D, D, lane);
      if (d2 < bd || (d2 == bd && c.i2 < best)) { bd = d2; best = c.i2; }
    }
    if (c.i3 < K && c.v3 >= c.v1 - 2.0f * delta) {
      const double d3 = exact_dist(xr, cb + (size_t)c.i3 * D, D, lane);
      if (d3 < bd || (d3 == bd && c.i3 < best)) { bd = d3; best = c.i3; }
    }
    if (lane == 0 && max_err) {
      // observed fast-pass error on the best candidate, in units of the bound's scale
      const float obs = fabsf(c.v1 - (float)(0.5 * (xx - d1))) / ((float)sqrt(xx) * cmax + cmax_half);
      atomicMax(max_err, __float_as_uint(obs));
    }
  } else {
    // not certified: four or more fast scores lie within the bound.  One warp re-scanning K centroids is a latency
    // chain of milliseconds that the whole grid then waits for (measured: 1-2 such rows per 65 536-row batch made
    // this kernel 5x longer), so the row is only queued here; vq_rescan_kernel resolves the queue with a CTA per row.
    if (lane == 0) {
      const unsigned int slot = atomicAdd(n_fallback, 1u);
      if (slot < (unsigned int)rescan_cap) {
        rescan_rows[slot] = r;
        rescan_thr[2 * slot] = c.v1;                                           // best fast score
        rescan_thr[2 * slot + 1] = (float)sqrt(xx) * cmax + cmax_half;         // scale of the error bounds
      }
    }
    return;
  }
  if (lane == 0) {
    if (out) out[r] = (int16_t)best;
    if (out32) out32[r] = best;
  }
}

// Exact resolution of the queued rows: one CTA per row (grid-stride over the queue), warp w scans the centroids
// k = 4 w, 4 w + 1, ... interleaved in groups of 4 (coalesced: lanes stride the dimension); fp32 filter against the
// threshold  best fast score - (eps_fast + eps_here) * scale  — the exact winner k* satisfies
// a_here(k*) >= exact(k*) - eps_here*scale >= exact(i1) - eps_here*scale >= v1 - (eps_fast + eps_here)*scale —
// fp64 re-score of what passes, smallest distance / smallest index wins, block-reduced through shared memory.
// The queue has one slot per row, so it cannot overflow (a codebook of duplicated centroids queues every row).
__global__ void __launch_bounds__(256)
vq_rescan_kernel(const float* __restrict__ x, int ldx, int M, int D, const float* __restrict__ cb,
                 const float* __restrict__ hn, int K, const unsigned int* __restrict__ n_queued,
                 const int* __restrict__ rescan_rows, const float* __restrict__ rescan_thr, int rescan_cap,
                 float eps_sum, int16_t* __restrict__ out, int32_t* __restrict__ out32) {
  __shared__ double s_d[8];
  __shared__ int s_i[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = min((int)*n_queued, rescan_cap);
  for (int q = blockIdx.x; q < n; q += gridDim.x) {
    const int r = rescan_rows[q];
    const float thr = rescan_thr[2 * q] - eps_sum * rescan_thr[2 * q + 1];
    const float* xr = x + (size_t)r * ldx;
    double bd = INFINITY;
    int bi = 0x7fffffff;
    for (int k0 = 4 * warp; k0 < K; k0 += 32) {
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int d = lane * 4; d < D; d += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + d);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (k0 + u < K) {
            const float4 cv = *reinterpret_cast<const float4*>(cb + (size_t)(k0 + u) * D + d);
            a4[u] = fmaf(xv.x, cv.x, a4[u]); a4[u] = fmaf(xv.y, cv.y, a4[u]);
            a4[u] = fmaf(xv.z, cv.z, a4[u]); a4[u] = fmaf(xv.w, cv.w, a4[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u;
        if (k >= K) break;
        const float a = warp_sum(a4[u]) - __ldg(hn + k);
        if (a >= thr) {
          const double dd = exact_dist(xr, cb + (size_t)k * D, D, lane);
          if (dd < bd || (dd == bd && k < bi)) { bd = dd; bi = k; }
        }
      }
    }
    if (lane == 0) { s_d[warp] = bd; s_i[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w)
        if (s_d[w] < bd || (s_d[w] == bd && s_i[w] < bi)) { bd = s_d[w]; bi = s_i[w]; }
      if (out) out[r] = (int16_t)bi;
      if (out32) out32[r] = bi;
    }
    __syncthreads();
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

int b2t_vq_scan_tensor(const void* A2, const void* C2, int M, int K, int Kpad, int D, const float* half_norm,
                       void* parts, int slices, cudaStream_t st);   // gemm_tc.cu

namespace {
struct VqWs {
  float* xn; __nv_bfloat16* a2; __nv_bfloat16* c2; float* hn; Cand* parts; float* cmax; unsigned int* nfb;
  unsigned int* maxerr; int* rescan_rows; float* rescan_thr; size_t total; int kpad; int slices;
};
VqWs vq_carve(void* base, int rows, int dim, int K) {
  VqWs w;
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? (void*)(p + off) : nullptr; off += align_up(bytes, 256); return r; };
  w.kpad = (K + 255) / 256 * 256;
  w.slices = w.kpad / 256 * 2;
  w.xn = (float*)take((size_t)rows * dim * 4);
  w.a2 = (__nv_bfloat16*)take((size_t)rows * dim * 4);
  w.c2 = (__nv_bfloat16*)take((size_t)K * dim * 4);
  w.hn = (float*)take((size_t)w.kpad * 4);
  w.parts = (Cand*)take((size_t)rows * w.slices * sizeof(Cand));
  float* misc = (float*)take(256);
  w.cmax = misc;
  w.nfb = (unsigned int*)misc + 4;
  w.maxerr = (unsigned int*)misc + 8;
  w.rescan_rows = (int*)take((size_t)rows * 4);      // queue of uncertified rows: one slot per row
  w.rescan_thr = (float*)take((size_t)rows * 8);
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t b2t_vq_workspace_bytes(int rows, int dim, int codebook_size) {
  return vq_carve(nullptr, rows, dim, codebook_size).total;
}

extern "C" int b2t_vq_argmin(const float* x, int ldx, int rows, int dim, const float* codebook,
                             const float* half_norm, int codebook_size, int apply_ln, int impl, int16_t* out,
                             int32_t* out_i32, void* workspace, size_t workspace_bytes, void* stream) {
  B2T_REQUIRE(x && codebook && (out || out_i32) && workspace, B2T_ERR_ARG, "b2t_vq_argmin: null argument");
  B2T_REQUIRE(dim % 32 == 0 && dim >= 32 && ldx % 4 == 0 && ldx >= dim && ((uintptr_t)x % 16) == 0 && ((uintptr_t)codebook % 16) == 0, B2T_ERR_ARG,
              "b2t_vq_argmin: dim must be a multiple of 32, ldx a multiple of 4, pointers 16-byte aligned (dim=%d ldx=%d)", dim, ldx);
  B2T_REQUIRE(codebook_size >= 1 && codebook_size <= 32768, B2T_ERR_ARG,
              "b2t_vq_argmin: codebook_size must be in [1, 32768] for int16 tokens (got %d)", codebook_size);
  B2T_REQUIRE(!apply_ln || dim == 1024, B2T_ERR_ARG, "b2t_vq_argmin: fused LayerNorm needs dim == 1024");
  const bool tensor_ok = dim % 64 == 0;
  B2T_REQUIRE(impl != B2T_IMPL_TENSOR || tensor_ok, B2T_ERR_ARG, "b2t_vq_argmin: tensor path needs dim %% 64 == 0");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (rows <= 0) return B2T_OK;
  VqWs w = vq_carve(workspace, rows, dim, codebook_size);
  B2T_REQUIRE(workspace_bytes >= w.total, B2T_ERR_WORKSPACE, "b2t_vq_argmin: workspace %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  const bool tensor = impl == B2T_IMPL_TENSOR || (impl == B2T_IMPL_AUTO && tensor_ok && codebook_size >= 64);

  const float* xs = x;
  int lds = ldx;
  if (apply_ln) {
    vq_ln_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, ldx, w.xn, rows);
    B2T_LAUNCH_CHECK();
    xs = w.xn; lds = 1024;
  }
  if (half_norm == nullptr) {
    half_norm_kernel<<<(codebook_size + 7) / 8, 256, 0, st>>>(codebook, codebook_size, dim, w.hn);
    B2T_LAUNCH_CHECK();
  } else {
    B2T_CUDA(cudaMemcpyAsync(w.hn, half_norm, (size_t)codebook_size * 4, cudaMemcpyDeviceToDevice, st));
  }
  max_reduce_kernel<<<1, 1024, 0, st>>>(w.hn, codebook_size, w.cmax);
  B2T_LAUNCH_CHECK();
  B2T_CUDA(cudaMemsetAsync(w.nfb, 0, 32, st));   // fallback counter + max observed error
  int slices = 1;
  float rel_eps;
  if (tensor) {
    if (w.kpad > codebook_size) {
      fill_inf_kernel<<<(w.kpad - codebook_size + 255) / 256, 256, 0, st>>>(w.hn, codebook_size, w.kpad);
      B2T_LAUNCH_CHECK();
    }
    split_rows_kernel<<<rows, 256, 0, st>>>(xs, lds, rows, dim, w.a2);
    B2T_LAUNCH_CHECK();
    split_rows_kernel<<<codebook_size, 256, 0, st>>>(codebook, dim, codebook_size, dim, w.c2);
    B2T_LAUNCH_CHECK();
    slices = w.slices;
    rc = b2t_vq_scan_tensor(w.a2, w.c2, rows, codebook_size, w.kpad, dim, w.hn, w.parts, slices, st);
    if (rc != B2T_OK) return rc;
    // bf16x3 product: dropped terms lo*lo and the third mantissa piece, each <= 2^-16 |x_d c_d|, plus
    // fp32 accumulation of 3D/16 MMA steps; 2^-14 leaves >= 8x head-room (checked by b2t_vq_debug_stats)
    rel_eps = 6.103515625e-5f;
  } else {
    vq_scan_simt_kernel<<<(rows + 63) / 64, 256, 0, st>>>(xs, lds, rows, dim, codebook, w.hn, codebook_size, w.parts);
    B2T_LAUNCH_CHECK();
    // sequential fp32 FMA dot of length D: |err| <= D * 2^-24 * sum|x_d c_d|; 2x margin
    rel_eps = 2.0f * (float)dim * 5.9604645e-8f;
  }
  vq_finalize_kernel<<<(rows + 7) / 8, 256, 0, st>>>(xs, lds, rows, dim, codebook, w.hn, codebook_size, w.parts,
                                                     slices, rel_eps, w.cmax, out, out_i32, w.nfb, w.maxerr,
                                                     w.rescan_rows, w.rescan_thr, rows);
  B2T_LAUNCH_CHECK();
  {
    // queued rows (device-side count; usually 0-2): a small fixed grid, a CTA per row
    const float eps_here = 2.0f * (float)dim * 5.9604645e-8f;      // sequential fp32 FMA dot of length D, 2x margin
    int grid = b2t_num_sms();
    vq_rescan_kernel<<<grid, 256, 0, st>>>(xs, lds, rows, dim, codebook, w.hn, codebook_size, w.nfb, w.rescan_rows,
                                           w.rescan_thr, rows, rel_eps + eps_here, out, out_i32);
    B2T_LAUNCH_CHECK();
  }
  return B2T_OK;
}

// Test/diagnostic hook: after a b2t_vq_argmin on `workspace` has completed, returns the number of rows
// that took the re-scan path and the largest observed |fast score - exact score| relative to the bound's
// scale (|x| |c|max + 0.5 |c|max^2) together with the rel_eps the call certified with.
extern "C" int b2t_vq_debug_stats(const void* workspace, int rows, int dim, int codebook_size,
                                  unsigned int* n_fallback_host, float* max_rel_err_host) {
  B2T_REQUIRE(workspace, B2T_ERR_ARG, "b2t_vq_debug_stats: null argument");
  VqWs w = vq_carve(const_cast<void*>(workspace), rows, dim, codebook_size);
  B2T_CUDA(cudaDeviceSynchronize());
  unsigned int tmp[2] = {0, 0};
  B2T_CUDA(cudaMemcpy(&tmp[0], w.nfb, 4, cudaMemcpyDeviceToHost));
  B2T_CUDA(cudaMemcpy(&tmp[1], w.maxerr, 4, cudaMemcpyDeviceToHost));
  if (n_fallback_host) *n_fallback_host = tmp[0];
  if (max_rel_err_host) { float f; memcpy(&f, &tmp[1], 4); *max_rel_err_host = f; }
  return B2T_OK;
}
