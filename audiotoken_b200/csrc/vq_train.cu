// Codebook training step: the EMA k-means update of a Euclidean VQ codebook.
//
// Reference: scripts/clustering/cluster_tokens.py:293-311 calls `VectorQuantize(dim, codebook_size, decay=0.8,
// commitment_weight=1)` (:142-147) in training mode on every batch of LayerNormed layer-19 embeddings; the class is
// third-party (`vector_quantize_pytorch`, unpinned, requirements.txt:10).  Its published training forward for a
// Euclidean codebook with ema_update, one head, no dead-code expiry (threshold_ema_dead_code = 0):
//     idx        = argmin_k |x - embed_k|                       (old codebook; b2t_vq_argmin)
//     quantize   = embed[idx];  commit = mean((quantize - x)^2) * commitment_weight
//     n_k        = #{r : idx_r = k};  s_k = sum_{idx_r = k} x_r
//     cluster_size <- cluster_size * decay + n   * (1 - decay)
//     embed_avg    <- embed_avg    * decay + s   * (1 - decay)
//     smoothed_k   = (cluster_size_k + eps) / (sum(cluster_size) + K eps) * sum(cluster_size)
//     embed_k      = embed_avg_k / smoothed_k
// The assignment is the exact tensor-core kernel of vq.cu; this file is the HBM-bound rest, written so that the
// result is **deterministic** (no floating-point atomics): a stable counting sort of the rows by centroid
// (per-block histograms -> column scan -> in-order scatter) followed by one CTA per centroid that sums its rows in
// row order, folds the EMA and the normalisation in and emits its share of the commitment loss.
#include "common.cuh"

namespace {

constexpr int kRowsPerBlock = 1024;

// hist[blk][k] = rows of block blk assigned to k.  Integer atomics: the result does not depend on their order.
__global__ void __launch_bounds__(256)
ema_hist_kernel(const int32_t* __restrict__ idx, int rows, int K, int32_t* __restrict__ hist) {
  const int blk = blockIdx.x;
  int32_t* h = hist + (size_t)blk * K;
  const int r0 = blk * kRowsPerBlock;
  for (int r = r0 + threadIdx.x; r < min(rows, r0 + kRowsPerBlock); r += blockDim.x) {
    const int k = __ldg(idx + r);
    if (k >= 0 && k < K) atomicAdd(h + k, 1);
  }
}

// Column scan: hist[blk][k] becomes the number of rows of centroid k in blocks < blk; count[k] = total.
__global__ void __launch_bounds__(256)
ema_colscan_kernel(int32_t* __restrict__ hist, int nblk, int K, int32_t* __restrict__ count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int run = 0;
  for (int b = 0; b < nblk; ++b) {
    const int v = hist[(size_t)b * K + k];
    hist[(size_t)b * K + k] = run;
    run += v;
  }
  count[k] = run;
}

// One block: exclusive scan of count -> start[0..K]; the EMA of cluster_size and the Laplace-smoothed sizes.
// Sums run over k in a fixed order (thread-strided partials, then a tree), so the result is reproducible.
__global__ void __launch_bounds__(1024)
ema_sizes_kernel(const int32_t* __restrict__ count, int K, float decay, float eps, float* __restrict__ cluster_size,
                 int32_t* __restrict__ start, float* __restrict__ smoothed) {
  __shared__ int s_part[1024];
  __shared__ double s_sum[1024];
  const int t = threadIdx.x;
  const int per = (K + 1023) / 1024;
  const int k0 = t * per, k1 = min(K, k0 + per);
  int c = 0;
  double tot = 0.0;
  for (int k = k0; k < k1; ++k) {
    c += count[k];
    const float cs = cluster_size[k] * decay + (float)count[k] * (1.0f - decay);
    cluster_size[k] = cs;
    tot += (double)cs;
  }
  s_part[t] = c;
  s_sum[t] = tot;
  __syncthreads();
  // Hillis-Steele inclusive scan of the per-thread counts; tree sum of the sizes
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = t >= o ? s_part[t - o] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  for (int o = 512; o > 0; o >>= 1) {
    if (t < o) s_sum[t] += s_sum[t + o];
    __syncthreads();
  }
  int run = t ? s_part[t - 1] : 0;
  const float total = (float)s_sum[0];
  for (int k = k0; k < k1; ++k) {
    start[k] = run;
    run += count[k];
    smoothed[k] = (cluster_size[k] + eps) / (total + (float)K * eps) * total;
  }
  if (t == 1023) start[K] = s_part[1023];
}

// Stable scatter: one warp per block of rows walks its rows in order, 32 at a time; rows of the same centroid
// inside a group are ranked by lane (match_any), the per-(block, centroid) cursor lives in hist[blk][k].
__global__ void __launch_bounds__(32)
ema_scatter_kernel(const int32_t* __restrict__ idx, int rows, int K, int32_t* __restrict__ hist,
                   const int32_t* __restrict__ start, int32_t* __restrict__ perm) {
  const int blk = blockIdx.x, lane = threadIdx.x;
  int32_t* cur = hist + (size_t)blk * K;
  const int r0 = blk * kRowsPerBlock, r1 = min(rows, r0 + kRowsPerBlock);
  for (int base = r0; base < r1; base += 32) {
    const int r = base + lane;
    const bool in = r < r1;
    int k = in ? __ldg(idx + r) : -1;
    if (k >= K) k = -1;
    const unsigned same = __match_any_sync(0xffffffffu, k);
    const int rank = __popc(same & ((1u << lane) - 1u));
    const bool leader = rank == 0;
    int c0 = 0;
    if (k >= 0 && leader) { c0 = cur[k]; cur[k] = c0 + __popc(same); }
    c0 = __shfl_sync(0xffffffffu, c0, __ffs(same) - 1);
    if (k >= 0) perm[__ldg(start + k) + c0 + rank] = r;
    __syncwarp();
  }
}

// One CTA per centroid, thread per 4 columns (dim <= 4096): s = sum of its rows in row order, the commitment-loss
// share sum |x - embed_old|^2, then the EMA and the normalised codebook row.
__global__ void __launch_bounds__(256)
ema_update_kernel(const float* __restrict__ x, int ldx, int dim, const int32_t* __restrict__ perm,
                  const int32_t* __restrict__ start, const float* __restrict__ smoothed, float decay,
                  float* __restrict__ codebook, float* __restrict__ embed_avg, double* __restrict__ loss_part) {
  const int k = blockIdx.x;
  const int a = __ldg(start + k), b = __ldg(start + k + 1);
  const float sm = __ldg(smoothed + k);
  double lacc = 0.0;
  for (int d0 = threadIdx.x * 4; d0 < dim; d0 += blockDim.x * 4) {
    const float4 c = *reinterpret_cast<const float4*>(codebook + (size_t)k * dim + d0);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    float l = 0.f;
    int j = a;
    for (; j + 1 < b; j += 2) {          // two rows in flight; the additions stay in row order
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + (size_t)__ldg(perm + j) * ldx + d0));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (size_t)__ldg(perm + j + 1) * ldx + d0));
      s.x += v0.x; s.y += v0.y; s.z += v0.z; s.w += v0.w;
      l += (v0.x - c.x) * (v0.x - c.x) + (v0.y - c.y) * (v0.y - c.y) + (v0.z - c.z) * (v0.z - c.z) + (v0.w - c.w) * (v0.w - c.w);
      s.x += v1.x; s.y += v1.y; s.z += v1.z; s.w += v1.w;
      l += (v1.x - c.x) * (v1.x - c.x) + (v1.y - c.y) * (v1.y - c.y) + (v1.z - c.z) * (v1.z - c.z) + (v1.w - c.w) * (v1.w - c.w);
    }
    if (j < b) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + (size_t)__ldg(perm + j) * ldx + d0));
      s.x += v0.x; s.y += v0.y; s.z += v0.z; s.w += v0.w;
      l += (v0.x - c.x) * (v0.x - c.x) + (v0.y - c.y) * (v0.y - c.y) + (v0.z - c.z) * (v0.z - c.z) + (v0.w - c.w) * (v0.w - c.w);
    }
    float4* pa = reinterpret_cast<float4*>(embed_avg + (size_t)k * dim + d0);
    float4 e = *pa;
    const float w = 1.0f - decay;
    e.x = e.x * decay + s.x * w; e.y = e.y * decay + s.y * w; e.z = e.z * decay + s.z * w; e.w = e.w * decay + s.w * w;
    *pa = e;
    *reinterpret_cast<float4*>(codebook + (size_t)k * dim + d0) = make_float4(e.x / sm, e.y / sm, e.z / sm, e.w / sm);
    lacc += (double)l;
  }
  // block sum of the loss share in a fixed order
  __shared__ double s_l[256];
  s_l[threadIdx.x] = lacc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_l[threadIdx.x] += s_l[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_part[k] = s_l[0];
}

__global__ void __launch_bounds__(1024)
ema_loss_kernel(const double* __restrict__ loss_part, int K, double denom, float weight, float* __restrict__ out) {
  __shared__ double s[1024];
  double a = 0.0;
  for (int k = threadIdx.x; k < K; k += 1024) a += loss_part[k];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)(s[0] / denom) * weight;
}

// quantize = embed_old[idx]  (the value VectorQuantize returns first; optional)
__global__ void __launch_bounds__(256)
ema_gather_kernel(const int32_t* __restrict__ idx, int rows, int dim, int K, const float* __restrict__ codebook,
                  float* __restrict__ out) {
  const int r = blockIdx.x;
  const int k = min(max(__ldg(idx + r), 0), K - 1);
  for (int d0 = threadIdx.x * 4; d0 < dim; d0 += blockDim.x * 4)
    *reinterpret_cast<float4*>(out + (size_t)r * dim + d0) = __ldg(reinterpret_cast<const float4*>(codebook + (size_t)k * dim + d0));
}

struct EmaWs {
  int32_t *hist, *count, *start, *perm;
  float* smoothed;
  double* loss_part;
  size_t total;
  int nblk;
};

EmaWs ema_carve(void* base, int rows, int K) {
  EmaWs w{};
  w.nblk = (rows + kRowsPerBlock - 1) / kRowsPerBlock;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return (char*)base + o; };
  w.hist = (int32_t*)take((size_t)w.nblk * K * 4);
  w.count = (int32_t*)take((size_t)K * 4);
  w.start = (int32_t*)take((size_t)(K + 1) * 4);
  w.perm = (int32_t*)take((size_t)rows * 4);
  w.smoothed = (float*)take((size_t)K * 4);
  w.loss_part = (double*)take((size_t)K * 8);
  w.total = off;
  return w;
}

}  // namespace

extern "C" size_t b2t_vq_ema_workspace_bytes(int rows, int dim, int codebook_size) {
  (void)dim;
  if (rows <= 0 || codebook_size <= 0) return 256;
  return ema_carve(nullptr, rows, codebook_size).total;
}

extern "C" int b2t_vq_ema_update(const float* x, int ldx, int rows, int dim, const int32_t* idx, float* codebook,
                                 float* embed_avg, float* cluster_size, int codebook_size, float decay, float eps,
                                 float commitment_weight, float* commit_loss, float* quantized, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  B2T_REQUIRE(x && idx && codebook && embed_avg && cluster_size && workspace, B2T_ERR_ARG, "b2t_vq_ema_update: null argument");
  B2T_REQUIRE(dim % 4 == 0 && dim >= 4 && dim <= 4096 && ldx % 4 == 0 && ldx >= dim && ((uintptr_t)x % 16) == 0 &&
                  ((uintptr_t)codebook % 16) == 0 && ((uintptr_t)embed_avg % 16) == 0,
              B2T_ERR_ARG, "b2t_vq_ema_update: dim must be a multiple of 4 (<= 4096), ldx a multiple of 4, pointers 16-byte aligned (dim=%d ldx=%d)", dim, ldx);
  B2T_REQUIRE(codebook_size >= 1 && codebook_size <= 32768, B2T_ERR_ARG, "b2t_vq_ema_update: codebook_size must be in [1, 32768] (got %d)", codebook_size);
  B2T_REQUIRE(decay >= 0.f && decay <= 1.f && eps > 0.f, B2T_ERR_ARG, "b2t_vq_ema_update: decay must be in [0, 1] and eps > 0");
  B2T_REQUIRE(rows >= 1, B2T_ERR_ARG, "b2t_vq_ema_update: an EMA step needs at least one row (got %d)", rows);
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  const int K = codebook_size;
  EmaWs w = ema_carve(workspace, rows, K);
  B2T_REQUIRE(workspace_bytes >= w.total, B2T_ERR_WORKSPACE, "b2t_vq_ema_update: workspace %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  if (quantized) {
    ema_gather_kernel<<<rows, 256, 0, st>>>(idx, rows, dim, K, codebook, quantized);
    B2T_LAUNCH_CHECK();
  }
  B2T_CUDA(cudaMemsetAsync(w.hist, 0, (size_t)w.nblk * K * 4, st));
  ema_hist_kernel<<<w.nblk, 256, 0, st>>>(idx, rows, K, w.hist);
  B2T_LAUNCH_CHECK();
  ema_colscan_kernel<<<(K + 255) / 256, 256, 0, st>>>(w.hist, w.nblk, K, w.count);
  B2T_LAUNCH_CHECK();
  ema_sizes_kernel<<<1, 1024, 0, st>>>(w.count, K, decay, eps, cluster_size, w.start, w.smoothed);
  B2T_LAUNCH_CHECK();
  ema_scatter_kernel<<<w.nblk, 32, 0, st>>>(idx, rows, K, w.hist, w.start, w.perm);
  B2T_LAUNCH_CHECK();
  ema_update_kernel<<<K, 256, 0, st>>>(x, ldx, dim, w.perm, w.start, w.smoothed, decay, codebook, embed_avg, w.loss_part);
  B2T_LAUNCH_CHECK();
  if (commit_loss) {
    ema_loss_kernel<<<1, 1024, 0, st>>>(w.loss_part, K, (double)rows * (double)dim, commitment_weight, commit_loss);
    B2T_LAUNCH_CHECK();
  }
  return B2T_OK;
}
