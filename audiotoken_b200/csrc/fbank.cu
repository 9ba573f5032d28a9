// Log-mel front end (reference audiotoken/processors.py): three kernels.
//   fbank_logmel_kernel : one warp per valid frame — load 400 samples, x2^15, DC removal,
//                         pre-emphasis, povey window, 512-point real FFT (as a 256-point complex
//                         FFT in shared memory), power, sparse mel triangles, floor, ln.
//   fbank_stats_kernel  : per clip / mel bin masked mean and biased variance (two pass).
//   fbank_stack_ln_kernel: normalise + stride-2 stacking + pad value 1.0 + LayerNorm(160).
// HBM-bound by design: wave is read once from DRAM (frame overlap is served by L1/L2), log-mel is
// written once; algorithmic bytes per audio-second = 16000*4 in + 100*80*4 out.
#include "common.cuh"

namespace {

constexpr int kFrame = 400, kHop = 160, kMel = 80, kMaxSpan = 32;
constexpr int kWarps = 6;      // 5 KB of FFT buffers per warp + 14 KB of tables stay below the 48 KB static limit

__device__ float2 g_tw512[256];  // exp(-2*pi*i*k/512), k < 256

__global__ void init_twiddles_kernel() {
  int k = threadIdx.x;
  double s, c;
  sincospi(-2.0 * (double)k / 512.0, &s, &c);
  g_tw512[k] = make_float2((float)c, (float)s);
}

__global__ void __launch_bounds__(kWarps * 32)
fbank_logmel_kernel(const float* __restrict__ wave, const int64_t* __restrict__ wave_off,
                    const int32_t* __restrict__ frame_off, int n_clips, int total_frames,
                    const float* __restrict__ window, const int32_t* __restrict__ mel_start,
                    const int32_t* __restrict__ mel_count, const float* __restrict__ mel_weight,
                    float* __restrict__ logmel, int mel_bf16) {
  __shared__ float2 s_tw[256];
  __shared__ float s_win[kFrame];
  __shared__ float s_melw[kMel * kMaxSpan];
  __shared__ int s_mstart[kMel], s_mcnt[kMel];
  __shared__ float2 s_z[kWarps][320];     // 256 points + padding (up to 4 per 16)
  __shared__ float2 s_zb[kWarps][320];

  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tw[i] = g_tw512[i];
  for (int i = threadIdx.x; i < kFrame; i += blockDim.x) s_win[i] = window[i];
  for (int i = threadIdx.x; i < kMel * kMaxSpan; i += blockDim.x) {
    float w = mel_weight[i];
    s_melw[i] = mel_bf16 ? bf16_round(w) : w;
  }
  for (int i = threadIdx.x; i < kMel; i += blockDim.x) {
    s_mstart[i] = mel_start[i];
    s_mcnt[i] = mel_count[i];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xs = reinterpret_cast<float*>(s_zb[warp]);   // frame / power spectrum; dead while the FFT ping-pongs
  float2* z = s_z[warp];
  float* zf = reinterpret_cast<float*>(z);

  for (int g = blockIdx.x * kWarps + warp; g < total_frames; g += gridDim.x * kWarps) {
    const int clip = find_segment(frame_off, n_clips, g);
    const int n = g - __ldg(frame_off + clip);
    const float* src = wave + __ldg(wave_off + clip) + (int64_t)n * kHop;

    // -- load, scale to int16 range (processors.py:155), remove DC (:168-169)
    float v[13];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 13; ++j) {
      int idx = lane + 32 * j;
      v[j] = idx < kFrame ? __ldg(src + idx) * 32768.0f : 0.f;
      sum += v[j];
    }
    sum = warp_sum(sum);
    const float mean = sum / (float)kFrame;
#pragma unroll
    for (int j = 0; j < 13; ++j) {
      int idx = lane + 32 * j;
      if (idx < kFrame) xs[idx] = v[j] - mean;
    }
    __syncwarp();

    // -- pre-emphasis (:171-173) and window (:175), zero-padded to 512 and packed as complex
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int idx = lane + 32 * j;
      float y = 0.f;
      if (idx < kFrame) {
        float x0 = xs[idx];
        if (idx > 0) y = __fsub_rn(x0, __fmul_rn(0.97f, xs[idx - 1]));
        else y = __fmul_rn(x0, 0.03f);
        y = __fmul_rn(y, s_win[idx]);
      }
      zf[idx] = y;
    }
    __syncwarp();

    // -- 256-point complex FFT: Stockham radix-4, 4 passes, natural-order output.  Reads of a pass are unit-stride
    //    across lanes; the scattered writes of the first two passes land in buffers padded by P elements per 16
    //    (P = 1, 4) so that every shared-memory access of the transform is bank-conflict free.
    {
      auto ph = [](int i, int P) { return i + P * (i >> 4); };
      auto pass = [&](const float2* src, int Ps, float2* dst, int Pd, int Ns) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = lane + 32 * u, k = j & (Ns - 1);
          float2 v0 = src[ph(j, Ps)], v1 = src[ph(j + 64, Ps)], v2 = src[ph(j + 128, Ps)], v3 = src[ph(j + 192, Ps)];
          if (Ns > 1) {
            const int m = k * (128 / Ns);                    // twiddle exp(-2 pi i r k / (4 Ns)) = tw512[r m]
            const float2 w1 = s_tw[m], w2 = s_tw[2 * m];
            float2 w3 = s_tw[(3 * m) & 255];
            if (3 * m >= 256) { w3.x = -w3.x; w3.y = -w3.y; }
            v1 = make_float2(v1.x * w1.x - v1.y * w1.y, v1.x * w1.y + v1.y * w1.x);
            v2 = make_float2(v2.x * w2.x - v2.y * w2.y, v2.x * w2.y + v2.y * w2.x);
            v3 = make_float2(v3.x * w3.x - v3.y * w3.y, v3.x * w3.y + v3.y * w3.x);
          }
          const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
          const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
          const int j0 = ((j - k) << 2) + k;
          dst[ph(j0, Pd)] = make_float2(s02.x + s13.x, s02.y + s13.y);
          dst[ph(j0 + Ns, Pd)] = make_float2(d02.x + d13.y, d02.y - d13.x);          // d02 - i d13
          dst[ph(j0 + 2 * Ns, Pd)] = make_float2(s02.x - s13.x, s02.y - s13.y);
          dst[ph(j0 + 3 * Ns, Pd)] = make_float2(d02.x - d13.y, d02.y + d13.x);      // d02 + i d13
        }
        __syncwarp();
      };
      float2* zb = s_zb[warp];
      pass(z, 0, zb, 1, 1);
      pass(zb, 1, z, 4, 4);
      pass(z, 4, zb, 0, 16);
      pass(zb, 0, z, 0, 64);
    }

    // -- split into the 512-point real spectrum, power (:177-181); bins 0..255 (Nyquist weight = 0)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      int k = lane + 32 * u;
      float2 zk = z[k];
      float2 zm = z[(256 - k) & 255];
      float ex = 0.5f * (zk.x + zm.x), ey = 0.5f * (zk.y - zm.y);
      float dx = 0.5f * (zk.x - zm.x), dy = 0.5f * (zk.y + zm.y);
      float ox = dy, oy = -dx;
      float2 w = s_tw[k];
      float xr = ex + (w.x * ox - w.y * oy);
      float xi = ey + (w.x * oy + w.y * ox);
      float p = xr * xr + xi * xi;
      xs[k] = mel_bf16 ? bf16_round(p) : p;
    }
    __syncwarp();

    // -- mel projection (:184), floor (:185), ln (:188)
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      int f = lane + 32 * u;
      if (f < kMel) {
        int st = s_mstart[f], cnt = s_mcnt[f];
        float acc = 0.f;
        for (int b = 0; b < cnt; ++b) acc = fmaf(xs[st + b], s_melw[f * kMaxSpan + b], acc);
        if (mel_bf16) acc = bf16_round(acc);
        acc = fmaxf(acc, 1.192092955078125e-07f);
        logmel[(size_t)g * kMel + f] = logf(acc);
      }
    }
    __syncwarp();
  }
}

// one block per clip; 320 threads = 4 frame lanes x 80 bins
__global__ void __launch_bounds__(320)
fbank_stats_kernel(const float* __restrict__ logmel, const int32_t* __restrict__ frame_off,
                   float* __restrict__ mean_out, float* __restrict__ std_out) {
  __shared__ double s_red[4][kMel];
  __shared__ float s_mean[kMel];
  const int clip = blockIdx.x;
  const int f0 = frame_off[clip], nf = frame_off[clip + 1] - f0;
  const int bin = threadIdx.x % kMel, fl = threadIdx.x / kMel;
  const float* base = logmel + (size_t)f0 * kMel + bin;
  double acc = 0.0;
  for (int n = fl; n < nf; n += 4) acc += (double)base[(size_t)n * kMel];
  s_red[fl][bin] = acc;
  __syncthreads();
  const double cnt = nf > 0 ? (double)nf : 1.0;  // clamp(min=1), processors.py:131
  if (fl == 0) s_mean[bin] = (float)((s_red[0][bin] + s_red[1][bin] + s_red[2][bin] + s_red[3][bin]) / cnt);
  __syncthreads();
  const float mu = s_mean[bin];
  acc = 0.0;
  for (int n = fl; n < nf; n += 4) {
    float d = base[(size_t)n * kMel] - mu;
    acc += (double)d * (double)d;
  }
  s_red[fl][bin] = acc;
  __syncthreads();
  if (fl == 0) {
    float var = (float)((s_red[0][bin] + s_red[1][bin] + s_red[2][bin] + s_red[3][bin]) / cnt);
    mean_out[clip * kMel + bin] = mu;
    std_out[clip * kMel + bin] = sqrtf(var + 1e-7f);
  }
}

// one warp per token row: 160 features = 5 per lane; LayerNorm(160) fused.
template <typename OutT>
__global__ void __launch_bounds__(256)
fbank_stack_ln_kernel(const float* __restrict__ logmel, const float* __restrict__ mean,
                      const float* __restrict__ std_, const int32_t* __restrict__ frame_off,
                      const int32_t* __restrict__ stack_frames, const int32_t* __restrict__ row_off,
                      int n_clips, int total_rows, const float* __restrict__ ln_w,
                      const float* __restrict__ ln_b, OutT* __restrict__ out,
                      float* __restrict__ features, uint8_t* __restrict__ row_valid) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= total_rows) return;
  const int clip = find_segment(row_off, n_clips, r);
  const int t = r - __ldg(row_off + clip);
  const int f0 = __ldg(frame_off + clip);
  const int ns = __ldg(stack_frames + clip);
  float v[5];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    int e = lane + 32 * j;
    int sub = e >= kMel ? 1 : 0, bin = e - sub * kMel;
    int fr = 2 * t + sub;
    float x = 1.0f;  // padding_value (processors.py:200-201)
    if (fr < ns) {
      float lm = __ldg(logmel + (size_t)(f0 + fr) * kMel + bin);
      x = (lm - __ldg(mean + clip * kMel + bin)) / __ldg(std_ + clip * kMel + bin);
    }
    v[j] = x;
    sum += x;
    if (features) features[(size_t)r * 160 + e] = x;
  }
  if (lane == 0 && row_valid) row_valid[r] = (2 * t < ns) ? 1 : 0;  // mask of sub-frame 0 (:204)
  sum = warp_sum(sum);
  const float mu = sum / 160.f;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) { float d = v[j] - mu; sq += d * d; }
  sq = warp_sum(sq);
  const float rstd = rsqrtf(sq / 160.f + 1e-5f);
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    int e = lane + 32 * j;
    st_act(out + (size_t)r * 160 + e, (v[j] - mu) * rstd * __ldg(ln_w + e) + __ldg(ln_b + e));
  }
}

int ensure_twiddles(cudaStream_t st) {
  static bool done[B2T_MAX_DEVICES] = {false};
  const int dev = b2t_device_index();
  if (done[dev]) return B2T_OK;
  init_twiddles_kernel<<<1, 256, 0, st>>>();
  B2T_LAUNCH_CHECK();
  done[dev] = true;
  return B2T_OK;
}

}  // namespace

extern "C" int b2t_fbank_logmel(const float* wave, const b2t_batch* b, const b2t_fbank_tables* t,
                                float* logmel, int mel_bf16, void* stream) {
  B2T_REQUIRE(wave && b && t && logmel, B2T_ERR_ARG, "b2t_fbank_logmel: null argument");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (b->total_frames <= 0) return B2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  rc = ensure_twiddles(st);
  if (rc != B2T_OK) return rc;
  int blocks = (b->total_frames + kWarps - 1) / kWarps;
  int cap = b2t_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  fbank_logmel_kernel<<<blocks, kWarps * 32, 0, st>>>(wave, b->wave_off, b->frame_off, b->n_clips,
                                                     b->total_frames, t->window, t->mel_start,
                                                     t->mel_count, t->mel_weight, logmel, mel_bf16);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}

extern "C" int b2t_fbank_stats(const float* logmel, const b2t_batch* b, float* mean, float* std_,
                               void* stream) {
  B2T_REQUIRE(logmel && b && mean && std_, B2T_ERR_ARG, "b2t_fbank_stats: null argument");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (b->n_clips <= 0) return B2T_OK;
  fbank_stats_kernel<<<b->n_clips, 320, 0, (cudaStream_t)stream>>>(logmel, b->frame_off, mean, std_);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}

extern "C" int b2t_fbank_stack_ln(const float* logmel, const float* mean, const float* std_,
                                  const b2t_batch* b, const float* ln_w, const float* ln_b,
                                  void* out, float* features, uint8_t* row_valid, int precision,
                                  void* stream) {
  B2T_REQUIRE(logmel && mean && std_ && b && ln_w && ln_b && out, B2T_ERR_ARG,
              "b2t_fbank_stack_ln: null argument");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (b->total_rows <= 0) return B2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (b->total_rows + 7) / 8;
  if (precision == B2T_PREC_BF16)
    fbank_stack_ln_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        logmel, mean, std_, b->frame_off, b->stack_frames, b->row_off, b->n_clips, b->total_rows,
        ln_w, ln_b, (__nv_bfloat16*)out, features, row_valid);
  else
    fbank_stack_ln_kernel<float><<<blocks, 256, 0, st>>>(
        logmel, mean, std_, b->frame_off, b->stack_frames, b->row_off, b->n_clips, b->total_rows,
        ln_w, ln_b, (float*)out, features, row_valid);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
