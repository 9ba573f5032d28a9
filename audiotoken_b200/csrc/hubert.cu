// The reference's own `semantic_s`: mHuBERT-base + k-means (reference audiotoken/encoder.py:60-108; HF
// transformers hubert/modeling_hubert.py — line numbers as cited in oracle/hubert.py).  SURVEY.md 8f rank 1.
//
// Ragged batch of un-padded, already normalised clips.  Stages (B2T_PREC_BF16 reproduces the rounding points of
// torch.amp.autocast, encoder.py:89; B2T_PREC_FP32 is the reference's CPU arithmetic on CUDA cores):
//   conv0 (1 -> 512, k 10, s 5, no bias) + GroupNorm(512 groups) + GELU     hubert_conv0_{stats,apply}_kernel
//       two passes over the waveform (4 B per sample each): per-channel sums over every conv0 frame that touches
//       the clip, divided by the frame count of the PADDED chunk (the reference pads every chunk to chunk_size and
//       the GroupNorm statistics run over the padding too — oracle/hubert.py::hidden_states_ragged); conv0 is
//       recomputed in the second pass instead of being stored.
//   conv1..6 (512 -> 512, k 3,3,3,3,2,2, s 2, GELU)                          b2t_gemm, one launch per layer
//       level l keeps clip c at row off0[c] >> l, so out row r reads in rows 2r .. 2r+k-1: the im2col matrix IS the
//       activation buffer read with a row stride of 2*512 elements and a row width of k*512 (overlapping-row TMA
//       tensor map on the tcgen05 path) — no im2col copy, no per-clip launches; rows between clips are slack.
//   LayerNorm(512) of the valid frames -> compact rows -> Linear 512 -> 768   hubert_gather_ln_kernel, b2t_gemm
//   positional conv (k 128, 16 groups, weight-normed) + GELU, x + pos, LayerNorm(768)
//       group-major zero-gapped copy of x (64 zero rows between clips = the conv's zero padding) so that group g is
//       the overlapping-row GEMM  [rows, 128 taps x 48]  x  [48 (padded to 128), 6144]^T;  hubert_pos_{scatter,finish}_kernel
//   n_layers post-LN transformer layers (12 heads x 64, FFN 3072, GELU)        b2t_gemm, b2t_attention (zero bias),
//                                                                              hubert_add_ln_kernel
//   affine-free LayerNorm(768) -> nearest k-means centre (exact)               b2t_vq_argmin
//
// Tensor names for b2t_hubert_set_tensor ("act" = bf16 in B2T_PREC_BF16, fp32 in B2T_PREC_FP32; the rest fp32):
//   fe.conv0.w [512,10] | fe.gn.w fe.gn.b [512] | fe.conv<l>.w [512, k*512] act, K index = tap*512 + c_in (l = 1..6)
//   fp.ln.w fp.ln.b [512] | fp.proj.w [768,512] act | fp.proj.b [768]
//   pos.w [16,128,6144] act (rows 48..127 of every group zero; K index = tap*48 + c_in) | pos.b [16,128]
//   enc.ln.w enc.ln.b [768] | attn.zero_bias [73,64] act (zeros: plain attention through the relative-key kernels)
//   L<i>.attn.wqkv [2304,768] act | .bqkv [2304] | .wo [768,768] act | .bo [768] | L<i>.ln.w/.b
//   L<i>.ffn.w1 [3072,768] act | .b1 [3072] | .w2 [768,3072] act | .b2 [768] | L<i>.final.ln.w/.b
//   codebook [K,768]
#include <map>
#include <string>
#include "common.cuh"

void b2t_reset_launch_count();

struct b2t_hubert_model {
  int n_layers;
  int codebook_size;
  int precision;
  std::map<std::string, const void*> t;
};

namespace {

constexpr int kC = 512, kK0 = 10, kS0 = 5, kTileF = 128, kHid = 768, kGroups = 16, kGC = 48, kPosK = 128, kPosN = 128;

template <bool kBF16>
B2T_DEVICE void conv0_load(const float* __restrict__ wave, long long base, int n, int f0, int nf, float* sx, int tid) {
  const int ns = nf * kS0 + (kK0 - kS0);
  for (int i = tid; i < ns; i += 256) {
    const int sidx = f0 * kS0 + i;
    sx[i] = sidx < n ? r16<kBF16>(wave[base + sidx]) : 0.f;      // zero padding of the chunk (datasets.py:99-103)
  }
}

// Wav2Vec2FeatureExtractor(do_normalize): (x - mean) / sqrt(var + 1e-7) over the clip (reference encoder.py:20-26; the
// batch reader applies it to every streamed chunk, datasets.py:75-79).  One CTA per clip, fp64 sums.
__global__ void __launch_bounds__(1024)
hubert_wave_stats_kernel(const float* __restrict__ wave, const int64_t* __restrict__ wave_off,
                         const int32_t* __restrict__ n_samples, float2* __restrict__ stats) {
  __shared__ double ss[32], sq[32];
  const int clip = blockIdx.x, n = n_samples[clip];
  const float* x = wave + wave_off[clip];
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const double v = x[i]; s += v; q += v * v; }
  s = warp_sum(s); q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s = ss[threadIdx.x]; q = sq[threadIdx.x];
    s = warp_sum(s); q = warp_sum(q);
    if (threadIdx.x == 0) {
      const double mean = s / n, var = fmax(q / n - mean * mean, 0.0);
      stats[clip] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-7)));
    }
  }
}
__global__ void __launch_bounds__(256)
hubert_wave_norm_kernel(const float* __restrict__ wave, const int64_t* __restrict__ wave_off, const int64_t* __restrict__ norm_off,
                        const int32_t* __restrict__ n_samples, const float2* __restrict__ stats, float* __restrict__ out) {
  const int clip = blockIdx.y, n = n_samples[clip];
  const float2 st = stats[clip];
  const float* x = wave + wave_off[clip];
  float* o = out + norm_off[clip];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) o[i] = (x[i] - st.x) * st.y;
}

// per-channel sum and sum of squares of conv0 over one 128-frame tile
template <bool kBF16>
__global__ void __launch_bounds__(256)
hubert_conv0_stats_kernel(const float* __restrict__ wave, const int64_t* __restrict__ wave_off,
                          const int32_t* __restrict__ n_samples, const int32_t* __restrict__ gn_count,
                          const int32_t* __restrict__ tile_clip, const int32_t* __restrict__ tile_f0,
                          const float* __restrict__ w0, float2* __restrict__ partial) {
  __shared__ float sx[kTileF * kS0 + 8];
  const int tid = threadIdx.x;
  const int clip = tile_clip[blockIdx.x], f0 = tile_f0[blockIdx.x], n = n_samples[clip];
  const int nstat = min((n + kS0 - 1) / kS0, gn_count[clip]);     // frames that touch a real sample, within the chunk
  const int nf = min(kTileF, nstat - f0);
  conv0_load<kBF16>(wave, wave_off[clip], n, f0, nf, sx, tid);
  float w[2][kK0];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int j = 0; j < kK0; ++j) w[c][j] = r16<kBF16>(w0[(2 * tid + c) * kK0 + j]);
  __syncthreads();
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  for (int f = 0; f < nf; ++f) {
    float h0 = 0.f, h1 = 0.f;
#pragma unroll
    for (int j = 0; j < kK0; ++j) { const float x = sx[f * kS0 + j]; h0 = fmaf(w[0][j], x, h0); h1 = fmaf(w[1][j], x, h1); }
    h0 = r16<kBF16>(h0); h1 = r16<kBF16>(h1);                     // conv1d output dtype under autocast
    s0 += h0; q0 = fmaf(h0, h0, q0); s1 += h1; q1 = fmaf(h1, h1, q1);
  }
  partial[(size_t)blockIdx.x * kC + 2 * tid] = make_float2(s0, q0);
  partial[(size_t)blockIdx.x * kC + 2 * tid + 1] = make_float2(s1, q1);
}

// tiles of a clip summed in tile order (fp64): mean and 1/sqrt(var + eps) per (clip, channel)
__global__ void __launch_bounds__(kC)
hubert_gn_finalize_kernel(const float2* __restrict__ partial, const int32_t* __restrict__ tile_first,
                          const int32_t* __restrict__ gn_count, float2* __restrict__ mean_rstd) {
  const int clip = blockIdx.x, c = threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int t = tile_first[clip]; t < tile_first[clip + 1]; ++t) {
    const float2 p = partial[(size_t)t * kC + c];
    s += p.x; q += p.y;
  }
  const double cnt = (double)gn_count[clip];
  const double mean = s / cnt;
  const double var = q / cnt - mean * mean;
  mean_rstd[(size_t)clip * kC + c] = make_float2((float)mean, (float)(1.0 / sqrt(fmax(var, 0.0) + 1e-5)));
}

// conv0 again, GroupNorm, GELU -> level-0 rows (channels-last)
template <typename ActT, bool kBF16>
__global__ void __launch_bounds__(256)
hubert_conv0_apply_kernel(const float* __restrict__ wave, const int64_t* __restrict__ wave_off,
                          const int32_t* __restrict__ n_samples, const int32_t* __restrict__ off0,
                          const int32_t* __restrict__ tile_clip, const int32_t* __restrict__ tile_f0,
                          const float* __restrict__ w0, const float2* __restrict__ mean_rstd,
                          const float* __restrict__ gamma, const float* __restrict__ beta, ActT* __restrict__ out) {
  __shared__ float sx[kTileF * kS0 + 8];
  const int tid = threadIdx.x;
  const int clip = tile_clip[blockIdx.x], f0 = tile_f0[blockIdx.x], n = n_samples[clip];
  const int nin = (n - kK0) / kS0 + 1;                            // frames whose receptive field lies inside the clip
  const int nf = min(kTileF, nin - f0);
  conv0_load<kBF16>(wave, wave_off[clip], n, f0, nf, sx, tid);
  float w[2][kK0], a[2], b[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int ch = 2 * tid + c;
#pragma unroll
    for (int j = 0; j < kK0; ++j) w[c][j] = r16<kBF16>(w0[ch * kK0 + j]);
    const float2 mr = mean_rstd[(size_t)clip * kC + ch];
    a[c] = mr.y * gamma[ch];                                      // y = (h - mean) * rstd * gamma + beta
    b[c] = beta[ch] - mr.x * a[c];
  }
  __syncthreads();
  ActT* o = out + ((size_t)off0[clip] + f0) * kC + 2 * tid;
  for (int f = 0; f < nf; ++f) {
    float h0 = 0.f, h1 = 0.f;
#pragma unroll
    for (int j = 0; j < kK0; ++j) { const float x = sx[f * kS0 + j]; h0 = fmaf(w[0][j], x, h0); h1 = fmaf(w[1][j], x, h1); }
    h0 = geluf_(fmaf(r16<kBF16>(h0), a[0], b[0]));                // GroupNorm and GELU run in fp32 under autocast
    h1 = geluf_(fmaf(r16<kBF16>(h1), a[1], b[1]));
    if constexpr (sizeof(ActT) == 2) {
      *reinterpret_cast<__nv_bfloat162*>(o + (size_t)f * kC) = __floats2bfloat162_rn(h0, h1);
    } else {
      *reinterpret_cast<float2*>(o + (size_t)f * kC) = make_float2(h0, h1);
    }
  }
}

// warp-per-row LayerNorm helpers: a row of W = 32 * 4 * NV values lives in NV float4 per lane
template <int NV>
B2T_DEVICE void ln_row(float4 (&v)[NV], const float* __restrict__ w, const float* __restrict__ b, int lane) {
  constexpr float inv = 1.0f / (128.0f * NV);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mu = warp_sum(sum) * inv;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j].x -= mu; v[j].y -= mu; v[j].z -= mu; v[j].w -= mu;
    sq += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float rstd = rsqrtf(warp_sum(sq) * inv + 1e-5f);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w != nullptr) { g = reinterpret_cast<const float4*>(w)[lane + 32 * j]; be = reinterpret_cast<const float4*>(b)[lane + 32 * j]; }
    v[j].x = v[j].x * rstd * g.x + be.x; v[j].y = v[j].y * rstd * g.y + be.y;
    v[j].z = v[j].z * rstd * g.z + be.z; v[j].w = v[j].w * rstd * g.w + be.w;
  }
}
template <typename T> B2T_DEVICE float4 ld4(const T* p);
template <> B2T_DEVICE float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> B2T_DEVICE float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
  return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}
template <typename T> B2T_DEVICE void st4(T* p, float4 v);
template <> B2T_DEVICE void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> B2T_DEVICE void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}

// valid frames of level 6 -> compact rows, LayerNorm(512) (feature_projection.layer_norm)
template <typename ActT>
__global__ void __launch_bounds__(256)
hubert_gather_ln_kernel(const ActT* __restrict__ act6, const int32_t* __restrict__ off0, const int32_t* __restrict__ row_off,
                        const int32_t* __restrict__ valid_rows, int n_clips, int rows, const float* __restrict__ w,
                        const float* __restrict__ b, ActT* __restrict__ out, float* __restrict__ tap_feats) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int clip = find_segment(row_off, n_clips, r);
  const bool padded = r - row_off[clip] >= valid_rows[clip];     // no feature row exists (and none is needed) for it
  const ActT* src = act6 + ((size_t)(off0[clip] >> 6) + (r - row_off[clip])) * kC;
  float4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = padded ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4<ActT>(src + 4 * (lane + 32 * j));
  if (tap_feats != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) st4<float>(tap_feats + (size_t)r * kC + 4 * (lane + 32 * j), v[j]);
  }
  ln_row<4>(v, w, b, lane);
#pragma unroll
  for (int j = 0; j < 4; ++j) st4<ActT>(out + (size_t)r * kC + 4 * (lane + 32 * j), v[j]);
}

// compact rows [M, 768] -> group-major zero-gapped buffer [16][pos_rows][48] (gaps are zeroed by a memset)
template <typename ActT>
__global__ void __launch_bounds__(256)
hubert_pos_scatter_kernel(const ActT* __restrict__ x, const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                          const int32_t* __restrict__ pos_off, int n_clips, int rows, int pos_rows, ActT* __restrict__ pbuf) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int clip = find_segment(row_off, n_clips, r);
  if (r - row_off[clip] >= valid_rows[clip]) return;             // padded frame: zero after the projection (:430-433)
  const size_t prow = (size_t)pos_off[clip] + (r - row_off[clip]);
  for (int c = threadIdx.x; c < kHid; c += blockDim.x) {
    const int g = c / kGC, j = c - g * kGC;
    pbuf[((size_t)g * pos_rows + prow) * kGC + j] = x[(size_t)r * kHid + c];
  }
}

// h = LayerNorm(x + pos) (encoder.layer_norm): pos comes from the 16 group GEMMs ([pos_rows, 16 * 128], GELU applied)
template <typename ActT, bool kBF16>
__global__ void __launch_bounds__(256)
hubert_pos_finish_kernel(const ActT* __restrict__ x, const ActT* __restrict__ pos_out, const int32_t* __restrict__ row_off,
                         const int32_t* __restrict__ valid_rows, const int32_t* __restrict__ pos_off, int n_clips, int rows, const float* __restrict__ w,
                         const float* __restrict__ b, float* __restrict__ h, ActT* __restrict__ hact) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int clip = find_segment(row_off, n_clips, r);
  // GEMM row R reads buffer rows R .. R + 127 = taps 0 .. 127 <-> frames t - 64 .. t + 63  =>  R = pos_off + t - 64
  const ActT* prow = pos_out + ((size_t)pos_off[clip] + (r - row_off[clip]) - kPosK / 2) * (kGroups * kPosN);
  const bool padded = r - row_off[clip] >= valid_rows[clip];     // x = 0 there; the positional conv still sees its neighbours
  float4 v[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int c = 4 * (lane + 32 * j);                           // 4 consecutive channels never straddle a group (48 % 4 == 0)
    const int g = c / kGC, cj = c - g * kGC;
    const float4 xv = padded ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4<ActT>(x + (size_t)r * kHid + c);
    const float4 pv = ld4<ActT>(prow + g * kPosN + cj);
    v[j] = make_float4(r16<kBF16>(xv.x + pv.x), r16<kBF16>(xv.y + pv.y), r16<kBF16>(xv.z + pv.z), r16<kBF16>(xv.w + pv.w));
  }
  ln_row<6>(v, w, b, lane);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    st4<float>(h + (size_t)r * kHid + 4 * (lane + 32 * j), v[j]);
    st4<ActT>(hact + (size_t)r * kHid + 4 * (lane + 32 * j), v[j]);
  }
}

// post-LN residual step: h <- LayerNorm(h + d; w, b); hact <- act(h).  w == nullptr: affine-free LayerNorm of h into
// `plain` (the tail in front of the quantiser), h untouched.
template <typename ActT>
__global__ void __launch_bounds__(256)
hubert_add_ln_kernel(float* __restrict__ h, const ActT* __restrict__ d, const float* __restrict__ w, const float* __restrict__ b,
                     ActT* __restrict__ hact, float* __restrict__ plain, int rows) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float4 v[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    v[j] = ld4<float>(h + (size_t)r * kHid + 4 * (lane + 32 * j));
    if (d != nullptr) {
      const float4 dv = ld4<ActT>(d + (size_t)r * kHid + 4 * (lane + 32 * j));
      v[j].x += dv.x; v[j].y += dv.y; v[j].z += dv.z; v[j].w += dv.w;
    }
  }
  ln_row<6>(v, w, b, lane);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    if (plain != nullptr) {
      st4<float>(plain + (size_t)r * kHid + 4 * (lane + 32 * j), v[j]);
    } else {
      st4<float>(h + (size_t)r * kHid + 4 * (lane + 32 * j), v[j]);
      st4<ActT>(hact + (size_t)r * kHid + 4 * (lane + 32 * j), v[j]);
    }
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Ws {
  float* wave_norm; float2* wave_stats; float2* partial; float2* mean_rstd; void* lvl[7]; void* lnf; void* hp; void* pbuf; void* pos_out; float* h; void* hact;
  void* big; void* att; void* d; float* plain; void* vq; size_t vq_bytes; size_t total;
};

Ws carve(void* base, const b2t_hubert_batch* b, int K, int precision) {
  const size_t act = precision == B2T_PREC_BF16 ? 2 : 4;
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? (void*)(p + off) : nullptr; off += align_up(bytes, 256); return r; };
  const size_t M = (size_t)b->total_rows;
  Ws w;
  w.wave_norm = (float*)take((size_t)b->total_samples * 4);
  w.wave_stats = (float2*)take((size_t)b->n_clips * 8);
  w.partial = (float2*)take((size_t)b->n_stat_tiles * kC * 8);
  w.mean_rstd = (float2*)take((size_t)b->n_clips * kC * 8);
  for (int l = 0; l < 7; ++l) w.lvl[l] = take((((size_t)b->level0_rows >> l) + 4) * kC * act);   // + slack: the last GEMM row reads k rows
  w.lnf = take(M * kC * act);
  w.hp = take(M * kHid * act);
  w.pbuf = take(((size_t)kGroups * b->pos_rows + kPosK) * kGC * act);
  w.pos_out = take((size_t)b->pos_rows * kGroups * kPosN * act);
  w.h = (float*)take(M * kHid * 4);
  w.hact = take(M * kHid * act);
  w.big = take(M * 3072 * act);
  w.att = take(M * kHid * act);
  w.d = take(M * kHid * act);
  w.plain = (float*)take(M * kHid * 4);
  w.vq_bytes = b2t_vq_workspace_bytes((int)M, kHid, K);
  w.vq = take(w.vq_bytes);
  w.total = off;
  return w;
}

}  // namespace

extern "C" b2t_hubert_model* b2t_hubert_create(int n_layers, int codebook_size, int precision) {
  if (n_layers < 0 || n_layers > 48 || codebook_size < 1 || codebook_size > 32768 ||
      (precision != B2T_PREC_BF16 && precision != B2T_PREC_FP32)) {
    b2t_set_error("b2t_hubert_create: bad arguments (n_layers=%d K=%d precision=%d)", n_layers, codebook_size, precision);
    return nullptr;
  }
  auto* m = new b2t_hubert_model();
  m->n_layers = n_layers; m->codebook_size = codebook_size; m->precision = precision;
  return m;
}

extern "C" void b2t_hubert_destroy(b2t_hubert_model* m) { delete m; }

extern "C" int b2t_hubert_set_tensor(b2t_hubert_model* m, const char* name, const void* ptr) {
  B2T_REQUIRE(m && name && ptr, B2T_ERR_ARG, "b2t_hubert_set_tensor: null argument");
  m->t[std::string(name)] = ptr;
  return B2T_OK;
}

extern "C" size_t b2t_hubert_workspace_bytes(const b2t_hubert_model* m, const b2t_hubert_batch* b) {
  if (!m || !b) return 0;
  return carve(nullptr, b, m->codebook_size, m->precision).total;
}

extern "C" int b2t_hubert_encode(const b2t_hubert_model* m, const float* wave, const b2t_hubert_batch* b, void* workspace,
                                 size_t workspace_bytes, int normalize, int16_t* tokens, int tap_layer, float* tap_out,
                                 float* tap_feats, void* stream) {
  B2T_REQUIRE(m && wave && b && workspace && tokens, B2T_ERR_ARG, "b2t_hubert_encode: null argument");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  b2t_reset_launch_count();
  const int M = b->total_rows;
  if (M <= 0 || b->n_clips <= 0) return B2T_OK;
  B2T_REQUIRE(b->level0_rows % 64 == 0 && b->pos_rows >= kPosK, B2T_ERR_ARG, "b2t_hubert_encode: bad level tables");
  Ws w = carve(workspace, b, m->codebook_size, m->precision);
  B2T_REQUIRE(workspace_bytes >= w.total, B2T_ERR_WORKSPACE, "b2t_hubert_encode: workspace %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  const int prec = m->precision;
  const bool bf = prec == B2T_PREC_BF16;
  const size_t act = bf ? 2 : 4;
  bool missing = false;
  std::string miss_name;
  auto T = [&](const std::string& n) -> const void* {
    auto it = m->t.find(n);
    if (it == m->t.end()) { if (!missing) miss_name = n; missing = true; return nullptr; }
    return it->second;
  };
#define RUN(call) do { int rc__ = (call); if (rc__ != B2T_OK) return rc__; } while (0)
#define NEED() B2T_REQUIRE(!missing, B2T_ERR_STATE, "b2t_hubert_encode: tensor '%s' not set", miss_name.c_str())
  auto gemm = [&](const void* A, int lda, int rows, const void* W, const void* bias, void* out, int ldo, int N, int K, int epi) -> int {
    b2t_gemm_args g{};
    g.A = A; g.lda = lda; g.W = W; g.bias = (const float*)bias; g.out = out; g.ldo = ldo; g.M = rows; g.N = N; g.K = K;
    g.epilogue = epi; g.alpha = 1.f; g.precision = prec; g.impl = bf ? B2T_IMPL_AUTO : B2T_IMPL_SIMT;
    return b2t_gemm(&g, stream);
  };

  // ---- per-clip waveform normalisation (skipped when the caller hands in processor output, as the reference's
  //      encoder operator receives it)
  const int64_t* woff = b->wave_off;
  if (normalize) {
    hubert_wave_stats_kernel<<<b->n_clips, 1024, 0, st>>>(wave, b->wave_off, b->n_samples, w.wave_stats);
    B2T_LAUNCH_CHECK();
    hubert_wave_norm_kernel<<<dim3(64, b->n_clips), 256, 0, st>>>(wave, b->wave_off, b->norm_off, b->n_samples, w.wave_stats, w.wave_norm);
    B2T_LAUNCH_CHECK();
    wave = w.wave_norm;
    woff = b->norm_off;
  }
  // ---- feature encoder
  const float* w0 = (const float*)T("fe.conv0.w"); const float* gw = (const float*)T("fe.gn.w"); const float* gb = (const float*)T("fe.gn.b");
  NEED();
  if (bf) {
    hubert_conv0_stats_kernel<true><<<b->n_stat_tiles, 256, 0, st>>>(wave, woff, b->n_samples, b->gn_count, b->stat_tile_clip,
                                                                     b->stat_tile_f0, w0, w.partial);
  } else {
    hubert_conv0_stats_kernel<false><<<b->n_stat_tiles, 256, 0, st>>>(wave, woff, b->n_samples, b->gn_count, b->stat_tile_clip,
                                                                      b->stat_tile_f0, w0, w.partial);
  }
  B2T_LAUNCH_CHECK();
  hubert_gn_finalize_kernel<<<b->n_clips, kC, 0, st>>>(w.partial, b->stat_tile_first, b->gn_count, w.mean_rstd);
  B2T_LAUNCH_CHECK();
  if (b->n_apply_tiles > 0) {
    if (bf) {
      hubert_conv0_apply_kernel<__nv_bfloat16, true><<<b->n_apply_tiles, 256, 0, st>>>(
          wave, woff, b->n_samples, b->off0, b->apply_tile_clip, b->apply_tile_f0, w0, w.mean_rstd, gw, gb, (__nv_bfloat16*)w.lvl[0]);
    } else {
      hubert_conv0_apply_kernel<float, false><<<b->n_apply_tiles, 256, 0, st>>>(
          wave, woff, b->n_samples, b->off0, b->apply_tile_clip, b->apply_tile_f0, w0, w.mean_rstd, gw, gb, (float*)w.lvl[0]);
    }
    B2T_LAUNCH_CHECK();
  }
  static const int kTaps[7] = {10, 3, 3, 3, 3, 2, 2};
  for (int l = 1; l <= 6; ++l) {
    const void* wl = T("fe.conv" + std::to_string(l) + ".w");
    NEED();
    RUN(gemm(w.lvl[l - 1], 2 * kC, b->level0_rows >> l, wl, nullptr, w.lvl[l], kC, kC, kTaps[l] * kC, B2T_EPI_BIAS_GELU));
  }
  // ---- feature projection on the valid frames
  {
    const float* lw = (const float*)T("fp.ln.w"); const float* lb = (const float*)T("fp.ln.b");
    const void* pw = T("fp.proj.w"); const void* pb = T("fp.proj.b");
    NEED();
    if (bf) hubert_gather_ln_kernel<__nv_bfloat16><<<(M + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)w.lvl[6], b->off0, b->row_off, b->attn.valid_rows, b->n_clips, M, lw, lb, (__nv_bfloat16*)w.lnf, tap_feats);
    else hubert_gather_ln_kernel<float><<<(M + 7) / 8, 256, 0, st>>>((const float*)w.lvl[6], b->off0, b->row_off, b->attn.valid_rows, b->n_clips, M, lw, lb, (float*)w.lnf, tap_feats);
    B2T_LAUNCH_CHECK();
    RUN(gemm(w.lnf, kC, M, pw, pb, w.hp, kHid, kHid, kC, B2T_EPI_BIAS));
  }
  // ---- positional conv embedding, x + pos, encoder.layer_norm
  {
    const void* pw = T("pos.w"); const float* pb = (const float*)T("pos.b");
    const float* lw = (const float*)T("enc.ln.w"); const float* lb = (const float*)T("enc.ln.b");
    NEED();
    B2T_CUDA(cudaMemsetAsync(w.pbuf, 0, ((size_t)kGroups * b->pos_rows + kPosK) * kGC * act, st));
    if (bf) hubert_pos_scatter_kernel<__nv_bfloat16><<<M, 256, 0, st>>>((const __nv_bfloat16*)w.hp, b->row_off, b->attn.valid_rows, b->pos_off, b->n_clips, M, b->pos_rows, (__nv_bfloat16*)w.pbuf);
    else hubert_pos_scatter_kernel<float><<<M, 256, 0, st>>>((const float*)w.hp, b->row_off, b->attn.valid_rows, b->pos_off, b->n_clips, M, b->pos_rows, (float*)w.pbuf);
    B2T_LAUNCH_CHECK();
    const int prows = b->pos_rows - kPosK + 1;                   // GEMM rows whose 128-row window lies inside the buffer
    for (int g = 0; g < kGroups; ++g)
      RUN(gemm((const uint8_t*)w.pbuf + (size_t)g * b->pos_rows * kGC * act, kGC, prows,
               (const uint8_t*)pw + (size_t)g * kPosN * kPosK * kGC * act, pb + g * kPosN,
               (uint8_t*)w.pos_out + (size_t)g * kPosN * act, kGroups * kPosN, kPosN, kPosK * kGC, B2T_EPI_BIAS_GELU));
    if (bf) hubert_pos_finish_kernel<__nv_bfloat16, true><<<(M + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)w.hp, (const __nv_bfloat16*)w.pos_out, b->row_off, b->attn.valid_rows, b->pos_off, b->n_clips, M, lw, lb, w.h, (__nv_bfloat16*)w.hact);
    else hubert_pos_finish_kernel<float, false><<<(M + 7) / 8, 256, 0, st>>>((const float*)w.hp, (const float*)w.pos_out, b->row_off, b->attn.valid_rows, b->pos_off, b->n_clips, M, lw, lb, w.h, (float*)w.hact);
    B2T_LAUNCH_CHECK();
  }
  auto add_ln = [&](const void* d, const float* lw, const float* lb, float* plain) -> int {
    if (bf) hubert_add_ln_kernel<__nv_bfloat16><<<(M + 7) / 8, 256, 0, st>>>(w.h, (const __nv_bfloat16*)d, lw, lb, (__nv_bfloat16*)w.hact, plain, M);
    else hubert_add_ln_kernel<float><<<(M + 7) / 8, 256, 0, st>>>(w.h, (const float*)d, lw, lb, (float*)w.hact, plain, M);
    B2T_LAUNCH_CHECK();
    return B2T_OK;
  };
  if (tap_layer == 0 && tap_out) B2T_CUDA(cudaMemcpyAsync(tap_out, w.h, (size_t)M * kHid * 4, cudaMemcpyDeviceToDevice, st));
  // ---- transformer (post-LN): hidden_states[i] = input of layer i
  const void* zero_bias = T("attn.zero_bias");
  NEED();
  for (int i = 0; i < m->n_layers; ++i) {
    const std::string L = "L" + std::to_string(i) + ".";
    const void* wqkv = T(L + "attn.wqkv"); const void* bqkv = T(L + "attn.bqkv"); const void* wo = T(L + "attn.wo"); const void* bo = T(L + "attn.bo");
    const float* l1w = (const float*)T(L + "ln.w"); const float* l1b = (const float*)T(L + "ln.b");
    const void* w1 = T(L + "ffn.w1"); const void* b1 = T(L + "ffn.b1"); const void* w2 = T(L + "ffn.w2"); const void* b2 = T(L + "ffn.b2");
    const float* l2w = (const float*)T(L + "final.ln.w"); const float* l2b = (const float*)T(L + "final.ln.b");
    NEED();
    RUN(gemm(w.hact, kHid, M, wqkv, bqkv, w.big, 3 * kHid, 3 * kHid, kHid, B2T_EPI_BIAS));
    RUN(b2t_attention(w.big, zero_bias, &b->attn, w.att, kHid / 64, prec, bf ? B2T_IMPL_AUTO : B2T_IMPL_SIMT, stream));
    RUN(gemm(w.att, kHid, M, wo, bo, w.d, kHid, kHid, kHid, B2T_EPI_BIAS));
    RUN(add_ln(w.d, l1w, l1b, nullptr));
    RUN(gemm(w.hact, kHid, M, w1, b1, w.big, 3072, 3072, kHid, B2T_EPI_BIAS_GELU));
    RUN(gemm(w.big, 3072, M, w2, b2, w.d, kHid, kHid, 3072, B2T_EPI_BIAS));
    RUN(add_ln(w.d, l2w, l2b, nullptr));
    if (tap_layer == i + 1 && tap_out) B2T_CUDA(cudaMemcpyAsync(tap_out, w.h, (size_t)M * kHid * 4, cudaMemcpyDeviceToDevice, st));
  }
  // ---- tail: affine-free LayerNorm -> nearest k-means centre (reference encoder.py:96-103)
  const void* cb = T("codebook");
  NEED();
  RUN(add_ln(nullptr, nullptr, nullptr, w.plain));
  RUN(b2t_vq_argmin(w.plain, kHid, M, kHid, (const float*)cb, nullptr, m->codebook_size, 0, bf ? B2T_IMPL_AUTO : B2T_IMPL_SIMT, tokens,
                    nullptr, w.vq, w.vq_bytes, stream));
  return B2T_OK;
#undef RUN
#undef NEED
}
