// Audio ingest on the device: PCM16 / fp32 decode, mono mix-down and sinc resampling in one pass.
// Replaces reference audiotoken/utils.py:26-44 `convert_audio` (stereo -> mean over channels, then
// torchaudio.transforms.Resample with its defaults) and the per-chunk Resample of the streaming reader (:98-99).
// torchaudio's formulation: rates reduced by their gcd to (orig, new); output sample j = n*new + i is the correlation of
// phase filter i (2*width + orig taps, most of them exactly zero) with the input window starting at n*orig - width.
// The host strips each phase filter to its non-zero support (`start[i]`, `count[i]`, <= 2*width + 2 taps); one thread
// per output sample, filter taps through the read-only cache, neighbouring threads read neighbouring input samples.
// HBM-bound: (2 or 4) * channels bytes in per input sample, 4 bytes out per output sample.
#include "common.cuh"

namespace {

template <typename T> B2T_DEVICE float pcm_to_float(T v);
template <> B2T_DEVICE float pcm_to_float<float>(float v) { return v; }
template <> B2T_DEVICE float pcm_to_float<int16_t>(int16_t v) { return (float)v * (1.0f / 32768.0f); }

template <typename T>
__global__ void __launch_bounds__(256)
resample_kernel(const T* __restrict__ in, long long in_len, int channels, long long ch_stride, long long t_stride,
                const float* __restrict__ taps, const int32_t* __restrict__ start, const int32_t* __restrict__ count,
                int max_taps, int orig, int new_, int width, float* __restrict__ out, long long out_len) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= out_len) return;
  const long long n = j / new_;
  const int i = (int)(j - n * new_);
  const int st = __ldg(start + i), cnt = __ldg(count + i);
  const float* h = taps + (size_t)i * max_taps;
  const long long base = n * orig - width + st;
  float acc = 0.f;
  for (int m = 0; m < cnt; ++m) {
    const long long t = base + m;
    if (t < 0 || t >= in_len) continue;
    float x = pcm_to_float<T>(in[t * t_stride]);
    if (channels == 2) x = (x + pcm_to_float<T>(in[t * t_stride + ch_stride])) * 0.5f;   // torch.mean over 2 channels
    acc = fmaf(__ldg(h + m), x, acc);
  }
  out[j] = acc;
}

}  // namespace

extern "C" int b2t_ingest_resample(const void* in, int in_is_int16, long long in_len, int channels, long long ch_stride,
                                   long long t_stride, const float* taps, const int32_t* start, const int32_t* count,
                                   int max_taps, int orig, int new_, int width, float* out, long long out_len,
                                   void* stream) {
  B2T_REQUIRE(in && taps && start && count && out, B2T_ERR_ARG, "b2t_ingest_resample: null argument");
  B2T_REQUIRE(channels == 1 || channels == 2, B2T_ERR_ARG, "b2t_ingest_resample: only mono or stereo audio is supported");
  B2T_REQUIRE(orig > 0 && new_ > 0 && max_taps > 0 && in_len >= 0 && out_len >= 0, B2T_ERR_ARG, "b2t_ingest_resample: bad sizes");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (out_len == 0) return B2T_OK;
  const long long blocks = (out_len + 255) / 256;
  B2T_REQUIRE(blocks < (1LL << 31), B2T_ERR_ARG, "b2t_ingest_resample: output too long");
  cudaStream_t st = (cudaStream_t)stream;
  if (in_is_int16)
    resample_kernel<int16_t><<<(unsigned)blocks, 256, 0, st>>>((const int16_t*)in, in_len, channels, ch_stride, t_stride, taps,
                                                              start, count, max_taps, orig, new_, width, out, out_len);
  else
    resample_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)in, in_len, channels, ch_stride, t_stride, taps, start,
                                                            count, max_taps, orig, new_, width, out, out_len);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
