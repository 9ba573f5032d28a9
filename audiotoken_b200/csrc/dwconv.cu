// Conformer convolution-module middle (HF modeling_wav2vec2_bert.py:213-221):
//   causal depthwise conv1d (k=31, left pad 30 zeros, per clip) -> LayerNorm(1024) -> swish.
// Work item = (clip, 64-row time tile) processed as four 16-row sub-tiles; 512 threads, each owns a channel
// pair: 31x2 weights (loaded once per work item, tap-major [31][1024] layout => coalesced) and 16x2
// accumulators live in registers, the 46 input rows of a sub-tile are streamed (coalesced 2 KB rows, the
// 30-row halo re-read hits L1/L2).  The LayerNorm needs the whole
// 1024-channel row, so row statistics go through a two-level (warp shuffle, shared memory) reduction.
// HBM-bound: 2 KB (bf16) read + 2 KB written per row.
#include "common.cuh"

namespace {

constexpr int kC = 1024, kK = 31, kTT = 16, kSub = 4, kThreads = 512;   // work item = kSub * kTT = 64 rows

template <typename T> B2T_DEVICE float2 ld2(const T* p);
template <> B2T_DEVICE float2 ld2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> B2T_DEVICE float2 ld2<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <typename T> B2T_DEVICE void st2(T* p, float a, float b);
template <> B2T_DEVICE void st2<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <> B2T_DEVICE void st2<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// packed fp32 FMA (Blackwell FFMA2): both channels of the thread in one instruction
B2T_DEVICE float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

// Sum 16 per-lane values over the 32 lanes of a warp with a halving butterfly: after step s every lane keeps
// half of the remaining values (31 shuffles instead of 16 x 5).  On return lane l holds in v[0] the total of
// value index (l & 15) — the bit-reversed bookkeeping is folded into which half is kept.
B2T_DEVICE float warp_sum16(float (&v)[16], int lane) {
  // step 1: lanes with bit 4 clear keep values 0..7, the others 8..15
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? v[i] : v[i + 8];
      const float keep = up ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4];
      const float keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2];
      const float keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];   // lane l: total of value index ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1)
}
B2T_DEVICE int warp_sum16_index(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

template <typename T, bool kBF16>
__global__ void __launch_bounds__(kThreads)
dwconv_ln_swish_kernel(const T* __restrict__ x, const float* __restrict__ w_dw,
                       const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                       const int32_t* __restrict__ row_off, const int32_t* __restrict__ ctile_clip,
                       const int32_t* __restrict__ ctile_t0, T* __restrict__ out) {
  __shared__ float s_red[16][kTT];
  __shared__ float s_stat[kTT];
  const int clip = ctile_clip[blockIdx.x], tile0 = ctile_t0[blockIdx.x];
  const int r0 = row_off[clip], rows = row_off[clip + 1] - r0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = tid * 2;

  float2 wv[kK];
#pragma unroll
  for (int k = 0; k < kK; ++k) {
    const float2 wk = __ldg(reinterpret_cast<const float2*>(w_dw + (size_t)k * kC + c));   // [31][1024]
    wv[k] = make_float2(r16<kBF16>(wk.x), r16<kBF16>(wk.y));
  }
  const float g0 = __ldg(ln_w + c), g1 = __ldg(ln_w + c + 1);
  const float b0 = __ldg(ln_b + c), b1 = __ldg(ln_b + c + 1);
#pragma unroll 1
  for (int sub = 0; sub < kSub; ++sub) {
  const int t0 = tile0 + sub * kTT;
  if (t0 >= rows) break;
  float2 av[kTT];
#pragma unroll
  for (int t = 0; t < kTT; ++t) av[t] = make_float2(0.f, 0.f);

  // input row (t0 - 30 + j), j = 0..45, contributes to output t with tap k = j - t  (0 <= k <= 30)
#pragma unroll
  for (int j = 0; j < kTT + kK - 1; ++j) {
    const int tr = t0 - (kK - 1) + j;
    float2 v = make_float2(0.f, 0.f);
    if (tr >= 0 && tr < rows) v = ld2<T>(x + (size_t)(r0 + tr) * kC + c);
#pragma unroll
    for (int t = 0; t < kTT; ++t) {
      const int k = j - t;
      if (k >= 0 && k < kK) av[t] = ffma2(wv[k], v, av[t]);
    }
  }
  // LayerNorm over channels, all 16 rows at once (two-pass: mean, then centred sum of squares)
  float a0[kTT], a1[kTT], part[kTT];
#pragma unroll
  for (int t = 0; t < kTT; ++t) {
    a0[t] = r16<kBF16>(av[t].x);
    a1[t] = r16<kBF16>(av[t].y);
    part[t] = a0[t] + a1[t];
  }
  const int ridx = warp_sum16_index(lane);
  float tot = warp_sum16(part, lane);
  if ((lane & 1) == 0) s_red[warp][ridx] = tot;
  __syncthreads();
  if (tid < kTT) {
    float s = 0.f;
#pragma unroll
    for (int wi = 0; wi < 16; ++wi) s += s_red[wi][tid];
    s_stat[tid] = s * (1.0f / kC);
  }
  __syncthreads();
  float mu[kTT];
#pragma unroll
  for (int t = 0; t < kTT; ++t) {
    mu[t] = s_stat[t];
    float d0 = a0[t] - mu[t], d1 = a1[t] - mu[t];
    part[t] = d0 * d0 + d1 * d1;
  }
  tot = warp_sum16(part, lane);
  __syncthreads();
  if ((lane & 1) == 0) s_red[warp][ridx] = tot;
  __syncthreads();
  if (tid < kTT) {
    float s = 0.f;
#pragma unroll
    for (int wi = 0; wi < 16; ++wi) s += s_red[wi][tid];
    s_stat[tid] = rsqrtf(s * (1.0f / kC) + 1e-5f);
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < kTT; ++t) {
    if (t0 + t < rows) {
      const float rs = s_stat[t];
      float y0 = (a0[t] - mu[t]) * rs * g0 + b0;
      float y1 = (a1[t] - mu[t]) * rs * g1 + b1;
      y0 = swishf_(y0);
      y1 = swishf_(y1);
      st2<T>(out + (size_t)(r0 + t0 + t) * kC + c, y0, y1);
    }
  }
  __syncthreads();   // s_stat / s_red are reused by the next sub-tile
  }
}

// ---- bf16 production kernel: rows staged once through a shared-memory ring by bulk async copies ----------------
// One persistent CTA per SM owns a CONTIGUOUS range of 64-row work items (clip order), so consecutive items of a
// clip continue in the ring and every input row is fetched from HBM once (30 extra rows only where a chain starts).
// Thread 0 issues cp.async.bulk (global -> shared, mbarrier complete_tx) for the 16 rows the NEXT sub-tile needs
// while all 512 threads run the 31-tap packed-FMA loop of the current one out of shared memory; the load latency the
// register-only kernel exposed is gone and the kernel runs at the fp32 FMA rate (31 FMA per output element).
constexpr int kRing = 64;
constexpr int kRowBytes = kC * 2;
constexpr int kRingSmem = kRing * kRowBytes + 64;

B2T_DEVICE void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
B2T_DEVICE void mbar_init_(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
B2T_DEVICE void mbar_expect_(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
B2T_DEVICE void mbar_wait_(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) b2t_trap_report("dwconv_ring mbar_wait ran out of its bound (bar smem address, parity)", bar, parity);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
dwconv_ring_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w_dw,
                   const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                   const int32_t* __restrict__ row_off, const int32_t* __restrict__ ctile_clip,
                   const int32_t* __restrict__ ctile_t0, int n_items, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ float s_red[16][kTT];
  __shared__ float s_stat[kTT];
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t bar_pro = ring_s + kRing * kRowBytes, bar_in0 = bar_pro + 8, bar_in1 = bar_pro + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = tid * 2;
  if (tid == 0) {
    mbar_init_(bar_pro, 1); mbar_init_(bar_in0, 1); mbar_init_(bar_in1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float2 wv[kK];
#pragma unroll
  for (int k = 0; k < kK; ++k) {
    const float2 wk = __ldg(reinterpret_cast<const float2*>(w_dw + (size_t)k * kC + c));
    wv[k] = make_float2(bf16_round(wk.x), bf16_round(wk.y));
  }
  const float g0 = __ldg(ln_w + c), g1 = __ldg(ln_w + c + 1);
  const float b0 = __ldg(ln_b + c), b1 = __ldg(ln_b + c + 1);
  __syncthreads();

  const int per = (n_items + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(n_items, i0 + per);
  int prev_clip = -1, prev_t0 = -1;
  uint32_t n_pro = 0, n_issued = 0, n_waited = 0;      // barrier use counts (identical in every thread)
#pragma unroll 1
  for (int item = i0; item < i1; ++item) {
    const int clip = ctile_clip[item], tile0 = ctile_t0[item];
    const int r0 = row_off[clip], rows = row_off[clip + 1] - r0;
    const bool chained = (clip == prev_clip && tile0 == prev_t0 + kSub * kTT);
    const bool next_chained = (item + 1 < i1) && ctile_clip[item + 1] == clip;
    prev_clip = clip; prev_t0 = tile0;
    if (!chained) {
      __syncthreads();                                   // nobody still reads the ring
      if (tid == 0) {
        const int first = max(0, tile0 - (kK - 1)), last = min(rows, tile0 + kTT);
        const uint32_t bytes = (uint32_t)(last - first) * kRowBytes;
        mbar_expect_(bar_pro, bytes);
        // ring slot of clip row t: (t - tile0 + 32) & 63; rows first..last-1 are contiguous in memory and in the ring
        bulk_g2s(ring_s + (uint32_t)((first - tile0 + 32) & (kRing - 1)) * kRowBytes, x + (size_t)(r0 + first) * kC, bytes, bar_pro);
      }
      if (tile0 == 0) {
        // rows -30..-1 of the clip are the causal zero padding: slots 2..31 of the ring
        uint4* z = reinterpret_cast<uint4*>(ring + 2 * kRowBytes);
        for (int i = tid; i < 30 * kRowBytes / 16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
      }
      mbar_wait_(bar_pro, n_pro & 1u);
      ++n_pro;
    }
#pragma unroll 1
    for (int sub = 0; sub < kSub; ++sub) {
      const int T = tile0 + sub * kTT;
      if (T >= rows) break;
      // prefetch the 16 rows the next sub-tile adds (also across a chained item boundary)
      const bool want_next = (T + kTT < rows) && (sub + 1 < kSub || next_chained);
      if (want_next) {
        if (tid == 0) {
          const int n = min(kTT, rows - (T + kTT));
          const uint32_t bar = (n_issued & 1u) ? bar_in1 : bar_in0;
          mbar_expect_(bar, (uint32_t)n * kRowBytes);
          bulk_g2s(ring_s + (uint32_t)((sub * kTT + kTT + 32) & (kRing - 1)) * kRowBytes, x + (size_t)(r0 + T + kTT) * kC,
                   (uint32_t)n * kRowBytes, bar);
        }
      }
      if (sub > 0 || chained) {                          // rows T..T+15 arrived with the chunk issued one step ago
        mbar_wait_((n_waited & 1u) ? bar_in1 : bar_in0, (n_waited >> 1) & 1u);
        ++n_waited;
      }
      if (want_next) ++n_issued;

      float2 av[kTT];
#pragma unroll
      for (int t = 0; t < kTT; ++t) av[t] = make_float2(0.f, 0.f);
      // input row (T - 30 + j), j = 0..45, contributes to output t with tap k = j - t  (0 <= k <= 30)
#pragma unroll
      for (int j = 0; j < kTT + kK - 1; ++j) {
        const uint32_t slot = (uint32_t)((sub * kTT + 2 + j) & (kRing - 1));
        const uint32_t raw = *reinterpret_cast<const uint32_t*>(ring + slot * kRowBytes + tid * 4);
        const float2 v = make_float2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xFFFF0000u));   // bf16x2 -> fp32
#pragma unroll
        for (int t = 0; t < kTT; ++t) {
          const int k = j - t;
          if (k >= 0 && k < kK) av[t] = ffma2(wv[k], v, av[t]);
        }
      }
      // LayerNorm over channels, all 16 rows at once (two-pass), then swish
      float a0[kTT], a1[kTT], part[kTT];
#pragma unroll
      for (int t = 0; t < kTT; ++t) {
        a0[t] = bf16_round(av[t].x);
        a1[t] = bf16_round(av[t].y);
        part[t] = a0[t] + a1[t];
      }
      const int ridx = warp_sum16_index(lane);
      float tot = warp_sum16(part, lane);
      if ((lane & 1) == 0) s_red[warp][ridx] = tot;
      __syncthreads();
      if (tid < kTT) {
        float sm = 0.f;
#pragma unroll
        for (int wi = 0; wi < 16; ++wi) sm += s_red[wi][tid];
        s_stat[tid] = sm * (1.0f / kC);
      }
      __syncthreads();
      float mu[kTT];
#pragma unroll
      for (int t = 0; t < kTT; ++t) {
        mu[t] = s_stat[t];
        const float d0 = a0[t] - mu[t], d1 = a1[t] - mu[t];
        part[t] = d0 * d0 + d1 * d1;
      }
      tot = warp_sum16(part, lane);
      __syncthreads();
      if ((lane & 1) == 0) s_red[warp][ridx] = tot;
      __syncthreads();
      if (tid < kTT) {
        float sm = 0.f;
#pragma unroll
        for (int wi = 0; wi < 16; ++wi) sm += s_red[wi][tid];
        s_stat[tid] = rsqrtf(sm * (1.0f / kC) + 1e-5f);
      }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < kTT; ++t) {
        if (T + t < rows) {
          const float rs = s_stat[t];
          const float y0 = fmaf(a0[t] - mu[t], rs * g0, b0), y1 = fmaf(a1[t] - mu[t], rs * g1, b1);
          // swish(y) = y * sigmoid(y) = 0.5 y (1 + tanh(0.5 y)): one MUFU per element; the result is rounded to bf16
          float t0, t1;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(0.5f * y0));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(0.5f * y1));
          const float h0 = 0.5f * y0, h1 = 0.5f * y1;
          st2<__nv_bfloat16>(out + (size_t)(r0 + T + t) * kC + c, fmaf(h0, t0, h0), fmaf(h1, t1, h1));
        }
      }
      __syncthreads();   // s_stat / s_red and the ring rows this sub-tile read may be reused from here on
    }
  }
}


// ---- round 2: the same ring, half the instructions ---------------------------------------------------------------
// dwconv_ring_kernel spends 1 723 warp instructions per 16-row sub-tile of which 496 are the packed FMAs (ncu: issue
// slots 61 % busy, 5 block barriers per sub-tile).  This version keeps the ring protocol and the tap loop (same FMA
// order => the same conv values) and rebuilds everything behind it:
//   * every value stays a packed f32x2 (64-bit register pair) from the tap loop to the store: rounding to bf16 is one
//     cvt.rn.bf16x2 + two unpack ops per row, LayerNorm scale/offset and swish are fma.rn.f32x2 / mul.f32x2;
//   * row statistics in ONE pass: per thread (sum a, sum a^2) of its two channels for 16 rows = 32 values, reduced over
//     the warp by a halving butterfly that leaves value l in lane l (31 shuffles), one shared-memory write per lane, ONE
//     block barrier per sub-tile (partials double-buffered by sub-tile parity), then every warp adds the 16 partial
//     vectors itself (lane l = value l): no second barrier, no 16-thread serial stage;
//   * y/2 = a * (rstd * g/2) + (b/2 - mean * rstd * g/2) and swish(y) = y/2 + y/2 * tanh(y/2): 3 packed FMAs + 2 MUFU
//     per row, stores predicated instead of branched.
// var = E[a^2] - mean^2 in fp32 (a is bf16-valued, 1024 channels): the result is rounded to bf16 again, the
// difference to the two-pass form is far below that rounding.
B2T_DEVICE uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
B2T_DEVICE float2 up2(uint64_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
B2T_DEVICE uint64_t fma2p(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
B2T_DEVICE uint64_t mul2p(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// packed bf16x2 word (low half = channel c, high half = channel c + 1) -> f32x2
B2T_DEVICE uint64_t bf2_to_f2(uint32_t raw) {
  uint64_t r;
  asm("{\n\t.reg .b32 lo, hi;\n\tshl.b32 lo, %1, 16;\n\tand.b32 hi, %1, 0xffff0000;\n\tmov.b64 %0, {lo, hi};\n\t}" : "=l"(r) : "r"(raw));
  return r;
}
B2T_DEVICE uint32_t f2_to_bf2(uint64_t v) {
  uint32_t r;
  asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tcvt.rn.bf16x2.f32 %0, hi, lo;\n\t}" : "=r"(r) : "l"(v));
  return r;
}

// kVar: bit 0 = tap loop on scalar FFMA instead of FFMA2; bits 1, 2 = measurement only (no LayerNorm tail / no tap loop)
template <int kVar>
__global__ void __launch_bounds__(kThreads, 1)
dwconv_ring2_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w_dw,
                    const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                    const int32_t* __restrict__ row_off, const int32_t* __restrict__ ctile_clip,
                    const int32_t* __restrict__ ctile_t0, int n_items, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ float s_part[2][16][32];                  // [sub-tile parity][warp][value]: value t = sum a, 16 + t = sum a^2 of row t
  __shared__ float2 s_ab[16][kTT];                     // per warp: (rstd, -mean * rstd) of the 16 rows
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t bar_pro = ring_s + kRing * kRowBytes, bar_in0 = bar_pro + 8, bar_in1 = bar_pro + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = tid * 2;
  if (tid == 0) {
    mbar_init_(bar_pro, 1); mbar_init_(bar_in0, 1); mbar_init_(bar_in1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint64_t wv[kK];
#pragma unroll
  for (int k = 0; k < kK; ++k) {
    const float2 wk = __ldg(reinterpret_cast<const float2*>(w_dw + (size_t)k * kC + c));
    wv[k] = pk2(bf16_round(wk.x), bf16_round(wk.y));
  }
  const uint64_t gh = pk2(0.5f * __ldg(ln_w + c), 0.5f * __ldg(ln_w + c + 1));     // g / 2
  const uint64_t bh = pk2(0.5f * __ldg(ln_b + c), 0.5f * __ldg(ln_b + c + 1));     // b / 2
  const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4, up2_ = lane & 2, up1 = lane & 1;
  __syncthreads();

  const int per = (n_items + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(n_items, i0 + per);
  int prev_clip = -1, prev_t0 = -1;
  uint32_t n_pro = 0, n_issued = 0, n_waited = 0, n_sub = 0;      // barrier use counts (identical in every thread)
#pragma unroll 1
  for (int item = i0; item < i1; ++item) {
    const int clip = ctile_clip[item], tile0 = ctile_t0[item];
    const int r0 = row_off[clip], rows = row_off[clip + 1] - r0;
    const bool chained = (clip == prev_clip && tile0 == prev_t0 + kSub * kTT);
    const bool next_chained = (item + 1 < i1) && ctile_clip[item + 1] == clip;
    prev_clip = clip; prev_t0 = tile0;
    if (!chained) {
      __syncthreads();                                   // nobody still reads the ring
      if (tid == 0) {
        const int first = max(0, tile0 - (kK - 1)), last = min(rows, tile0 + kTT);
        const uint32_t bytes = (uint32_t)(last - first) * kRowBytes;
        mbar_expect_(bar_pro, bytes);
        bulk_g2s(ring_s + (uint32_t)((first - tile0 + 32) & (kRing - 1)) * kRowBytes, x + (size_t)(r0 + first) * kC, bytes, bar_pro);
      }
      if (tile0 == 0) {
        uint4* z = reinterpret_cast<uint4*>(ring + 2 * kRowBytes);
        for (int i = tid; i < 30 * kRowBytes / 16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
      }
      mbar_wait_(bar_pro, n_pro & 1u);
      ++n_pro;
    }
#pragma unroll 1
    for (int sub = 0; sub < kSub; ++sub) {
      const int T = tile0 + sub * kTT;
      if (T >= rows) break;
      // prefetch the 16 rows the next sub-tile adds (also across a chained item boundary); every thread is past the
      // block barrier of the previous sub-tile, i.e. past its last read of the slots this overwrites
      const bool want_next = (T + kTT < rows) && (sub + 1 < kSub || next_chained);
      if (want_next && tid == 0) {
        const int n = min(kTT, rows - (T + kTT));
        const uint32_t bar = (n_issued & 1u) ? bar_in1 : bar_in0;
        mbar_expect_(bar, (uint32_t)n * kRowBytes);
        bulk_g2s(ring_s + (uint32_t)((sub * kTT + kTT + 32) & (kRing - 1)) * kRowBytes, x + (size_t)(r0 + T + kTT) * kC,
                 (uint32_t)n * kRowBytes, bar);
      }
      if (sub > 0 || chained) {                          // rows T..T+15 arrived with the chunk issued one step ago
        mbar_wait_((n_waited & 1u) ? bar_in1 : bar_in0, (n_waited >> 1) & 1u);
        ++n_waited;
      }
      if (want_next) ++n_issued;

      uint64_t av[kTT];
#pragma unroll
      for (int t = 0; t < kTT; ++t) av[t] = 0ull;
      // input row (T - 30 + j), j = 0..45, contributes to output t with tap k = j - t  (0 <= k <= 30)
#pragma unroll
      for (int j = 0; j < kTT + kK - 1; ++j) {
        const uint32_t slot = (uint32_t)((sub * kTT + 2 + j) & (kRing - 1));
        const uint64_t v = bf2_to_f2(*reinterpret_cast<const uint32_t*>(ring + slot * kRowBytes + tid * 4));
        if (kVar & 4) {
          if (j == 0) av[0] = v;
          continue;
        }
#pragma unroll
        for (int t = 0; t < kTT; ++t) {
          const int k = j - t;
          if (k >= 0 && k < kK) {
            if (kVar & 1) {
              const float2 a = up2(av[t]), w = up2(wv[k]), vv = up2(v);
              av[t] = pk2(fmaf(w.x, vv.x, a.x), fmaf(w.y, vv.y, a.y));
            } else {
              av[t] = fma2p(wv[k], v, av[t]);
            }
          }
        }
      }
      if (kVar & 2) {
        __nv_bfloat16* orow = out + (size_t)(r0 + T) * kC + c;
#pragma unroll
        for (int t = 0; t < kTT; ++t)
          if (t < rows - T) *reinterpret_cast<uint32_t*>(orow + (size_t)t * kC) = f2_to_bf2(av[t]);
        __syncthreads();
        continue;
      }
      // conv output rounded to bf16 (the value LayerNorm sees under autocast), statistics, first butterfly step fused:
      // lanes 0..15 keep the sums, lanes 16..31 the sums of squares of the 16 rows
      float w16[16];
#pragma unroll
      for (int t = 0; t < kTT; ++t) {
        av[t] = bf2_to_f2(f2_to_bf2(av[t]));
        const float2 a = up2(av[t]);
        const float s1 = a.x + a.y, s2 = fmaf(a.y, a.y, a.x * a.x);
        w16[t] = (up16 ? s2 : s1) + __shfl_xor_sync(0xffffffffu, up16 ? s1 : s2, 16);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        w16[i] = (up8 ? w16[i + 8] : w16[i]) + __shfl_xor_sync(0xffffffffu, up8 ? w16[i] : w16[i + 8], 8);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        w16[i] = (up4 ? w16[i + 4] : w16[i]) + __shfl_xor_sync(0xffffffffu, up4 ? w16[i] : w16[i + 4], 4);
#pragma unroll
      for (int i = 0; i < 2; ++i)
        w16[i] = (up2_ ? w16[i + 2] : w16[i]) + __shfl_xor_sync(0xffffffffu, up2_ ? w16[i] : w16[i + 2], 2);
      const float mine = (up1 ? w16[1] : w16[0]) + __shfl_xor_sync(0xffffffffu, up1 ? w16[0] : w16[1], 1);
      // lane l now holds the warp's total of value l (l < 16: sum a of row l, l >= 16: sum a^2 of row l - 16)
      const uint32_t par = n_sub & 1u;
      ++n_sub;
      s_part[par][warp][lane] = mine;
      __syncthreads();                                   // the only block barrier of the sub-tile
      float tot = 0.f;
#pragma unroll
      for (int wi = 0; wi < 16; ++wi) tot += s_part[par][wi][lane];
      const float sq = __shfl_down_sync(0xffffffffu, tot, 16);
      if (lane < kTT) {
        const float mean = tot * (1.0f / kC);
        const float var = fmaxf(fmaf(-mean, mean, sq * (1.0f / kC)), 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        s_ab[warp][lane] = make_float2(rstd, -mean * rstd);
      }
      __syncwarp();
      __nv_bfloat16* orow = out + (size_t)(r0 + T) * kC + c;
      const int nvalid = rows - T;
#pragma unroll
      for (int t = 0; t < kTT; ++t) {
        const float2 ab = s_ab[warp][t];
        const uint64_t scale = mul2p(pk2(ab.x, ab.x), gh);           // rstd * g / 2
        const uint64_t off = fma2p(pk2(ab.y, ab.y), gh, bh);         // b / 2 - mean * rstd * g / 2
        const uint64_t h = fma2p(av[t], scale, off);                 // y / 2
        const float2 hf = up2(h);
        float t0, t1;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(hf.x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(hf.y));
        const uint32_t o = f2_to_bf2(fma2p(h, pk2(t0, t1), h));      // swish(y) = y/2 + y/2 * tanh(y/2)
        if (t < nvalid) *reinterpret_cast<uint32_t*>(orow + (size_t)t * kC) = o;
      }
      // s_ab[warp] is rewritten only behind the next sub-tile's block barrier
    }
  }
}

}  // namespace

int b2t_dwconv_mma_launch(const void* x, const float* w_dw, const float* ln_w, const float* ln_b, const b2t_batch* b, void* out,
                          int variant, cudaStream_t st);   // dwconv_mma.cu

// b2t_set_option("dwconv_ring", v): 0 = direct loads, 1 = round-1 ring kernel, 2 = ring + lean LayerNorm tail (default),
// 3 = the same on scalar FFMA, 4-6 = measurement variants (no tail / no taps: wrong results by construction),
// 7 / 8 = tensor-core formulation (dwconv_mma.cu: m16n8k16 / m16n8k8 steps).  Measured per 64 806-row launch, stand-alone:
// 259 / 171 / 162 / 167 us and 149 / 145 us; in the power-capped pipeline the step time is the same with 2 and 8.
int g_dwconv_ring = 2;

extern "C" int b2t_dwconv_ln_swish(const void* x, const float* w_dw, const float* ln_weight,
                                   const float* ln_bias, const b2t_batch* b, void* out,
                                   int precision, void* stream) {
  B2T_REQUIRE(x && w_dw && ln_weight && ln_bias && b && out, B2T_ERR_ARG, "b2t_dwconv_ln_swish: null argument");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (b->n_ctiles <= 0) return B2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == B2T_PREC_BF16 && (g_dwconv_ring == 7 || g_dwconv_ring == 8))
    return b2t_dwconv_mma_launch(x, w_dw, ln_weight, ln_bias, b, out, g_dwconv_ring - 7, st);
  if (precision == B2T_PREC_BF16 && g_dwconv_ring >= 2) {
    int grid = b2t_num_sms();
    if (b->n_ctiles < grid) grid = b->n_ctiles;
#define B2T_RING2(V)                                                                                                  \
  do {                                                                                                                \
    B2T_SMEM_OPT_IN(kRingSmem, dwconv_ring2_kernel<V>);                                                               \
    dwconv_ring2_kernel<V><<<grid, kThreads, kRingSmem, st>>>((const __nv_bfloat16*)x, w_dw, ln_weight, ln_bias,      \
        b->row_off, b->ctile_clip, b->ctile_t0, b->n_ctiles, (__nv_bfloat16*)out);                                    \
  } while (0)
    switch (g_dwconv_ring) {
      case 2: B2T_RING2(0); break;
      case 3: B2T_RING2(1); break;
      case 4: B2T_RING2(2); break;      // measurement variants: wrong results by construction
      case 5: B2T_RING2(3); break;
      case 6: B2T_RING2(4); break;
      default: B2T_REQUIRE(false, B2T_ERR_ARG, "dwconv_ring: unknown variant");
    }
#undef B2T_RING2
  } else if (precision == B2T_PREC_BF16 && g_dwconv_ring) {
    B2T_SMEM_OPT_IN(kRingSmem, dwconv_ring_kernel);
    int grid = b2t_num_sms();
    if (b->n_ctiles < grid) grid = b->n_ctiles;
    dwconv_ring_kernel<<<grid, kThreads, kRingSmem, st>>>((const __nv_bfloat16*)x, w_dw, ln_weight, ln_bias, b->row_off,
                                                         b->ctile_clip, b->ctile_t0, b->n_ctiles, (__nv_bfloat16*)out);
  } else if (precision == B2T_PREC_BF16)
    dwconv_ln_swish_kernel<__nv_bfloat16, true><<<b->n_ctiles, kThreads, 0, st>>>(
        (const __nv_bfloat16*)x, w_dw, ln_weight, ln_bias, b->row_off, b->ctile_clip, b->ctile_t0, (__nv_bfloat16*)out);
  else
    dwconv_ln_swish_kernel<float, false><<<b->n_ctiles, kThreads, 0, st>>>(
        (const float*)x, w_dw, ln_weight, ln_bias, b->row_off, b->ctile_clip, b->ctile_t0, (float*)out);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
