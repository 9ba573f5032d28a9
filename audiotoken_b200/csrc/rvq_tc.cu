// Residual VQ (EnCodec quantizer.encode; reference audiotoken/encoder.py:50-52, mirrored by transformers
// models/encodec/modeling_encodec.py:364-369, 424-438) on tensor cores with an exact result.
//
// One CTA owns 128 frames for all n_q stages.  Each of the 128 "row" threads keeps its frame's fp32 residual in
// registers.  Per stage:
//   1. the row threads split the residual into an error-compensated bf16 pair (hi, lo) and write it as two K-major
//      SWIZZLE_128B operand tiles into shared memory;
//   2. warp 1 issues tcgen05.mma for  r.E^T ~= hi.Ehi + lo.Ehi + hi.Elo  against the stage's codebook, streamed by
//      warp 0 through a TMA ring in 256-code tiles; accumulators are double-buffered in TMEM;
//   3. the row threads read the scores with tcgen05.ld, subtract 0.5|e|^2 and keep the best three and the fourth
//      score (chunks whose maximum cannot enter the list are skipped) — the [frames, 1024] matrix never exists;
//   4. the winner is certified with the error bound delta (|score - exact| <= delta): if the runner-up is more than
//      2*delta behind, it is the exact argmin; otherwise the candidates inside the band are re-scored in fp64, and
//      if even the fourth score is inside the band the row is re-scanned exhaustively (never observed);
//   5. r -= E[winner] in fp32, exactly as the reference does.
#include "tc_ptx.cuh"
#include "gemm_epilogue.cuh"
#include "vq_cand.cuh"

namespace {

constexpr int kRows = 128, kD = 128, kCodes = 1024, kBN = 256;
constexpr int kRingStages = 3;
constexpr int kTileA = kRows * kBK * 2;          // 16 KB: one 64-wide k-block of hi or lo
constexpr int kTileB = kBN * kBK * 2;            // 32 KB
constexpr int kSmemA = 4 * kTileA;               // hi kb0, hi kb1, lo kb0, lo kb1
constexpr int kSmemBytes = kSmemA + kRingStages * kTileB + 256 + 1024;
constexpr int kThreadsRvq = 64 + 128;
constexpr float kRelEps = 6.103515625e-5f;       // 2^-14 (bf16x3, see vq.cu)

__device__ __noinline__ double exact_dist128(const float* __restrict__ r, const float* __restrict__ e) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 8
  for (int d = 0; d < kD; d += 4) {
    const float4 ev = __ldg(reinterpret_cast<const float4*>(e + d));
    const double t0 = (double)r[d] - (double)ev.x, t1 = (double)r[d + 1] - (double)ev.y;
    const double t2 = (double)r[d + 2] - (double)ev.z, t3 = (double)r[d + 3] - (double)ev.w;
    a0 += t0 * t0; a1 += t1 * t1; a2 += t2 * t2; a3 += t3 * t3;
  }
  return (a0 + a1) + (a2 + a3);
}

// Rare path (the runner-up's fast score is within 2*delta of the best): exact decision in fp64.  `r` is a copy of
// the row's residual in local memory, so the hot loop keeps its registers.
__device__ __noinline__ int rvq_resolve(const float* __restrict__ r, const float* __restrict__ E,
                                        const float* __restrict__ hn, Cand cand, float delta,
                                        unsigned int* __restrict__ stats) {
  int best = cand.i1;
  if (!(cand.v4 >= cand.v1 - 2.0f * delta)) {
    if (stats) atomicAdd(stats, 1u);
    double bd = exact_dist128(r, E + (size_t)cand.i1 * kD);
    const double d2 = exact_dist128(r, E + (size_t)cand.i2 * kD);
    if (d2 < bd || (d2 == bd && cand.i2 < best)) { bd = d2; best = cand.i2; }
    if (cand.v3 >= cand.v1 - 2.0f * delta) {
      const double d3 = exact_dist128(r, E + (size_t)cand.i3 * kD);
      if (d3 < bd || (d3 == bd && cand.i3 < best)) { bd = d3; best = cand.i3; }
    }
    return best;
  }
  if (stats) atomicAdd(stats + 1, 1u);
  float run = -INFINITY;
  double bd = INFINITY;
  best = 0;
  for (int code = 0; code < kCodes; ++code) {
    const float* e = E + (size_t)code * kD;
    float a = 0.f;
#pragma unroll 8
    for (int d = 0; d < kD; ++d) a = fmaf(r[d], __ldg(e + d), a);
    a -= __ldg(hn + code);
    if (a >= run - 2.0f * delta) {
      const double dd = exact_dist128(r, e);
      if (dd < bd) { bd = dd; best = code; }
    }
    run = fmaxf(run, a);
  }
  return best;
}

__global__ void __launch_bounds__(kThreadsRvq, 1)
rvq_tc_kernel(const __grid_constant__ CUtensorMap map_c2 /* [n_q*1024, 256] bf16 = hi | lo */,
              const float* __restrict__ emb, int rows, const float* __restrict__ codebooks,
              const float* __restrict__ half_norm, const float* __restrict__ cmax_half, int n_q,
              int16_t* __restrict__ codes, unsigned int* __restrict__ stats /* [2]: fp64 re-scores, re-scans */) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + kSmemA;
  const uint32_t bars = base + kSmemA + kRingStages * kTileB;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kRingStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kRingStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kRingStages + 2 + a); };
  const uint32_t aready_bar = bars + 8u * (2 * kRingStages + 4);
  const uint32_t tmem_slot = bars + 8u * (2 * kRingStages + 5);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + kSmemA + kRingStages * kTileB + 8 * (2 * kRingStages + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTilesPerStage = kCodes / kBN;     // 4

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRingStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    mbar_init(aready_bar, 4);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_c2);
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer: per 256-code tile the k-blocks  Ehi[0:64], Ehi[64:128], Elo[0:64], Elo[64:128] =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int q = 0; q < n_q; ++q)
        for (int j = 0; j < kTilesPerStage; ++j)
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), kTileB);
            tma_load_2d(sB + stage * kTileB, &map_c2, full_bar(stage), kb * kBK, q * kCodes + j * kBN);
            if (++stage == kRingStages) { stage = 0; phase ^= 1u; }
          }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kRows, kBN);
      int stage = 0; uint32_t phase = 0;
      int g = 0;
      for (int q = 0; q < n_q; ++q) {
        mbar_wait(aready_bar, (uint32_t)(q & 1));          // operand tiles of this stage are in shared memory
        tc_fence_after();
        for (int j = 0; j < kTilesPerStage; ++j, ++g) {
          const int acc = g & 1;
          mbar_wait(tempty_bar(acc), (uint32_t)(((g >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kBN);
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint64_t db = make_smem_desc(sB + stage * kTileB);
            const int kk = kb & 1;
            const uint64_t da_hi = make_smem_desc(sA + kk * kTileA);
            const uint64_t da_lo = make_smem_desc(sA + (2 + kk) * kTileA);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16(tmem_d, da_hi + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            if (kb < 2) {                                  // codebook hi also meets the residual's lo part
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                umma_bf16(tmem_d, da_lo + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
            }
            umma_commit(empty_bar(stage));
            if (++stage == kRingStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(tfull_bar(acc));
        }
      }
    }
  } else {
    // ===== row threads: TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    const int rl = quad * 32 + lane;                       // row inside the CTA = TMEM lane
    const int row = blockIdx.x * kRows + rl;
    const bool live = row < rows;
    float r[kD];
    if (live) {
#pragma unroll
      for (int d = 0; d < kD; d += 4) {
        const float4 v = *reinterpret_cast<const float4*>(emb + (size_t)row * kD + d);
        r[d] = v.x; r[d + 1] = v.y; r[d + 2] = v.z; r[d + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int d = 0; d < kD; ++d) r[d] = 0.f;
    }
    const uint32_t row_off = (uint32_t)((rl >> 3) * 1024 + (rl & 7) * 128);
    int g = 0;
    for (int q = 0; q < n_q; ++q) {
      // 1. split the residual: hi = bf16(r), lo = bf16(r - hi)  ->  swizzled K-major operand tiles
      float rr = 0.f;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x0 = r[kb * 64 + c * 8 + 2 * i], x1 = r[kb * 64 + c * 8 + 2 * i + 1];
            rr = fmaf(x0, x0, rr); rr = fmaf(x1, x1, rr);
            const float h0 = bf16_round(x0), h1 = bf16_round(x1);
            hi[i] = pack2_bf16(h0, h1);
            lo[i] = pack2_bf16(x0 - h0, x1 - h1);
          }
          const uint32_t off = row_off + (uint32_t)((c ^ (rl & 7)) * 16);
          *reinterpret_cast<uint4*>(base_ptr + kb * kTileA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(base_ptr + (2 + kb) * kTileA + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(aready_bar);
      // 2./3. scores of the 4 code tiles -> best three + fourth score
      const float* hn = half_norm + (size_t)q * kCodes;
      Cand cand = cand_empty();
      for (int j = 0; j < kTilesPerStage; ++j, ++g) {
        const int acc = g & 1;
        mbar_wait(tfull_bar(acc), (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kBN);
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32_nowait(taddr + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          const int col = j * kBN + c * 32;
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 h = __ldg(reinterpret_cast<const float4*>(hn + col + i));
            const float s0 = __uint_as_float(v[i]) - h.x, s1 = __uint_as_float(v[i + 1]) - h.y;
            const float s2 = __uint_as_float(v[i + 2]) - h.z, s3 = __uint_as_float(v[i + 3]) - h.w;
            v[i] = __float_as_uint(s0); v[i + 1] = __float_as_uint(s1); v[i + 2] = __float_as_uint(s2); v[i + 3] = __float_as_uint(s3);
            mx = fmaxf(mx, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
          }
          if (mx > cand.v4) {
#pragma unroll
            for (int i = 0; i < 32; ++i) cand_insert_ordered(cand, __uint_as_float(v[i]), col + i);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      // 4. certify / re-score
      const float* E = codebooks + (size_t)q * kCodes * kD;
      const float ch = __ldg(cmax_half + q);
      const float delta = kRelEps * (sqrtf(rr) * sqrtf(2.0f * ch) + ch);
      int best = cand.i1;
      if (live && cand.v2 >= cand.v1 - 2.0f * delta) {
        float rc[kD];
#pragma unroll
        for (int d = 0; d < kD; ++d) rc[d] = r[d];
        best = rvq_resolve(rc, E, hn, cand, delta, stats);
      }
      if (live) codes[(size_t)q * rows + row] = (int16_t)best;
      // 5. residual update in fp32, as the reference: r = r - E[idx]
      if (q + 1 < n_q) {
        const float* e = E + (size_t)(live ? best : 0) * kD;
#pragma unroll
        for (int d = 0; d < kD; d += 4) {
          const float4 ev = __ldg(reinterpret_cast<const float4*>(e + d));
          r[d] -= ev.x; r[d + 1] -= ev.y; r[d + 2] -= ev.z; r[d + 3] -= ev.w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

// c2: bf16 [n_q_total * 1024, 256] (hi | lo of the fp32 codebooks); stats: optional device uint32[2]
int b2t_rvq_tensor(const float* emb, int rows, const void* c2, int n_q_total, const float* codebooks,
                   const float* half_norm, const float* cmax_half, int n_q, int16_t* codes, unsigned int* stats,
                   cudaStream_t st) {
  if (rows <= 0) return B2T_OK;
  CUtensorMap map;
  int rc = make_map(&map, c2, n_q_total * kCodes, 2 * kD, 2 * kD, kBN);
  if (rc != B2T_OK) return rc;
  static bool cfg = false;
  if (!cfg) { B2T_CUDA(cudaFuncSetAttribute(rvq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)); cfg = true; }
  rvq_tc_kernel<<<(rows + kRows - 1) / kRows, kThreadsRvq, kSmemBytes, st>>>(map, emb, rows, codebooks, half_norm, cmax_half,
                                                                          n_q, codes, stats);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
