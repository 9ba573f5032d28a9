// Residual VQ (EnCodec quantizer.encode; reference audiotoken/encoder.py:50-52, mirrored by transformers
// models/encodec/modeling_encodec.py:364-369, 424-438) on tensor cores with an exact result.
//
// One CTA owns 128 frames for all n_q stages.  Each of the 128 "row" threads keeps its frame's fp32 residual in
// registers.  Per stage:
//   1. the row threads split the residual into an error-compensated bf16 pair (hi, lo) and write it as two K-major
//      SWIZZLE_128B operand tiles into shared memory;
//   2. warp 1 issues tcgen05.mma for  r.E^T ~= hi.Ehi + lo.Ehi + hi.Elo  against the stage's codebook, streamed by
//      warp 0 through a TMA ring in 256-code tiles; accumulators are double-buffered in TMEM;
//   3. the row threads read the scores with tcgen05.ld, subtract 0.5|e|^2 and keep the best two (score, index) pairs
//      and the third score with branch-free selects — the [frames, 1024] matrix never exists;
//   4. the winner is certified with the error bound delta (|score - exact| <= delta): if the runner-up is more than
//      2*delta behind, it is the exact argmin; otherwise the two candidates are re-scored in fp64, and if even the
//      third score is inside the band the row is re-scanned exhaustively (about 1e-4 of the rows);
//   5. r -= E[winner] in fp32, exactly as the reference does.
// Clusters of 4 CTAs share the codebook stream: each CTA fetches a quarter of every tile and TMA-multicasts it to all
// four, so the L2 -> SM traffic (512 KB per stage and CTA otherwise) drops by 4x.
#include "tc_ptx.cuh"
#include "gemm_epilogue.cuh"
#include "vq_cand.cuh"

namespace {

// one row per lane: 256-bit loads fetch whole 32-byte sectors (a 16-byte load per lane uses half of each)
B2T_DEVICE void ldg8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

constexpr int kRows = 128, kD = 128, kCodes = 1024, kBN = 256;
constexpr int kRingStages = 3;
constexpr int kCluster = 4;                       // CTAs sharing one codebook stream (TMA multicast)
constexpr int kTileA = kRows * kBK * 2;          // 16 KB: one 64-wide k-block of hi or lo
constexpr int kTileB = kBN * kBK * 2;            // 32 KB
constexpr int kSmemA = 4 * kTileA;               // hi kb0, hi kb1, lo kb0, lo kb1
constexpr int kScratch = 4 * kD * 4;              // per row warp: one residual row for the cooperative slow path
constexpr int kSmemBytes = kSmemA + kRingStages * kTileB + 256 + kScratch + 1024;
constexpr int kThreadsRvq = 64 + 128;
constexpr float kRelEps = 6.103515625e-5f * 1.0625f;   // 2^-14 (bf16x3, see vq.cu) + 2^-18 key truncation

B2T_DEVICE double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Exact decision for ONE row whose fast scores could not certify the winner, executed by the whole warp.
// rs: the row's fp32 residual in shared memory (128 floats).  kind 0: the two leaders i1, i2 (different chunks, nothing
// else inside the band); kind 1: all 32 codes of chunk `chunk`; kind 2: exhaustive scan with an fp32 filter.
__device__ __noinline__ int rvq_resolve_warp(const float* __restrict__ rs, const float* __restrict__ E,
                                             const float* __restrict__ hn, int kind, int i1, int i2, int chunk,
                                             float delta, int lane) {
  if (kind == 0) {
    // lanes split the 128 dimensions, fp64 partial sums, warp reduction
    const float4 rv = *reinterpret_cast<const float4*>(rs + 4 * lane);
    const float4 e1 = __ldg(reinterpret_cast<const float4*>(E + (size_t)i1 * kD + 4 * lane));
    const float4 e2 = __ldg(reinterpret_cast<const float4*>(E + (size_t)i2 * kD + 4 * lane));
    double a = 0.0, b = 0.0, t;
    t = (double)rv.x - (double)e1.x; a += t * t; t = (double)rv.y - (double)e1.y; a += t * t;
    t = (double)rv.z - (double)e1.z; a += t * t; t = (double)rv.w - (double)e1.w; a += t * t;
    t = (double)rv.x - (double)e2.x; b += t * t; t = (double)rv.y - (double)e2.y; b += t * t;
    t = (double)rv.z - (double)e2.z; b += t * t; t = (double)rv.w - (double)e2.w; b += t * t;
    a = warp_sum_d(a); b = warp_sum_d(b);
    return (b < a || (b == a && i2 < i1)) ? i2 : i1;
  }
  // lanes split the codes; every lane scores its codes against the whole residual
  const int n_codes = kind == 1 ? 32 : kCodes;
  const int first = kind == 1 ? chunk * 32 : 0;
  double bd = INFINITY;
  int bi = 0x7fffffff;
  float run = -INFINITY;
  if (kind == 2) {
    // fp32 scores first: only codes within 2*delta' of the best fp32 score can win (delta' covers fp32 rounding too)
    for (int c = lane; c < n_codes; c += 32) {
      const float* e = E + (size_t)c * kD;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
      for (int d = 0; d < kD; d += 4) {
        const float4 ev = __ldg(reinterpret_cast<const float4*>(e + d));
        const float4 rv = *reinterpret_cast<const float4*>(rs + d);
        a0 = fmaf(rv.x, ev.x, a0); a1 = fmaf(rv.y, ev.y, a1); a2 = fmaf(rv.z, ev.z, a2); a3 = fmaf(rv.w, ev.w, a3);
      }
      run = fmaxf(run, (a0 + a1) + (a2 + a3) - __ldg(hn + c));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) run = fmaxf(run, __shfl_xor_sync(0xffffffffu, run, o));
  }
  for (int c = first + lane; c < first + n_codes; c += 32) {
    const float* e = E + (size_t)c * kD;
    if (kind == 2) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
      for (int d = 0; d < kD; d += 4) {
        const float4 ev = __ldg(reinterpret_cast<const float4*>(e + d));
        const float4 rv = *reinterpret_cast<const float4*>(rs + d);
        a0 = fmaf(rv.x, ev.x, a0); a1 = fmaf(rv.y, ev.y, a1); a2 = fmaf(rv.z, ev.z, a2); a3 = fmaf(rv.w, ev.w, a3);
      }
      if (!((a0 + a1) + (a2 + a3) - __ldg(hn + c) >= run - 2.0f * delta)) continue;
    }
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
    for (int d = 0; d < kD; d += 4) {
      const float4 ev = __ldg(reinterpret_cast<const float4*>(e + d));
      const float4 rv = *reinterpret_cast<const float4*>(rs + d);
      const double t0 = (double)rv.x - (double)ev.x, t1 = (double)rv.y - (double)ev.y;
      const double t2 = (double)rv.z - (double)ev.z, t3 = (double)rv.w - (double)ev.w;
      a0 += t0 * t0; a1 += t1 * t1; a0 += t2 * t2; a1 += t3 * t3;
    }
    const double dd = a0 + a1;
    if (dd < bd) { bd = dd; bi = c; }            // ascending c per lane: the first index wins ties
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
  }
  return bi;
}

__global__ void __launch_bounds__(kThreadsRvq, 1)
rvq_tc_kernel(const __grid_constant__ CUtensorMap map_c2 /* [n_q*1024, 256] bf16 = hi | lo */,
              const float* __restrict__ emb, int rows, const float* __restrict__ codebooks,
              const float* __restrict__ half_norm, const float* __restrict__ cmax_half, int n_q,
              int16_t* __restrict__ codes, unsigned int* __restrict__ stats /* [2]: fp64 re-scores, re-scans */, int dbg) {
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

  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < kRingStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kCluster); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    mbar_init(aready_bar, 4);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_c2);
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  // barrier inits must be visible cluster-wide before a peer multicasts into / arrives on this CTA
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer: per 256-code tile the k-blocks  Ehi[0:64], Ehi[64:128], Elo[0:64], Elo[64:128];
    //       this CTA fetches codes [rank*64, rank*64+64) of the tile and multicasts them to the whole cluster =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int q = 0; q < n_q; ++q)
        for (int j = 0; j < kTilesPerStage; ++j)
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);       // all CTAs of the cluster released this slot
            mbar_expect_tx(full_bar(stage), kTileB);
            tma_load_2d_mc(sB + stage * kTileB + cta_rank * (kTileB / kCluster), &map_c2, full_bar(stage), kb * kBK,
                           q * kCodes + j * kBN + (int)cta_rank * (kBN / kCluster), (uint16_t)((1u << kCluster) - 1));
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
            umma_commit_mc(empty_bar(stage), (uint16_t)((1u << kCluster) - 1));
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
      for (int d = 0; d < kD; d += 8) {
        float v[8];
        ldg8(emb + (size_t)row * kD + d, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) r[d + e] = v[e];
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
      // Per 32-code chunk the best two scores are tracked with 3 min/max per element on KEYS: the score with its low
      // 5 mantissa bits replaced by (31 - position), so value and position travel in one register (the 2^-18
      // relative truncation is added to delta).  Chunk results are merged into the best two (key, chunk) pairs and
      // the third-best key g3 of the stage.
      float g1 = -INFINITY, g2 = -INFINITY, g3 = -INFINITY;
      int q1 = 0, q2 = 0;
      for (int j = 0; j < kTilesPerStage; ++j, ++g) {
        const int acc = g & 1;
        mbar_wait(tfull_bar(acc), (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kBN);
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          if (dbg & 2) break;
          uint32_t v[32];
          tmem_ld_32x32_nowait(taddr + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          const int chunk = j * (kBN / 32) + c;
          const float* hc = hn + chunk * 32;
          float c1 = -INFINITY, c2 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 h = __ldg(reinterpret_cast<const float4*>(hc + i));
            const float sc[4] = {__uint_as_float(v[i]) - h.x, __uint_as_float(v[i + 1]) - h.y,
                                 __uint_as_float(v[i + 2]) - h.z, __uint_as_float(v[i + 3]) - h.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float k = __uint_as_float((__float_as_uint(sc[u]) & 0xFFFFFFE0u) | (uint32_t)(31 - i - u));
              const float t = fminf(c1, k);
              c1 = fmaxf(c1, k);
              c2 = fmaxf(c2, t);
            }
          }
          // merge (c1 >= c2) into (g1 >= g2 >= g3); strict '>' keeps the earlier chunk among equal keys
          {
            const bool a1 = c1 > g1, a2 = c1 > g2;
            const float n3 = a2 ? g2 : fmaxf(g3, c1);
            const float n2 = a1 ? g1 : (a2 ? c1 : g2);
            const int nq2 = a1 ? q1 : (a2 ? chunk : q2);
            g1 = a1 ? c1 : g1; q1 = a1 ? chunk : q1;
            g2 = n2; q2 = nq2; g3 = n3;
            const bool b2 = c2 > g2;                       // c2 <= c1, so it can at best become the runner-up
            g3 = b2 ? g2 : fmaxf(g3, c2);
            q2 = b2 ? chunk : q2;
            g2 = b2 ? c2 : g2;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      Cand cand = cand_empty();
      cand.v1 = g1; cand.i1 = q1 * 32 + 31 - (int)(__float_as_uint(g1) & 31u);
      cand.v2 = g2; cand.i2 = q2 * 32 + 31 - (int)(__float_as_uint(g2) & 31u);
      cand.v3 = g3; cand.i3 = (q1 == q2) ? q1 : -1;        // both leaders in one chunk: its third is unknown
      // 4. certify / re-score
      const float* E = codebooks + (size_t)q * kCodes * kD;
      const float ch = __ldg(cmax_half + q);
      const float delta = kRelEps * (sqrtf(rr) * sqrtf(2.0f * ch) + ch);
      int best = cand.i1;
      {
        // rows the bound cannot certify are resolved exactly, one at a time, by the whole warp
        const bool need = live && !(dbg & 1) && cand.v2 >= cand.v1 - 2.0f * delta;
        unsigned int todo = __ballot_sync(0xffffffffu, need);
        float* rs = reinterpret_cast<float*>(base_ptr + kSmemA + kRingStages * kTileB + 256) + (warp - 2) * kD;
        while (todo) {
          const int owner = __ffs(todo) - 1;
          todo &= todo - 1;
          if (lane == owner) {
#pragma unroll
            for (int d = 0; d < kD; d += 4) *reinterpret_cast<float4*>(rs + d) = make_float4(r[d], r[d + 1], r[d + 2], r[d + 3]);
          }
          const int kind_l = (cand.v3 >= cand.v1 - 2.0f * delta) ? 2 : (cand.i3 >= 0 ? 1 : 0);
          const int kind = __shfl_sync(0xffffffffu, kind_l, owner);
          const int o1 = __shfl_sync(0xffffffffu, cand.i1, owner), o2 = __shfl_sync(0xffffffffu, cand.i2, owner);
          const int oc = __shfl_sync(0xffffffffu, cand.i3, owner);
          const float od = __shfl_sync(0xffffffffu, delta, owner);
          __syncwarp();
          const int res = rvq_resolve_warp(rs, E, hn, kind, o1, o2, oc, od, lane);
          if (lane == owner) { best = res; if (stats) atomicAdd(stats + (kind == 2 ? 1 : 0), 1u); }
          __syncwarp();
        }
      }
      if (live) codes[(size_t)q * rows + row] = (int16_t)best;
      // 5. residual update in fp32, as the reference: r = r - E[idx]
      if (q + 1 < n_q && !(dbg & 4)) {
        const float* e = E + (size_t)(live ? best : 0) * kD;
#pragma unroll
        for (int d = 0; d < kD; d += 8) {
          float v[8];
          ldg8(e + d, v);
#pragma unroll
          for (int u = 0; u < 8; ++u) r[d + u] -= v[u];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  // peers may still multicast into / arrive on this CTA's shared memory until they are done too
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

int g_rvq_dbg = 0;   // b2t_set_option("rvq_dbg", bits): timing experiments only (results are wrong when set)

// c2: bf16 [n_q_total * 1024, 256] (hi | lo of the fp32 codebooks); stats: optional device uint32[2]
int b2t_rvq_tensor(const float* emb, int rows, const void* c2, int n_q_total, const float* codebooks,
                   const float* half_norm, const float* cmax_half, int n_q, int16_t* codes, unsigned int* stats,
                   cudaStream_t st) {
  if (rows <= 0) return B2T_OK;
  CUtensorMap map;
  int rc = make_map(&map, c2, n_q_total * kCodes, 2 * kD, 2 * kD, kBN / kCluster);
  if (rc != B2T_OK) return rc;
  B2T_SMEM_OPT_IN(kSmemBytes, rvq_tc_kernel);
  const int ctas = ((rows + kRows - 1) / kRows + kCluster - 1) / kCluster * kCluster;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(ctas); lc.blockDim = dim3(kThreadsRvq); lc.dynamicSmemBytes = kSmemBytes; lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  lc.attrs = attr; lc.numAttrs = 1;
  B2T_CUDA(cudaLaunchKernelEx(&lc, rvq_tc_kernel, map, emb, rows, codebooks, half_norm, cmax_half, n_q, codes, stats, g_rvq_dbg));
  b2t_count_launch();
  return B2T_OK;
}
