// tcgen05 / TMEM / TMA GEMM for sm_100a:  out = A[M,K] (bf16, K-major) . W[N,K]^T (bf16, K-major)
// with fp32 accumulation in tensor memory and the fused epilogues of gemm_epilogue.cuh.
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0     TMA producer  : cp.async.bulk.tensor 2D loads of a 128x64 A tile and a BNx64 W tile
//                              (SWIZZLE_128B) into a kStages-deep shared-memory ring, mbarrier
//                              complete_tx signalling.
//   warp 1     MMA issuer    : one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                              (M=128, N=BN, K=16) x4 per 64-wide k-block; tcgen05.commit frees the
//                              smem slot / publishes the accumulator.  Also owns the TMEM allocation.
//   warps 2-9  epilogue      : tcgen05.ld 32x32b.x32 (TMEM lane quadrant = warp % 4; the two warps of a
//                              quadrant split the columns), bias / activation / residual / GLU, direct
//                              16-byte global stores.  Two warps per scheduler hide the TMEM/ALU latency.
// The accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Tiles are visited n-fastest so concurrently running CTAs share the A tile in L2.
#include <string>
#include "tc_ptx.cuh"
#include "gemm_epilogue.cuh"
#include "vq_cand.cuh"

constexpr int kEpiArgmax = 100;   // internal epilogue: per-row top-3 of (acc - half_norm[col]) (vq.cu)

namespace {

constexpr int kBM = 128;
constexpr int kEpiWarps = 8;                    // 2 per TMEM lane quadrant, each owns half of the columns
constexpr int kThreads = 64 + 32 * kEpiWarps;

// Residual epilogue with coalesced global traffic.  After tcgen05.ld each lane owns one accumulator
// row (32 consecutive columns); writing that straight to the fp32 stream touches 32 different cache
// lines per instruction.  Instead the delta alpha*r16(acc+bias) is transposed through a per-warp
// 16x32 fp32 shared-memory tile (16-byte chunks XOR-swizzled by row) so that 8 lanes cover one
// 128-byte row segment and a warp instruction moves 4 full lines:  x = resid; x += delta; resid = x.
B2T_DEVICE void resid_chunk_coalesced(const EpiParams& p, float4* stg, int row_base, int col0, int lane,
                                      const float (&acc)[32]) {
  float d[32];
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
      d[i] = p.alpha * bf16_round(acc[i] + b.x); d[i + 1] = p.alpha * bf16_round(acc[i + 1] + b.y);
      d[i + 2] = p.alpha * bf16_round(acc[i + 2] + b.z); d[i + 3] = p.alpha * bf16_round(acc[i + 3] + b.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = p.alpha * bf16_round(acc[i]);
  }
  const int q = lane & 7;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if ((lane >> 4) == half) {
      const int rl = lane & 15;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        stg[rl * 8 + (i ^ (rl & 7))] = make_float4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
    }
    __syncwarp();
    float4 x[4];
    float* g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rl = j * 4 + (lane >> 3);
      const int grow = row_base + half * 16 + rl;
      g[j] = grow < p.M ? p.resid + (size_t)grow * p.N + col0 + q * 4 : nullptr;
      if (g[j]) x[j] = *reinterpret_cast<const float4*>(g[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rl = j * 4 + (lane >> 3);
      const float4 dd = stg[rl * 8 + (q ^ (rl & 7))];
      if (g[j]) {
        x[j].x += dd.x; x[j].y += dd.y; x[j].z += dd.z; x[j].w += dd.w;
        if (p.round_resid) {
          x[j].x = bf16_round(x[j].x); x[j].y = bf16_round(x[j].y); x[j].z = bf16_round(x[j].z); x[j].w = bf16_round(x[j].w);
        }
        *reinterpret_cast<float4*>(g[j]) = x[j];
      }
    }
    __syncwarp();
  }
}

// bf16 epilogues (BIAS / SWISH / GLU): values of 32 accumulator columns -> packed bf16 (uint32 pairs)
#ifndef B2T_GEMM_DIRECT_STORE
#define B2T_GEMM_DIRECT_STORE 1
#endif
#ifndef B2T_GEMM_SWISH_POLY
#define B2T_GEMM_SWISH_POLY 1
#endif
template <int EPI>
B2T_DEVICE void epi_pack32(const EpiParams& p, int col0, const float (&acc)[32], uint32_t* pk) {
  if constexpr (EPI == B2T_EPI_BIAS_SWISH && B2T_GEMM_SWISH_POLY != 0) {
    // two XU-pipe operations per element instead of 3.5: the bf16 rounding of the linear output (an autocast rounding
    // point) is ONE packed conversion per pair, unpacked with a shift and a mask; swish2 = polynomial 2^x + MUFU.RCP
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
      const uint32_t r0 = pack2_bf16(acc[i] + b.x, acc[i + 1] + b.y), r1 = pack2_bf16(acc[i + 2] + b.z, acc[i + 3] + b.w);
      const float2 s0 = swish2(make_float2(__uint_as_float(r0 << 16), __uint_as_float(r0 & 0xffff0000u)));
      const float2 s1 = swish2(make_float2(__uint_as_float(r1 << 16), __uint_as_float(r1 & 0xffff0000u)));
      pk[i >> 1] = pack2_bf16(s0.x, s0.y);
      pk[(i >> 1) + 1] = pack2_bf16(s1.x, s1.y);
    }
    return;
  }
  float v[32];
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
      v[i] = bf16_round(acc[i] + b.x); v[i + 1] = bf16_round(acc[i + 1] + b.y);
      v[i + 2] = bf16_round(acc[i + 2] + b.z); v[i + 3] = bf16_round(acc[i + 3] + b.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round(acc[i]);
  }
  if constexpr (EPI == B2T_EPI_GLU) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      pk[i] = pack2_bf16(v[4 * i] * sigmoidf_(v[4 * i + 1]), v[4 * i + 2] * sigmoidf_(v[4 * i + 3]));
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = v[2 * i], b = v[2 * i + 1];
      if constexpr (EPI == B2T_EPI_BIAS_SWISH) { a = swishf_(a); b = swishf_(b); }
      if constexpr (EPI == B2T_EPI_BIAS_GELU) { a = geluf_(a); b = geluf_(b); }
      pk[i] = pack2_bf16(a, b);
    }
  }
}

// Coalesced store of a 32-row x 128-byte bf16 block held one row per lane (8 x uint4): transposed through the
// per-warp 16 x 128 B staging tile (16-byte chunks XOR-swizzled by row) so that 8 lanes cover one row and every
// store instruction writes 4 full 128-byte lines instead of 32 partial ones.
B2T_DEVICE void store_rows_coalesced(uint4* stg, __nv_bfloat16* out, int ldo, int M, int row_base, int col, int lane,
                                     const uint4 (&pk)[8]) {
  const int q = lane & 7;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if ((lane >> 4) == half) {
      const int rl = lane & 15;
#pragma unroll
      for (int i = 0; i < 8; ++i) stg[rl * 8 + (i ^ (rl & 7))] = pk[i];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rl = j * 4 + (lane >> 3);
      const int grow = row_base + half * 16 + rl;
      const uint4 v = stg[rl * 8 + (q ^ (rl & 7))];
      if (grow < M) *reinterpret_cast<uint4*>(out + (size_t)grow * ldo + col + q * 8) = v;
    }
    __syncwarp();
  }
}

// nearest-centroid epilogue: score = acc - 0.5|c|^2 (p.bias = half norms, +inf beyond the codebook)
B2T_DEVICE void argmax_chunk(const EpiParams& p, Cand& cand, int col0, const float (&acc)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 h = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
    cand_insert_ordered(cand, acc[i] - h.x, col0 + i);
    cand_insert_ordered(cand, acc[i + 1] - h.y, col0 + i + 1);
    cand_insert_ordered(cand, acc[i + 2] - h.z, col0 + i + 2);
    cand_insert_ordered(cand, acc[i + 3] - h.w, col0 + i + 3);
  }
}

template <int BN, bool kPair = false>
struct SmemLayout {
  static constexpr int kStageA = kBM * kBK * 2;                       // 16 KB
  static constexpr int kStageB = (kPair ? BN / 2 : BN) * kBK * 2;     // 32 KB (BN=256), 16 KB per CTA in pair mode
  static constexpr int kStages = (BN == 256 && !kPair) ? 4 : 6;
  static constexpr int kTileBytes = kStages * (kStageA + kStageB);
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kEpiStage = 2048;                 // per epilogue warp: 16 rows x 32 fp32 (transposes)
  static constexpr int kEpiOff = kTileBytes + 256;       // barriers live in [kTileBytes, kTileBytes+256)
  static constexpr int kTotal = kEpiOff + kEpiWarps * kEpiStage + 1024;  // +1024 for manual alignment
};

// kSplit3: A and W hold [hi | lo] bf16 halves (each Kd = K/3 wide); the K loop runs the three products
// hi*hi, lo*hi, hi*lo back to back into one accumulator (error-compensated bf16x3 dot product).
// kMC ("pair mode"): clusters of 2 CTAs (one TPC) run tcgen05.mma.cta_group::2 on a 256 x BN tile: CTA r holds
// rows r*128.. of A and rows r*128.. of the W tile in ITS shared memory, the leader CTA (rank 0) issues the MMAs
// for both, each CTA's TMEM receives its own 128 accumulator rows and each CTA runs its own epilogue.  With one
// CTA per tile the tensor core's operand reads (96 B/clk) plus the TMA fill (96 B/clk) exceed the 128 B/clk of
// shared-memory bandwidth; in pair mode both drop to 64 B/clk.  Barrier protocol: every TMA load signals the
// LEADER's full barrier; the leader's tcgen05.commit (multicast) releases the stage / publishes the accumulator
// in both CTAs; both CTAs' epilogue warps arrive on the leader's tmem-empty barrier.
template <int BN, int EPI, bool kSplit3 = false, bool kMC = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               int K, EpiParams p) {
  using L = SmemLayout<BN, kMC>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + kStages * L::kStageA;
  const uint32_t bars = base + L::kTileBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = p.N / BN;
  const int tiles_m = (p.M + kBM - 1) / kBM;
  uint32_t cta_rank = 0;
  if constexpr (kMC) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  // work items: single tiles, or (kMC) pairs of m-tiles handled by one cluster
  const int w_first = kMC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int w_stride = kMC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int w_count = kMC ? ((tiles_m + 1) / 2) * tiles_n : tiles_m * tiles_n;
  auto tile_m0 = [&](int w) { return (kMC ? 2 * (w / tiles_n) + (int)cta_rank : w / tiles_n) * kBM; };
  auto tile_n0 = [&](int w) { return (w % tiles_n) * BN; };
  const int num_kb = (K + kBK - 1) / kBK;
  constexpr uint32_t kTmemCols = 2 * BN;   // 512 (BN=256) or 256 (BN=128): powers of two

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kMC ? 2 * kEpiWarps : kEpiWarps); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w); }
  if (warp == 1) { if constexpr (kMC) tmem_alloc_pair(tmem_slot, kTmemCols); else tmem_alloc(tmem_slot, kTmemCols); }
  tc_fence_before();
  __syncthreads();
  if constexpr (kMC) {   // barrier inits must be visible to the peer before it multicasts into this CTA
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = w_first; t < w_count; t += w_stride) {
        const int m0 = tile_m0(t), n0 = tile_n0(t);
        for (int kb = 0; kb < num_kb; ++kb) {
          int ka = kb * kBK, kw = kb * kBK;
          if constexpr (kSplit3) {
            const int nkd = num_kb / 3, seg = kb / nkd, j = kb - seg * nkd;
            ka = (seg == 1 ? nkd : 0) * kBK + j * kBK;     // A: hi, lo, hi
            kw = (seg == 2 ? nkd : 0) * kBK + j * kBK;     // W: hi, hi, lo
          }
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if constexpr (kMC) {
            // both CTAs fill their own shared memory but signal the leader's barrier (it expects all 4 loads)
            const uint32_t lead_full = mapa_rank0(full_bar(stage));
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * (L::kStageA + L::kStageB));
            tma_load_2d_pair(sA + stage * L::kStageA, &map_a, lead_full, ka, m0);
            tma_load_2d_pair(sB + stage * L::kStageB, &map_w, lead_full, kw, n0 + (int)cta_rank * (BN / 2));
          } else {
            mbar_expect_tx(full_bar(stage), L::kStageA + L::kStageB);
            tma_load_2d(sA + stage * L::kStageA, &map_a, full_bar(stage), ka, m0);
            tma_load_2d(sB + stage * L::kStageB, &map_w, full_bar(stage), kw, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (pair mode: the leader CTA issues for both) =====
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc(kMC ? 2 * kBM : kBM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = w_first; t < w_count; t += w_stride) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc(sA + stage * L::kStageA);
          const uint64_t db = make_smem_desc(sB + stage * L::kStageB);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the >>4 field
            if constexpr (kMC) umma_bf16_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (kMC) umma_commit_pair(empty_bar(stage), (uint16_t)3);   // frees the slot in both CTAs
          else umma_commit(empty_bar(stage));   // implies tcgen05.fence::before_thread_sync
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if constexpr (kMC) umma_commit_pair(tfull_bar(acc), (uint16_t)3);
        else umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane quadrant = warp % 4; warps sharing a quadrant split the columns =====
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;                       // 0 .. kEpiWarps/4 - 1
    constexpr int kChunks = BN / 32 / (kEpiWarps / 4);      // 32-column chunks per warp
    float4* stg = reinterpret_cast<float4*>(smem_raw + (base - smem_u32(smem_raw)) + L::kEpiOff + (warp - 2) * L::kEpiStage);
    static_assert(kChunks % 2 == 0, "chunks are processed in pairs");
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = w_first; t < w_count; t += w_stride) {
      const int m0 = tile_m0(t), n0 = tile_n0(t);
      if constexpr (EPI == B2T_EPI_RESID) {
        // pull the residual region of the NEXT tile of this CTA into L2 while this tile is processed:
        // the read-modify-write below is otherwise a chain of exposed DRAM round trips
        const int tn = (t == w_first) ? t : t + w_stride;   // first tile: prefetch itself
        for (int tp = tn; tp <= t + w_stride && tp < w_count; tp += w_stride) {
          const int pr = tile_m0(tp) + quad * 32 + lane;
          if (pr < p.M) {
            const float* base_p = p.resid + (size_t)pr * p.N + tile_n0(tp) + part * kChunks * 32;
#pragma unroll
            for (int i = 0; i < kChunks; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(base_p + i * 32));
          }
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + part * kChunks * 32);
      Cand cand = cand_empty();
      uint32_t glu_pk[32];      // packed bf16 outputs of one store unit (32 words = 8 x uint4 = 128 B per row)
#pragma unroll
      for (int c = 0; c < kChunks; c += 2) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32_nowait(taddr + (uint32_t)(c * 32), r0);
        tmem_ld_32x32_nowait(taddr + (uint32_t)(c * 32 + 32), r1);
        tmem_ld_wait();
        if (c + 2 >= kChunks) {
          // the accumulator is now in registers: hand the TMEM buffer back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            #ifdef B2T_GEMM_RELEASE_ARRIVE
            if constexpr (kMC) mbar_arrive_cluster(mapa_rank0(tempty_bar(acc)));
#else
            if constexpr (kMC) mbar_arrive_cluster_relaxed(mapa_rank0(tempty_bar(acc)));
#endif
              // TMEM hand-back: the accumulator reads have completed, nothing in memory to publish (the release form cost a MEMBAR.ALL.GPU per tile: 12 % of the kernel's stall samples)
            else mbar_arrive(tempty_bar(acc));
          }
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]);
        constexpr bool kPacked = (EPI == B2T_EPI_BIAS || EPI == B2T_EPI_BIAS_SWISH || EPI == B2T_EPI_BIAS_GELU || (EPI == B2T_EPI_GLU && BN == 256));
        const int col_a = n0 + (part * kChunks + c) * 32;
        if constexpr (kPacked) {
          // 64 accumulator columns -> 128 B (BIAS/SWISH) or 64 B (GLU) of bf16 per row, stored coalesced
          constexpr int kW = (EPI == B2T_EPI_GLU) ? 8 : 16;          // uint32 words produced per 32 columns
          epi_pack32<EPI>(p, col_a, v, glu_pk + ((EPI == B2T_EPI_GLU) ? c * kW : 0));
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r1[i]);
          epi_pack32<EPI>(p, col_a + 32, v, glu_pk + ((EPI == B2T_EPI_GLU) ? c * kW : 0) + kW);
          if (EPI != B2T_EPI_GLU || c + 2 >= kChunks) {
            uint4 pk4[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk4[i] = make_uint4(glu_pk[4 * i], glu_pk[4 * i + 1], glu_pk[4 * i + 2], glu_pk[4 * i + 3]);
            const int ocol = (EPI == B2T_EPI_GLU) ? (n0 + part * kChunks * 32) / 2 : col_a;
#if B2T_GEMM_DIRECT_STORE
            // row-per-lane 256-bit stores (whole 32-byte sectors): no staging through shared memory, no warp syncs in the
            // epilogue chain.  The epilogue warps never wait for the MMA in the K = 1024 shapes (ncu: no samples on their
            // accumulator-full wait): per-thread latency, not bandwidth, bounds those GEMMs.
            if (row < p.M) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + ocol;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             ::"l"(dst + 16 * i), "r"(pk4[2 * i].x), "r"(pk4[2 * i].y), "r"(pk4[2 * i].z), "r"(pk4[2 * i].w),
                               "r"(pk4[2 * i + 1].x), "r"(pk4[2 * i + 1].y), "r"(pk4[2 * i + 1].z), "r"(pk4[2 * i + 1].w)
                             : "memory");
            }
#else
            store_rows_coalesced(reinterpret_cast<uint4*>(stg), reinterpret_cast<__nv_bfloat16*>(p.out), p.ldo, p.M,
                                 m0 + quad * 32, ocol, lane, pk4);
#endif
          }
        } else {
          if constexpr (EPI == kEpiArgmax) {
            argmax_chunk(p, cand, col_a, v);
          } else if constexpr (EPI == B2T_EPI_RESID) {
            resid_chunk_coalesced(p, stg, m0 + quad * 32, col_a, lane, v);
          } else {
            epilogue_store<EPI, true, 32>(p, row, col_a, v);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r1[i]);
          if constexpr (EPI == kEpiArgmax) {
            argmax_chunk(p, cand, col_a + 32, v);
          } else if constexpr (EPI == B2T_EPI_RESID) {
            resid_chunk_coalesced(p, stg, m0 + quad * 32, col_a + 32, lane, v);
          } else {
            epilogue_store<EPI, true, 32>(p, row, col_a + 32, v);
          }
        }
      }
      if constexpr (EPI == kEpiArgmax) {
        // one partial record per (row, column slice): slice = n-tile * (warps per quadrant) + part
        if (row < p.M)
          reinterpret_cast<Cand*>(p.out)[(size_t)row * p.ldo + (n0 / BN) * (kEpiWarps / 4) + part] = cand;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kMC) {   // the peer may still multicast into / arrive on this CTA's shared memory until it is done too
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kMC) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// cluster-of-2 launch of a multicast kernel instance
template <typename Kern>
int launch_cluster2(Kern kern, int pairs, int smem, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mw, int K,
                    const EpiParams& p) {
  int grid = b2t_num_sms() & ~1;
  if (2 * pairs < grid) grid = 2 * pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  B2T_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mw, K, p));
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}

bool g_multicast = true;   // b2t_set_option("gemm_multicast", 0/1)

template <int BN, int EPI>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mw, const b2t_gemm_args* a, const EpiParams& p, cudaStream_t st) {
  using L = SmemLayout<BN>;
  const int tiles_m = (a->M + kBM - 1) / kBM, tiles_n = a->N / BN;
  if constexpr (BN == 256) {
    if (g_multicast && tiles_m >= 2) {
      using LP = SmemLayout<BN, true>;
      B2T_SMEM_OPT_IN(LP::kTotal, gemm_tc_kernel<BN, EPI, false, true>);
      return launch_cluster2(gemm_tc_kernel<BN, EPI, false, true>, ((tiles_m + 1) / 2) * tiles_n, LP::kTotal, st, ma, mw, a->K, p);
    }
  }
  B2T_SMEM_OPT_IN(L::kTotal, gemm_tc_kernel<BN, EPI>);
  const int tiles = tiles_m * tiles_n;
  int grid = b2t_num_sms();
  if (tiles < grid) grid = tiles;
  gemm_tc_kernel<BN, EPI><<<grid, kThreads, L::kTotal, st>>>(ma, mw, a->K, p);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}

template <int BN>
int dispatch_epi(const CUtensorMap& ma, const CUtensorMap& mw, const b2t_gemm_args* a, const EpiParams& p, cudaStream_t st) {
  switch (a->epilogue) {
    case B2T_EPI_BIAS: return launch_tc<BN, B2T_EPI_BIAS>(ma, mw, a, p, st);
    case B2T_EPI_BIAS_SWISH: return launch_tc<BN, B2T_EPI_BIAS_SWISH>(ma, mw, a, p, st);
    case B2T_EPI_RESID: return launch_tc<BN, B2T_EPI_RESID>(ma, mw, a, p, st);
    case B2T_EPI_GLU: return launch_tc<BN, B2T_EPI_GLU>(ma, mw, a, p, st);
    case B2T_EPI_BIAS_MASK: return launch_tc<BN, B2T_EPI_BIAS_MASK>(ma, mw, a, p, st);
    case B2T_EPI_BIAS_GELU: return launch_tc<BN, B2T_EPI_BIAS_GELU>(ma, mw, a, p, st);
  }
  b2t_set_error("b2t_gemm: unknown epilogue %d", a->epilogue);
  return B2T_ERR_ARG;
}

}  // namespace

// Fast pass of the nearest-centroid search on tensor cores (vq.cu).
//   A2 [M, 2*D] bf16 = x_hi | x_lo,  C2 [K, 2*D] bf16 = c_hi | c_lo,  half_norm [Kpad] (+inf padding),
//   parts [M, slices] Cand with slices = (Kpad / 256) * 2.
int b2t_vq_scan_tensor(const void* A2, const void* C2, int M, int K, int Kpad, int D, const float* half_norm,
                       void* parts, int slices, cudaStream_t st) {
  using L = SmemLayout<256>;
  CUtensorMap ma, mw;
  int rc = make_map(&ma, A2, M, 2 * D, 2 * D, kBM);
  if (rc != B2T_OK) return rc;
  rc = make_map(&mw, C2, K, 2 * D, 2 * D, 256);
  if (rc != B2T_OK) return rc;
  B2T_SMEM_OPT_IN(L::kTotal, gemm_tc_kernel<256, kEpiArgmax, true>);
  EpiParams p{half_norm, parts, slices, nullptr, nullptr, M, Kpad, 1.f, 0};
  const int tiles = ((M + kBM - 1) / kBM) * (Kpad / 256);
  int grid = b2t_num_sms();
  if (tiles < grid) grid = tiles;
  gemm_tc_kernel<256, kEpiArgmax, true><<<grid, kThreads, L::kTotal, st>>>(ma, mw, 3 * D, p);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}

int b2t_test_trap(int light);      // api.cu
extern int g_attn_heads_per_cta;   // attention_tc.cu
extern int g_attn_two_pass;
extern int g_attn_ctas;
extern int g_attn_poly_exp;
extern bool g_rvq_tensor;          // acoustic.cu
extern int g_rvq_dbg;              // rvq_tc.cu
void b2t_seanet_set_sub_frames(int n);   // seanet_tc.cu
void b2t_seanet_set_lstm_pdl(int on);
void b2t_seanet_set_lstm_overlap(int on);
void b2t_seanet_set_l0_fused(int on);
extern int g_dwconv_ring;          // dwconv.cu
extern bool g_ffn_resid_epilogue;  // pipeline.cu

extern "C" int b2t_set_option(const char* name, int value) {
  B2T_REQUIRE(name, B2T_ERR_ARG, "b2t_set_option: null name");
  if (std::string(name) == "gemm_multicast") { g_multicast = value != 0; return B2T_OK; }
  if (std::string(name) == "ffn_resid_epilogue") { g_ffn_resid_epilogue = value != 0; return B2T_OK; }
  if (std::string(name) == "test_trap") { return value ? b2t_test_trap(value == 2) : B2T_OK; }   // 2: through the mapped host record
  if (std::string(name) == "debug_sync") { b2t_set_debug_sync(value); return B2T_OK; }
  if (std::string(name) == "dwconv_ring") { g_dwconv_ring = value; return B2T_OK; }
  if (std::string(name) == "seanet_l0_fused") { b2t_seanet_set_l0_fused(value); return B2T_OK; }
  if (std::string(name) == "lstm_pdl") { b2t_seanet_set_lstm_pdl(value); return B2T_OK; }
  if (std::string(name) == "lstm_overlap") { b2t_seanet_set_lstm_overlap(value); return B2T_OK; }
  if (std::string(name) == "seanet_sub_frames") { b2t_seanet_set_sub_frames(value); return B2T_OK; }
  if (std::string(name) == "rvq_dbg") { g_rvq_dbg = value; return B2T_OK; }
  if (std::string(name) == "rvq_tensor") { g_rvq_tensor = value != 0; return B2T_OK; }
  if (std::string(name) == "attn_two_pass") { g_attn_two_pass = value; return B2T_OK; }
  if (std::string(name) == "attn_ctas") { g_attn_ctas = value; return B2T_OK; }
  if (std::string(name) == "attn_poly_exp") { g_attn_poly_exp = value; return B2T_OK; }
  if (std::string(name) == "attn_heads_per_cta") {
    B2T_REQUIRE(value == 1 || value == 2 || value == 4 || value == 8 || value == 16, B2T_ERR_ARG, "attn_heads_per_cta must divide 16");
    g_attn_heads_per_cta = value;
    return B2T_OK;
  }
  b2t_set_error("b2t_set_option: unknown option '%s'", name);
  return B2T_ERR_ARG;
}

int b2t_gemm_tensor(const b2t_gemm_args* a, const EpiParams& p, cudaStream_t st) {
  B2T_REQUIRE(a->N % 128 == 0, B2T_ERR_ARG, "b2t_gemm(tensor): N must be a multiple of 128 (N=%d)", a->N);
  B2T_REQUIRE(((uintptr_t)a->A % 16) == 0 && ((uintptr_t)a->W % 16) == 0, B2T_ERR_ARG,
              "b2t_gemm(tensor): A and W must be 16-byte aligned");
  CUtensorMap ma, mw;
  const int bn = (a->N % 256 == 0) ? 256 : 128;
  int rc = make_map(&ma, a->A, a->M, a->K, a->lda, kBM);
  if (rc != B2T_OK) return rc;
  const bool mc = bn == 256 && g_multicast && (a->M + kBM - 1) / kBM >= 2;
  rc = make_map(&mw, a->W, a->N, a->K, a->K, mc ? bn / 2 : bn);
  if (rc != B2T_OK) return rc;
  if (bn == 256) return dispatch_epi<256>(ma, mw, a, p, st);
  return dispatch_epi<128>(ma, mw, a, p, st);
}
