// Acoustic path on tensor cores (bf16 operands, fp32 accumulation = the reference's GPU numerics under
// torch.amp.autocast('cuda', bfloat16), audiotoken/encoder.py:45-52):
//
//   seanet_conv0_kernel   Conv(1->32, k7): CUDA cores, one thread per (row, 8 channels); HBM-bound.
//   seanet_tc_kernel      every other contraction of the SEANet encoder as a tcgen05 GEMM:
//                           * causal convs read their im2col matrix straight from the channels-last activation
//                             buffer through an OVERLAPPING-ROW tensor map (row stride = stride*C_in elements,
//                             row width = k*C_in), so no im2col copy and no per-tap loop exists;
//                           * the residual block's 1x1 convs (shortcut on x, k1 on ELU(h)) are one GEMM over the
//                             combined rows [x | ELU(h)];
//                           * reflect padding lives in per-clip halo rows that the PRODUCER's epilogue fills
//                             (row t in 1..H is mirrored to row -t; ELU commutes with the mirror);
//                           * the LSTM step is a GEMM [x_t | h_{t-1}] . [W_ih | W_hh]^T over all still-active clips
//                             with the cell update fused into the epilogue (gate-interleaved weight rows).
//                         Structure as gemm_tc.cu: persistent CTAs, warp 0 = TMA producer (SWIZZLE_128B, 4-stage
//                         ring), warp 1 = tcgen05.mma issuer, fp32 accumulators double-buffered in TMEM, 8 epilogue
//                         warps (tcgen05.ld, one accumulator row per lane).
//
// Row spaces.  Clip c has F_c frames (off4 = prefix sum).  Level l (rows per frame R = 320,160,40,8,1) stores clip
// c at padded row  R*(off4[c]-off4[c0]) + H*(c-c0)  with H halo rows first (H = 2,4,5,8 = the next stride; 6 at
// level 4 for the k7 conv).  A GEMM's M index runs over such a space; the epilogue maps m -> (clip, t) by binary
// search and writes to the output level's space.  Requires every clip length to be a multiple of 320 samples
// (the host zero-extends; exact because all convs are causal — SURVEY A.6).
#include <vector>
#include "tc_ptx.cuh"
#include "gemm_epilogue.cuh"
#include "seanet_tc.h"

void b2t_acoustic_mark(int cls, int begin, cudaStream_t st);   // acoustic.cu: 0 front end, 1 LSTM, 2 final conv, 3 RVQ

namespace {

constexpr int kBM = 128;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStages = 4;

// ELU(alpha = 1) with one MUFU: exp(x) = ex2(x * log2 e); every consumer rounds the result to bf16
B2T_DEVICE float elu_fast(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return x > 0.f ? x : e - 1.0f;
}
B2T_DEVICE float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
B2T_DEVICE float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

B2T_DEVICE void tmem_ld_32x32_x16_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

// ---- epilogue parameter blocks ---------------------------------------------------------------------
struct ConvEpi {
  const int32_t* off4;       // [n_clips + 1] frame prefix sums
  const int32_t* rank;       // [n_clips] position of the clip in the length-sorted order (time-major output)
  const int32_t* toff;       // [t_max + 1] prefix sums of the active-clip counts (time-major output)
  const float* bias;         // [N]
  __nv_bfloat16* out_raw; int ld_raw;    // bf16(acc + bias)            (nullable)
  __nv_bfloat16* out_elu; int ld_elu;    // bf16(ELU(bf16(acc + bias)))  (nullable)
  float* out_f32; int ld_f32;            // acc + bias in fp32           (nullable)
  int c0, nsub;              // clips [c0, c0 + nsub) make up the M space
  int r, h_in;               // M space: clip starts at r*(off4[c]-off4[c0]) + h_in*(c-c0), first h_in rows are halo
  int h_out, c0_out;         // output space: row = r*(off4[c]-off4[c0_out]) + h_out*(c-c0_out) + h_out + t
  int mirror;                // mirror rows 1..h_out of out_elu into the halo (and zero what a short clip leaves)
  int tm_out;                // out_raw row = toff[t] + rank[c]  (time-major, for the LSTM)
};

struct LstmEpi {
  const float* bias;         // [2048], gate-interleaved: index 4*u + g
  const int32_t* order;      // [n_clips] clip ids, longest first
  const int32_t* off4;
  float* c;                  // [n_clips, 512] cell state, indexed by sorted position
  __nv_bfloat16* h_out;      // time-major [total4, 512]
  const __nv_bfloat16* skip; // time-major x4 (last layer) or null
  __nv_bfloat16* s2e;        // clip-major padded (halo 6) ELU(h + skip) for the final conv (last layer) or null
  int t, toff_t, n_active;
};

struct RowInfo { int clip, t, len; bool valid; };

B2T_DEVICE RowInfo map_row(const ConvEpi& p, int m, int M) {
  RowInfo ri{0, -1, 0, false};
  if (m >= M) return ri;
  const int f0 = __ldg(p.off4 + p.c0);
  int lo = 0, hi = p.nsub;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    const int start = p.r * (__ldg(p.off4 + p.c0 + mid) - f0) + p.h_in * mid;
    if (start <= m) lo = mid; else hi = mid;
  }
  ri.clip = p.c0 + lo;
  const int fa = __ldg(p.off4 + ri.clip), fb = __ldg(p.off4 + ri.clip + 1);
  ri.t = m - (p.r * (fa - f0) + p.h_in * lo) - p.h_in;
  ri.len = p.r * (fb - fa);
  ri.valid = ri.t >= 0 && ri.t < ri.len;
  return ri;
}

// CW bf16 values of one row.  Rows are written one per lane, so a 16-byte store would dirty half a 32-byte sector
// per instruction (ncu: 47 % excessive sectors); sm_100's 256-bit store writes whole sectors.  dst is 32-byte aligned
// whenever CW is a multiple of 16 (column offsets are multiples of 16 elements, row pitches multiples of 32 bytes).
template <int CW>
B2T_DEVICE void store_bf16(__nv_bfloat16* dst, const uint32_t* pk) {
  if constexpr (CW % 16 == 0) {
#pragma unroll
    for (int i = 0; i < CW / 16; ++i)
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   ::"l"(dst + 16 * i), "r"(pk[8 * i]), "r"(pk[8 * i + 1]), "r"(pk[8 * i + 2]), "r"(pk[8 * i + 3]),
                     "r"(pk[8 * i + 4]), "r"(pk[8 * i + 5]), "r"(pk[8 * i + 6]), "r"(pk[8 * i + 7]) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < CW / 8; ++i)
      *reinterpret_cast<uint4*>(dst + 8 * i) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
  }
}

// one chunk of CW accumulator columns [col, col + CW) of the row described by ri
template <int CW>
B2T_DEVICE void conv_epilogue_chunk(const ConvEpi& p, const RowInfo& ri, int col, const uint32_t* acc) {
  if (!ri.valid) return;
  float v[CW];
#pragma unroll
  for (int i = 0; i < CW; i += 4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
    v[i] = __uint_as_float(acc[i]) + b.x; v[i + 1] = __uint_as_float(acc[i + 1]) + b.y;
    v[i + 2] = __uint_as_float(acc[i + 2]) + b.z; v[i + 3] = __uint_as_float(acc[i + 3]) + b.w;
  }
  const long long orow = (long long)p.r * (__ldg(p.off4 + ri.clip) - __ldg(p.off4 + p.c0_out)) +
                         (long long)p.h_out * (ri.clip - p.c0_out) + p.h_out + ri.t;
  if (p.out_f32) {
    float* o = p.out_f32 + orow * p.ld_f32 + col;
#pragma unroll
    for (int i = 0; i < CW; i += 8)
      asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   ::"l"(o + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3]), "f"(v[i + 4]), "f"(v[i + 5]),
                     "f"(v[i + 6]), "f"(v[i + 7]) : "memory");
  }
  uint32_t pk[CW / 2];
  if (p.out_raw) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) pk[i] = pack2_bf16(v[2 * i], v[2 * i + 1]);
    const long long rr = p.tm_out ? (long long)__ldg(p.toff + ri.t) + __ldg(p.rank + ri.clip) : orow;
    store_bf16<CW>(p.out_raw + rr * p.ld_raw + col, pk);
  }
  if (p.out_elu) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i)
      pk[i] = pack2_bf16(elu_fast(v[2 * i]), elu_fast(v[2 * i + 1]));
    __nv_bfloat16* o = p.out_elu + orow * p.ld_elu + col;
    store_bf16<CW>(o, pk);
    if (p.mirror) {
      if (ri.t >= 1 && ri.t <= p.h_out) store_bf16<CW>(o - 2LL * ri.t * p.ld_elu, pk);
      if (ri.t == 0 && ri.len <= p.h_out) {        // short clip: halo rows the mirror does not reach read as zero
        uint32_t z[CW / 2];
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) z[i] = 0u;
        for (int j = ri.len; j <= p.h_out; ++j) store_bf16<CW>(o - (long long)j * p.ld_elu, z);
      }
    }
  }
}

// LSTM cell for 8 hidden units (32 gate columns, order i f g o per unit) of sorted clip b
B2T_DEVICE void lstm_epilogue_chunk(const LstmEpi& p, int b, int col, const uint32_t* acc) {
  if (b >= p.n_active) return;
  const int u0 = col >> 2;
  float* cp = p.c + (size_t)b * 512 + u0;
  float cs[8];
  if (p.t > 0) {     // one clip per lane: 256-bit accesses move whole 32-byte sectors
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(cs[0]), "=f"(cs[1]), "=f"(cs[2]), "=f"(cs[3]), "=f"(cs[4]), "=f"(cs[5]), "=f"(cs[6]), "=f"(cs[7])
                 : "l"(cp) : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  }
  float h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4 * j));
    const float gi = __uint_as_float(acc[4 * j]) + bb.x, gf = __uint_as_float(acc[4 * j + 1]) + bb.y;
    const float gg = __uint_as_float(acc[4 * j + 2]) + bb.z, go = __uint_as_float(acc[4 * j + 3]) + bb.w;
    cs[j] = sigmoid_fast(gf) * cs[j] + sigmoid_fast(gi) * tanh_fast(gg);
    h[j] = sigmoid_fast(go) * tanh_fast(cs[j]);
  }
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(cp), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pk[j] = pack2_bf16(h[2 * j], h[2 * j + 1]);
  const size_t trow = (size_t)p.toff_t + b;
  store_bf16<8>(p.h_out + trow * 512 + u0, pk);
  if (p.s2e) {
    const uint4 sk = *reinterpret_cast<const uint4*>(p.skip + trow * 512 + u0);
    const uint32_t sw[4] = {sk.x, sk.y, sk.z, sk.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 s2 = *reinterpret_cast<const __nv_bfloat162*>(&sw[j]);
      const float a = bf16_round(bf16_round(h[2 * j]) + __low2float(s2));
      const float c = bf16_round(bf16_round(h[2 * j + 1]) + __high2float(s2));
      pk[j] = pack2_bf16(elu_fast(a), elu_fast(c));
    }
    const int clip = __ldg(p.order + b);
    const int fa = __ldg(p.off4 + clip), len = __ldg(p.off4 + clip + 1) - fa;
    __nv_bfloat16* o = p.s2e + ((size_t)fa + 6 * (size_t)clip + 6 + p.t) * 512 + u0;
    store_bf16<8>(o, pk);
    if (p.t >= 1 && p.t <= 6) store_bf16<8>(o - 2LL * p.t * 512, pk);
    if (p.t == 0 && len <= 6) {
      const uint32_t z[4] = {0u, 0u, 0u, 0u};
      for (int j = len; j <= 6; ++j) store_bf16<8>(o - (long long)j * 512, z);
    }
  }
}

template <int BN>
struct Smem {
  static constexpr int kStageA = kBM * kBK * 2;
  static constexpr int kStageB = BN * kBK * 2;
  static constexpr int kTileBytes = kStages * (kStageA + kStageB);
  static constexpr int kTotal = kTileBytes + 256 + 1024;
};

// out[m, n] = sum_k A[m, k] W[n, k];  A's k-blocks [0, kb0) come from map_a0 (rows row0 + m), the following kb1
// blocks from map_a1 (rows row1 + m); W k-block index runs over both.
template <int BN, typename Epi>
__global__ void __launch_bounds__(kThreads, BN <= 64 ? 2 : 1)
seanet_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_w, int kb0, int kb1, int row0, int row1, int M, int N, Epi p,
                 int pdl) {
  // pdl != 0 (LSTM step chain, programmatic dependent launch): this grid may start while the previous step is
  // still running.  Everything that does not depend on it — barrier/TMEM set-up and the x_t half of the K loop
  // (operands from map_a0 and the weights) — runs ahead; h_{t-1} (map_a1) and the cell state are touched only
  // after griddepcontrol.wait.
  using L = Smem<BN>;
  constexpr bool kLstm = std::is_same<Epi, LstmEpi>::value;
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
  const int tiles_n = N / BN;
  const int tiles = ((M + kBM - 1) / kBM) * tiles_n;
  const int num_kb = kb0 + kb1;
  constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpiWarps); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a0); tma_prefetch_desc(&map_a1); tma_prefetch_desc(&map_w); }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Producer and MMA warps run their loops warp-wide (uniform control flow) and let one elected lane issue: inside an
  // `if (lane == 0)` region ptxas wraps every TMA / tcgen05 instruction in an elect-and-retry loop with R2UR copies,
  // and for the narrow convolutions (a handful of MMAs per tile) that issue latency was a large part of a tile.
  if (warp == 0) {
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    bool waited = !pdl;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * kBM, n0 = (t % tiles_n) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (!waited && kb >= kb0) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; }
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (leader) {
          mbar_expect_tx(full_bar(stage), L::kStageA + L::kStageB);
          if (kb < kb0) tma_load_2d(sA + stage * L::kStageA, &map_a0, full_bar(stage), kb * kBK, row0 + m0);
          else tma_load_2d(sA + stage * L::kStageA, &map_a1, full_bar(stage), (kb - kb0) * kBK, row1 + m0);
          tma_load_2d(sB + stage * L::kStageB, &map_w, full_bar(stage), kb * kBK, n0);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc(kBM, BN);
    const uint64_t da0 = make_smem_desc(sA), db0 = make_smem_desc(sB);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint64_t da = da0 + (uint64_t)(stage * (L::kStageA >> 4));
          const uint64_t db = db0 + (uint64_t)(stage * (L::kStageB >> 4));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    __syncwarp();
  } else {
    // epilogue: TMEM lane quadrant = warp % 4; the two warps of a quadrant split the columns when BN >= 32
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;
    constexpr int kParts = BN >= 32 ? 2 : 1;
    constexpr int kColsPerPart = BN / kParts;
    constexpr int kCW = (kColsPerPart >= 32 && BN > 64) ? 32 : 16;   // narrow chunks keep BN <= 64 within 96 registers
    constexpr int kChunks = kColsPerPart / kCW;
    int acc = 0; uint32_t acc_phase = 0;
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * kBM, n0 = (t % tiles_n) * BN;
      const int m = m0 + quad * 32 + lane;
      RowInfo ri{0, -1, 0, false};
      if constexpr (!kLstm) { if (part < kParts) ri = map_row(p, m, M); }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (part < kParts) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + part * kColsPerPart);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          uint32_t r[32];
          if constexpr (kCW == 32) tmem_ld_32x32_nowait(taddr + (uint32_t)(c * 32), r);
          else tmem_ld_32x32_x16_nowait(taddr + (uint32_t)(c * 16), r);
          tmem_ld_wait();
          if (c == kChunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          const int col = n0 + part * kColsPerPart + c * kCW;
          if constexpr (kLstm) lstm_epilogue_chunk(p, m, col, r);
          else conv_epilogue_chunk<kCW>(p, ri, col, r);
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");   // transitivity: no grid of the chain exits early
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// Conv(1 -> 32, k = 7) on the raw waveform: 4 threads per output row (8 channels each, weights in registers), 512
// rows per block.  xh row = [x (32) | ELU(h) (16)] (ld 48), xe = ELU(x) (ld 32) with the two mirrored halo rows.
constexpr int kConv0Rows = 512;
__global__ void __launch_bounds__(256, 3)
seanet_conv0_kernel(const float* __restrict__ wave, const int64_t* __restrict__ wave_off,
                    const int32_t* __restrict__ true_len, const int32_t* __restrict__ off4, int c0, int nsub, int M,
                    const float* __restrict__ w /*[32][16]*/, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ xh, __nv_bfloat16* __restrict__ xe) {
  __shared__ int s_first;
  const int cg = (threadIdx.x & 3) * 8;
  float wr[8][7], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = __ldg(bias + cg + c);
#pragma unroll
    for (int j = 0; j < 7; ++j) wr[c][j] = __ldg(w + (cg + c) * 16 + j);
  }
  const int m_block = blockIdx.x * kConv0Rows;
  const int f0 = __ldg(off4 + c0);
  auto start = [&](int c) { return 320 * (__ldg(off4 + c0 + c) - f0) + 2 * c; };   // first (halo) row of clip c0 + c
  if (threadIdx.x == 0) {
    int lo = 0, hi = nsub;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (start(mid) <= m_block) lo = mid; else hi = mid; }
    s_first = lo;
  }
  __syncthreads();
  int c = s_first;
#pragma unroll 1
  for (int it = 0; it < kConv0Rows / 64; ++it) {
    const int m = m_block + it * 64 + (threadIdx.x >> 2);
    if (m >= M) break;
    while (c + 1 < nsub && start(c + 1) <= m) ++c;
    const int t = m - start(c) - 2;
    if (t < 0) continue;                                   // halo row: written by the mirror below
    const int clip = c0 + c;
    const float* x = wave + wave_off[clip];
    const int tl = true_len[clip];
    float s[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      int i = t + j - 6;
      i = i < 0 ? -i : i;                                  // reflect padding of the 6 left samples
      s[j] = i < tl ? __ldg(x + i) : 0.f;
    }
    uint32_t raw[4], el[4];
    float v[8];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float a = br[ch];
#pragma unroll
      for (int j = 0; j < 7; ++j) a = fmaf(wr[ch][j], s[j], a);
      v[ch] = a;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      raw[i] = pack2_bf16(v[2 * i], v[2 * i + 1]);
      el[i] = pack2_bf16(elu_fast(bf16_round(v[2 * i])), elu_fast(bf16_round(v[2 * i + 1])));
    }
    store_bf16<8>(xh + (size_t)m * 48 + cg, raw);
    store_bf16<8>(xe + (size_t)m * 32 + cg, el);
    if (t >= 1 && t <= 2) store_bf16<8>(xe + ((size_t)m - 2 * t) * 32 + cg, el);
  }
}

// ---- level 0 fused: waveform -> conv0 -> ELU -> k3 -> ELU -> [shortcut | k1] -> ELU, one kernel --------------------
// The 24 kHz level moves 45 % of the front end's bytes when run as separate kernels (x, ELU x, ELU h, y round-trip
// HBM at 64-96 B per row each).  Here the only HBM traffic is 4 B of waveform in and 64 B of ELU(y) out per row:
//   builder warps (8)  conv0 on CUDA cores straight from the waveform for the three tap times of every row (reflect
//                      padding by index), written as the two swizzled UMMA operand tiles  A1 = [ELU x(t-2) | ELU x(t-1) |
//                      ELU x(t)] (K = 96) and A2 = [x(t) | . ] (K = 48) of a double-buffered pair;
//   MMA warp           D1 = A1 . W3^T (N = 16), later D2 = A2 . [Wsc | Wk1]^T (N = 32), accumulators in TMEM;
//   epilogue warps (4) E1: h = ELU(D1 + b3) -> bf16 into columns 32..47 of A2 (shared memory, never HBM);
//                      E2: ELU(D2 + b) -> global (+ halo mirror).  E1 of tile n+1 runs before E2 of tile n so the
//                      MMA round trip is hidden.
//   edge warp          the taps a row cannot get from its neighbours' registers: the two rows at the start of a tile
//                      and at the start of a clip (reflect) need x(t-1), x(t-2) evaluated on their own — at most 18
//                      (row, tap, channel half) tasks per tile, one per lane, next to the builders instead of inside
//                      builder warp 0 (where they tripled the critical path); it also looks the next tile up.
constexpr int kL0Builders = 8;
constexpr int kL0EdgeWarp = kL0Builders, kL0Epi0 = kL0Builders + 1, kL0MmaWarp = kL0Builders + 5;
constexpr int kL0Threads = 32 * (kL0Builders + 1 + 4 + 1);
constexpr int kL0A1 = 2 * kBM * 128;          // two 64-wide k-blocks, 32 KB
constexpr int kL0A2 = kBM * 128;              // one k-block, 16 KB
constexpr int kL0Smem = 2 * (kL0A1 + kL0A2) + 4096 + 4096 + 1024 + 1024 + 1024;   // operands, W3, Wres, w0, barrier block, slack

// what the builder threads need to know about a 128-row tile: the clip of its first row and the next one
struct L0TileInfo { long long woff0, woff1; int t0, len0, len1, tl0, tl1, has1; };

struct L0Params {
  const float* wave; const int64_t* wave_off; const int32_t* true_len; const int32_t* off4;
  const float* w0; const float* b0;           // conv0 fp32 [32][16], [32]
  const float* b3; const float* bres;         // biases of k3 (16) and of shortcut + k1 (32)
  __nv_bfloat16* ye;                          // [P_0 rows, 32] ELU(y) with mirrored halo (H = 2)
  int c0, nsub, M;
  long long* dbg;                             // optional timeline of CTA 0 (developer tool): [tile][8] clock64 stamps
};

__global__ void __launch_bounds__(kL0Threads, 2)
seanet_l0_kernel(const __grid_constant__ CUtensorMap map_w3, const __grid_constant__ CUtensorMap map_wres, L0Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  // layout: A1[0] A1[1] A2[0] A2[1] W3(4 KB) Wres(4 KB) w0(1 KB) barriers
  const uint32_t oA1 = 0, oA2 = 2 * kL0A1, oW3 = oA2 + 2 * kL0A2, oWr = oW3 + 4096, oW0 = oWr + 4096, oBar = oW0 + 1024;
  const uint32_t bars = base + oBar;
  auto a1_full = [&](int b) { return bars + 8u * b; };
  auto a2_full = [&](int b) { return bars + 8u * (2 + b); };
  auto a_empty = [&](int b) { return bars + 8u * (4 + b); };
  auto d1_full = [&](int b) { return bars + 8u * (6 + b); };
  auto d1_empty = [&](int b) { return bars + 8u * (8 + b); };
  auto d2_full = [&](int b) { return bars + 8u * (10 + b); };
  auto d2_empty = [&](int b) { return bars + 8u * (12 + b); };
  const uint32_t w_bar = bars + 8u * 14, tmem_slot = bars + 8u * 15;
  auto ti_full = [&](int b) { return bars + 320u + 8u * b; };      // tile look-up n is in tinfo[n & 7]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_total = (p.M + kBM - 1) / kBM;
  const int my_tiles = blockIdx.x < n_tiles_total ? (n_tiles_total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(a1_full(b), kL0Builders + 1); mbar_init(a2_full(b), 4); mbar_init(a_empty(b), 1);
      mbar_init(d1_full(b), 1); mbar_init(d1_empty(b), 4); mbar_init(d2_full(b), 1); mbar_init(d2_empty(b), 4);
    }
    mbar_init(w_bar, 1);
    mbar_init(ti_full(0), 1); mbar_init(ti_full(1), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  // zero the operand buffers once (K padding columns are never written again) and stage the conv0 weights
  for (int i = threadIdx.x; i < (2 * (kL0A1 + kL0A2)) / 16; i += kL0Threads) reinterpret_cast<uint4*>(bp)[i] = make_uint4(0u, 0u, 0u, 0u);
  float* sw0 = reinterpret_cast<float*>(bp + oW0);
  for (int i = threadIdx.x; i < 256; i += kL0Threads) {
    const int c = i >> 3, j = i & 7;
    sw0[i] = j < 7 ? __ldg(p.w0 + c * 16 + j) : __ldg(p.b0 + c);
  }
  float* sb3 = reinterpret_cast<float*>(bp + oBar + 128);      // 16 + 32 biases behind the barriers
  if (threadIdx.x < 16) sb3[threadIdx.x] = __ldg(p.b3 + threadIdx.x);
  else if (threadIdx.x < 48) sb3[threadIdx.x] = __ldg(p.bres + threadIdx.x - 16);
  if (warp == kL0MmaWarp) tmem_alloc(tmem_slot, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(bp + oBar + 8 * 15);

  // Tile look-ups (first row of the tile -> clip, time, the clip after it) are produced two tiles ahead by the MMA
  // thread in its idle time (slot n & 7, mbarrier ti_full[n & 1]); builders, the edge warp and the E2 epilogue derive
  // every row of the tile from them (a tile spans at most two clips: >= 320 rows each), so no thread runs a binary
  // search or a chain of dependent global loads per row.
  L0TileInfo* tinfo = reinterpret_cast<L0TileInfo*>(bp + oBar + 512);
  auto derive = [&](const L0TileInfo& ti, int i) {
    RowInfo ri;
    const int t = ti.t0 + i;
    const bool second = t >= ti.len0;
    ri.t = second ? t - ti.len0 - 2 : t;                    // the next clip starts with its 2 halo rows
    ri.len = second ? ti.len1 : ti.len0;
    ri.clip = second ? 1 : 0;
    ri.valid = i >= 0 && i < kBM && ri.t >= 0 && ri.t < ri.len && (!second || ti.has1);
    return ri;
  };

  if (warp <= kL0EdgeWarp) {
    // ===== builders (warps 0-7): thread pair per row; half = channels [16*half, 16*half + 16) =====
    // ===== edge warp (warp 8): one (row, tap, half) task per lane =====
    const bool edge = warp == kL0EdgeWarp;
    // conv0 at time u for 16 channels [16*half, 16*half + 16): raw bf16 and ELU bf16 (2 x 16-byte chunks each)
    auto conv0_at = [&](const float* x, int tl, int half, int u, uint32_t (&raw)[8], uint32_t (&el)[8]) {
      float sj[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        int idx = u + j - 6;
        idx = idx < 0 ? -idx : idx;                         // reflect padding of conv0 (on the waveform)
        sj[j] = idx < tl ? __ldg(x + idx) : 0.f;
      }
#pragma unroll
      for (int cc = 0; cc < 16; cc += 2) {
        float v2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float4 wa = *reinterpret_cast<const float4*>(sw0 + (half * 16 + cc + e) * 8);
          const float4 wb = *reinterpret_cast<const float4*>(sw0 + (half * 16 + cc + e) * 8 + 4);
          float a = wb.w;
          a = fmaf(wa.x, sj[0], a); a = fmaf(wa.y, sj[1], a); a = fmaf(wa.z, sj[2], a); a = fmaf(wa.w, sj[3], a);
          a = fmaf(wb.x, sj[4], a); a = fmaf(wb.y, sj[5], a); a = fmaf(wb.z, sj[6], a);
          v2[e] = a;
        }
        raw[cc >> 1] = pack2_bf16(v2[0], v2[1]);
        el[cc >> 1] = pack2_bf16(elu_fast(v2[0]), elu_fast(v2[1]));
      }
    };
    // ELU x(time) lands in tile row `row`, tap `tap`: chunks 4*tap + 2*half (+1) of that row
    auto put_tap = [&](uint8_t* a1, int half, int row, int tap, const uint32_t (&el)[8]) {
      const uint32_t ro = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int chunk = 4 * tap + 2 * half + h2;
        const uint32_t off = (uint32_t)(chunk >> 3) * (kBM * 128) + ro + (uint32_t)(((chunk & 7) ^ (row & 7)) * 16);
        *reinterpret_cast<uint4*>(a1 + off) = make_uint4(el[4 * h2], el[4 * h2 + 1], el[4 * h2 + 2], el[4 * h2 + 3]);
      }
    };
    for (int n = 0; n < my_tiles; ++n) {
      const int buf = n & 1;
      if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0 && n < 64) p.dbg[n * 8 + 0] = clock64();
      mbar_wait(ti_full(buf), (uint32_t)((n >> 1) & 1));
      const L0TileInfo ti = tinfo[n & 7];
      uint8_t* a1 = bp + oA1 + buf * kL0A1;
      uint8_t* a2 = bp + oA2 + buf * kL0A2;
      if (!edge) {
        const int i = threadIdx.x & 127, half = threadIdx.x >> 7;
        const uint32_t row_off = (uint32_t)((i >> 3) * 1024 + (i & 7) * 128);
        const RowInfo ri = derive(ti, i);
        mbar_wait(a_empty(buf), (uint32_t)(((n >> 1) & 1) ^ 1));
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0 && n < 64) p.dbg[n * 8 + 1] = clock64();
        if (ri.valid) {
          const float* x = p.wave + (ri.clip ? ti.woff1 : ti.woff0);
          const int tl = ri.clip ? ti.tl1 : ti.tl0;
          uint32_t raw[8], el[8];
          conv0_at(x, tl, half, ri.t, raw, el);
          // x(t) is tap 2 of row t, tap 1 of row t+1 and tap 0 of row t+2: computed once, written three times
          put_tap(a1, half, i, 2, el);
          if (i + 1 < kBM && ri.t + 1 < ri.len) put_tap(a1, half, i + 1, 1, el);
          if (i + 2 < kBM && ri.t + 2 < ri.len) put_tap(a1, half, i + 2, 0, el);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int chunk = 2 * half + h2;
            *reinterpret_cast<uint4*>(a2 + row_off + (uint32_t)((chunk ^ (i & 7)) * 16)) =
                make_uint4(raw[4 * h2], raw[4 * h2 + 1], raw[4 * h2 + 2], raw[4 * h2 + 3]);
          }
        }
      } else {
        // taps whose producer row lies in the previous tile or before the clip start (reflect: x(-d) = x(d)):
        // row i needs tap 2-d evaluated on its own iff i < d or t(i) < d.  Candidates: the first two rows of the
        // tile, of the first clip (when the tile starts inside its halo) and of the second clip; lane = (group,
        // (row, d) in {(0,1), (0,2), (1,2)}, half).  Coinciding candidates write identical values.
        const int grp = lane / 6, sub = (lane % 6) >> 1, half = lane & 1;
        const int rel = sub == 2 ? 1 : 0, d = sub == 0 ? 1 : 2;
        const int first = grp == 0 ? 0 : grp == 1 ? -ti.t0 : ti.len0 + 2 - ti.t0;   // row with i == 0 / t == 0
        const int i = first + rel;
        const bool cand = lane < 18 && (grp != 1 || ti.t0 < 0);
        const RowInfo ri = derive(ti, cand ? i : -1);
        mbar_wait(a_empty(buf), (uint32_t)(((n >> 1) & 1) ^ 1));
        if (ri.valid && (i < d || ri.t < d)) {
          const float* x = p.wave + (ri.clip ? ti.woff1 : ti.woff0);
          const int tl = ri.clip ? ti.tl1 : ti.tl0;
          const int u = ri.t - d;
          uint32_t raw[8], el[8];
          conv0_at(x, tl, half, u < 0 ? -u : u, raw, el);
          put_tap(a1, half, i, 2 - d, el);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_full(buf));
      if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0 && n < 64) p.dbg[n * 8 + 2] = clock64();
    }
  } else if (warp < kL0MmaWarp) {
    // ===== epilogue warps: TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    const int i = quad * 32 + lane;
    const uint32_t row_off = (uint32_t)((i >> 3) * 1024 + (i & 7) * 128);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    auto e1 = [&](int n) {
      const int buf = n & 1;
      mbar_wait(d1_full(buf), (uint32_t)((n >> 1) & 1));
      if (p.dbg && blockIdx.x == 0 && warp == kL0Epi0 && lane == 0 && n < 64) p.dbg[n * 8 + 3] = clock64();
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32_x16_nowait(lane_addr + (uint32_t)(buf * 16), r);
      tmem_ld_wait();
      tc_fence_before();
      uint32_t pk[8];
#pragma unroll
      for (int c = 0; c < 16; c += 2)
        pk[c >> 1] = pack2_bf16(elu_fast(__uint_as_float(r[c]) + sb3[c]), elu_fast(__uint_as_float(r[c + 1]) + sb3[c + 1]));
      uint8_t* a2 = bp + oA2 + buf * kL0A2;
      *reinterpret_cast<uint4*>(a2 + row_off + (uint32_t)((4 ^ (i & 7)) * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(a2 + row_off + (uint32_t)((5 ^ (i & 7)) * 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(a2_full(buf)); mbar_arrive(d1_empty(buf)); }
      if (p.dbg && blockIdx.x == 0 && warp == kL0Epi0 && lane == 0 && n < 64) p.dbg[n * 8 + 4] = clock64();
    };
    auto e2 = [&](int n) {
      const int buf = n & 1;
      const int m = ((int)blockIdx.x + n * (int)gridDim.x) * kBM + i;
      const RowInfo ri = derive(tinfo[n & 7], m < p.M ? i : -1);   // slot n & 7 is rewritten for tile n + 8, after E1(n + 5)
      mbar_wait(d2_full(buf), (uint32_t)((n >> 1) & 1));
      if (p.dbg && blockIdx.x == 0 && warp == kL0Epi0 && lane == 0 && n < 64) p.dbg[n * 8 + 5] = clock64();
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32_nowait(lane_addr + (uint32_t)(32 + buf * 32), r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty(buf));
      if (ri.valid) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2)
          pk[c >> 1] = pack2_bf16(elu_fast(__uint_as_float(r[c]) + sb3[16 + c]), elu_fast(__uint_as_float(r[c + 1]) + sb3[16 + c + 1]));
        __nv_bfloat16* o = p.ye + (size_t)m * 32;            // same row space in and out (level 0, H = 2)
        store_bf16<32>(o, pk);
        if (ri.t >= 1 && ri.t <= 2) store_bf16<32>(o - 2 * ri.t * 32, pk);
      }
      if (p.dbg && blockIdx.x == 0 && warp == kL0Epi0 && lane == 0 && n < 64) p.dbg[n * 8 + 6] = clock64();
    };
    if (my_tiles > 0) e1(0);
    for (int n = 0; n < my_tiles; ++n) {
      if (n + 1 < my_tiles) e1(n + 1);
      e2(n);
    }
  } else {
    // ===== MMA warp =====
    if (lane == 0) {
      const uint32_t sW3 = base + oW3, sWr = base + oWr;
      mbar_expect_tx(w_bar, 4096 + 4096);
      tma_load_2d(sW3, &map_w3, w_bar, 0, 0);
      tma_load_2d(sW3 + 2048, &map_w3, w_bar, 64, 0);
      tma_load_2d(sWr, &map_wres, w_bar, 0, 0);
      mbar_wait(w_bar, 0);
      constexpr uint32_t idesc1 = make_idesc(kBM, 16), idesc2 = make_idesc(kBM, 32);
      // incremental look-up: tiles of a CTA only move forward through the clip table
      const int f0 = __ldg(p.off4 + p.c0);
      int lk = 0;                                            // clip (relative to c0) of the previous look-up
      auto clip_row0 = [&](int c) { return 320 * (__ldg(p.off4 + p.c0 + c) - f0) + 2 * c; };   // first (halo) row of clip c
      auto lookup = [&](int n) {
        const int m0 = ((int)blockIdx.x + n * (int)gridDim.x) * kBM;
        int steps = 0;
        while (lk + 1 < p.nsub && clip_row0(lk + 1) <= m0) {
          ++lk;
          if (++steps == 8) {                                // many tiny clips: finish with a binary search
            int lo = lk, hi = p.nsub;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (clip_row0(mid) <= m0) lo = mid; else hi = mid; }
            lk = lo;
            break;
          }
        }
        const int clip = p.c0 + lk;
        const int fa = __ldg(p.off4 + clip), fb = __ldg(p.off4 + clip + 1);
        L0TileInfo ti;
        ti.t0 = m0 - (320 * (fa - f0) + 2 * lk) - 2; ti.len0 = 320 * (fb - fa);
        ti.woff0 = __ldg(p.wave_off + clip); ti.tl0 = __ldg(p.true_len + clip);
        ti.has1 = lk + 1 < p.nsub;
        ti.woff1 = ti.has1 ? __ldg(p.wave_off + clip + 1) : 0;
        ti.tl1 = ti.has1 ? __ldg(p.true_len + clip + 1) : 0;
        ti.len1 = ti.has1 ? 320 * (__ldg(p.off4 + clip + 2) - fb) : 0;
        tinfo[n & 7] = ti;
        mbar_arrive(ti_full(n & 1));                         // release: the slot is written before the arrival
      };
      if (my_tiles > 0) lookup(0);
      if (my_tiles > 1) lookup(1);
      auto mma2 = [&](int k) {
        const int buf = k & 1;
        mbar_wait(a2_full(buf), (uint32_t)((k >> 1) & 1));
        mbar_wait(d2_empty(buf), (uint32_t)(((k >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint64_t da = make_smem_desc(base + oA2 + buf * kL0A2), db = make_smem_desc(sWr);
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
          umma_bf16(tmem_base + (uint32_t)(32 + buf * 32), da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc2, kk ? 1u : 0u);
        umma_commit(d2_full(buf));
        umma_commit(a_empty(buf));
      };
      for (int n = 0; n < my_tiles; ++n) {
        const int buf = n & 1;
        mbar_wait(a1_full(buf), (uint32_t)((n >> 1) & 1));
        mbar_wait(d1_empty(buf), (uint32_t)(((n >> 1) & 1) ^ 1));
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = make_smem_desc(base + oA1 + buf * kL0A1 + kb * (kBM * 128));
          const uint64_t db = make_smem_desc(sW3 + kb * 2048);
#pragma unroll
          for (int kk = 0; kk < (kb ? 2 : 4); ++kk)            // K = 96: the second k-block holds 32 real columns
            umma_bf16(tmem_base + (uint32_t)(buf * 16), da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc1, (kb | kk) ? 1u : 0u);
        }
        umma_commit(d1_full(buf));
        if (p.dbg && blockIdx.x == 0 && n < 64) p.dbg[n * 8 + 7] = clock64();
        // D2 of the same tile as soon as E1 has produced ELU h: the operand buffers go back to the builders one
        // MMA round trip after D1 instead of one whole tile later (the builders are the long pole, not this warp)
        mma2(n);
        if (n + 2 < my_tiles) lookup(n + 2);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kL0MmaWarp) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// ---- host ----------------------------------------------------------------------------------------
int make_map_k(CUtensorMap* map, const void* ptr, long long rows, int K, long long ld_elems, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  B2T_REQUIRE(fn != nullptr, B2T_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2T_REQUIRE(r == CUDA_SUCCESS, B2T_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld K=%d ld=%lld", (int)r, rows, K, ld_elems);
  return B2T_OK;
}

template <int BN, typename Epi>
int launch_bn(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, int kb0, int kb1, int row0, int row1,
              int M, int N, const Epi& p, cudaStream_t st, int pdl = 0) {
  using L = Smem<BN>;
  B2T_SMEM_OPT_IN(L::kTotal, seanet_tc_kernel<BN, Epi>);
  const int tiles = ((M + kBM - 1) / kBM) * (N / BN);
  if (tiles <= 0) return B2T_OK;
  int occ = (227 * 1024) / (L::kTotal + 1024);
  const int tmem_occ = 512 / ((2 * BN < 32) ? 32 : 2 * BN);
  if (occ > tmem_occ) occ = tmem_occ;
  if (occ > (BN <= 64 ? 2 : 1)) occ = (BN <= 64 ? 2 : 1);   // register budget (__launch_bounds__)
  if (occ < 1) occ = 1;
  int grid = b2t_num_sms() * occ;
  if (tiles < grid) grid = tiles;
  if (pdl) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(kThreads); lc.dynamicSmemBytes = L::kTotal; lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = 1;
    B2T_CUDA(cudaLaunchKernelEx(&lc, seanet_tc_kernel<BN, Epi>, a0, a1, w, kb0, kb1, row0, row1, M, N, p, pdl));
    b2t_count_launch();
    return B2T_OK;
  }
  seanet_tc_kernel<BN, Epi><<<grid, kThreads, L::kTotal, st>>>(a0, a1, w, kb0, kb1, row0, row1, M, N, p, 0);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}

int launch_conv(const CUtensorMap& a, const CUtensorMap& w, int K, int M, int N, const ConvEpi& p, cudaStream_t st) {
  const int kb = (K + kBK - 1) / kBK;
  switch (N) {
    case 16: return launch_bn<16, ConvEpi>(a, a, w, kb, 0, 0, 0, M, N, p, st);
    case 32: return launch_bn<32, ConvEpi>(a, a, w, kb, 0, 0, 0, M, N, p, st);
    case 64: return launch_bn<64, ConvEpi>(a, a, w, kb, 0, 0, 0, M, N, p, st);
    case 128: return launch_bn<128, ConvEpi>(a, a, w, kb, 0, 0, 0, M, N, p, st);
    default:
      B2T_REQUIRE(N % 256 == 0, B2T_ERR_ARG, "seanet_tc: unsupported N=%d", N);
      return launch_bn<256, ConvEpi>(a, a, w, kb, 0, 0, 0, M, N, p, st);
  }
}

constexpr int kR[5] = {320, 160, 40, 8, 1};      // rows per frame at level l
constexpr int kH[5] = {2, 4, 5, 8, 6};           // halo rows in front of every clip at level l
constexpr int kC[5] = {32, 64, 128, 256, 512};
constexpr int kS[4] = {2, 4, 5, 8};
constexpr int kGuard = 16;                       // rows in front of every buffer (windows of row 0 reach back)

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct TcWs {
  __nv_bfloat16* xe[4]; __nv_bfloat16* xh[4]; __nv_bfloat16* ye[4];   // sub-batch buffers (point past the guard)
  __nv_bfloat16* x4t; __nv_bfloat16* s1t; __nv_bfloat16* h2t; __nv_bfloat16* s2e;
  float* emb; float* c; float* c2;              // c / c2: cell state of LSTM layer 1 / 2
  size_t total;
};

TcWs tc_carve(void* base, long long sub_frames, int sub_clips, long long total4, int n_clips) {
  TcWs w{};
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* r = p ? p + off : nullptr; off += align256(bytes); return r; };
  for (int l = 0; l < 4; ++l) {
    const size_t rows = (size_t)kR[l] * sub_frames + (size_t)kH[l] * sub_clips + kGuard + kBM;
    const size_t g = (size_t)kGuard;
    uint8_t* a = take(rows * kC[l] * 2); uint8_t* b = take(rows * (kC[l] * 3 / 2) * 2); uint8_t* c = take(rows * kC[l] * 2);
    w.xe[l] = a ? (__nv_bfloat16*)a + g * kC[l] : nullptr;
    w.xh[l] = b ? (__nv_bfloat16*)b + g * (kC[l] * 3 / 2) : nullptr;
    w.ye[l] = c ? (__nv_bfloat16*)c + g * kC[l] : nullptr;
  }
  const size_t t4 = (size_t)total4 + kBM;
  w.x4t = (__nv_bfloat16*)take(t4 * 512 * 2);
  w.s1t = (__nv_bfloat16*)take(t4 * 512 * 2);
  w.h2t = (__nv_bfloat16*)take(t4 * 512 * 2);
  uint8_t* s = take((t4 + 6 * (size_t)n_clips + kGuard) * 512 * 2);
  w.s2e = s ? (__nv_bfloat16*)s + (size_t)kGuard * 512 : nullptr;
  w.emb = (float*)take((size_t)total4 * 128 * 4);
  w.c = (float*)take((size_t)n_clips * 512 * 4);
  w.c2 = (float*)take((size_t)n_clips * 512 * 4);
  w.total = off;
  return w;
}

// greedy split of the clip list into front-end sub-batches of at most kSubFrames frames
long long* g_l0_dbg = nullptr;    // b2t_seanet_set_l0_dbg(device buffer of 64 x 8 int64) — developer timeline
int g_l0_fused = 1;               // b2t_set_option("seanet_l0_fused", 0/1)
int g_lstm_pdl = 1;               // b2t_set_option("lstm_pdl", 0/1)
int g_lstm_overlap = 1;           // b2t_set_option("lstm_overlap", 0/1): layer 2 follows layer 1 on a second stream
long long g_sub_frames = 24576;   // b2t_set_option("seanet_sub_frames", n)
struct SubBatch { int c0, c1; long long frames; };
std::vector<SubBatch> split_clips(const int32_t* frames_host, int n, long long* max_frames, int* max_clips) {
  std::vector<SubBatch> v;
  long long mf = 0; int mc = 0;
  int c = 0;
  while (c < n) {
    SubBatch s{c, c, 0};
    while (s.c1 < n && (s.c1 == s.c0 || s.frames + frames_host[s.c1] <= g_sub_frames)) { s.frames += frames_host[s.c1]; ++s.c1; }
    if (s.frames > mf) mf = s.frames;
    if (s.c1 - s.c0 > mc) mc = s.c1 - s.c0;
    v.push_back(s);
    c = s.c1;
  }
  *max_frames = mf; *max_clips = mc;
  return v;
}

}  // namespace

void b2t_seanet_set_sub_frames(int n) { if (n > 0) g_sub_frames = n; }
void b2t_seanet_set_lstm_pdl(int on) { g_lstm_pdl = on != 0; }
void b2t_seanet_set_lstm_overlap(int on) { g_lstm_overlap = on != 0; }
void b2t_seanet_set_l0_fused(int on) { g_l0_fused = on != 0; }
extern "C" void b2t_seanet_set_l0_dbg(long long* p) { g_l0_dbg = p; }

size_t b2t_seanet_tc_workspace_bytes(const b2t_acoustic_batch* b) {
  if (!b->frames_host) return 0;
  long long mf; int mc;
  split_clips(b->frames_host, b->n_clips, &mf, &mc);
  return tc_carve(nullptr, mf, mc, b->total[4], b->n_clips).total;
}

// Encoder on tensor cores: wave -> emb fp32 [total4, 128] (dense frame order).  `T(name)` resolves model tensors.
int b2t_seanet_tc_encode(const SeanetTcWeights& wt, const float* wave, const b2t_acoustic_batch* b, void* workspace,
                         size_t workspace_bytes, float* emb, const int32_t* active_host, cudaStream_t st) {
  B2T_REQUIRE(b->frames_host && b->rank && b->toff, B2T_ERR_ARG, "seanet_tc: batch lacks frames_host / rank / toff");
  B2T_REQUIRE(b->aligned320, B2T_ERR_ARG, "seanet_tc: every clip length must be a multiple of 320 samples");
  long long mf; int mc;
  std::vector<SubBatch> subs = split_clips(b->frames_host, b->n_clips, &mf, &mc);
  TcWs w = tc_carve(workspace, mf, mc, b->total[4], b->n_clips);
  B2T_REQUIRE(workspace_bytes >= w.total, B2T_ERR_WORKSPACE, "seanet_tc: workspace %zu < %zu", workspace_bytes, w.total);
  B2T_REQUIRE((long long)kR[0] * mf + (long long)kH[0] * mc < (1LL << 31) - 256, B2T_ERR_ARG, "seanet_tc: sub-batch too large");
  const int total4 = b->total[4], n = b->n_clips;
#define RUN(call) do { int rc__ = (call); if (rc__ != B2T_OK) return rc__; } while (0)

  // ---- strided-conv front end, one sub-batch of clips at a time (levels 0-3 reuse the same buffers) ----
  b2t_acoustic_mark(0, 1, st);
  for (const SubBatch& sb : subs) {
    const int ns = sb.c1 - sb.c0;
    long long Ml[5];
    for (int l = 0; l < 5; ++l) Ml[l] = (long long)kR[l] * sb.frames + (long long)kH[l] * ns;
    if (g_l0_fused) {
      const int M0 = (int)Ml[0];
      CUtensorMap m3, mr;
      RUN(make_map_k(&m3, wt.k3_w[0], 16, 128, 128, 16));
      RUN(make_map_k(&mr, wt.res_w[0], 32, 64, 64, 32));
      L0Params lp{wave, b->wave_off, b->true_len, b->off[4], wt.conv0_w, wt.conv0_b, wt.k3_b[0], wt.res_b[0], w.ye[0], sb.c0, ns, M0,
                  g_l0_dbg};
      B2T_SMEM_OPT_IN(kL0Smem, seanet_l0_kernel);
      int grid = 2 * b2t_num_sms();
      const int tiles0 = (M0 + kBM - 1) / kBM;
      if (tiles0 < grid) grid = tiles0;
      seanet_l0_kernel<<<grid, kL0Threads, kL0Smem, st>>>(m3, mr, lp);
      B2T_LAUNCH_CHECK();
    } else {
      const int M0 = (int)Ml[0];
      seanet_conv0_kernel<<<(M0 + kConv0Rows - 1) / kConv0Rows, 256, 0, st>>>(wave, b->wave_off, b->true_len, b->off[4], sb.c0, ns, M0,
                                                            wt.conv0_w, wt.conv0_b, w.xh[0], w.xe[0]);
      B2T_LAUNCH_CHECK();
    }
    for (int l = 0; l < 4; ++l) {
      const int C = kC[l], s = kS[l], M = (int)Ml[l];
      CUtensorMap ma, mw;
      ConvEpi e{};
      e.off4 = b->off[4]; e.rank = b->rank; e.toff = b->toff; e.c0 = sb.c0; e.nsub = ns; e.c0_out = sb.c0;
      const bool fused0 = (l == 0 && g_l0_fused);        // level 0: conv0 + residual block ran as one kernel
      // (1) ELU -> Conv(C -> C/2, k3): window of row m = rows m-2 .. m of xe (3C contiguous elements)
      if (!fused0) {
      RUN(make_map_k(&ma, w.xe[l] - 2 * C, M, 3 * C, C, kBM));
      RUN(make_map_k(&mw, wt.k3_w[l], C / 2, wt.k3_kpad[l], wt.k3_kpad[l], C / 2));
      e.bias = wt.k3_b[l]; e.out_raw = nullptr; e.out_f32 = nullptr;
      e.out_elu = w.xh[l] + C; e.ld_elu = C * 3 / 2; e.r = kR[l]; e.h_in = kH[l]; e.h_out = kH[l]; e.mirror = 0; e.tm_out = 0;
      RUN(launch_conv(ma, mw, 3 * C, M, C / 2, e, st));
      // (2) shortcut(x) + Conv(C/2 -> C, k1)(ELU(h)) as one GEMM over [x | ELU(h)], then ELU for the strided conv
      RUN(make_map_k(&ma, w.xh[l], M, C * 3 / 2, C * 3 / 2, kBM));
      RUN(make_map_k(&mw, wt.res_w[l], C, wt.res_kpad[l], wt.res_kpad[l], C > 256 ? 256 : C));
      e.bias = wt.res_b[l]; e.out_elu = w.ye[l]; e.ld_elu = C; e.mirror = 1;
      RUN(launch_conv(ma, mw, C * 3 / 2, M, C, e, st));
      }
      e.out_raw = nullptr; e.out_f32 = nullptr; e.tm_out = 0;
      // (3) Conv(C -> 2C, k = 2s, stride s): super-rows of s input rows; window of output row m' = super-rows m'-1, m'
      const int Mo = (int)(Ml[l] / s);        // = R[l+1]*frames + 1*ns  (one halo super-row per clip)
      RUN(make_map_k(&ma, w.ye[l] - s * C, Mo, 2 * s * C, (long long)s * C, kBM));
      RUN(make_map_k(&mw, wt.down_w[l], 2 * C, wt.down_kpad[l], wt.down_kpad[l], 2 * C > 256 ? 256 : 2 * C));
      e.bias = wt.down_b[l]; e.r = kR[l + 1]; e.h_in = 1; e.mirror = 1;
      if (l < 3) {
        e.out_raw = w.xh[l + 1]; e.ld_raw = 3 * C; e.out_elu = w.xe[l + 1]; e.ld_elu = 2 * C; e.h_out = kH[l + 1];
      } else {
        e.out_raw = w.x4t; e.ld_raw = 512; e.out_elu = nullptr; e.tm_out = 1; e.h_out = 0; e.mirror = 0;
      }
      RUN(launch_conv(ma, mw, 2 * s * C, Mo, 2 * C, e, st));
    }
  }

  b2t_acoustic_mark(0, 0, st);
  b2t_acoustic_mark(1, 1, st);
  // ---- LSTM: per step one GEMM [x_t | h_{t-1}] . [W_ih | W_hh]^T with the cell update in the epilogue ----
  std::vector<int> toff(b->t_max + 1, 0);
  for (int t = 0; t < b->t_max; ++t) toff[t + 1] = toff[t] + active_host[t];
  // One step kernel has 80 CTAs (10 row tiles x 8 column tiles at 1250 clips) on 148 SMs and the 2 x t_max steps are
  // two dependent chains — but layer 2 at step t only needs layer 1 up to step t.  With lstm_overlap the second layer
  // runs on a side stream kChunk steps behind the first (one event per chunk), so two step kernels are resident at a
  // time; each layer keeps its own cell-state buffer.
  constexpr int kChunk = 32;
  // side stream and events belong to a device: one set per (host thread, device)
  static thread_local cudaStream_t side_of[64] = {};
  static thread_local std::vector<cudaEvent_t> evs_of[64];
  int cur_dev = 0;
  B2T_CUDA(cudaGetDevice(&cur_dev));
  cudaStream_t& side = side_of[cur_dev & 63];
  std::vector<cudaEvent_t>& evs = evs_of[cur_dev & 63];
  const bool overlap = g_lstm_overlap != 0 && b->t_max > kChunk;
  if (overlap && !side) B2T_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  CUtensorMap mx[2], mh[2], mw[2];
  LstmEpi ep[2];
  for (int layer = 0; layer < 2; ++layer) {
    const __nv_bfloat16* xin = layer == 0 ? w.x4t : w.s1t;
    __nv_bfloat16* hout = layer == 0 ? w.s1t : w.h2t;
    RUN(make_map_k(&mx[layer], xin, (long long)total4 + kBM, 512, 512, kBM));
    RUN(make_map_k(&mh[layer], hout, (long long)total4 + kBM, 512, 512, kBM));
    RUN(make_map_k(&mw[layer], wt.lstm_w[layer], 2048, 1024, 1024, 256));
    LstmEpi e{};
    e.bias = wt.lstm_b[layer]; e.order = b->order; e.off4 = b->off[4]; e.c = layer == 0 ? w.c : w.c2; e.h_out = hout;
    e.skip = layer == 1 ? w.x4t : nullptr; e.s2e = layer == 1 ? w.s2e : nullptr;
    ep[layer] = e;
  }
  int t_end = 0;
  while (t_end < b->t_max && active_host[t_end] > 0) ++t_end;
  auto run_steps = [&](int layer, int ta, int tb, cudaStream_t s, bool first_plain) -> int {
    for (int t = ta; t < tb; ++t) {
      LstmEpi e = ep[layer];
      e.t = t; e.toff_t = toff[t]; e.n_active = active_host[t];
      const int pdl = (t > 0 && !(first_plain && t == ta)) ? g_lstm_pdl : 0;
      RUN((launch_bn<256, LstmEpi>(mx[layer], mh[layer], mw[layer], 8, t > 0 ? 8 : 0, toff[t], t > 0 ? toff[t - 1] : 0,
                                   active_host[t], 2048, e, s, pdl)));
    }
    return B2T_OK;
  };
  if (!overlap) {
    RUN(run_steps(0, 0, t_end, st, false));
    RUN(run_steps(1, 0, t_end, st, false));
  } else {
    const int n_chunks = (t_end + kChunk - 1) / kChunk;
    while ((int)evs.size() < n_chunks + 1) {
      cudaEvent_t ev;
      B2T_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      evs.push_back(ev);
    }
    for (int k = 0; k < n_chunks; ++k) {
      const int ta = k * kChunk, tb = std::min(t_end, ta + kChunk);
      RUN(run_steps(0, ta, tb, st, false));
      B2T_CUDA(cudaEventRecord(evs[k], st));
      B2T_CUDA(cudaStreamWaitEvent(side, evs[k], 0));
      RUN(run_steps(1, ta, tb, side, true));      // the first step after an event wait is a plain (fully ordered) launch
    }
    B2T_CUDA(cudaEventRecord(evs[n_chunks], side));
    B2T_CUDA(cudaStreamWaitEvent(st, evs[n_chunks], 0));
  }

  b2t_acoustic_mark(1, 0, st);
  b2t_acoustic_mark(2, 1, st);
  // ---- ELU -> Conv(512 -> 128, k7) over the clip-major padded rows (halo 6) -> dense fp32 embeddings ----
  {
    CUtensorMap ma, mw;
    const long long M = (long long)total4 + 6LL * n;
    RUN(make_map_k(&ma, w.s2e - 6 * 512, M, 7 * 512, 512, kBM));
    RUN(make_map_k(&mw, wt.final_w, 128, 7 * 512, 7 * 512, 128));
    ConvEpi e{};
    e.off4 = b->off[4]; e.rank = b->rank; e.toff = b->toff; e.c0 = 0; e.nsub = n; e.c0_out = 0;
    e.bias = wt.final_b; e.out_f32 = emb ? emb : w.emb; e.ld_f32 = 128; e.r = 1; e.h_in = 6; e.h_out = 0;
    RUN(launch_conv(ma, mw, 7 * 512, (int)M, 128, e, st));
  }
  b2t_acoustic_mark(2, 0, st);
#undef RUN
  return B2T_OK;
}

float* b2t_seanet_tc_emb(void* workspace, const b2t_acoustic_batch* b) {
  long long mf; int mc;
  split_clips(b->frames_host, b->n_clips, &mf, &mc);
  return tc_carve(workspace, mf, mc, b->total[4], b->n_clips).emb;
}
