// Relative-key flash attention on tcgen05 / TMEM / TMA (sm_100a)
// (reference audiotoken/modeling_wav2vec2_bert.py:37-77; see attention.cu for the maths).
//
// One CTA = one (clip, 128-query tile, head); 192 threads:
//   warp 0    TMA producer : Q tile + distance embedding E once, then a 3-stage ring of 128-key K and V tiles
//                            (all boxes 64 x rows out of the packed qkv matrix, SWIZZLE_128B).
//   warp 1    MMA issuer   : R = Q.E^T (128x80), then per key tile S = Q.K^T (128x128x64, 4 tcgen05.mma) and
//                            O_tile = P.V (128x64x128, 8 tcgen05.mma, V consumed MN-major straight from the TMA
//                            layout); S and O_tile are double-buffered in TMEM so QK^T of tile i+1 overlaps the
//                            softmax of tile i.
//   warps 2-5 softmax      : thread = query row = TMEM lane.  tcgen05.ld the S row, scale + relative-key bias
//                            (gathered from the thread's own R row only inside the diagonal band, a per-row
//                            constant elsewhere) + key mask, row max / exp2 / row sum without any shuffle,
//                            P -> bf16 -> shared memory in the K-major SWIZZLE_128B layout the PV MMA reads,
//                            O accumulation in registers with the deferred rescale of online softmax.
// Nothing of size T x T or T x 73 touches HBM.
#include "tc_ptx.cuh"

namespace {

constexpr int kHeads = 16, kHD = 64, kRel = 73, kLeft = 64, kRight = 8;
constexpr int kQKV = 3 * kHeads * kHD, kH = kHeads * kHD;
constexpr int kQT = 128, kKT = 128, kStages = 3;
constexpr int kThreadsAttn = 192;

struct AttnSmem {
  static constexpr int kQ = 0;                                  // 16 KB
  static constexpr int kKV = kQ + kQT * 128;                    // kStages x (K 16 KB + V 16 KB)
  static constexpr int kP = kKV + kStages * 2 * kKT * 128;      // 2 x 32 KB
  static constexpr int kE = kP + 2 * 2 * kQT * 128;             // 80 x 128 B = 10 KB (1024-aligned)
  static constexpr int kR = kE + 80 * 128;                      // [128][80] bf16 = 20 KB
  static constexpr int kBars = kR + kQT * 80 * 2;
  static constexpr int kTotal = kBars + 256 + 1024;
};
// barrier slots (8 bytes each)
enum { B_QFULL = 0, B_RFULL, B_KVFULL, B_KVEMPTY = B_KVFULL + kStages, B_SFULL = B_KVEMPTY + kStages, B_SEMPTY = B_SFULL + 2,
       B_PFULL = B_SEMPTY + 2, B_PEMPTY = B_PFULL + 2, B_PVFULL = B_PEMPTY + 2, B_PVEMPTY = B_PVFULL + 2, B_COUNT = B_PVEMPTY + 2 };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier area");

B2T_DEVICE float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
B2T_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
B2T_DEVICE void tmem_ld_32x32_x16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(kThreadsAttn, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_e,
                    const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                    const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0,
                    __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + AttnSmem::kQ, sKV = base + AttnSmem::kKV, sP = base + AttnSmem::kP, sE = base + AttnSmem::kE;
  __nv_bfloat16* sR = reinterpret_cast<__nv_bfloat16*>(gbase + AttnSmem::kR);
  const uint32_t bars = base + AttnSmem::kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gbase + AttnSmem::kBars + 8 * B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip = qtile_clip[blockIdx.x], q0 = qtile_q0[blockIdx.x], head = blockIdx.y;
  const int r0 = row_off[clip], rows = row_off[clip + 1] - r0, nkeys = valid_rows[clip];
  const int nkt = (nkeys + kKT - 1) / kKT;

  if (threadIdx.x == 0) {
    mbar_init(bar(B_QFULL), 1); mbar_init(bar(B_RFULL), 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(B_KVFULL + s), 1); mbar_init(bar(B_KVEMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(B_SFULL + b), 1); mbar_init(bar(B_SEMPTY + b), 4);
      mbar_init(bar(B_PFULL + b), 4); mbar_init(bar(B_PEMPTY + b), 1);
      mbar_init(bar(B_PVFULL + b), 1); mbar_init(bar(B_PVEMPTY + b), 4);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_qkv); tma_prefetch_desc(&map_e); }
  if (warp == 1) tmem_alloc(bars + 8u * B_COUNT, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tS = tmem_base, tPV = tmem_base + 256, tR = tmem_base + 384;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(bar(B_QFULL), kQT * 128 + 80 * 128);
      tma_load_2d(sQ, &map_qkv, bar(B_QFULL), head * kHD, r0 + q0);
      tma_load_2d(sE, &map_e, bar(B_QFULL), 0, 0);
      for (int i = 0; i < nkt; ++i) {
        const int st = i % kStages;
        mbar_wait(bar(B_KVEMPTY + st), ((i / kStages) & 1) ^ 1u);
        mbar_expect_tx(bar(B_KVFULL + st), 2 * kKT * 128);
        tma_load_2d(sKV + st * 2 * kKT * 128, &map_qkv, bar(B_KVFULL + st), kH + head * kHD, r0 + i * kKT);
        tma_load_2d(sKV + st * 2 * kKT * 128 + kKT * 128, &map_qkv, bar(B_KVFULL + st), 2 * kH + head * kHD, r0 + i * kKT);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, kKT);          // S = Q K^T   (both K-major)
      constexpr uint32_t idesc_r = make_idesc(128, 80);           // R = Q E^T
      constexpr uint32_t idesc_o = make_idesc(128, kHD, 1);       // O = P V     (V MN-major)
      mbar_wait(bar(B_QFULL), 0);
      tc_fence_after();
      const uint64_t dq = make_smem_desc(sQ), de = make_smem_desc(sE);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tR, dq + (uint64_t)(2 * k), de + (uint64_t)(2 * k), idesc_r, k != 0);
      umma_commit(bar(B_RFULL));
      auto issue_pv = [&](int j) {
        const int st = j % kStages, b = j & 1;
        mbar_wait(bar(B_PFULL + b), (j >> 1) & 1);
        mbar_wait(bar(B_PVEMPTY + b), ((j >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t sv = sKV + st * 2 * kKT * 128 + kKT * 128;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t dp = make_smem_desc(sP + b * 2 * kQT * 128 + (kk >> 2) * kQT * 128) + (uint64_t)(2 * (kk & 3));
          const uint64_t dv = make_smem_desc(sv + kk * 16 * 128);
          umma_bf16(tPV + (uint32_t)(b * kHD), dp, dv, idesc_o, kk != 0);
        }
        umma_commit(bar(B_PVFULL + b));
        umma_commit(bar(B_PEMPTY + b));
        umma_commit(bar(B_KVEMPTY + st));
      };
      for (int i = 0; i < nkt; ++i) {
        const int st = i % kStages, b = i & 1;
        mbar_wait(bar(B_KVFULL + st), (i / kStages) & 1);
        mbar_wait(bar(B_SEMPTY + b), ((i >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint64_t dk = make_smem_desc(sKV + st * 2 * kKT * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tS + (uint32_t)(b * kKT), dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
        umma_commit(bar(B_SFULL + b));
        if (i > 0) issue_pv(i - 1);
      }
      issue_pv(nkt - 1);
    }
  } else {
    // ===== softmax / output warps: thread = query row = TMEM lane =====
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int qpos = q0 + r;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    constexpr float kScale = 0.125f * 1.4426950408889634f;   // log2 domain

    // R row -> bf16 (the reference's einsum output dtype) -> shared memory (this thread's row only)
    mbar_wait(bar(B_RFULL), 0);
    tc_fence_after();
    {
      uint32_t a[32], b2[32], c[16];
      tmem_ld_32x32_nowait(tR + lane_base, a);
      tmem_ld_32x32_nowait(tR + lane_base + 32, b2);
      tmem_ld_32x32_x16_nowait(tR + lane_base + 64, c);
      tmem_ld_wait();
      __nv_bfloat16* rr = sR + r * 80;
#pragma unroll
      for (int i = 0; i < 32; ++i) { rr[i] = __float2bfloat16_rn(__uint_as_float(a[i])); rr[32 + i] = __float2bfloat16_rn(__uint_as_float(b2[i])); }
#pragma unroll
      for (int i = 0; i < 16; ++i) rr[64 + i] = __float2bfloat16_rn(__uint_as_float(c[i]));
    }
    const __nv_bfloat16* myR = sR + r * 80;
    const float rl = __bfloat162float(myR[0]) * kScale, rrt = __bfloat162float(myR[kRel - 1]) * kScale;

    float m = -INFINITY, l = 0.f, corr_prev = 1.f;
    float o[kHD];
#pragma unroll
    for (int d = 0; d < kHD; ++d) o[d] = 0.f;

    auto absorb_pv = [&](int j) {   // O = O * corr_j + PV_j
      const int b = j & 1;
      mbar_wait(bar(B_PVFULL + b), (j >> 1) & 1);
      tc_fence_after();
      uint32_t x[32], y[32];
      tmem_ld_32x32_nowait(tPV + lane_base + (uint32_t)(b * kHD), x);
      tmem_ld_32x32_nowait(tPV + lane_base + (uint32_t)(b * kHD) + 32, y);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_PVEMPTY + b));
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        o[d] = fmaf(o[d], corr_prev, __uint_as_float(x[d]));
        o[32 + d] = fmaf(o[32 + d], corr_prev, __uint_as_float(y[d]));
      }
    };

    for (int i = 0; i < nkt; ++i) {
      const int b = i & 1, k0 = i * kKT;
      mbar_wait(bar(B_SFULL + b), (i >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tS + lane_base + (uint32_t)(b * kKT);
      const int dlo = k0 - qpos, dhi = k0 + kKT - 1 - qpos;
      const int mode = dhi <= -kLeft ? 0 : (dlo >= kRight ? 1 : 2);   // all-left / all-right / diagonal band
      const float cb = mode == 0 ? rl : rrt;
      auto score = [&](float s, int kj) -> float {
        float t;
        if (mode == 2) {
          const int idx = max(-kLeft, min(kRight, kj - qpos)) + kLeft;
          t = (s + __bfloat162float(myR[idx])) * kScale;
        } else {
          t = fmaf(s, kScale, cb);
        }
        return kj < nkeys ? t : -INFINITY;
      };
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kKT; c += 64) {
        uint32_t x[32], y[32];
        tmem_ld_32x32_nowait(ts + c, x);
        tmem_ld_32x32_nowait(ts + c + 32, y);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          mx = fmaxf(mx, score(__uint_as_float(x[e]), k0 + c + e));
          mx = fmaxf(mx, score(__uint_as_float(y[e]), k0 + c + 32 + e));
        }
      }
      const float mn = fmaxf(m, mx);
      const float corr = ex2a(m - mn);
      m = mn;
      // pass 2: P = 2^(t - m) -> bf16 -> shared memory (K-major, 128B swizzle, two 64-key halves)
      mbar_wait(bar(B_PEMPTY + b), ((i >> 1) & 1) ^ 1u);
      float ls = 0.f;
      uint8_t* pbuf = gbase + AttnSmem::kP + b * 2 * kQT * 128 + r * 128;
#pragma unroll 1
      for (int c = 0; c < kKT; c += 64) {
        uint32_t x[32], y[32];
        tmem_ld_32x32_nowait(ts + c, x);
        tmem_ld_32x32_nowait(ts + c + 32, y);
        tmem_ld_wait();
        float p[64];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          p[e] = ex2a(score(__uint_as_float(x[e]), k0 + c + e) - mn);
          p[32 + e] = ex2a(score(__uint_as_float(y[e]), k0 + c + 32 + e) - mn);
        }
#pragma unroll
        for (int e = 0; e < 64; ++e) ls += p[e];
        uint8_t* half = pbuf + (c >> 6) * kQT * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4 v;
          v.x = pack_bf16x2(p[ch * 8 + 0], p[ch * 8 + 1]); v.y = pack_bf16x2(p[ch * 8 + 2], p[ch * 8 + 3]);
          v.z = pack_bf16x2(p[ch * 8 + 4], p[ch * 8 + 5]); v.w = pack_bf16x2(p[ch * 8 + 6], p[ch * 8 + 7]);
          *reinterpret_cast<uint4*>(half + ((ch ^ (r & 7)) << 4)) = v;
        }
      }
      l = l * corr + ls;
      // S buffer drained, P visible to the tensor core (generic -> async proxy)
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(B_SEMPTY + b)); mbar_arrive(bar(B_PFULL + b)); }
      if (i > 0) absorb_pv(i - 1);
      corr_prev = corr;
    }
    absorb_pv(nkt - 1);
    if (qpos < rows) {
      const float inv = 1.0f / l;
      uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(r0 + qpos) * kH + head * kHD);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 v;
        v.x = pack_bf16x2(o[ch * 8 + 0] * inv, o[ch * 8 + 1] * inv); v.y = pack_bf16x2(o[ch * 8 + 2] * inv, o[ch * 8 + 3] * inv);
        v.z = pack_bf16x2(o[ch * 8 + 4] * inv, o[ch * 8 + 5] * inv); v.w = pack_bf16x2(o[ch * 8 + 6] * inv, o[ch * 8 + 7] * inv);
        dst[ch] = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// host entry used by b2t_relkey_attention (attention.cu)
int b2t_attention_tensor_tc(const void* qkv, const void* dist_emb, const b2t_batch* b, void* out, cudaStream_t st) {
  B2T_REQUIRE(b->n_qtiles128 > 0 && b->qtile128_clip && b->qtile128_q0, B2T_ERR_ARG,
              "b2t_relkey_attention(tcgen05): the batch has no 128-row query tiles");
  CUtensorMap mq, me;
  int rc = make_map(&mq, qkv, b->total_rows, kQKV, kQKV, 128);
  if (rc != B2T_OK) return rc;
  rc = make_map(&me, dist_emb, kRel, kHD, kHD, 80);
  if (rc != B2T_OK) return rc;
  static bool cfg = false;
  if (!cfg) {
    B2T_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal));
    cfg = true;
  }
  dim3 grid(b->n_qtiles128, kHeads);
  attention_tc_kernel<<<grid, kThreadsAttn, AttnSmem::kTotal, st>>>(mq, me, b->row_off, b->valid_rows, b->qtile128_clip,
                                                                   b->qtile128_q0, (__nv_bfloat16*)out);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
