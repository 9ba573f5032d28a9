// Relative-key flash attention on tcgen05 / TMEM / TMA (sm_100a)
// (reference audiotoken/modeling_wav2vec2_bert.py:37-77; see attention.cu for the maths).
//
// Two kernels share this file's layout and helpers:
//   attention_tc2_kernel  (default)  two-pass, fixed row bound: QK^T once for the row maxima, a second time for
//                                    P = 2^(score - bound); O accumulated in TMEM; TMA / S-issue / PV-issue in three
//                                    converged warps.  Described in front of the kernel below.
//   attention_tc_kernel   (attn_two_pass = 0)  the online-softmax predecessor, kept for A/B runs:
//
// One CTA = one (clip, 128-query tile, head); 320 threads:
//   warp 0    TMA producer : Q tile + distance embedding E once, then a ring of 64-key K and V tiles
//                            (all boxes 64 x rows out of the packed qkv matrix, SWIZZLE_128B).
//   warp 1    MMA issuer   : R = Q.E^T (128x80), then per key tile S = Q.K^T (128x64x64, 4 tcgen05.mma) and
//                            O_tile = P.V (128x64x64, 4 tcgen05.mma, V consumed MN-major straight from the TMA
//                            layout); S and O_tile are double-buffered in TMEM so QK^T of tile i+1 overlaps the
//                            softmax of tile i.
//   warps 2-9 softmax      : thread = (query row = TMEM lane, 32-key half).  tcgen05.ld the S half row, scale + relative-key bias
//                            (gathered from the thread's own R row only inside the diagonal band, a per-row
//                            constant elsewhere) + key mask, row max / exp2 / row sum without any shuffle,
//                            P -> bf16 -> shared memory in the K-major SWIZZLE_128B layout the PV MMA reads,
//                            O accumulation in registers with the deferred rescale of online softmax.
// Nothing of size T x T or T x 73 touches HBM.
#include "tc_ptx.cuh"

namespace {

constexpr int kHeads = 16, kHD = 64, kRel = 73, kLeft = 64, kRight = 8;
constexpr int kQKV = 3 * kHeads * kHD, kH = kHeads * kHD;
#ifndef B2T_ATTN_WIDE
#define B2T_ATTN_WIDE 0
#endif
#if B2T_ATTN_WIDE
// wide configuration (kept as an experiment, measured SLOWER: 212 / 294 TFLOP/s on 10 s / 30 s clips against
// 259 / 340 for the default): 128-key tiles, one CTA per SM, 16 softmax warps (4 per TMEM lane quadrant, 32 keys
// each); TMEM: S 2 x 128 + PV 2 x 64 columns.  Two co-resident CTAs hide the hand-off latencies better than
// doubling the work per hand-off.
constexpr int kQT = 128, kKT = 128, kStages = 2;
constexpr int kSoftmaxWarps = 16;
constexpr int kCtasPerSm = 1, kTmemCols = 512;
#else
constexpr int kQT = 128, kKT = 64, kStages = 2;
constexpr int kSoftmaxWarps = 8;                 // two warps per TMEM lane quadrant: each owns a 32-key slice of S
#ifndef B2T_ATTN_CTAS
#define B2T_ATTN_CTAS 2
#endif
constexpr int kCtasPerSm = B2T_ATTN_CTAS, kTmemCols = 256;
#endif

#ifndef B2T_ATTN_POLL_SLEEP
#define B2T_ATTN_POLL_SLEEP 0                   // ns the polling TMA producer sleeps when none of its streams can advance
#endif
constexpr int kPBuf = kQT * kKT * 2;             // one P buffer: kKT / 64 K-major blocks of 128 rows x 128 B
constexpr int kThreadsAttn = 64 + 32 * kSoftmaxWarps;

// Two CTAs are resident per SM (<= 113 KB shared memory, 256 TMEM columns, 320 threads each): while one CTA
// sits in a TMEM-load / barrier / proxy-fence latency the other one computes — the two softmax pipelines
// interleave without any explicit ping-pong protocol.
constexpr int kRS = 82;                          // row stride of the R tile in shared memory (bf16 elements)
struct AttnSmem {
  static constexpr int kQ = 0;                                  // 16 KB
  static constexpr int kKV = kQ + kQT * 128;                    // kStages x (K 8 KB + V 8 KB)
  static constexpr int kP = kKV + kStages * 2 * kKT * 128;      // 2 P buffers; buffer 1 first stages E (80 x 128 B)
  static constexpr int kE = kP + kPBuf;                         // = P buffer 1 (E is dead once R has been computed)
  static constexpr int kR = kP + 2 * kPBuf;                     // [128][kRS] bf16: rows 41 words apart, so that a warp's per-row
                                                                // accesses (lane = row) spread over all 32 banks (80 gave 8-way conflicts)
  static constexpr int kMax = (kR + kQT * kRS * 2 + 1023) & ~1023;   // row-max exchange [2 slots][kWG][128] fp32 (+ the ones block)
  static constexpr int kSum = kMax + 2 * 4 * kQT * 4;           // final row-sum exchange [kWG][128] fp32
  static constexpr int kBars = kSum + 4 * kQT * 4;
  static constexpr int kTotal = kBars + 256 + 1024;
};
static_assert(kCtasPerSm * AttnSmem::kTotal <= 227 * 1024, "shared memory per SM");
// barrier slots (8 bytes each)
enum { B_EFULL = 0, B_QFULL, B_QEMPTY = B_QFULL + 2, B_RFULL = B_QEMPTY + 2, B_REMPTY, B_KVFULL, B_KVEMPTY = B_KVFULL + kStages,
       B_SFULL = B_KVEMPTY + kStages, B_SEMPTY = B_SFULL + 2, B_PFULL = B_SEMPTY + 2, B_PEMPTY = B_PFULL + 2,
       B_PVFULL = B_PEMPTY + 2, B_PVEMPTY = B_PVFULL + 2, B_COUNT = B_PVEMPTY + 2 };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier area");

// named barrier 1 + quad among the warps that own the same 32 query rows (immediate ids: the kernel then reserves
// 5 hardware barriers instead of all 16)
template <int kCount>
B2T_DEVICE void row_barrier(int quad) {
  switch (quad) {
    case 0: asm volatile("bar.sync 1, %0;" ::"n"(kCount) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;" ::"n"(kCount) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;" ::"n"(kCount) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;" ::"n"(kCount) : "memory"); break;
  }
}
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
B2T_DEVICE void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32_nowait(taddr, r); }
B2T_DEVICE void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x32_x16_nowait(taddr, r); }

// The CTA is persistent over the 16 heads of its (clip, query tile): barriers, the TMEM allocation and E are
// set up once, the producer prefetches the next head's Q and K/V while the softmax warps finish the current
// head.  g = head * nkt + i is the running key-tile counter that drives every ring / phase.
__global__ void __launch_bounds__(kThreadsAttn, kCtasPerSm)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv,
                    const __grid_constant__ CUtensorMap map_e,
                    const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                    const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0,
                    __nv_bfloat16* __restrict__ out, int hpc /* heads per CTA: blockIdx.y selects the group */) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + AttnSmem::kQ, sKV = base + AttnSmem::kKV, sP = base + AttnSmem::kP, sE = base + AttnSmem::kE;
  __nv_bfloat16* sR = reinterpret_cast<__nv_bfloat16*>(gbase + AttnSmem::kR);
  const uint32_t bars = base + AttnSmem::kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gbase + AttnSmem::kBars + 8 * B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip = qtile_clip[blockIdx.x], q0 = qtile_q0[blockIdx.x];
  const int r0 = row_off[clip], rows = row_off[clip + 1] - r0, nkeys = valid_rows[clip];
  const int nkt = (nkeys + kKT - 1) / kKT;
  const int head0 = blockIdx.y * hpc;

  if (threadIdx.x == 0) {
    mbar_init(bar(B_EFULL), 1); mbar_init(bar(B_RFULL), 1); mbar_init(bar(B_REMPTY), kSoftmaxWarps);
    for (int s = 0; s < 2; ++s) { mbar_init(bar(B_QFULL + s), 1); mbar_init(bar(B_QEMPTY + s), 1); }
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(B_KVFULL + s), 1); mbar_init(bar(B_KVEMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(B_SFULL + b), 1); mbar_init(bar(B_SEMPTY + b), kSoftmaxWarps);
      mbar_init(bar(B_PFULL + b), kSoftmaxWarps); mbar_init(bar(B_PEMPTY + b), 1);
      mbar_init(bar(B_PVFULL + b), 1); mbar_init(bar(B_PVEMPTY + b), kSoftmaxWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_qkv); tma_prefetch_desc(&map_kv); tma_prefetch_desc(&map_e); }
  if (warp == 1) tmem_alloc(bars + 8u * B_COUNT, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // S double buffer [0,128), PV double buffer [128,256); R (80 columns) borrows the PV region before the first PV MMA
  const uint32_t tS = tmem_base, tPV = tmem_base + 2 * kKT, tR = tmem_base + 2 * kKT;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(bar(B_EFULL), 80 * 128);
      tma_load_2d(sE, &map_e, bar(B_EFULL), 0, 0);
      for (int h = 0; h < hpc; ++h) {
        const int qb = 0, head = head0 + h;
        mbar_wait(bar(B_QEMPTY + qb), (h & 1) ^ 1u);
        mbar_expect_tx(bar(B_QFULL + qb), kQT * 128);
        tma_load_2d(sQ + qb * kQT * 128, &map_qkv, bar(B_QFULL + qb), head * kHD, r0 + q0);
        for (int i = 0; i < nkt; ++i) {
          const int g = h * nkt + i, st = g % kStages;
          mbar_wait(bar(B_KVEMPTY + st), ((g / kStages) & 1) ^ 1u);
          mbar_expect_tx(bar(B_KVFULL + st), 2 * kKT * 128);
          tma_load_2d(sKV + st * 2 * kKT * 128, &map_kv, bar(B_KVFULL + st), kH + head * kHD, r0 + i * kKT);
          tma_load_2d(sKV + st * 2 * kKT * 128 + kKT * 128, &map_kv, bar(B_KVFULL + st), 2 * kH + head * kHD, r0 + i * kKT);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, kKT);          // S = Q K^T   (both K-major)
      constexpr uint32_t idesc_r = make_idesc(128, 80);           // R = Q E^T
      constexpr uint32_t idesc_o = make_idesc(128, kHD, 1);       // O = P V     (V MN-major)
      const uint64_t de = make_smem_desc(sE);
      auto issue_pv = [&](int g) {
        const int st = g % kStages, b = g & 1;
        mbar_wait(bar(B_PFULL + b), (g >> 1) & 1);
        mbar_wait(bar(B_PVEMPTY + b), ((g >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t sv = sKV + st * 2 * kKT * 128 + kKT * 128;
#pragma unroll
        for (int kk = 0; kk < kKT / 16; ++kk) {
          const uint64_t dp = make_smem_desc(sP + b * kPBuf + (kk >> 2) * (kQT * 128)) + (uint64_t)(2 * (kk & 3));
          const uint64_t dv = make_smem_desc(sv + kk * 16 * 128);
          umma_bf16(tPV + (uint32_t)(b * kHD), dp, dv, idesc_o, kk != 0);
        }
        umma_commit(bar(B_PVFULL + b));
        umma_commit(bar(B_PEMPTY + b));
        umma_commit(bar(B_KVEMPTY + st));
      };
      mbar_wait(bar(B_EFULL), 0);
      for (int h = 0; h < hpc; ++h) {
        const int qb = 0;
        mbar_wait(bar(B_QFULL + qb), h & 1);
        mbar_wait(bar(B_REMPTY), (h & 1) ^ 1u);          // softmax warps have read the previous head's R
        tc_fence_after();
        const uint64_t dq = make_smem_desc(sQ + qb * kQT * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tR, dq + (uint64_t)(2 * k), de + (uint64_t)(2 * k), idesc_r, k != 0);
        umma_commit(bar(B_RFULL));
        for (int i = 0; i < nkt; ++i) {
          const int g = h * nkt + i, st = g % kStages, b = g & 1;
          mbar_wait(bar(B_KVFULL + st), (g / kStages) & 1);
          mbar_wait(bar(B_SEMPTY + b), ((g >> 1) & 1) ^ 1u);
          tc_fence_after();
          const uint64_t dk = make_smem_desc(sKV + st * 2 * kKT * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS + (uint32_t)(b * kKT), dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
          umma_commit(bar(B_SFULL + b));
          if (i == nkt - 1) umma_commit(bar(B_QEMPTY + qb));     // last use of this head's Q
          if (g > 0) issue_pv(g - 1);
        }
      }
      issue_pv(hpc * nkt - 1);
    }
  } else {
    // ===== softmax / output warps: thread = (query row = TMEM lane, key slice wg of kKW keys) =====
    constexpr int kWG = kSoftmaxWarps / 4;          // warps per TMEM lane quadrant
    constexpr int kKW = kKT / kWG;                  // keys of each S tile handled by one thread (32 or 64)
    constexpr int kDW = kHD / kWG;                  // head dims of the output handled by one thread
    const int quad = warp & 3;
    const int wg = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int qpos = q0 + r;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    constexpr float kScale = 0.125f * 1.4426950408889634f;   // log2 domain
    float* smax = reinterpret_cast<float*>(gbase + AttnSmem::kMax);   // [slot][wg][row]
    float* ssum = reinterpret_cast<float*>(gbase + AttnSmem::kSum);
    const __nv_bfloat16* myR = sR + r * kRS;

    float corr_prev = 1.f;
    float o[kDW];
    // O = O * corr + PV of global tile g; on the first tile of a head the accumulator restarts
    auto absorb_pv = [&](int g, bool first) {
      const int b = g & 1;
      mbar_wait(bar(B_PVFULL + b), (g >> 1) & 1);
      tc_fence_after();
      uint32_t x[kDW];
      tmem_ld_cols(tPV + lane_base + (uint32_t)(b * kHD + wg * kDW), x);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_PVEMPTY + b));
#pragma unroll
      for (int d = 0; d < kDW; ++d) o[d] = first ? __uint_as_float(x[d]) : fmaf(o[d], corr_prev, __uint_as_float(x[d]));
    };

    for (int h = 0; h < hpc; ++h) {
      // R row -> bf16 (the reference's einsum output dtype) -> shared memory; the warps of a row split the columns
      mbar_wait(bar(B_RFULL), h & 1);
      tc_fence_after();
      {
        __nv_bfloat16* rr = sR + r * kRS;
        if (wg < 2) {
          uint32_t a[32];
          tmem_ld_32x32_nowait(tR + lane_base + wg * 32, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[wg * 32 + i] = __float2bfloat16_rn(__uint_as_float(a[i]));
        }
        if (wg == kWG - 1) {
          uint32_t c[16];
          tmem_ld_32x32_x16_nowait(tR + lane_base + 64, c);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) rr[64 + i] = __float2bfloat16_rn(__uint_as_float(c[i]));
        }
      }
      tc_fence_before();
      row_barrier<32 * kWG>(quad);   // only the warps sharing this row   // every R row is complete
      if (lane == 0) mbar_arrive(bar(B_REMPTY));
      const float rl = __bfloat162float(myR[0]) * kScale, rrt = __bfloat162float(myR[kRel - 1]) * kScale;
      float m = -INFINITY, l = 0.f;

      for (int i = 0; i < nkt; ++i) {
        const int g = h * nkt + i, b = g & 1, k0 = i * kKT + wg * kKW;      // first key of this thread's slice
        mbar_wait(bar(B_SFULL + b), (g >> 1) & 1);
        tc_fence_after();
        float t[kKW];
        {
          const uint32_t ts = tS + lane_base + (uint32_t)(b * kKT + wg * kKW);
          uint32_t x[kKW / 32][32];
#pragma unroll
          for (int u = 0; u < kKW / 32; ++u) tmem_ld_32x32_nowait(ts + 32 * u, x[u]);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < kKW / 32; ++u)
#pragma unroll
            for (int e = 0; e < 32; ++e) t[32 * u + e] = __uint_as_float(x[u][e]);
        }
        // the S slice is in registers now: hand the buffer back to the MMA warp early
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_SEMPTY + b));
        // scores in the log2 domain: t = (q.k + bias) * log2e/8.  Outside the diagonal band the bias is one
        // constant per row, so the row maximum is taken on the raw q.k and scale, bias and -max fold into a
        // single FMA in front of the exp2.
        const int dlo = k0 - qpos, dhi = k0 + kKW - 1 - qpos;
        const bool band = !(dhi <= -kLeft || dlo >= kRight);
        const float cb = dhi <= -kLeft ? rl : rrt;
        if (band) {
#pragma unroll
          for (int e = 0; e < kKW; ++e) {
            const int idx = max(-kLeft, min(kRight, dlo + e)) + kLeft;
            t[e] = t[e] + __bfloat162float(myR[idx]);
          }
        }
        if (k0 + kKW > nkeys) {
#pragma unroll
          for (int e = 0; e < kKW; ++e) if (k0 + e >= nkeys) t[e] = -INFINITY;
        }
        float mx = -INFINITY;
#pragma unroll
        for (int e = 0; e < kKW; ++e) mx = fmaxf(mx, t[e]);
        mx = band ? mx * kScale : fmaf(mx, kScale, cb);          // this slice's maximum in the log2 domain
        // combine the row maximum with the warps that own the other key slices of this row
        smax[((g & 1) * kWG + wg) * kQT + r] = mx;
        row_barrier<32 * kWG>(quad);   // only the warps sharing this row
#pragma unroll
        for (int u = 0; u < kWG; ++u) mx = fmaxf(mx, smax[((g & 1) * kWG + u) * kQT + r]);
        const float mn = fmaxf(m, mx);
        const float corr = ex2a(m - mn);
        m = mn;
        const float off = (band ? 0.f : cb) - mn;                  // p = 2^(raw * kScale + off)
        // P -> bf16 -> shared memory (K-major, 128B swizzle)
        mbar_wait(bar(B_PEMPTY + b), ((g >> 1) & 1) ^ 1u);
        float ls = 0.f;
        const int kcol = wg * kKW;                                   // key column within the tile
        uint8_t* half = gbase + AttnSmem::kP + b * kPBuf + (kcol >> 6) * (kQT * 128) + r * 128;
        const int ch0 = (kcol & 63) >> 3;
#pragma unroll
        for (int ch = 0; ch < kKW / 8; ++ch) {
          float pv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { pv[e] = ex2a(fmaf(t[ch * 8 + e], kScale, off)); ls += pv[e]; }
          uint4 v;
          v.x = pack_bf16x2(pv[0], pv[1]); v.y = pack_bf16x2(pv[2], pv[3]);
          v.z = pack_bf16x2(pv[4], pv[5]); v.w = pack_bf16x2(pv[6], pv[7]);
          *reinterpret_cast<uint4*>(half + (((ch0 + ch) ^ (r & 7)) << 4)) = v;
        }
        l = l * corr + ls;
        fence_proxy_async();          // P visible to the tensor core (generic -> async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PFULL + b));
        if (i > 0) absorb_pv(g - 1, i == 1);
        corr_prev = corr;
      }
      absorb_pv(h * nkt + nkt - 1, nkt == 1);
      // row sum = sum over the key slices (same running maximum in all warps of a row)
      ssum[wg * kQT + r] = l;
      row_barrier<32 * kWG>(quad);   // only the warps sharing this row
      l = 0.f;
#pragma unroll
      for (int u = 0; u < kWG; ++u) l += ssum[u * kQT + r];
      if (qpos < rows) {
        const float inv = 1.0f / l;
        // one row per lane: 256-bit stores write whole 32-byte sectors (kDW is a multiple of 16)
        __nv_bfloat16* dst = out + (size_t)(r0 + qpos) * kH + (head0 + h) * kHD + wg * kDW;
#pragma unroll
        for (int ch = 0; ch < kDW / 16; ++ch) {
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = pack_bf16x2(o[ch * 16 + 2 * e] * inv, o[ch * 16 + 2 * e + 1] * inv);
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"l"(dst + 16 * ch), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                       : "memory");
        }
      }
      // ssum is rewritten only after the next head's R barrier; sR likewise (bar.sync at the top of the loop)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


#if !B2T_ATTN_WIDE
// ---- two-pass variant ------------------------------------------------------------------------------------
// Online softmax costs this kernel a cross-warp max exchange, a correction factor and a TMEM read + rescale of the
// output accumulator per key tile — more instructions and hand-offs than the exponentials themselves, while the
// tensor pipe idles (13 % busy).  Here the row maximum is fixed BEFORE any exponential is taken:
//   pass 1  S = Q.K^T for every key tile (tensor core), softmax warps keep only the running maximum of their slice;
//           row bound m = (max raw score + max relative-key bias of the row) / 8 >= every biased score of the row
//           (a bound instead of the exact maximum is enough: P stays <= 1 and bf16 keeps fp32's exponent range);
//   pass 2  S again (QK^T is recomputed: +50 % tensor work on an idle pipe), P = 2^(score - m) -> bf16 -> shared
//           memory, O += P.V accumulated in TMEM across all key tiles; O is read once per head, the row sums are
//           combined once per head.
// Per key tile a softmax thread runs LDTM + max (pass 1) and LDTM + FMA/EX2/ADD + pack + store (pass 2): no
// exchange, no correction, no accumulator traffic.  K/V tiles move through a ring of four 8 KB slots (pass 1 uses
// one per tile, pass 2 two), the S and P double buffers are shared by both passes.
namespace tp {
enum { EFULL = 0, QFULL, RFULL, OFULL, KVFULL, KVEMPTY = KVFULL + 4, SFULL = KVEMPTY + 4, SEMPTY = SFULL + 2,
       PFULL = SEMPTY + 2, PEMPTY = PFULL + 2, COUNT = PEMPTY + 2 };
constexpr int kSlots = 4, kSlotBytes = kKT * 128;
static_assert(COUNT * 8 + 8 <= 256, "barrier area");
static_assert(kSlots * kSlotBytes == kStages * 2 * kKT * 128, "the slot ring reuses the K/V stage area");
}  // namespace tp

// kOnesSum: the row sums come from the tensor core as well — a second accumulator L += P . 1 (N = 16 MMAs against a
//           constant [16 keys x 16] block whose first column is one), i.e. the sum of exactly the bf16 values that
//           enter P.V; removes 32 FADD per thread and tile and the cross-warp sum exchange.
//           Measured SLOWER (404 vs 437 TFLOP/s on 30 s clips): the four extra MMAs lengthen the PV issue stream,
//           which is worth more than the 32 FADD they save in the (parallel) softmax warps.  Kept as an option.
constexpr int kThreadsAttn2 = 96 + 32 * kSoftmaxWarps;          // TMA, S-issue and PV-issue warps + softmax warps
template <bool kOnesSum>
__global__ void __launch_bounds__(kThreadsAttn2, kCtasPerSm)
attention_tc2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv,
                     const __grid_constant__ CUtensorMap map_e,
                     const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                     const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0,
                     __nv_bfloat16* __restrict__ out, long long* __restrict__ dbg /* developer timeline of one CTA, or null */,
                     int H /* hidden width = 64 * heads (grid.y = heads); qkv rows are q | k | v, each H wide */) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + AttnSmem::kQ, sKV = base + AttnSmem::kKV, sP = base + AttnSmem::kP, sE = base + AttnSmem::kE;
  __nv_bfloat16* sR = reinterpret_cast<__nv_bfloat16*>(gbase + AttnSmem::kR);
  const uint32_t bars = base + AttnSmem::kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gbase + AttnSmem::kBars + 8 * tp::COUNT);
  // timeline: CTA (300, 5) — a mid-grid CTA so that both CTAs of the SM are in steady state; stamps of warp 2 lane 0
  const bool tl = dbg != nullptr && blockIdx.x == 300 && blockIdx.y == 5 && threadIdx.x == 96;
  int tln = 0;
#ifdef B2T_ATTN_TIMELINE
  auto stamp = [&]() { if (tl && tln < 128) dbg[tln++] = clock64(); };
#else
  auto stamp = [&]() { (void)tl; (void)tln; };
#endif
  stamp();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip = qtile_clip[blockIdx.x], q0 = qtile_q0[blockIdx.x];
  const int r0 = row_off[clip], rows = row_off[clip + 1] - r0, nkeys = valid_rows[clip];
  const int nkt = (nkeys + kKT - 1) / kKT;
  const int head = blockIdx.y;

  if (threadIdx.x == 0) {
    mbar_init(bar(tp::EFULL), 1); mbar_init(bar(tp::QFULL), 1); mbar_init(bar(tp::RFULL), 1); mbar_init(bar(tp::OFULL), 1);
    for (int s = 0; s < tp::kSlots; ++s) { mbar_init(bar(tp::KVFULL + s), 1); mbar_init(bar(tp::KVEMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(tp::SFULL + b), 1); mbar_init(bar(tp::SEMPTY + b), kSoftmaxWarps);
      mbar_init(bar(tp::PFULL + b), kSoftmaxWarps); mbar_init(bar(tp::PEMPTY + b), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_qkv); tma_prefetch_desc(&map_kv); tma_prefetch_desc(&map_e); }
  if (warp == 1) tmem_alloc(bars + 8u * tp::COUNT, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // S double buffer [0,128); O accumulator [128,192); R (80 columns) borrows the O region until the first PV MMA
  const uint32_t tS = tmem_base, tO = tmem_base + 2 * kKT, tR = tmem_base + 2 * kKT;
  const uint32_t tL = tmem_base + 2 * kKT + kHD;   // [192,208): row sums (column 0), kOnesSum only
  const uint32_t sOnes = base + AttnSmem::kMax + 2048;          // 2 KB, 1024-aligned, behind the max exchange slots

  // Ring position p (0 .. 3 nkt - 1): K tiles of pass 1, then K(0) V(0) K(1) V(1) ...; slot p & 3, use count p >> 2.
  // The three single-thread instruction streams that feed the tensor core run in DIFFERENT warps: warp 0 issues the
  // TMA loads, warp 1 the S MMAs, warp 2 the PV MMAs.  One warp doing all of it was the critical path of the kernel
  // (clock64 timeline: ~1500 clocks of issue latency per key tile against ~900 clocks of softmax work); every
  // instruction added to one of these streams still shows up in the kernel time.
  if (warp == 0) {
    // ===== TMA producer: whole warp waits (uniform control flow), one elected lane issues =====
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(bar(tp::EFULL), 80 * 128);
      tma_load_2d(sE, &map_e, bar(tp::EFULL), 0, 0);
      mbar_expect_tx(bar(tp::QFULL), kQT * 128);
      tma_load_2d(sQ, &map_qkv, bar(tp::QFULL), head * kHD, r0 + q0);
    }
    const int total = 3 * nkt;
    for (int c = 0; c < total; ++c) {
      const uint32_t st = (uint32_t)c & 3u;
      const int t2 = c - nkt;
      const int tile = t2 < 0 ? c : (t2 >> 1);
      const int col = (t2 >= 0 && (t2 & 1)) ? 2 * H : H;
      mbar_wait(bar(tp::KVEMPTY) + 8u * st, (((uint32_t)c >> 2) & 1u) ^ 1u);
      if (leader) {
        mbar_expect_tx(bar(tp::KVFULL) + 8u * st, tp::kSlotBytes);
        tma_load_2d(sKV + st * tp::kSlotBytes, &map_kv, bar(tp::KVFULL) + 8u * st, col + head * kHD, r0 + tile * kKT);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===== PV issuer =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_o = make_idesc(128, kHD, 1);       // O += P V    (V MN-major)
    constexpr uint32_t idesc_l = make_idesc(128, 16, 1);        // L += P 1
    const uint64_t dkv = make_smem_desc(sKV), dp0 = make_smem_desc(sP), dones = make_smem_desc(sOnes);
    for (int ip = 0; ip < nkt; ++ip) {
      const uint32_t pb = (uint32_t)ip & 1u, pos = (uint32_t)(nkt + 2 * ip + 1), st = pos & 3u;
      const uint32_t bp = bar(tp::PFULL) + 8u * pb, bv = bar(tp::KVFULL) + 8u * st;
      const uint32_t pp = ((uint32_t)ip >> 1) & 1u, pv = (pos >> 2) & 1u;
      // One wait after the other.  Probing both barriers in one loop iteration (two mbarrier.try_wait in flight in
      // the warp, results AND-ed) is logically equivalent and ~1 % faster, but on the 8-GPU nodes it made about one
      // CTA in 3 million proceed early (a wrong 128-row tile, or a launch failure): tools/kernel_soak.py, round-2 notes
      // in profiles/.  Rule for this library: at most one potentially-blocking try_wait in flight per warp.
      mbar_wait(bp, pp);
      mbar_wait(bv, pv);
      tc_fence_after();
      if (leader) {
        const uint64_t dv = dkv + (uint64_t)(st * (tp::kSlotBytes >> 4)), dp = dp0 + (uint64_t)(pb * (kPBuf >> 4));
#pragma unroll
        for (int kk = 0; kk < kKT / 16; ++kk)
          umma_bf16(tO, dp + (uint64_t)(2 * kk), dv + (uint64_t)(kk * (16 * 128 >> 4)), idesc_o, (ip | kk) != 0);
        if constexpr (kOnesSum) {
#pragma unroll
          for (int kk = 0; kk < kKT / 16; ++kk) umma_bf16(tL, dp + (uint64_t)(2 * kk), dones, idesc_l, (ip | kk) != 0);
        }
        umma_commit(bar(tp::PEMPTY) + 8u * pb);
        umma_commit(bar(tp::KVEMPTY) + 8u * st);
        if (ip == nkt - 1) umma_commit(bar(tp::OFULL));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== S issuer (whole warp waits, one elected lane issues) =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc(128, kKT);          // S = Q K^T   (both K-major)
    constexpr uint32_t idesc_r = make_idesc(128, 80);           // R = Q E^T
    const uint64_t de = make_smem_desc(sE), dq = make_smem_desc(sQ), dkv = make_smem_desc(sKV);
    mbar_wait(bar(tp::EFULL), 0);
    mbar_wait(bar(tp::QFULL), 0);
    tc_fence_after();
    if (leader) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tR, dq + (uint64_t)(2 * k), de + (uint64_t)(2 * k), idesc_r, k != 0);
      umma_commit(bar(tp::RFULL));
    }
    // S tile j (j < nkt: pass 1, else pass 2) from ring position pos
    auto issue_s = [&](uint32_t j, uint32_t pos) {
      const uint32_t b = j & 1u, st = pos & 3u;
      const uint32_t bk = bar(tp::KVFULL) + 8u * st, bs = bar(tp::SEMPTY) + 8u * b;
      const uint32_t pk = (pos >> 2) & 1u, ps = ((j >> 1) & 1u) ^ 1u;
      mbar_wait(bk, pk);                            // one wait after the other (see the PV issuer)
      mbar_wait(bs, ps);
      tc_fence_after();
      if (leader) {
        const uint64_t dk = dkv + (uint64_t)(st * (tp::kSlotBytes >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tS + b * kKT, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
        umma_commit(bar(tp::SFULL) + 8u * b);
        umma_commit(bar(tp::KVEMPTY) + 8u * st);
      }
    };
    for (int j = 0; j < nkt; ++j) issue_s((uint32_t)j, (uint32_t)j);
    for (int i = 0; i < nkt; ++i) issue_s((uint32_t)(nkt + i), (uint32_t)(nkt + 2 * i));
    __syncwarp();
  } else {
    // ===== softmax / output warps: thread = (query row = TMEM lane, key slice wg of kKW keys) =====
    constexpr int kWG = kSoftmaxWarps / 4;          // warps per TMEM lane quadrant
    constexpr int kKW = kKT / kWG;                  // keys of each S tile handled by one thread
    constexpr int kDW = kHD / kWG;                  // head dims of the output handled by one thread
    static_assert(kWG == 2 && kKW == 32 && kDW == 32, "two-pass kernel: 8 softmax warps, 64-key tiles");
    const int quad = warp & 3;
    const int wg = (warp - 3) >> 2;
    const int r = quad * 32 + lane;
    const int qpos = q0 + r;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    constexpr float kScale = 0.125f * 1.4426950408889634f;   // log2 domain
    float* smax = reinterpret_cast<float*>(gbase + AttnSmem::kMax);   // [slot][wg][row]
    float* ssum = reinterpret_cast<float*>(gbase + AttnSmem::kSum);
    const __nv_bfloat16* myR = sR + r * kRS;

    // R row -> bf16 (the reference's einsum output dtype) -> shared memory; the two warps of a row split the
    // columns and keep the largest bias of their part
    stamp();                                        // [1] set-up done
    mbar_wait(bar(tp::RFULL), 0);
    stamp();                                        // [2] R ready
    tc_fence_after();
    float bmax = -INFINITY;
    {
      uint32_t* rr = reinterpret_cast<uint32_t*>(sR + r * kRS);     // bf16 pairs, 4-byte aligned (kRS is even)
      uint32_t a[32];
      tmem_ld_32x32_nowait(tR + lane_base + wg * 32, a);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(a[i]), __uint_as_float(a[i + 1]));
        rr[(wg * 32 + i) >> 1] = *reinterpret_cast<const uint32_t*>(&v);
        bmax = fmaxf(bmax, fmaxf(__low2float(v), __high2float(v)));
      }
      if (wg == kWG - 1) {
        uint32_t c[16];
        tmem_ld_32x32_x16_nowait(tR + lane_base + 64, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(c[i]), __uint_as_float(c[i + 1]));
          rr[(64 + i) >> 1] = *reinterpret_cast<const uint32_t*>(&v);
          if (64 + i < kRel) bmax = fmaxf(bmax, __low2float(v));
          if (64 + i + 1 < kRel) bmax = fmaxf(bmax, __high2float(v));
        }
      }
    }
    tc_fence_before();
    smax[(0 * kWG + wg) * kQT + r] = bmax;
    if constexpr (kOnesSum) {
      // [16 keys x 64] bf16 block in the MN-major SWIZZLE_128B layout of a V tile: element (key k, n = 0) = 1.
      // Written long before the first PV MMA; every writer's fence.proxy.async in front of its first PFULL arrival
      // publishes it to the tensor core.
      uint32_t* ob = reinterpret_cast<uint32_t*>(gbase + AttnSmem::kMax + 2048);
      const int tid = threadIdx.x - 96;             // 0 .. 255: 2 of the 512 words each
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int w = 2 * tid + j;                  // key row w >> 5; (k, n = 0) sits in chunk (0 ^ (k & 7)) of its row
        ob[w] = ((w & 31) == (((w >> 5) & 7) << 2)) ? 0x00003F80u : 0u;
      }
    }
    row_barrier<32 * kWG>(quad);                    // every R row is complete, both bias maxima are visible
    bmax = fmaxf(smax[(0 * kWG + 0) * kQT + r], smax[(0 * kWG + 1) * kQT + r]);
    const float rl = __bfloat162float(myR[0]) * kScale, rrt = __bfloat162float(myR[kRel - 1]) * kScale;
    stamp();                                        // [3] R in shared memory

    // S slice of tile j (running tile counter over both passes) into registers, buffer handed back at once
    auto load_s = [&](int j, float (&t)[kKW]) {
      const int b = j & 1;
      mbar_wait(bar(tp::SFULL + b), (j >> 1) & 1);
      tc_fence_after();
      uint32_t x[32];
      tmem_ld_32x32_nowait(tS + lane_base + (uint32_t)(b * kKT + wg * kKW), x);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(x[e]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(tp::SEMPTY + b));
    };

    // ---- pass 1: largest raw score of the row ----
    float mx = -INFINITY;
    for (int i = 0; i < nkt; ++i) {
      const int k0 = i * kKT + wg * kKW;
      float t[kKW];
      load_s(i, t);
      if (k0 + kKW > nkeys) {
#pragma unroll
        for (int e = 0; e < kKW; ++e) if (k0 + e >= nkeys) t[e] = -INFINITY;
      }
      float m0 = fmaxf(t[0], t[1]), m1 = fmaxf(t[2], t[3]);
#pragma unroll
      for (int e = 4; e < kKW; e += 4) { m0 = fmaxf(m0, fmaxf(t[e], t[e + 1])); m1 = fmaxf(m1, fmaxf(t[e + 2], t[e + 3])); }
      mx = fmaxf(mx, fmaxf(m0, m1));
      stamp();                                      // [4 + i] pass-1 tile done
    }
    smax[(1 * kWG + wg) * kQT + r] = mx;
    row_barrier<32 * kWG>(quad);
    mx = fmaxf(smax[(1 * kWG + 0) * kQT + r], smax[(1 * kWG + 1) * kQT + r]);
    const float m = (mx + bmax) * kScale;           // >= every biased score of this row, log2 domain

    // ---- pass 2: P = 2^(score - m), O += P.V on the tensor core ----
    float l = 0.f;
    for (int i = 0; i < nkt; ++i) {
      const int pb = i & 1, k0 = i * kKT + wg * kKW;
      float t[kKW];
      load_s(nkt + i, t);
      stamp();                                      // [4 + nkt + 3 i] S slice in registers
      // scores in the log2 domain: (q.k + bias) * log2e/8.  Outside the diagonal band the bias is one constant per
      // row, so scale, bias and -m fold into a single FMA in front of the exp2.
      const int dlo = k0 - qpos, dhi = k0 + kKW - 1 - qpos;
      const bool band = !(dhi <= -kLeft || dlo >= kRight);
      const float cb = dhi <= -kLeft ? rl : rrt;
      if (band) {
#pragma unroll
        for (int e = 0; e < kKW; ++e) {
          const int idx = max(-kLeft, min(kRight, dlo + e)) + kLeft;
          t[e] = t[e] + __bfloat162float(myR[idx]);
        }
      }
      if (k0 + kKW > nkeys) {
#pragma unroll
        for (int e = 0; e < kKW; ++e) if (k0 + e >= nkeys) t[e] = -INFINITY;
      }
      const float off = (band ? 0.f : cb) - m;                   // p = 2^(raw * kScale + off)
      float ls0 = 0.f, ls1 = 0.f;
      const int kcol = wg * kKW;                                   // key column within the tile
      uint8_t* half = gbase + AttnSmem::kP + pb * kPBuf + r * 128;
      const int ch0 = kcol >> 3;
      uint4 v[kKW / 8];                                            // the whole slice first, the buffer wait as late as possible
#pragma unroll
      for (int ch = 0; ch < kKW / 8; ++ch) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) pv[e] = ex2a(fmaf(t[ch * 8 + e], kScale, off));
        if constexpr (!kOnesSum) {
          ls0 += (pv[0] + pv[1]) + (pv[2] + pv[3]);
          ls1 += (pv[4] + pv[5]) + (pv[6] + pv[7]);
        }
        v[ch].x = pack_bf16x2(pv[0], pv[1]); v[ch].y = pack_bf16x2(pv[2], pv[3]);
        v[ch].z = pack_bf16x2(pv[4], pv[5]); v[ch].w = pack_bf16x2(pv[6], pv[7]);
      }
      mbar_wait(bar(tp::PEMPTY + pb), ((i >> 1) & 1) ^ 1u);
      stamp();                                      // [.. + 1] P buffer free
#pragma unroll
      for (int ch = 0; ch < kKW / 8; ++ch) *reinterpret_cast<uint4*>(half + (((ch0 + ch) ^ (r & 7)) << 4)) = v[ch];
      l += ls0 + ls1;
      fence_proxy_async();          // P visible to the tensor core (generic -> async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(tp::PFULL + pb));
      stamp();                                      // [.. + 2] P handed over
    }

    // ---- output: O / l ----
    if constexpr (!kOnesSum) {
      ssum[wg * kQT + r] = l;
      row_barrier<32 * kWG>(quad);
      l = ssum[0 * kQT + r] + ssum[1 * kQT + r];
    }
    mbar_wait(bar(tp::OFULL), 0);
    stamp();                                        // O complete
    tc_fence_after();
    uint32_t x[kDW];
    tmem_ld_32x32_nowait(tO + lane_base + (uint32_t)(wg * kDW), x);
    if constexpr (kOnesSum) {
      uint32_t ls[16];
      tmem_ld_32x32_x16_nowait(tL + lane_base, ls);
      tmem_ld_wait();
      l = __uint_as_float(ls[0]);
    }
    tmem_ld_wait();
    tc_fence_before();
    if (qpos < rows) {
      const float inv = 1.0f / l;
      // one row per lane: 256-bit stores write whole 32-byte sectors
      __nv_bfloat16* dst = out + (size_t)(r0 + qpos) * H + head * kHD + wg * kDW;
#pragma unroll
      for (int ch = 0; ch < kDW / 16; ++ch) {
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          w[e] = pack_bf16x2(__uint_as_float(x[ch * 16 + 2 * e]) * inv, __uint_as_float(x[ch * 16 + 2 * e + 1]) * inv);
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"l"(dst + 16 * ch), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
      }
    }
    stamp();                                        // output stored
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  stamp();                                          // CTA done
}

// bounded wait whose time-out is recorded in mapped host memory (no ABI call in register-critical code, see common.cuh)
B2T_DEVICE void mbar_wait_rec(uint32_t bar, uint32_t parity, unsigned* rec, uint32_t code, uint32_t bars) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) b2t_trap_record(rec, code, (bar - bars) >> 3, parity);
  }
}

// Thread layout of the single-pass kernels: warpgroup 0 = TMA producer, S issuer, PV issuer and one idle warp; warpgroups 1 and 2 =
// the eight softmax warps.  The CTA is launched with 80 registers per thread (two CTAs per SM); warpgroup 0 then gives
// registers back (setmaxnreg.dec 24) and the softmax warpgroups take them (setmaxnreg.inc 104): at 80 registers the
// softmax loop spilled its loop-invariant state and re-read it (LDL / S2R: long- and short-scoreboard stalls on 20 % of
// the loop's samples in the ncu source view).
constexpr int kThreadsAttn3 = 128 + 32 * kSoftmaxWarps;
constexpr int kRegsIssue = 24, kRegsSoftmax = 104;
static_assert(128 * kRegsIssue + 32 * kSoftmaxWarps * kRegsSoftmax <= kThreadsAttn3 * 80, "register pool of the CTA");
template <int kRegs> B2T_DEVICE void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> B2T_DEVICE void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// =====================================================================================================================
// attention_tc3_kernel (attn_two_pass = 2): ONE pass over the keys, two accumulators, lazy integer rescaling.
//
// The two-pass kernel above computes Q K^T twice to know the row bound before the first exponential.  Here every S tile
// is computed once.  As in the two-pass kernel a thread owns (query row, 32-key half h of every 64-key tile) — but each
// half has its OWN running bound m_h, row sum l_h and accumulator O_h in TMEM: the P.V product of a tile is split along
// the keys, the first two k16 MMAs (keys 0..31) accumulate into O_0, the last two (keys 32..63) into O_1 — the same four
// MMAs as before, so the split is free, and no maximum is ever exchanged between threads (S 2 x 64 + O 2 x 64 = 256
// TMEM columns).  m_h is an INTEGER in the log2 domain (ceil of the largest biased score seen) and only moves when a tile
// exceeds it by more than 8: then the owning warp rescales its 32 rows of O_h in TMEM by the exact power of two
// (tcgen05.ld / st; the wait for the P buffer already guarantees that the previous P.V is complete).
// P = 2^(score - m_h) <= 2^8 is exact in its exponent, so the bf16 rounding of P does not depend on the rescaling
// schedule.  At the end  O = (O_0 2^(m_0 - M) + O_1 2^(m_1 - M)) / (l_0 2^(m_0 - M) + l_1 2^(m_1 - M)).
// Against the two-pass kernel: half the S MMAs, half the K loads, one TMEM read of S instead of two.
// =====================================================================================================================
B2T_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
B2T_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreadsAttn3, kCtasPerSm)
attention_tc3_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv,
                     const __grid_constant__ CUtensorMap map_e,
                     const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                     const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0,
                     __nv_bfloat16* __restrict__ out, int H, long long* __restrict__ dbg /* developer timeline, or null */,
                     unsigned* __restrict__ trap_rec) {
  static_assert(kKT == 64 && kSoftmaxWarps == 8 && kTmemCols == 256, "single-pass kernel: 64-key tiles, 2 x 4 softmax warps");
  constexpr uint32_t kTrapSite = 0x300u;           // trap record: 0x300 | warp (mbarrier wait: a = barrier index, b = parity), 0x380 = TMA polling loop
#ifdef B2T_ATTN_TIMELINE
  // clock64 stamps of one mid-grid CTA: softmax thread 96 -> dbg[0..255], S issuer (thread 32) -> dbg[256..383],
  // PV issuer (thread 64) -> dbg[384..511], TMA producer (thread 0) -> dbg[512..639]
  const bool tlc = dbg != nullptr && blockIdx.x == 300 && blockIdx.y == 5;
  int tln = 0;
  auto stamp = [&](int base_, int cap) { if (tlc && tln < cap) dbg[base_ + tln++] = clock64(); };
#else
  auto stamp = [&](int, int) { (void)dbg; };
#endif
#ifdef B2T_ATTN_TIMELINE
  // per-CTA log (SM id, first and last clock of the CTA) behind the 1024 timeline words: gaps between CTAs of one SM
  long long* cta_log = dbg ? dbg + 1024 + 3 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (cta_log && threadIdx.x == 0) {
    uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    cta_log[0] = smid; cta_log[1] = clock64();
  }
#endif
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + AttnSmem::kQ, sKV = base + AttnSmem::kKV, sP = base + AttnSmem::kP, sE = base + AttnSmem::kE;
  __nv_bfloat16* sR = reinterpret_cast<__nv_bfloat16*>(gbase + AttnSmem::kR);
  const uint32_t bars = base + AttnSmem::kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  auto wait = [&](uint32_t b_, uint32_t parity_) { mbar_wait_rec(b_, parity_, trap_rec, kTrapSite | (threadIdx.x >> 5), bars); };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gbase + AttnSmem::kBars + 8 * tp::COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip = qtile_clip[blockIdx.x], q0 = qtile_q0[blockIdx.x];
  const int r0 = row_off[clip], rows = row_off[clip + 1] - r0, nkeys = valid_rows[clip];
  const int nkt = (nkeys + kKT - 1) / kKT;
  const int head = blockIdx.y;

  if (threadIdx.x == 0) {
    mbar_init(bar(tp::EFULL), 1); mbar_init(bar(tp::QFULL), 1); mbar_init(bar(tp::RFULL), 1); mbar_init(bar(tp::OFULL), 1);
    for (int s = 0; s < tp::kSlots; ++s) { mbar_init(bar(tp::KVFULL + s), 1); mbar_init(bar(tp::KVEMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(tp::SFULL + b), 1); mbar_init(bar(tp::SEMPTY + b), kSoftmaxWarps);
      mbar_init(bar(tp::PFULL + b), kSoftmaxWarps); mbar_init(bar(tp::PEMPTY + b), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_qkv); tma_prefetch_desc(&map_kv); tma_prefetch_desc(&map_e); }
  if (warp == 1) tmem_alloc(bars + 8u * tp::COUNT, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // S double buffer [0,128); accumulators O_0 [128,192) and O_1 [192,256); R (80 columns) borrows [128,208) until the first PV
  const uint32_t tS = tmem_base, tO = tmem_base + 2 * kKT, tR = tmem_base + 2 * kKT;

  if (warp == 0) {
    reg_dec<kRegsIssue>();
    // ===== TMA producer =====
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(bar(tp::EFULL), 80 * 128);
      tma_load_2d(sE, &map_e, bar(tp::EFULL), 0, 0);
      mbar_expect_tx(bar(tp::QFULL), kQT * 128);
      tma_load_2d(sQ, &map_qkv, bar(tp::QFULL), head * kHD, r0 + q0);
    }
    // K(j) lives in slot j & 1, V(j) in slot 2 + (j & 1); the two streams are issued independently (non-blocking probes):
    // an in-order K0 V0 K1 V1 stream would hold the next K load back until P.V two tiles earlier has released its V slot,
    // i.e. K would be requested barely one tile ahead of its use and the load latency would sit on the critical path.
    int kn = 0, vn = 0;
    uint32_t spins = 0;
    while (kn < nkt || vn < nkt) {
      bool progress = false;
      if (kn < nkt) {
        const uint32_t st = (uint32_t)kn & 1u;
        const bool ok = mbar_test_wait(bar(tp::KVEMPTY) + 8u * st, (((uint32_t)kn >> 1) & 1u) ^ 1u);
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          if (leader) {
            mbar_expect_tx(bar(tp::KVFULL) + 8u * st, tp::kSlotBytes);
            tma_load_2d(sKV + st * tp::kSlotBytes, &map_kv, bar(tp::KVFULL) + 8u * st, H + head * kHD, r0 + kn * kKT);
          }
          ++kn; progress = true;
          if (threadIdx.x == 0) stamp(512, 128);                  // K(kn) issued
        }
      }
      if (vn < nkt) {
        const uint32_t st = 2u + ((uint32_t)vn & 1u);
        const bool ok = mbar_test_wait(bar(tp::KVEMPTY) + 8u * st, (((uint32_t)vn >> 1) & 1u) ^ 1u);
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          if (leader) {
            mbar_expect_tx(bar(tp::KVFULL) + 8u * st, tp::kSlotBytes);
            tma_load_2d(sKV + st * tp::kSlotBytes, &map_kv, bar(tp::KVFULL) + 8u * st, 2 * H + head * kHD, r0 + vn * kKT);
          }
          ++vn; progress = true;
        }
      }
      if (progress) spins = 0;
      else if (B2T_ATTN_POLL_SLEEP > 0 ? (__nanosleep(B2T_ATTN_POLL_SLEEP), ++spins > (1u << 24)) : (++spins > (1u << 28))) b2t_trap_record(trap_rec, kTrapSite | 0x80u, (unsigned)kn, (unsigned)vn);
    }
    __syncwarp();
  } else if (warp == 2) {
    reg_dec<kRegsIssue>();
    // ===== PV issuer: keys 0..31 of every tile accumulate into O_0, keys 32..63 into O_1 =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_o = make_idesc(128, kHD, 1);       // O += P V    (V MN-major)
    const uint64_t dkv = make_smem_desc(sKV), dp0 = make_smem_desc(sP);
    for (int ip = 0; ip < nkt; ++ip) {
      const uint32_t pb = (uint32_t)ip & 1u, st = 2u + ((uint32_t)ip & 1u);
      if (threadIdx.x == 64) stamp(384, 128);                     // [3 ip] before the waits
      wait(bar(tp::PFULL) + 8u * pb, ((uint32_t)ip >> 1) & 1u);       // one wait after the other (see the two-pass kernel)
      if (threadIdx.x == 64) stamp(384, 128);                     // [3 ip + 1] P ready
      wait(bar(tp::KVFULL) + 8u * st, ((uint32_t)ip >> 1) & 1u);
      if (threadIdx.x == 64) stamp(384, 128);                     // [3 ip + 2] V landed
      tc_fence_after();
      if (leader) {
        const uint64_t dv = dkv + (uint64_t)(st * (tp::kSlotBytes >> 4)), dp = dp0 + (uint64_t)(pb * (kPBuf >> 4));
#pragma unroll
        for (int kk = 0; kk < kKT / 16; ++kk)
#ifdef B2T_TC3_ONEACC
          umma_bf16(tO, dp + (uint64_t)(2 * kk), dv + (uint64_t)(kk * (16 * 128 >> 4)), idesc_o, (ip != 0 || kk != 0) ? 1u : 0u);
#else
          umma_bf16(tO + (kk >> 1) * kHD, dp + (uint64_t)(2 * kk), dv + (uint64_t)(kk * (16 * 128 >> 4)), idesc_o,
                    (ip != 0 || (kk & 1) != 0) ? 1u : 0u);
#endif
        umma_commit(bar(tp::PEMPTY) + 8u * pb);
        umma_commit(bar(tp::KVEMPTY) + 8u * st);
        if (ip == nkt - 1) umma_commit(bar(tp::OFULL));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    reg_dec<kRegsIssue>();
    // ===== S issuer =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc(128, kKT);          // S = Q K^T   (both K-major)
    constexpr uint32_t idesc_r = make_idesc(128, 80);           // R = Q E^T
    const uint64_t de = make_smem_desc(sE), dq = make_smem_desc(sQ), dkv = make_smem_desc(sKV);
    wait(bar(tp::EFULL), 0);
    wait(bar(tp::QFULL), 0);
    tc_fence_after();
    if (leader) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tR, dq + (uint64_t)(2 * k), de + (uint64_t)(2 * k), idesc_r, k != 0);
      umma_commit(bar(tp::RFULL));
    }
    for (int j = 0; j < nkt; ++j) {
      const uint32_t b = (uint32_t)j & 1u, st = (uint32_t)j & 1u;
      if (threadIdx.x == 32) stamp(256, 128);                     // [3 j] before the waits
      wait(bar(tp::KVFULL) + 8u * st, ((uint32_t)j >> 1) & 1u);
      if (threadIdx.x == 32) stamp(256, 128);                     // [3 j + 1] K landed
      wait(bar(tp::SEMPTY) + 8u * b, (((uint32_t)j >> 1) & 1u) ^ 1u);
      if (threadIdx.x == 32) stamp(256, 128);                     // [3 j + 2] S buffer free
      tc_fence_after();
      if (leader) {
        const uint64_t dk = dkv + (uint64_t)(st * (tp::kSlotBytes >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tS + b * kKT, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
        umma_commit(bar(tp::SFULL) + 8u * b);
        umma_commit(bar(tp::KVEMPTY) + 8u * st);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    reg_dec<kRegsIssue>();                          // idle member of warpgroup 0
  } else {
    // ===== softmax / output warps: thread = (query row = TMEM lane, 32-key half wg) with its own bound / sum / accumulator =====
    reg_inc<kRegsSoftmax>();
    constexpr int kWG = 2, kKW = 32;
    const int quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int r = quad * 32 + lane;
    const int qpos = q0 + r;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    constexpr float kScale = 0.125f * 1.4426950408889634f;   // log2 domain
    constexpr float kTau = 8.0f;
    float* smax = reinterpret_cast<float*>(gbase + AttnSmem::kMax);   // [wg][row]
    float* ssum = reinterpret_cast<float*>(gbase + AttnSmem::kSum);
    const __nv_bfloat16* myR = sR + r * kRS;

    // R row -> bf16 (the reference's einsum output dtype) -> shared memory; the two warps of a row split the columns
    wait(bar(tp::RFULL), 0);
    tc_fence_after();
    {
      uint32_t* rr = reinterpret_cast<uint32_t*>(sR + r * kRS);
      uint32_t a[32];
      tmem_ld_32x32_nowait(tR + lane_base + wg * 32, a);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 2) rr[(wg * 32 + i) >> 1] = pack_bf16x2(__uint_as_float(a[i]), __uint_as_float(a[i + 1]));
      if (wg == kWG - 1) {
        uint32_t c[16];
        tmem_ld_32x32_x16_nowait(tR + lane_base + 64, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 2) rr[(64 + i) >> 1] = pack_bf16x2(__uint_as_float(c[i]), __uint_as_float(c[i + 1]));
      }
    }
    tc_fence_before();
    row_barrier<32 * kWG>(quad);                    // every R row is complete (and no R read is outstanding: O may be written)
    const float rl = __bfloat162float(myR[0]) * kScale, rrt = __bfloat162float(myR[kRel - 1]) * kScale;

    // "no key seen yet" is a large negative FINITE bound: 2^(-inf + 1e30) = 0 needs no special case in the loop, and
    // 2^(kNone - m) = 0 in the final merge
    constexpr float kNone = -1.0e30f;
    float m_run = kNone, l = 0.f;
    // byte offsets of this thread's four 16-byte chunks inside a P buffer (128-byte swizzle: chunk ^ (row & 7))
    const uint32_t pofs = (uint32_t)(AttnSmem::kP + r * 128 + ((((wg * kKW) >> 3) ^ (r & 4)) << 4));
    const uint32_t plo = (uint32_t)(r & 3) << 4;
    for (int i = 0; i < nkt; ++i) {
      const int b = i & 1, k0 = i * kKT + wg * kKW;
      if (threadIdx.x == 128) stamp(0, 256);                       // [5 i] tile start
      wait(bar(tp::SFULL + b), (i >> 1) & 1);
      if (threadIdx.x == 128) stamp(0, 256);                       // [5 i + 1] S ready
      tc_fence_after();
      float t[kKW];
      {
        uint32_t x[32];
        tmem_ld_32x32_nowait(tS + lane_base + (uint32_t)(b * kKT + wg * kKW), x);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(x[e]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(tp::SEMPTY + b));            // S slice in registers: hand the buffer back at once
      const int dlo = k0 - qpos, dhi = k0 + kKW - 1 - qpos;
      const bool band = !(dhi <= -kLeft || dlo >= kRight);        // outside the diagonal band the bias is one constant per row
      const float cb = dhi <= -kLeft ? rl : rrt;
      if (band) {
#pragma unroll
        for (int e = 0; e < kKW; ++e) t[e] += __bfloat162float(myR[max(-kLeft, min(kRight, dlo + e)) + kLeft]);
      }
      if (k0 + kKW > nkeys) {
#pragma unroll
        for (int e = 0; e < kKW; ++e) if (k0 + e >= nkeys) t[e] = -INFINITY;
      }
      float m0 = fmaxf(t[0], t[1]), m1 = fmaxf(t[2], t[3]);
#pragma unroll
      for (int e = 4; e < kKW; e += 4) { m0 = fmaxf(m0, fmaxf(t[e], t[e + 1])); m1 = fmaxf(m1, fmaxf(t[e + 2], t[e + 3])); }
      const float mt = fmaxf(m0, m1);
      const float mts = band ? mt * kScale : fmaf(mt, kScale, cb);           // largest biased score of the slice, log2 domain
      const bool grow = mts > m_run + kTau;                                   // false for a fully masked slice (mts = -inf)
      const float m_new = grow ? ceilf(mts) : m_run;
      const float off = (band ? 0.f : cb) - m_new;                            // p = 2^(raw * kScale + off)
      // scale + offset and the row sums as packed f32x2 operations (half the issue slots: the loop is issue-bound)
      const float2 sc2 = make_float2(kScale, kScale), off2 = make_float2(off, off);
      float2 ls0 = make_float2(0.f, 0.f), ls1 = ls0;
      uint4 v[kKW / 8];
#pragma unroll
      for (int ch = 0; ch < kKW / 8; ++ch) {
        float2 pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 xe = ffma2(make_float2(t[ch * 8 + 2 * e], t[ch * 8 + 2 * e + 1]), sc2, off2);
#if defined(B2T_TC3_NOEXP)
          pv[e] = make_float2(xe.x * 0.001f, xe.y * 0.001f);          // timing experiment only (wrong results)
#else
          pv[e] = make_float2(ex2a(xe.x), ex2a(xe.y));
#endif
        }
        ls0 = fadd2(ls0, fadd2(pv[0], pv[1]));
        ls1 = fadd2(ls1, fadd2(pv[2], pv[3]));
        v[ch].x = pack_bf16x2(pv[0].x, pv[0].y); v[ch].y = pack_bf16x2(pv[1].x, pv[1].y);
        v[ch].z = pack_bf16x2(pv[2].x, pv[2].y); v[ch].w = pack_bf16x2(pv[3].x, pv[3].y);
      }
      const int pb = i & 1;
      if (threadIdx.x == 128) stamp(0, 256);                       // [5 i + 2] exponentials done
      wait(bar(tp::PEMPTY + pb), ((i >> 1) & 1) ^ 1u);        // P buffer free == every earlier P.V is complete
      if (threadIdx.x == 128) stamp(0, 256);                       // [5 i + 3] P buffer free
      const float m_old = m_run;
      m_run = m_new;
      uint8_t* half = gbase + pofs + pb * kPBuf;
#pragma unroll
      for (int ch = 0; ch < kKW / 8; ++ch) *reinterpret_cast<uint4*>(half + (((uint32_t)ch << 4) ^ plo)) = v[ch];
      // the bound of some row of this warp moved: rescale the warp's 32 rows of O_wg by the exact power of two (rows
      // whose bound stayed: factor 1).  Done here, after the P tile has left the registers; P.V(i) cannot start before
      // this warp's arrival below.  PEMPTY[pb] covered P.V(i-2); P.V(i-1) writes both accumulators too.
      const bool resc = grow && m_old != kNone;
      if (__any_sync(0xffffffffu, resc)) {
        wait(bar(tp::PEMPTY + (pb ^ 1)), ((i - 1) >> 1) & 1);
        tc_fence_after();
        const float f = resc ? ex2a(m_old - m_new) : 1.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t o[32];
          tmem_ld_32x32_nowait(tO + lane_base + (uint32_t)(wg * kHD + 32 * h), o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
          tmem_st_32x32(tO + lane_base + (uint32_t)(wg * kHD + 32 * h), o);
        }
        tmem_st_wait();
        tc_fence_before();
        l *= f;
      }
      l += (ls0.x + ls0.y) + (ls1.x + ls1.y);
      fence_proxy_async();          // P visible to the tensor core (generic -> async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(tp::PFULL + pb));
      if (threadIdx.x == 128) stamp(0, 256);                       // [5 i + 4] P handed over
    }
    if (threadIdx.x == 128) stamp(0, 256);

    // ---- output: merge the two accumulators
    smax[wg * kQT + r] = m_run;
    ssum[wg * kQT + r] = l;
    row_barrier<32 * kWG>(quad);
    const float ma = smax[r], mb = smax[kQT + r];
    const float mm = fmaxf(ma, mb);                                          // finite: key 0 of the clip is always valid
    const float fa = ex2a(ma - mm), fb = ex2a(mb - mm);                      // 2^(kNone - mm) = 0 for a half that never saw a key
    const float inv = 1.0f / (ssum[r] * fa + ssum[kQT + r] * fb);
    const float ga = fa * inv, gb = fb * inv;
    wait(bar(tp::OFULL), 0);
    if (threadIdx.x == 128) stamp(0, 256);
    tc_fence_after();
    __nv_bfloat16* dst = out + (size_t)(r0 + qpos) * H + head * kHD + wg * 32;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t xa[16], xb[16];
      tmem_ld_32x32_x16_nowait(tO + lane_base + (uint32_t)(wg * 32 + 16 * ch), xa);
      tmem_ld_32x32_x16_nowait(tO + lane_base + (uint32_t)(kHD + wg * 32 + 16 * ch), xb);
      tmem_ld_wait();
      uint32_t w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        w[e] = pack_bf16x2(fmaf(__uint_as_float(xb[2 * e]), gb, __uint_as_float(xa[2 * e]) * ga),
                           fmaf(__uint_as_float(xb[2 * e + 1]), gb, __uint_as_float(xa[2 * e + 1]) * ga));
      if (qpos < rows)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"l"(dst + 16 * ch), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
#ifdef B2T_ATTN_TIMELINE
  if (cta_log && threadIdx.x == 0) cta_log[2] = clock64();
#endif
}

// =====================================================================================================================
// attention_tc4_kernel (attn_two_pass = 3, default): the single-pass kernel above made PERSISTENT.
//
// Measured on attention_tc3_kernel (per-CTA clock log, tools/attn_timeline3.py): a CTA costs 1770 clocks per 64-key tile
// plus 12 000 clocks that do not depend on the clip length — launch gap (1850), barrier / TMEM set-up, the E and Q loads,
// R = Q.E^T and its conversion, pipeline fill, the drain of the last P.V, the output.  For a 10 s clip (8 key tiles) that
// fixed part is 46 % of the CTA.  Here 2 x #SM CTAs stay resident and walk over the (query tile, head) work items
// (item = blockIdx.x + k gridDim.x, head-major, heaviest clips first), and every role runs ahead across item boundaries:
//   * the TMA warp keeps three independent streams — next Q (as soon as the last S MMA of the current item has
//     retired), K ring, V ring — with running slot counters;
//   * R = Q.E^T is no longer a set-up phase with its own TMEM columns: it is issued as two PSEUDO TILES of the S ring
//     (N = 64 and N = 16 MMAs against E, which now stays in shared memory), so the S issuer computes the next item's R
//     while the softmax warps finish the current item;
//   * the softmax warps convert those pseudo tiles into the bf16 R rows between their last P hand-off and the wait for
//     the last P.V of the item — the drain latency of one item hides the set-up of the next;
//   * no extra barrier protects O: the first P hand-off of the next item is issued by every softmax warp after its
//     tcgen05.ld of O has completed.
// Barriers and TMEM are set up once per CTA.  Shared memory: Q 16 K, K/V ring 32 K, P 2 x 16 K, E 10 K, R 128 x 74 bf16.
// =====================================================================================================================
namespace t4 {
enum { EFULL = 0, QFULL, QEMPTY, OFULL, KVFULL, KVEMPTY = KVFULL + 4, SFULL = KVEMPTY + 4, SEMPTY = SFULL + 2,
       PFULL = SEMPTY + 2, PEMPTY = PFULL + 2, COUNT = PEMPTY + 2 };
constexpr int kRS = 74;                                       // R row stride (bf16): 37 words, odd -> conflict-free rows
constexpr int kQ = 0;                                         // 16 KB
constexpr int kKV = kQ + kQT * 128;                           // 4 slots x 8 KB: K in slots 0/1, V in slots 2/3
constexpr int kP = kKV + 4 * kKT * 128;                       // 2 P buffers
constexpr int kE = kP + 2 * kPBuf;                            // 80 x 128 B, resident
constexpr int kR = kE + 80 * 128;
constexpr int kML = kR + kQT * kRS * 2;                       // bound / row-sum exchange [2][kWG][128] fp32
constexpr int kBars = kML + 2 * 2 * kQT * 4;
constexpr int kTotal = kBars + 256 + 1024;
constexpr int kSlotBytes = kKT * 128;
static_assert(COUNT * 8 + 8 <= 256, "barrier area");
static_assert(kE % 1024 == 0 && kP % 1024 == 0 && kKV % 1024 == 0, "swizzled tiles are 1024-byte aligned");
static_assert(2 * (kTotal + 1024) <= 228 * 1024, "two CTAs per SM");
struct Item { int r0, q0, rows, nkeys, nkt, head; };
}  // namespace t4

B2T_DEVICE t4::Item t4_item(int it, int n_qtiles, const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                            const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0) {
  t4::Item w;
  const int x = it % n_qtiles;
  w.head = it / n_qtiles;
  const int clip = qtile_clip[x];
  w.q0 = qtile_q0[x];
  w.r0 = row_off[clip];
  w.rows = row_off[clip + 1] - w.r0;
  w.nkeys = valid_rows[clip];
  w.nkt = (w.nkeys + kKT - 1) / kKT;
  return w;
}

__global__ void __launch_bounds__(kThreadsAttn3, kCtasPerSm)
attention_tc4_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv,
                     const __grid_constant__ CUtensorMap map_e,
                     const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                     const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0,
                     __nv_bfloat16* __restrict__ out, int H, int n_qtiles, int n_items, unsigned* __restrict__ trap_rec) {
  static_assert(kKT == 64 && kSoftmaxWarps == 8 && kTmemCols == 256, "single-pass kernel: 64-key tiles, 2 x 4 softmax warps");
  constexpr uint32_t kTrapSite = 0x400u;           // trap record: 0x400 | warp (mbarrier wait: a = barrier index, b = parity), 0x480 = TMA polling loop
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + t4::kQ, sKV = base + t4::kKV, sP = base + t4::kP, sE = base + t4::kE;
  __nv_bfloat16* sR = reinterpret_cast<__nv_bfloat16*>(gbase + t4::kR);
  const uint32_t bars = base + t4::kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  auto wait = [&](uint32_t b_, uint32_t parity_) { mbar_wait_rec(b_, parity_, trap_rec, kTrapSite | (threadIdx.x >> 5), bars); };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gbase + t4::kBars + 8 * t4::COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(bar(t4::EFULL), 1); mbar_init(bar(t4::QFULL), 1); mbar_init(bar(t4::QEMPTY), 1); mbar_init(bar(t4::OFULL), 1);
    for (int s = 0; s < 4; ++s) { mbar_init(bar(t4::KVFULL + s), 1); mbar_init(bar(t4::KVEMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(t4::SFULL + b), 1); mbar_init(bar(t4::SEMPTY + b), kSoftmaxWarps);
      mbar_init(bar(t4::PFULL + b), kSoftmaxWarps); mbar_init(bar(t4::PEMPTY + b), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_qkv); tma_prefetch_desc(&map_kv); tma_prefetch_desc(&map_e); }
  if (warp == 1) tmem_alloc(bars + 8u * t4::COUNT, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tS = tmem_base, tO = tmem_base + 2 * kKT;       // S ring [0,128); accumulators O_0 [128,192), O_1 [192,256)

  if (warp == 0) {
    reg_dec<kRegsIssue>();
    // ===== TMA producer: three independent streams (next Q, K ring, V ring), non-blocking probes =====
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(bar(t4::EFULL), 80 * 128);
      tma_load_2d(sE, &map_e, bar(t4::EFULL), 0, 0);
    }
    int itq = blockIdx.x, itk = blockIdx.x, itv = blockIdx.x;
    uint32_t qn = 0, kg = 0, vg = 0;                              // running counts: Q loads, K tiles, V tiles
    int kn = 0, vn = 0;                                           // tile inside the K / V stream's current item
    t4::Item wk, wv;
    if (itk < n_items) { wk = t4_item(itk, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0); wv = wk; }
    uint32_t spins = 0;
    while (itq < n_items || itk < n_items || itv < n_items) {
      bool progress = false;
      if (itq < n_items) {
        const bool ok = mbar_test_wait(bar(t4::QEMPTY), (qn & 1u) ^ 1u);          // the previous item's last S MMA has retired
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          const t4::Item w = t4_item(itq, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0);
          if (leader) {
            mbar_expect_tx(bar(t4::QFULL), kQT * 128);
            tma_load_2d(sQ, &map_qkv, bar(t4::QFULL), w.head * kHD, w.r0 + w.q0);
          }
          ++qn; itq += G; progress = true;
        }
      }
      if (itk < n_items) {
        const uint32_t st = kg & 1u;
        const bool ok = mbar_test_wait(bar(t4::KVEMPTY) + 8u * st, ((kg >> 1) & 1u) ^ 1u);
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          if (leader) {
            mbar_expect_tx(bar(t4::KVFULL) + 8u * st, t4::kSlotBytes);
            tma_load_2d(sKV + st * t4::kSlotBytes, &map_kv, bar(t4::KVFULL) + 8u * st, H + wk.head * kHD, wk.r0 + kn * kKT);
          }
          ++kg; progress = true;
          if (++kn == wk.nkt) {
            kn = 0; itk += G;
            if (itk < n_items) wk = t4_item(itk, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0);
          }
        }
      }
      if (itv < n_items) {
        const uint32_t st = 2u + (vg & 1u);
        const bool ok = mbar_test_wait(bar(t4::KVEMPTY) + 8u * st, ((vg >> 1) & 1u) ^ 1u);
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          if (leader) {
            mbar_expect_tx(bar(t4::KVFULL) + 8u * st, t4::kSlotBytes);
            tma_load_2d(sKV + st * t4::kSlotBytes, &map_kv, bar(t4::KVFULL) + 8u * st, 2 * H + wv.head * kHD, wv.r0 + vn * kKT);
          }
          ++vg; progress = true;
          if (++vn == wv.nkt) {
            vn = 0; itv += G;
            if (itv < n_items) wv = t4_item(itv, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0);
          }
        }
      }
      if (progress) spins = 0;
      else if (B2T_ATTN_POLL_SLEEP > 0 ? (__nanosleep(B2T_ATTN_POLL_SLEEP), ++spins > (1u << 24)) : (++spins > (1u << 28))) b2t_trap_record(trap_rec, kTrapSite | 0x80u, kg, vg);
    }
    __syncwarp();
  } else if (warp == 1) {
    reg_dec<kRegsIssue>();
    // ===== S issuer: per item two pseudo tiles (R = Q.E^T, columns 0..63 and 64..79), then S = Q.K^T per key tile =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc(128, kKT);          // S = Q K^T, R[:, 0:64] = Q E[0:64]^T   (both K-major)
    constexpr uint32_t idesc_r = make_idesc(128, 16);           // R[:, 64:80]
    const uint64_t de = make_smem_desc(sE), dq = make_smem_desc(sQ), dkv = make_smem_desc(sKV);
    wait(bar(t4::EFULL), 0);
    uint32_t sg = 0, kg = 0, n = 0;
    for (int it = blockIdx.x; it < n_items; it += G, ++n) {
      const int x = it % n_qtiles;
      const int nkt = (valid_rows[qtile_clip[x]] + kKT - 1) / kKT;
      wait(bar(t4::QFULL), n & 1u);
      for (int ps = 0; ps < 2; ++ps, ++sg) {
        const uint32_t b = sg & 1u;
        wait(bar(t4::SEMPTY) + 8u * b, ((sg >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (leader) {
          const uint64_t dep = de + (uint64_t)(ps * (64 * 128 >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tS + b * kKT, dq + (uint64_t)(2 * k), dep + (uint64_t)(2 * k), ps == 0 ? idesc_s : idesc_r, k != 0);
          umma_commit(bar(t4::SFULL) + 8u * b);
        }
      }
      for (int j = 0; j < nkt; ++j, ++sg, ++kg) {
        const uint32_t b = sg & 1u, st = kg & 1u;
        wait(bar(t4::KVFULL) + 8u * st, (kg >> 1) & 1u);          // one wait after the other (see the two-pass kernel)
        wait(bar(t4::SEMPTY) + 8u * b, ((sg >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (leader) {
          const uint64_t dk = dkv + (uint64_t)(st * (t4::kSlotBytes >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS + b * kKT, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
          umma_commit(bar(t4::SFULL) + 8u * b);
          umma_commit(bar(t4::KVEMPTY) + 8u * st);
          if (j == nkt - 1) umma_commit(bar(t4::QEMPTY));                // Q may be overwritten by the next item's tile
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    reg_dec<kRegsIssue>();
    // ===== PV issuer: keys 0..31 of every tile accumulate into O_0, keys 32..63 into O_1 =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_o = make_idesc(128, kHD, 1);       // O += P V    (V MN-major)
    const uint64_t dkv = make_smem_desc(sKV), dp0 = make_smem_desc(sP);
    uint32_t pg = 0;
    for (int it = blockIdx.x; it < n_items; it += G) {
      const int x = it % n_qtiles;
      const int nkt = (valid_rows[qtile_clip[x]] + kKT - 1) / kKT;
      for (int ip = 0; ip < nkt; ++ip, ++pg) {
        const uint32_t pb = pg & 1u, st = 2u + (pg & 1u);
        // PFULL of an item's first tile also says: every softmax warp has finished reading the previous item's O
        wait(bar(t4::PFULL) + 8u * pb, (pg >> 1) & 1u);
        wait(bar(t4::KVFULL) + 8u * st, (pg >> 1) & 1u);
        tc_fence_after();
        if (leader) {
          const uint64_t dv = dkv + (uint64_t)(st * (t4::kSlotBytes >> 4)), dp = dp0 + (uint64_t)(pb * (kPBuf >> 4));
#pragma unroll
          for (int kk = 0; kk < kKT / 16; ++kk)
            umma_bf16(tO + (kk >> 1) * kHD, dp + (uint64_t)(2 * kk), dv + (uint64_t)(kk * (16 * 128 >> 4)), idesc_o,
                      (ip != 0 || (kk & 1) != 0) ? 1u : 0u);
          umma_commit(bar(t4::PEMPTY) + 8u * pb);
          umma_commit(bar(t4::KVEMPTY) + 8u * st);
          if (ip == nkt - 1) umma_commit(bar(t4::OFULL));
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    reg_dec<kRegsIssue>();                          // idle member of warpgroup 0
  } else {
    // ===== softmax / output warps: thread = (query row = TMEM lane, 32-key half wg) with its own bound / sum / accumulator =====
    reg_inc<kRegsSoftmax>();
    constexpr int kWG = 2, kKW = 32;
    const int quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    constexpr float kScale = 0.125f * 1.4426950408889634f;   // log2 domain
    constexpr float kTau = 8.0f;
    constexpr float kNone = -1.0e30f;                        // "no key seen yet": finite, so 2^(-inf + 1e30) = 0 without a special case
    float* sml = reinterpret_cast<float*>(gbase + t4::kML);  // [m | l][wg][row]
    const __nv_bfloat16* myR = sR + r * t4::kRS;
    const uint32_t pofs = (uint32_t)(t4::kP + r * 128 + ((((wg * kKW) >> 3) ^ (r & 4)) << 4));
    const uint32_t plo = (uint32_t)(r & 3) << 4;
    uint32_t sg = 0, pg = 0, n = 0;

    // the two pseudo tiles of an item -> this thread's part of the bf16 R row (the reference's einsum output dtype)
    auto take_r = [&]() {
      uint32_t* rr = reinterpret_cast<uint32_t*>(sR + r * t4::kRS);
      {
        const uint32_t b = sg & 1u;
        wait(bar(t4::SFULL) + 8u * b, (sg >> 1) & 1u);
        tc_fence_after();
        uint32_t a[32];
        tmem_ld_32x32_nowait(tS + lane_base + b * kKT + wg * 32, a);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(t4::SEMPTY) + 8u * b);
#pragma unroll
        for (int i = 0; i < 32; i += 2) rr[(wg * 32 + i) >> 1] = pack_bf16x2(__uint_as_float(a[i]), __uint_as_float(a[i + 1]));
        ++sg;
      }
      {
        const uint32_t b = sg & 1u;
        wait(bar(t4::SFULL) + 8u * b, (sg >> 1) & 1u);
        tc_fence_after();
        if (wg == kWG - 1) {
          uint32_t c[16];
          tmem_ld_32x32_x16_nowait(tS + lane_base + b * kKT, c);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 10; i += 2) rr[(64 + i) >> 1] = pack_bf16x2(__uint_as_float(c[i]), __uint_as_float(c[i + 1]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(t4::SEMPTY) + 8u * b);
        ++sg;
      }
    };

    int it = blockIdx.x;
    t4::Item w;
    if (it < n_items) {
      w = t4_item(it, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0);
      take_r();
      row_barrier<32 * kWG>(quad);
    }
    while (it < n_items) {
      const int qpos = w.q0 + r, nkeys = w.nkeys;
      const float rl = __bfloat162float(myR[0]) * kScale, rrt = __bfloat162float(myR[kRel - 1]) * kScale;
      float m_run = kNone, l = 0.f;
      for (int i = 0; i < w.nkt; ++i, ++sg, ++pg) {
        const uint32_t b = sg & 1u;
        const int k0 = i * kKT + wg * kKW;
        wait(bar(t4::SFULL) + 8u * b, (sg >> 1) & 1u);
        tc_fence_after();
        float t[kKW];
        {
          uint32_t x[32];
          tmem_ld_32x32_nowait(tS + lane_base + b * kKT + wg * kKW, x);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(x[e]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(t4::SEMPTY) + 8u * b);       // S slice in registers: hand the buffer back at once
        const int dlo = k0 - qpos, dhi = k0 + kKW - 1 - qpos;
        const bool band = !(dhi <= -kLeft || dlo >= kRight);        // outside the diagonal band the bias is one constant per row
        const float cb = dhi <= -kLeft ? rl : rrt;
        if (band) {
#pragma unroll
          for (int e = 0; e < kKW; ++e) t[e] += __bfloat162float(myR[max(-kLeft, min(kRight, dlo + e)) + kLeft]);
        }
        if (k0 + kKW > nkeys) {
#pragma unroll
          for (int e = 0; e < kKW; ++e) if (k0 + e >= nkeys) t[e] = -INFINITY;
        }
        float m0 = fmaxf(t[0], t[1]), m1 = fmaxf(t[2], t[3]);
#pragma unroll
        for (int e = 4; e < kKW; e += 4) { m0 = fmaxf(m0, fmaxf(t[e], t[e + 1])); m1 = fmaxf(m1, fmaxf(t[e + 2], t[e + 3])); }
        const float mt = fmaxf(m0, m1);
        const float mts = band ? mt * kScale : fmaf(mt, kScale, cb);           // largest biased score of the slice, log2 domain
        const bool grow = mts > m_run + kTau;                                   // false for a fully masked slice (mts = -inf)
        const float m_new = grow ? ceilf(mts) : m_run;
        const float off = (band ? 0.f : cb) - m_new;                            // p = 2^(raw * kScale + off)
        const float2 sc2 = make_float2(kScale, kScale), off2 = make_float2(off, off);
        float2 ls0 = make_float2(0.f, 0.f), ls1 = ls0;
        uint4 v[kKW / 8];
#pragma unroll
        for (int ch = 0; ch < kKW / 8; ++ch) {
          float2 pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 xe = ffma2(make_float2(t[ch * 8 + 2 * e], t[ch * 8 + 2 * e + 1]), sc2, off2);
            pv[e] = make_float2(ex2a(xe.x), ex2a(xe.y));
          }
          ls0 = fadd2(ls0, fadd2(pv[0], pv[1]));
          ls1 = fadd2(ls1, fadd2(pv[2], pv[3]));
          v[ch].x = pack_bf16x2(pv[0].x, pv[0].y); v[ch].y = pack_bf16x2(pv[1].x, pv[1].y);
          v[ch].z = pack_bf16x2(pv[2].x, pv[2].y); v[ch].w = pack_bf16x2(pv[3].x, pv[3].y);
        }
        const uint32_t pb = pg & 1u;
        wait(bar(t4::PEMPTY) + 8u * pb, ((pg >> 1) & 1u) ^ 1u);   // P buffer free == P.V two tiles back (and all before) complete
        const float m_old = m_run;
        m_run = m_new;
        uint8_t* half = gbase + pofs + pb * kPBuf;
#pragma unroll
        for (int ch = 0; ch < kKW / 8; ++ch) *reinterpret_cast<uint4*>(half + (((uint32_t)ch << 4) ^ plo)) = v[ch];
        // the bound of some row of this warp moved: rescale the warp's 32 rows of O_wg by the exact power of two (rows
        // whose bound stayed: factor 1).  P.V(i) cannot start before this warp's arrival below; P.V(i-1) must have retired.
        const bool resc = grow && m_old != kNone;                    // never on the first tile of an item
        if (__any_sync(0xffffffffu, resc)) {
          wait(bar(t4::PEMPTY) + 8u * (pb ^ 1u), ((pg - 1) >> 1) & 1u);
          tc_fence_after();
          const float f = resc ? ex2a(m_old - m_new) : 1.0f;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t o[32];
            tmem_ld_32x32_nowait(tO + lane_base + (uint32_t)(wg * kHD + 32 * h), o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tmem_st_32x32(tO + lane_base + (uint32_t)(wg * kHD + 32 * h), o);
          }
          tmem_st_wait();
          tc_fence_before();
          l *= f;
        }
        l += (ls0.x + ls0.y) + (ls1.x + ls1.y);
        fence_proxy_async();          // P visible to the tensor core (generic -> async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(t4::PFULL) + 8u * pb);
      }

      // ---- item end.  Publish (bound, sum); once both warps of the row quad are past their last R read, turn the NEXT
      //      item's pseudo tiles into R rows (this is what the drain of the last P.V hides); then merge and store O.
      sml[wg * kQT + r] = m_run;
      sml[2 * kQT + wg * kQT + r] = l;
      const int it_next = it + G;
      const t4::Item cur = w;
      if (it_next < n_items) w = t4_item(it_next, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0);
      row_barrier<32 * kWG>(quad);
      const float ma = sml[r], mb = sml[kQT + r];
      const float la = sml[2 * kQT + r], lb = sml[3 * kQT + r];
      if (it_next < n_items) take_r();
      const float mm = fmaxf(ma, mb);                                          // > kNone: key 0 of the clip is always valid
      const float fa = ex2a(ma - mm), fb = ex2a(mb - mm);                      // 2^(kNone - mm) = 0 for a half that never saw a key
      const float inv = 1.0f / (la * fa + lb * fb);
      const float ga = fa * inv, gb = fb * inv;
      wait(bar(t4::OFULL), n & 1u);
      tc_fence_after();
      __nv_bfloat16* dst = out + (size_t)(cur.r0 + qpos) * H + cur.head * kHD + wg * 32;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t xa[16], xb[16];
        tmem_ld_32x32_x16_nowait(tO + lane_base + (uint32_t)(wg * 32 + 16 * ch), xa);
        tmem_ld_32x32_x16_nowait(tO + lane_base + (uint32_t)(kHD + wg * 32 + 16 * ch), xb);
        tmem_ld_wait();
        uint32_t wv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          wv[e] = pack_bf16x2(fmaf(__uint_as_float(xb[2 * e]), gb, __uint_as_float(xa[2 * e]) * ga),
                              fmaf(__uint_as_float(xb[2 * e + 1]), gb, __uint_as_float(xa[2 * e + 1]) * ga));
        if (qpos < cur.rows)
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"l"(dst + 16 * ch), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]), "r"(wv[4]), "r"(wv[5]), "r"(wv[6]), "r"(wv[7])
                       : "memory");
      }
      tc_fence_before();               // the O reads are complete before this warp's next P hand-off lets P.V overwrite O
      row_barrier<32 * kWG>(quad);     // the next item's R rows are complete; the (bound, sum) slots may be rewritten
      it = it_next; ++n;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =====================================================================================================================
// attention_tc5_kernel (attn_two_pass = 4): the single-pass kernel with P IN TENSOR MEMORY.
//
// ncu on attention_tc3_kernel: the softmax warps wait for a free P buffer on 45 % of the tiles — the two-slot V ring can
// only request V(i) once P.V(i-2) has retired, which leaves one tile period for a TMA round trip — and a tile moves
// 80 KB through shared memory (operand reads 48 KB, P write 16 KB, TMA 16 KB).  Here the bf16 P tile never touches
// shared memory: a softmax thread writes its 32 probabilities as 16 packed words with tcgen05.st INTO THE S BUFFER it
// has just read (its own 32 columns: no other thread's unread scores are touched) and P.V is issued with the A operand
// in tensor memory (tcgen05.mma [d], [a_tmem], b_desc: keys 0..31 at columns 0..15, keys 32..63 at columns 32..47 of
// the buffer; one k16 MMA reads 8 columns).  The S buffer is handed back by the COMMIT of P.V(i) instead of by the
// softmax warps, so S(i+2) is issued one P.V later than before — still a whole softmax period ahead of its use.
// The 32 KB of shared memory this frees turn the K and V rings into 4 + 4 slots (requests run three tiles ahead).
// No generic->async proxy fence, no P-buffer barrier, no 16-byte swizzled stores in the loop.
// =====================================================================================================================
namespace t5 {
enum { EFULL = 0, QFULL, RFULL, OFULL, KFULL, KEMPTY = KFULL + 4, VFULL = KEMPTY + 4, VEMPTY = VFULL + 4, SFULL = VEMPTY + 4,
       PFULL = SFULL + 2, PVDONE = PFULL + 2, COUNT = PVDONE + 2 };
constexpr int kRing = 4;
constexpr int kRS = 74;                                       // R row stride (bf16): 37 words, odd -> conflict-free rows
constexpr int kQ = 0;                                         // 16 KB
constexpr int kK = kQ + kQT * 128;                            // 4 slots x 8 KB
constexpr int kV = kK + kRing * kKT * 128;                    // 4 slots x 8 KB
constexpr int kE = kV + kRing * kKT * 128;                    // 80 x 128 B
constexpr int kR = kE + 80 * 128;
constexpr int kML = kR + kQT * kRS * 2;                       // bound / row-sum exchange [2][kWG][128] fp32
constexpr int kBars = kML + 2 * 2 * kQT * 4;
constexpr int kTotal = kBars + 256 + 1024;
constexpr int kSlotBytes = kKT * 128;
static_assert(COUNT * 8 + 8 <= 256, "barrier area");
static_assert(kE % 1024 == 0 && kK % 1024 == 0 && kV % 1024 == 0, "swizzled tiles are 1024-byte aligned");
static_assert(2 * (kTotal + 1024) <= 228 * 1024, "two CTAs per SM");
}  // namespace t5

// D[tmem] (+)= A[tmem] . B[smem]: the A operand (128 rows = lanes, two bf16 per 32-bit column) is read from tensor memory
B2T_DEVICE void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
B2T_DEVICE void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <int kPoly>
__global__ void __launch_bounds__(kThreadsAttn3, kCtasPerSm)
attention_tc5_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv,
                     const __grid_constant__ CUtensorMap map_e,
                     const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                     const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0,
                     __nv_bfloat16* __restrict__ out, int H, unsigned* __restrict__ trap_rec) {
  static_assert(kKT == 64 && kSoftmaxWarps == 8 && kTmemCols == 256, "single-pass kernel: 64-key tiles, 2 x 4 softmax warps");
  constexpr uint32_t kTrapSite = 0x500u;           // trap record: 0x500 | warp (mbarrier wait: a = barrier index, b = parity), 0x580 = TMA polling loop
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + t5::kQ, sK = base + t5::kK, sV = base + t5::kV, sE = base + t5::kE;
  __nv_bfloat16* sR = reinterpret_cast<__nv_bfloat16*>(gbase + t5::kR);
  const uint32_t bars = base + t5::kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  auto wait = [&](uint32_t b_, uint32_t parity_) { mbar_wait_rec(b_, parity_, trap_rec, kTrapSite | (threadIdx.x >> 5), bars); };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gbase + t5::kBars + 8 * t5::COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip = qtile_clip[blockIdx.x], q0 = qtile_q0[blockIdx.x];
  const int r0 = row_off[clip], rows = row_off[clip + 1] - r0, nkeys = valid_rows[clip];
  const int nkt = (nkeys + kKT - 1) / kKT;
  const int head = blockIdx.y;

  if (threadIdx.x == 0) {
    mbar_init(bar(t5::EFULL), 1); mbar_init(bar(t5::QFULL), 1); mbar_init(bar(t5::RFULL), 1); mbar_init(bar(t5::OFULL), 1);
    for (int s = 0; s < t5::kRing; ++s) {
      mbar_init(bar(t5::KFULL + s), 1); mbar_init(bar(t5::KEMPTY + s), 1);
      mbar_init(bar(t5::VFULL + s), 1); mbar_init(bar(t5::VEMPTY + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(t5::SFULL + b), 1); mbar_init(bar(t5::PFULL + b), kSoftmaxWarps); mbar_init(bar(t5::PVDONE + b), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_qkv); tma_prefetch_desc(&map_kv); tma_prefetch_desc(&map_e); }
  if (warp == 1) tmem_alloc(bars + 8u * t5::COUNT, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // S / P double buffer [0,128); accumulators O_0 [128,192) and O_1 [192,256); R (80 columns) borrows [128,208) until the first PV
  const uint32_t tS = tmem_base, tO = tmem_base + 2 * kKT, tR = tmem_base + 2 * kKT;

  if (warp == 0) {
    reg_dec<kRegsIssue>();
    // ===== TMA producer: E, Q, then the K ring and the V ring as two independent streams (non-blocking probes) =====
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(bar(t5::EFULL), 80 * 128);
      tma_load_2d(sE, &map_e, bar(t5::EFULL), 0, 0);
      mbar_expect_tx(bar(t5::QFULL), kQT * 128);
      tma_load_2d(sQ, &map_qkv, bar(t5::QFULL), head * kHD, r0 + q0);
    }
    int kn = 0, vn = 0;
    uint32_t spins = 0;
    while (kn < nkt || vn < nkt) {
      bool progress = false;
      if (kn < nkt) {
        const uint32_t st = (uint32_t)kn & 3u;
        const bool ok = mbar_test_wait(bar(t5::KEMPTY) + 8u * st, (((uint32_t)kn >> 2) & 1u) ^ 1u);
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          if (leader) {
            mbar_expect_tx(bar(t5::KFULL) + 8u * st, t5::kSlotBytes);
            tma_load_2d(sK + st * t5::kSlotBytes, &map_kv, bar(t5::KFULL) + 8u * st, H + head * kHD, r0 + kn * kKT);
          }
          ++kn; progress = true;
        }
      }
      if (vn < nkt) {
        const uint32_t st = (uint32_t)vn & 3u;
        const bool ok = mbar_test_wait(bar(t5::VEMPTY) + 8u * st, (((uint32_t)vn >> 2) & 1u) ^ 1u);
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          if (leader) {
            mbar_expect_tx(bar(t5::VFULL) + 8u * st, t5::kSlotBytes);
            tma_load_2d(sV + st * t5::kSlotBytes, &map_kv, bar(t5::VFULL) + 8u * st, 2 * H + head * kHD, r0 + vn * kKT);
          }
          ++vn; progress = true;
        }
      }
      if (progress) spins = 0;
      else if (++spins > (1u << 28)) b2t_trap_record(trap_rec, kTrapSite | 0x80u, (unsigned)kn, (unsigned)vn);
    }
    __syncwarp();
  } else if (warp == 2) {
    reg_dec<kRegsIssue>();
    // ===== PV issuer: A = P from tensor memory; keys 0..31 of every tile accumulate into O_0, keys 32..63 into O_1 =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_o = make_idesc(128, kHD, 1);       // O += P V    (P K-major in TMEM, V MN-major in shared memory)
    const uint64_t dv0 = make_smem_desc(sV);
    for (int ip = 0; ip < nkt; ++ip) {
      const uint32_t pb = (uint32_t)ip & 1u, st = (uint32_t)ip & 3u;
      wait(bar(t5::PFULL) + 8u * pb, ((uint32_t)ip >> 1) & 1u);          // one wait after the other (see the two-pass kernel)
      wait(bar(t5::VFULL) + 8u * st, ((uint32_t)ip >> 2) & 1u);
      tc_fence_after();
      if (leader) {
        const uint64_t dv = dv0 + (uint64_t)(st * (t5::kSlotBytes >> 4));
#pragma unroll
        for (int kk = 0; kk < kKT / 16; ++kk)
          umma_bf16_ts(tO + (kk >> 1) * kHD, tS + pb * kKT + (kk >> 1) * 32 + (kk & 1) * 8, dv + (uint64_t)(kk * (16 * 128 >> 4)), idesc_o,
                       (ip != 0 || (kk & 1) != 0) ? 1u : 0u);
        umma_commit(bar(t5::PVDONE) + 8u * pb);                           // the S / P buffer is free again
        umma_commit(bar(t5::VEMPTY) + 8u * st);
        if (ip == nkt - 1) umma_commit(bar(t5::OFULL));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    reg_dec<kRegsIssue>();
    // ===== S issuer =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc(128, kKT);          // S = Q K^T   (both K-major)
    constexpr uint32_t idesc_r = make_idesc(128, 80);           // R = Q E^T
    const uint64_t de = make_smem_desc(sE), dq = make_smem_desc(sQ), dk0 = make_smem_desc(sK);
    wait(bar(t5::EFULL), 0);
    wait(bar(t5::QFULL), 0);
    tc_fence_after();
    if (leader) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tR, dq + (uint64_t)(2 * k), de + (uint64_t)(2 * k), idesc_r, k != 0);
      umma_commit(bar(t5::RFULL));
    }
    for (int j = 0; j < nkt; ++j) {
      const uint32_t b = (uint32_t)j & 1u, st = (uint32_t)j & 3u;
      wait(bar(t5::KFULL) + 8u * st, ((uint32_t)j >> 2) & 1u);
      wait(bar(t5::PVDONE) + 8u * b, (((uint32_t)j >> 1) & 1u) ^ 1u);    // P.V(j-2) has consumed the P tile that lives in this buffer
      tc_fence_after();
      if (leader) {
        const uint64_t dk = dk0 + (uint64_t)(st * (t5::kSlotBytes >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tS + b * kKT, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
        umma_commit(bar(t5::SFULL) + 8u * b);
        umma_commit(bar(t5::KEMPTY) + 8u * st);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    reg_dec<kRegsIssue>();                          // idle member of warpgroup 0
  } else {
    // ===== softmax / output warps: thread = (query row = TMEM lane, 32-key half wg) with its own bound / sum / accumulator =====
    reg_inc<kRegsSoftmax>();
    constexpr int kWG = 2, kKW = 32;
    const int quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int r = quad * 32 + lane;
    const int qpos = q0 + r;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    constexpr float kScale = 0.125f * 1.4426950408889634f;   // log2 domain
    constexpr float kTau = 8.0f;
    constexpr float kNone = -1.0e30f;                        // "no key seen yet": finite, so 2^(-inf + 1e30) = 0 without a special case
    float* sml = reinterpret_cast<float*>(gbase + t5::kML);  // [m | l][wg][row]
    const __nv_bfloat16* myR = sR + r * t5::kRS;

    // R row -> bf16 (the reference's einsum output dtype) -> shared memory; the two warps of a row split the columns
    wait(bar(t5::RFULL), 0);
    tc_fence_after();
    {
      uint32_t* rr = reinterpret_cast<uint32_t*>(sR + r * t5::kRS);
      uint32_t a[32];
      tmem_ld_32x32_nowait(tR + lane_base + wg * 32, a);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 2) rr[(wg * 32 + i) >> 1] = pack_bf16x2(__uint_as_float(a[i]), __uint_as_float(a[i + 1]));
      if (wg == kWG - 1) {
        uint32_t c[16];
        tmem_ld_32x32_x16_nowait(tR + lane_base + 64, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 10; i += 2) rr[(64 + i) >> 1] = pack_bf16x2(__uint_as_float(c[i]), __uint_as_float(c[i + 1]));
      }
    }
    tc_fence_before();
    row_barrier<32 * kWG>(quad);                    // every R row is complete (and no R read is outstanding: O may be written)
    const float rl = __bfloat162float(myR[0]) * kScale, rrt = __bfloat162float(myR[kRel - 1]) * kScale;

    float m_run = kNone, l = 0.f;
    const uint32_t tSP = tS + lane_base + (uint32_t)(wg * kKW);          // this thread's S slice; its P words go to the first 16 columns of it
    for (int i = 0; i < nkt; ++i) {
      const uint32_t b = (uint32_t)i & 1u;
      const int k0 = i * kKT + wg * kKW;
      wait(bar(t5::SFULL) + 8u * b, ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      float t[kKW];
      {
        uint32_t x[32];
        tmem_ld_32x32_nowait(tSP + b * kKT, x);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(x[e]);
      }
      const int dlo = k0 - qpos, dhi = k0 + kKW - 1 - qpos;
      const bool band = !(dhi <= -kLeft || dlo >= kRight);        // outside the diagonal band the bias is one constant per row
      const float cb = dhi <= -kLeft ? rl : rrt;
      if (band) {
#pragma unroll
        for (int e = 0; e < kKW; ++e) t[e] += __bfloat162float(myR[max(-kLeft, min(kRight, dlo + e)) + kLeft]);
      }
      if (k0 + kKW > nkeys) {
#pragma unroll
        for (int e = 0; e < kKW; ++e) if (k0 + e >= nkeys) t[e] = -INFINITY;
      }
      float m0 = fmaxf(t[0], t[1]), m1 = fmaxf(t[2], t[3]);
#pragma unroll
      for (int e = 4; e < kKW; e += 4) { m0 = fmaxf(m0, fmaxf(t[e], t[e + 1])); m1 = fmaxf(m1, fmaxf(t[e + 2], t[e + 3])); }
      const float mt = fmaxf(m0, m1);
      const float mts = band ? mt * kScale : fmaf(mt, kScale, cb);           // largest biased score of the slice, log2 domain
      const bool grow = mts > m_run + kTau;                                   // false for a fully masked slice (mts = -inf)
      const float m_new = grow ? ceilf(mts) : m_run;
      const float off = (band ? 0.f : cb) - m_new;                            // p = 2^(raw * kScale + off)
      const float2 sc2 = make_float2(kScale, kScale), off2 = make_float2(off, off);
      float2 ls0 = make_float2(0.f, 0.f), ls1 = ls0;
      uint32_t v[kKW / 2];
#pragma unroll
      for (int ch = 0; ch < kKW / 8; ++ch) {
        float2 pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 xe = ffma2(make_float2(t[ch * 8 + 2 * e], t[ch * 8 + 2 * e + 1]), sc2, off2);
          // kPoly of the 4 pairs of a chunk take the polynomial (FMA pipe), the others MUFU.EX2: the XU pipe is the
          // busiest unit of the loop (60 % of peak over the whole kernel, saturated inside this block)
          pv[e] = e < kPoly ? ex2_poly2(xe) : make_float2(ex2a(xe.x), ex2a(xe.y));
        }
        ls0 = fadd2(ls0, fadd2(pv[0], pv[1]));
        ls1 = fadd2(ls1, fadd2(pv[2], pv[3]));
#pragma unroll
        for (int e = 0; e < 4; ++e) v[ch * 4 + e] = pack_bf16x2(pv[e].x, pv[e].y);
      }
      tmem_st_32x32_x16(tSP + b * kKT, v);                          // P over the (already read) first half of this thread's S slice
      const float m_old = m_run;
      m_run = m_new;
      // the bound of some row of this warp moved: rescale the warp's 32 rows of O_wg by the exact power of two (rows
      // whose bound stayed: factor 1).  P.V(i) cannot start before this warp's arrival below; P.V(i-1) must have retired
      // (S(i) being complete only says that P.V(i-2) has).
      const bool resc = grow && m_old != kNone;                     // never on the first tile
      if (__any_sync(0xffffffffu, resc)) {
        wait(bar(t5::PVDONE) + 8u * (b ^ 1u), ((uint32_t)(i - 1) >> 1) & 1u);
        tc_fence_after();
        const float f = resc ? ex2a(m_old - m_new) : 1.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t o[32];
          tmem_ld_32x32_nowait(tO + lane_base + (uint32_t)(wg * kHD + 32 * h), o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
          tmem_st_32x32(tO + lane_base + (uint32_t)(wg * kHD + 32 * h), o);
        }
        l *= f;
      }
      l += (ls0.x + ls0.y) + (ls1.x + ls1.y);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(t5::PFULL) + 8u * b);
    }

    // ---- output: merge the two accumulators
    sml[wg * kQT + r] = m_run;
    sml[2 * kQT + wg * kQT + r] = l;
    row_barrier<32 * kWG>(quad);
    const float ma = sml[r], mb = sml[kQT + r];
    const float mm = fmaxf(ma, mb);                                          // > kNone: key 0 of the clip is always valid
    const float fa = ex2a(ma - mm), fb = ex2a(mb - mm);                      // 2^(kNone - mm) = 0 for a half that never saw a key
    const float inv = 1.0f / (sml[2 * kQT + r] * fa + sml[3 * kQT + r] * fb);
    const float ga = fa * inv, gb = fb * inv;
    wait(bar(t5::OFULL), 0);
    tc_fence_after();
    __nv_bfloat16* dst = out + (size_t)(r0 + qpos) * H + head * kHD + wg * 32;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t xa[16], xb[16];
      tmem_ld_32x32_x16_nowait(tO + lane_base + (uint32_t)(wg * 32 + 16 * ch), xa);
      tmem_ld_32x32_x16_nowait(tO + lane_base + (uint32_t)(kHD + wg * 32 + 16 * ch), xb);
      tmem_ld_wait();
      uint32_t w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        w[e] = pack_bf16x2(fmaf(__uint_as_float(xb[2 * e]), gb, __uint_as_float(xa[2 * e]) * ga),
                           fmaf(__uint_as_float(xb[2 * e + 1]), gb, __uint_as_float(xa[2 * e + 1]) * ga));
      if (qpos < rows)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"l"(dst + 16 * ch), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =====================================================================================================================
// attention_tc6_kernel (attn_two_pass = 6): persistent over (query tile, head) items AND P in tensor memory.
//
// attention_tc5_kernel still pays ~10 000 clocks per CTA that do not depend on the clip length (10 s clip: 8 key tiles x
// 1 360 clocks of loop against 10 400 of launch gap, parameter loads, barrier / TMEM set-up, E and Q loads, R, pipeline
// fill and drain).  Here 2 x #SM CTAs walk over the items (item = blockIdx.x + k gridDim.x, head-major, heaviest clips
// first) exactly as attention_tc4_kernel does — R as two pseudo tiles of the S ring, next Q requested when the last S MMA
// of an item retires, K / V rings running ahead across item boundaries, the next item's R rows converted while the last
// P.V of the current one drains — with the P-in-TMEM data path of attention_tc5_kernel.  An S buffer is released by the
// P.V commit when it held a key tile and by the eight softmax warps (RDONE) when it held a pseudo tile; every role
// prefetches the next item's parameters one item ahead, and the TMA warp's K stream hands them to its V and Q streams
// through a small shared-memory FIFO.
// =====================================================================================================================
namespace t6 {
enum { EFULL = 0, QFULL, QEMPTY, OFULL, KFULL, KEMPTY = KFULL + 4, VFULL = KEMPTY + 4, VEMPTY = VFULL + 4, SFULL = VEMPTY + 4,
       PFULL = SFULL + 2, PVDONE = PFULL + 2, RDONE = PVDONE + 2, COUNT = RDONE + 2 };
constexpr int kFifo = t5::kBars + 256;                        // 8 x int4 item descriptors (TMA warp only)
constexpr int kTotal = kFifo + 128 + 1024;
constexpr int kRegsIssue = 40, kRegsSoftmax = 96;             // the three-stream TMA warp does not fit 24 registers
static_assert(COUNT * 8 + 8 <= 256, "barrier area");
static_assert(2 * (kTotal + 1024) <= 228 * 1024, "two CTAs per SM");
static_assert(128 * kRegsIssue + 32 * kSoftmaxWarps * kRegsSoftmax <= kThreadsAttn3 * 80, "register pool of the CTA");
}  // namespace t6

template <int kPoly>
__global__ void __launch_bounds__(kThreadsAttn3, kCtasPerSm)
attention_tc6_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv,
                     const __grid_constant__ CUtensorMap map_e,
                     const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                     const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0,
                     __nv_bfloat16* __restrict__ out, int H, int n_qtiles, int n_items, unsigned* __restrict__ trap_rec
#ifdef B2T_ATTN_TIMELINE
                     , long long* __restrict__ dbg
#endif
                     ) {
#ifdef B2T_ATTN_TIMELINE
  // clock64 stamps of softmax thread 128 of CTA 37, items 3..6: 16 stamps per item (see the STAMP sites)
  int tln = 0;
#define STAMP() do { if (dbg != nullptr && blockIdx.x == 37 && threadIdx.x == 128 && n >= 3 && n < 7 && tln < 1000) dbg[tln++] = clock64(); } while (0)
#else
#define STAMP() do { } while (0)
#endif
  static_assert(kKT == 64 && kSoftmaxWarps == 8 && kTmemCols == 256, "single-pass kernel: 64-key tiles, 2 x 4 softmax warps");
  constexpr uint32_t kTrapSite = 0x600u;           // trap record: 0x600 | warp (mbarrier wait: a = barrier index, b = parity), 0x680 = TMA polling loop
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + t5::kQ, sK = base + t5::kK, sV = base + t5::kV, sE = base + t5::kE;
  __nv_bfloat16* sR = reinterpret_cast<__nv_bfloat16*>(gbase + t5::kR);
  const uint32_t bars = base + t5::kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  auto wait = [&](uint32_t b_, uint32_t parity_) { mbar_wait_rec(b_, parity_, trap_rec, kTrapSite | (threadIdx.x >> 5), bars); };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gbase + t5::kBars + 8 * t6::COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int n_my = (n_items - (int)blockIdx.x + G - 1) / G;     // items of this CTA (>= 1: the grid never exceeds n_items)

  if (threadIdx.x == 0) {
    mbar_init(bar(t6::EFULL), 1); mbar_init(bar(t6::QFULL), 1); mbar_init(bar(t6::QEMPTY), 1); mbar_init(bar(t6::OFULL), 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(bar(t6::KFULL + s), 1); mbar_init(bar(t6::KEMPTY + s), 1);
      mbar_init(bar(t6::VFULL + s), 1); mbar_init(bar(t6::VEMPTY + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(t6::SFULL + b), 1); mbar_init(bar(t6::PFULL + b), kSoftmaxWarps);
      mbar_init(bar(t6::PVDONE + b), 1); mbar_init(bar(t6::RDONE + b), kSoftmaxWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_qkv); tma_prefetch_desc(&map_kv); tma_prefetch_desc(&map_e); }
  if (warp == 1) tmem_alloc(bars + 8u * t6::COUNT, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tS = tmem_base, tO = tmem_base + 2 * kKT;       // S / P ring [0,128); accumulators O_0 [128,192), O_1 [192,256)

  if (warp == 0) {
    reg_dec<t6::kRegsIssue>();
    // ===== TMA producer: next Q, K ring, V ring as three independent streams over the item sequence (non-blocking probes).
    //       The K stream leads; it loads every item's parameters (prefetched one item ahead) and leaves {r0, q0, nkt, head}
    //       in an 8-entry FIFO for the V and Q streams. =====
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(bar(t6::EFULL), 80 * 128);
      tma_load_2d(sE, &map_e, bar(t6::EFULL), 0, 0);
    }
    int4* fifo = reinterpret_cast<int4*>(gbase + t6::kFifo);
    int ok_ = 0, ov_ = 0, oq_ = 0;                                // item ordinals of the K / V / Q streams
    uint32_t qn = 0, kg = 0, vg = 0;                              // running counts: Q loads, K tiles, V tiles
    int kn = 0, vn = 0;                                           // tile inside the K / V stream's current item
    t4::Item wk = t4_item(blockIdx.x, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0), wkn = wk;
    if (1 < n_my) wkn = t4_item(blockIdx.x + G, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0);
    if (leader) fifo[0] = make_int4(wk.r0, wk.q0, wk.nkt, wk.head);
    __syncwarp();
    int vr0 = wk.r0, vnkt = wk.nkt, vhead = wk.head;
    uint32_t spins = 0;
    while (oq_ < n_my || ok_ < n_my || ov_ < n_my) {
      bool progress = false;
      const int pushed = min(ok_, n_my - 1);                      // highest ordinal whose descriptor is in the FIFO
      if (oq_ < n_my && oq_ <= pushed) {
        const bool ok = mbar_test_wait(bar(t6::QEMPTY), (qn & 1u) ^ 1u);          // the previous item's last S MMA has retired
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          const int4 e = fifo[oq_ & 7];
          if (leader) {
            mbar_expect_tx(bar(t6::QFULL), kQT * 128);
            tma_load_2d(sQ, &map_qkv, bar(t6::QFULL), e.w * kHD, e.x + e.y);
          }
          ++qn; ++oq_; progress = true;
        }
      }
      if (ok_ < n_my) {
        const uint32_t st = kg & 3u;
        const bool ok = mbar_test_wait(bar(t6::KEMPTY) + 8u * st, ((kg >> 2) & 1u) ^ 1u);
        // moving on to the next item overwrites FIFO entry (ok_ + 1) & 7: both followers must be past ordinal ok_ - 7
        const bool room = kn + 1 < wk.nkt || ok_ + 1 - min(oq_, ov_) < 8;
        if (__shfl_sync(0xffffffffu, (int)(ok && room), 0)) {
          if (leader) {
            mbar_expect_tx(bar(t6::KFULL) + 8u * st, t5::kSlotBytes);
            tma_load_2d(sK + st * t5::kSlotBytes, &map_kv, bar(t6::KFULL) + 8u * st, H + wk.head * kHD, wk.r0 + kn * kKT);
          }
          ++kg; progress = true;
          if (++kn == wk.nkt) {
            kn = 0; ++ok_;
            if (ok_ < n_my) {
              wk = wkn;
              if (leader) fifo[ok_ & 7] = make_int4(wk.r0, wk.q0, wk.nkt, wk.head);
              __syncwarp();
              if (ok_ + 1 < n_my) wkn = t4_item(blockIdx.x + (ok_ + 1) * G, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0);
            }
          }
        }
      }
      if (ov_ < n_my && ov_ <= pushed) {
        const uint32_t st = vg & 3u;
        const bool ok = mbar_test_wait(bar(t6::VEMPTY) + 8u * st, ((vg >> 2) & 1u) ^ 1u);
        if (__shfl_sync(0xffffffffu, (int)ok, 0)) {
          if (leader) {
            mbar_expect_tx(bar(t6::VFULL) + 8u * st, t5::kSlotBytes);
            tma_load_2d(sV + st * t5::kSlotBytes, &map_kv, bar(t6::VFULL) + 8u * st, 2 * H + vhead * kHD, vr0 + vn * kKT);
          }
          ++vg; progress = true;
          if (++vn == vnkt) {
            vn = 0; ++ov_;
            if (ov_ < n_my && ov_ <= min(ok_, n_my - 1)) { const int4 e = fifo[ov_ & 7]; vr0 = e.x; vnkt = e.z; vhead = e.w; }
            else vnkt = 0;                                        // descriptor not pushed yet: fetched below before the next V tile
          }
        }
      }
      if (ov_ < n_my && vnkt == 0 && ov_ <= min(ok_, n_my - 1)) { const int4 e = fifo[ov_ & 7]; vr0 = e.x; vnkt = e.z; vhead = e.w; progress = true; }
      if (progress) spins = 0;
      else if (++spins > (1u << 28)) b2t_trap_record(trap_rec, kTrapSite | 0x80u, kg, vg);
    }
    __syncwarp();
  } else if (warp == 1) {
    reg_dec<t6::kRegsIssue>();
    // ===== S issuer: per item two pseudo tiles (R = Q.E^T, columns 0..63 and 64..79), then S = Q.K^T per key tile =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc(128, kKT);          // S = Q K^T, R[:, 0:64] = Q E[0:64]^T   (both K-major)
    constexpr uint32_t idesc_r = make_idesc(128, 16);           // R[:, 64:80]
    const uint64_t de = make_smem_desc(sE), dq = make_smem_desc(sQ), dk0 = make_smem_desc(sK);
    wait(bar(t6::EFULL), 0);
    uint32_t sg = 0, kg = 0, n = 0;
    uint32_t cp0 = 0, cp1 = 0, cr0 = 0, cr1 = 0, prev0 = 0, prev1 = 0;   // per S buffer: key tiles / pseudo tiles issued, kind of the last occupant
    auto release_wait = [&](uint32_t b) {                         // the previous occupant of buffer b has been consumed
      const uint32_t prev = b ? prev1 : prev0;
      if (prev == 1u) wait(bar(t6::PVDONE) + 8u * b, ((b ? cp1 : cp0) - 1u) & 1u);
      else if (prev == 2u) wait(bar(t6::RDONE) + 8u * b, ((b ? cr1 : cr0) - 1u) & 1u);
    };
    int nkt_next = (valid_rows[qtile_clip[(int)blockIdx.x % n_qtiles]] + kKT - 1) / kKT;
    for (int it = blockIdx.x; it < n_items; it += G, ++n) {
      const int nkt = nkt_next;
      if (it + G < n_items) nkt_next = (valid_rows[qtile_clip[(it + G) % n_qtiles]] + kKT - 1) / kKT;
      wait(bar(t6::QFULL), n & 1u);
      for (int ps = 0; ps < 2; ++ps, ++sg) {
        const uint32_t b = sg & 1u;
        release_wait(b);
        tc_fence_after();
        if (leader) {
          const uint64_t dep = de + (uint64_t)(ps * (64 * 128 >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tS + b * kKT, dq + (uint64_t)(2 * k), dep + (uint64_t)(2 * k), ps == 0 ? idesc_s : idesc_r, k != 0);
          umma_commit(bar(t6::SFULL) + 8u * b);
        }
        if (b) { ++cr1; prev1 = 2u; } else { ++cr0; prev0 = 2u; }
      }
      for (int j = 0; j < nkt; ++j, ++sg, ++kg) {
        const uint32_t b = sg & 1u, st = kg & 3u;
        wait(bar(t6::KFULL) + 8u * st, (kg >> 2) & 1u);           // one wait after the other (see the two-pass kernel)
        release_wait(b);
        tc_fence_after();
        if (leader) {
          const uint64_t dk = dk0 + (uint64_t)(st * (t5::kSlotBytes >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS + b * kKT, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
          umma_commit(bar(t6::SFULL) + 8u * b);
          umma_commit(bar(t6::KEMPTY) + 8u * st);
          if (j == nkt - 1) umma_commit(bar(t6::QEMPTY));          // Q may be overwritten by the next item's tile
        }
        if (b) { ++cp1; prev1 = 1u; } else { ++cp0; prev0 = 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    reg_dec<t6::kRegsIssue>();
    // ===== PV issuer: A = P from tensor memory; keys 0..31 of every tile accumulate into O_0, keys 32..63 into O_1 =====
    const bool leader = elect_one();
    constexpr uint32_t idesc_o = make_idesc(128, kHD, 1);       // O += P V    (P K-major in TMEM, V MN-major in shared memory)
    const uint64_t dv0 = make_smem_desc(sV);
    uint32_t sg = 0, vg = 0, cp0 = 0, cp1 = 0;
    int nkt_next = (valid_rows[qtile_clip[(int)blockIdx.x % n_qtiles]] + kKT - 1) / kKT;
    for (int it = blockIdx.x; it < n_items; it += G) {
      const int nkt = nkt_next;
      if (it + G < n_items) nkt_next = (valid_rows[qtile_clip[(it + G) % n_qtiles]] + kKT - 1) / kKT;
      sg += 2;                                                    // the item's two pseudo tiles
      for (int ip = 0; ip < nkt; ++ip, ++sg, ++vg) {
        const uint32_t b = sg & 1u, st = vg & 3u;
        // PFULL of an item's first tile also says: every softmax warp has finished reading the previous item's O
        wait(bar(t6::PFULL) + 8u * b, (b ? cp1 : cp0) & 1u);
        wait(bar(t6::VFULL) + 8u * st, (vg >> 2) & 1u);
        tc_fence_after();
        if (leader) {
          const uint64_t dv = dv0 + (uint64_t)(st * (t5::kSlotBytes >> 4));
#pragma unroll
          for (int kk = 0; kk < kKT / 16; ++kk)
            umma_bf16_ts(tO + (kk >> 1) * kHD, tS + b * kKT + (kk >> 1) * 32 + (kk & 1) * 8, dv + (uint64_t)(kk * (16 * 128 >> 4)), idesc_o,
                         (ip != 0 || (kk & 1) != 0) ? 1u : 0u);
          umma_commit(bar(t6::PVDONE) + 8u * b);
          umma_commit(bar(t6::VEMPTY) + 8u * st);
          if (ip == nkt - 1) umma_commit(bar(t6::OFULL));
        }
        if (b) ++cp1; else ++cp0;
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    reg_dec<t6::kRegsIssue>();                      // idle member of warpgroup 0
  } else {
    // ===== softmax / output warps: thread = (query row = TMEM lane, 32-key half wg) with its own bound / sum / accumulator =====
    reg_inc<t6::kRegsSoftmax>();
    constexpr int kWG = 2, kKW = 32;
    const int quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    constexpr float kScale = 0.125f * 1.4426950408889634f;   // log2 domain
    constexpr float kTau = 8.0f;
    constexpr float kNone = -1.0e30f;                        // "no key seen yet": finite, so 2^(-inf + 1e30) = 0 without a special case
    float* sml = reinterpret_cast<float*>(gbase + t5::kML);  // [m | l][wg][row]
    const __nv_bfloat16* myR = sR + r * t5::kRS;
    const uint32_t tSP = tS + lane_base + (uint32_t)(wg * kKW);          // this thread's S slice; its P words go to the first 16 columns of it
    uint32_t sg = 0, n = 0, cp0 = 0, cp1 = 0;

    // the two pseudo tiles of an item -> this thread's part of the bf16 R row (the reference's einsum output dtype)
    auto take_r = [&]() {
      uint32_t* rr = reinterpret_cast<uint32_t*>(sR + r * t5::kRS);
      {
        const uint32_t b = sg & 1u;
        wait(bar(t6::SFULL) + 8u * b, (sg >> 1) & 1u);
        tc_fence_after();
        uint32_t a[32];
        tmem_ld_32x32_nowait(tSP + b * kKT, a);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(t6::RDONE) + 8u * b);
#pragma unroll
        for (int i = 0; i < 32; i += 2) rr[(wg * 32 + i) >> 1] = pack_bf16x2(__uint_as_float(a[i]), __uint_as_float(a[i + 1]));
        ++sg;
      }
      {
        const uint32_t b = sg & 1u;
        wait(bar(t6::SFULL) + 8u * b, (sg >> 1) & 1u);
        tc_fence_after();
        if (wg == kWG - 1) {
          uint32_t c[16];
          tmem_ld_32x32_x16_nowait(tS + lane_base + b * kKT, c);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 10; i += 2) rr[(64 + i) >> 1] = pack_bf16x2(__uint_as_float(c[i]), __uint_as_float(c[i + 1]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(t6::RDONE) + 8u * b);
        ++sg;
      }
    };

    int it = blockIdx.x;
    t4::Item w = t4_item(it, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0);
    take_r();
    row_barrier<32 * kWG>(quad);
    while (it < n_items) {
      const int it_next = it + G;
      t4::Item wn = w;                                            // next item's parameters: in flight during this item
      if (it_next < n_items) wn = t4_item(it_next, n_qtiles, row_off, valid_rows, qtile_clip, qtile_q0);
      const int qpos = w.q0 + r, nkeys = w.nkeys;
      const float rl = __bfloat162float(myR[0]) * kScale, rrt = __bfloat162float(myR[kRel - 1]) * kScale;
      float m_run = kNone, l = 0.f;
      STAMP();                                                    // [0] item start
      for (int i = 0; i < w.nkt; ++i, ++sg) {
        const uint32_t b = sg & 1u;
        const int k0 = i * kKT + wg * kKW;
        wait(bar(t6::SFULL) + 8u * b, (sg >> 1) & 1u);
        if (i < 3) STAMP();                                       // [1..3] S ready, tiles 0..2
        tc_fence_after();
        float t[kKW];
        {
          uint32_t x[32];
          tmem_ld_32x32_nowait(tSP + b * kKT, x);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = __uint_as_float(x[e]);
        }
        const int dlo = k0 - qpos, dhi = k0 + kKW - 1 - qpos;
        const bool band = !(dhi <= -kLeft || dlo >= kRight);        // outside the diagonal band the bias is one constant per row
        const float cb = dhi <= -kLeft ? rl : rrt;
        if (band) {
#pragma unroll
          for (int e = 0; e < kKW; ++e) t[e] += __bfloat162float(myR[max(-kLeft, min(kRight, dlo + e)) + kLeft]);
        }
        if (k0 + kKW > nkeys) {
#pragma unroll
          for (int e = 0; e < kKW; ++e) if (k0 + e >= nkeys) t[e] = -INFINITY;
        }
        float m0 = fmaxf(t[0], t[1]), m1 = fmaxf(t[2], t[3]);
#pragma unroll
        for (int e = 4; e < kKW; e += 4) { m0 = fmaxf(m0, fmaxf(t[e], t[e + 1])); m1 = fmaxf(m1, fmaxf(t[e + 2], t[e + 3])); }
        const float mt = fmaxf(m0, m1);
        const float mts = band ? mt * kScale : fmaf(mt, kScale, cb);           // largest biased score of the slice, log2 domain
        const bool grow = mts > m_run + kTau;                                   // false for a fully masked slice (mts = -inf)
        const float m_new = grow ? ceilf(mts) : m_run;
        const float off = (band ? 0.f : cb) - m_new;                            // p = 2^(raw * kScale + off)
        const float2 sc2 = make_float2(kScale, kScale), off2 = make_float2(off, off);
        float2 ls0 = make_float2(0.f, 0.f), ls1 = ls0;
        uint32_t v[kKW / 2];
#pragma unroll
        for (int ch = 0; ch < kKW / 8; ++ch) {
          float2 pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 xe = ffma2(make_float2(t[ch * 8 + 2 * e], t[ch * 8 + 2 * e + 1]), sc2, off2);
            pv[e] = e < kPoly ? ex2_poly2(xe) : make_float2(ex2a(xe.x), ex2a(xe.y));
          }
          ls0 = fadd2(ls0, fadd2(pv[0], pv[1]));
          ls1 = fadd2(ls1, fadd2(pv[2], pv[3]));
#pragma unroll
          for (int e = 0; e < 4; ++e) v[ch * 4 + e] = pack_bf16x2(pv[e].x, pv[e].y);
        }
        tmem_st_32x32_x16(tSP + b * kKT, v);                          // P over the (already read) first half of this thread's S slice
        const float m_old = m_run;
        m_run = m_new;
        const bool resc = grow && m_old != kNone;                     // never on the first tile of an item
        if (__any_sync(0xffffffffu, resc)) {
          // P.V(i-1) (the other buffer's latest key tile) must have retired before O is rescaled
          wait(bar(t6::PVDONE) + 8u * (b ^ 1u), ((b ? cp0 : cp1) - 1u) & 1u);
          tc_fence_after();
          const float f = resc ? ex2a(m_old - m_new) : 1.0f;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t o[32];
            tmem_ld_32x32_nowait(tO + lane_base + (uint32_t)(wg * kHD + 32 * h), o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tmem_st_32x32(tO + lane_base + (uint32_t)(wg * kHD + 32 * h), o);
          }
          l *= f;
        }
        l += (ls0.x + ls0.y) + (ls1.x + ls1.y);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(t6::PFULL) + 8u * b);
        if (b) ++cp1; else ++cp0;
      }

      // ---- item end.  Publish (bound, sum); once both warps of the row quad are past their last R read, turn the NEXT
      //      item's pseudo tiles into R rows (this is what the drain of the last P.V hides); then merge and store O.
      STAMP();                                                    // [4] loop end
      sml[wg * kQT + r] = m_run;
      sml[2 * kQT + wg * kQT + r] = l;
      row_barrier<32 * kWG>(quad);
      STAMP();                                                    // [5] both warps of the quad done
      const float ma = sml[r], mb = sml[kQT + r];
      const float la = sml[2 * kQT + r], lb = sml[3 * kQT + r];
      if (it_next < n_items) take_r();
      STAMP();                                                    // [6] next item's R rows taken
      const float mm = fmaxf(ma, mb);                                          // > kNone: key 0 of the clip is always valid
      const float fa = ex2a(ma - mm), fb = ex2a(mb - mm);                      // 2^(kNone - mm) = 0 for a half that never saw a key
      const float inv = 1.0f / (la * fa + lb * fb);
      const float ga = fa * inv, gb = fb * inv;
      wait(bar(t6::OFULL), n & 1u);
      STAMP();                                                    // [7] last P.V retired
      tc_fence_after();
      __nv_bfloat16* dst = out + (size_t)(w.r0 + qpos) * H + w.head * kHD + wg * 32;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t xa[16], xb[16];
        tmem_ld_32x32_x16_nowait(tO + lane_base + (uint32_t)(wg * 32 + 16 * ch), xa);
        tmem_ld_32x32_x16_nowait(tO + lane_base + (uint32_t)(kHD + wg * 32 + 16 * ch), xb);
        tmem_ld_wait();
        uint32_t wv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          wv[e] = pack_bf16x2(fmaf(__uint_as_float(xb[2 * e]), gb, __uint_as_float(xa[2 * e]) * ga),
                              fmaf(__uint_as_float(xb[2 * e + 1]), gb, __uint_as_float(xa[2 * e + 1]) * ga));
        if (qpos < w.rows)
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"l"(dst + 16 * ch), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]), "r"(wv[4]), "r"(wv[5]), "r"(wv[6]), "r"(wv[7])
                       : "memory");
      }
      STAMP();                                                    // [8] O stored
      tc_fence_before();               // the O reads are complete before this warp's next P hand-off lets P.V overwrite O
      row_barrier<32 * kWG>(quad);     // the next item's R rows are complete; the (bound, sum) slots may be rewritten
      STAMP();                                                    // [9] item done
      it = it_next; ++n; w = wn;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}
#endif  // !B2T_ATTN_WIDE

}  // namespace

int g_attn_heads_per_cta = 1;   // kept for the option plumbing; the kernel needs 1 (E and R borrow per-head buffers)
long long* g_attn_dbg = nullptr;   // b2t_attention_set_dbg(device buffer of 128 int64): developer timeline
extern "C" void b2t_attention_set_dbg(long long* p) { g_attn_dbg = p; }
int g_attn_poly_exp = 1;        // b2t_set_option("attn_poly_exp", 0/1): one pair in four of the 2^x of the P-in-TMEM kernels on the FMA pipe
int g_attn_ctas = 0;            // b2t_set_option("attn_ctas", n): cap the persistent kernel's grid (tests: many items per CTA)
int g_attn_two_pass = 4;        // b2t_set_option("attn_two_pass", v): 4 = single pass with P in tensor memory (default), 3 = persistent single pass, 2 = single pass, 1 = two-pass fixed bound, 0 = online softmax, 5 = two-pass + tensor-core row sums

// host entry used by b2t_relkey_attention (attention.cu)
int b2t_attention_tensor_tc(const void* qkv, const void* dist_emb, const b2t_batch* b, void* out, int heads, cudaStream_t st) {
  const int H = heads * kHD;
  B2T_REQUIRE(b->n_qtiles128 > 0 && b->qtile128_clip && b->qtile128_q0, B2T_ERR_ARG,
              "b2t_relkey_attention(tcgen05): the batch has no 128-row query tiles");
  CUtensorMap mq, mk, me;
  int rc = make_map(&mq, qkv, b->total_rows, 3 * H, 3 * H, kQT);
  if (rc != B2T_OK) return rc;
  rc = make_map(&mk, qkv, b->total_rows, 3 * H, 3 * H, kKT);
  if (rc != B2T_OK) return rc;
  rc = make_map(&me, dist_emb, kRel, kHD, kHD, 80);
  if (rc != B2T_OK) return rc;
  B2T_SMEM_OPT_IN(AttnSmem::kTotal, attention_tc_kernel);
  const int hpc = 1;
  dim3 grid(b->n_qtiles128, heads / hpc);
#if !B2T_ATTN_WIDE
  if (g_attn_two_pass) {
    // g_attn_two_pass bits: 1 = two-pass, 4 = row sums on the tensor core as well
    B2T_SMEM_OPT_IN(AttnSmem::kTotal, attention_tc2_kernel<false>);
    B2T_SMEM_OPT_IN(AttnSmem::kTotal, attention_tc2_kernel<true>);
    if (g_attn_two_pass == 2) {          // single pass, split accumulators
      B2T_SMEM_OPT_IN(AttnSmem::kTotal, attention_tc3_kernel);
      attention_tc3_kernel<<<grid, kThreadsAttn3, AttnSmem::kTotal, st>>>(mq, mk, me, b->row_off, b->valid_rows, b->qtile128_clip,
                                                                         b->qtile128_q0, (__nv_bfloat16*)out, H, g_attn_dbg, b2t_trap_rec());
      B2T_LAUNCH_CHECK();
      return B2T_OK;
    }
    if (g_attn_two_pass == 4) {          // single pass, P in tensor memory
      auto kern5 = g_attn_poly_exp ? attention_tc5_kernel<1> : attention_tc5_kernel<0>;
      B2T_SMEM_OPT_IN(t5::kTotal, attention_tc5_kernel<0>);
      B2T_SMEM_OPT_IN(t5::kTotal, attention_tc5_kernel<1>);
      kern5<<<grid, kThreadsAttn3, t5::kTotal, st>>>(mq, mk, me, b->row_off, b->valid_rows, b->qtile128_clip, b->qtile128_q0,
                                                    (__nv_bfloat16*)out, H, b2t_trap_rec());
      B2T_LAUNCH_CHECK();
      return B2T_OK;
    }
    if (g_attn_two_pass == 6) {          // single pass, P in tensor memory, persistent over (query tile, head) items
      auto kern6 = g_attn_poly_exp ? attention_tc6_kernel<1> : attention_tc6_kernel<0>;
      B2T_SMEM_OPT_IN(t6::kTotal, attention_tc6_kernel<0>);
      B2T_SMEM_OPT_IN(t6::kTotal, attention_tc6_kernel<1>);
      const int n_items = b->n_qtiles128 * heads;
      int ctas = std::min(n_items, kCtasPerSm * b2t_num_sms());
      if (g_attn_ctas > 0) ctas = std::min(ctas, g_attn_ctas);
      kern6<<<ctas, kThreadsAttn3, t6::kTotal, st>>>(mq, mk, me, b->row_off, b->valid_rows, b->qtile128_clip, b->qtile128_q0,
                                                    (__nv_bfloat16*)out, H, b->n_qtiles128, n_items, b2t_trap_rec()
#ifdef B2T_ATTN_TIMELINE
                                                    , g_attn_dbg
#endif
                                                    );
      B2T_LAUNCH_CHECK();
      return B2T_OK;
    }
    if (g_attn_two_pass == 3) {          // single pass, persistent over (query tile, head) items
      B2T_SMEM_OPT_IN(t4::kTotal, attention_tc4_kernel);
      const int n_items = b->n_qtiles128 * heads;
      int ctas = std::min(n_items, kCtasPerSm * b2t_num_sms());
      if (g_attn_ctas > 0) ctas = std::min(ctas, g_attn_ctas);
      attention_tc4_kernel<<<ctas, kThreadsAttn3, t4::kTotal, st>>>(mq, mk, me, b->row_off, b->valid_rows, b->qtile128_clip,
                                                                   b->qtile128_q0, (__nv_bfloat16*)out, H, b->n_qtiles128, n_items, b2t_trap_rec());
      B2T_LAUNCH_CHECK();
      return B2T_OK;
    }
    auto kern = (g_attn_two_pass & 4) ? attention_tc2_kernel<true> : attention_tc2_kernel<false>;
    kern<<<grid, kThreadsAttn2, AttnSmem::kTotal, st>>>(mq, mk, me, b->row_off, b->valid_rows, b->qtile128_clip, b->qtile128_q0,
                                                       (__nv_bfloat16*)out, g_attn_dbg, H);
    B2T_LAUNCH_CHECK();
    return B2T_OK;
  }
#endif
  B2T_REQUIRE(heads == kHeads, B2T_ERR_ARG, "b2t_attention: the online-softmax tcgen05 kernel (attn_two_pass = 0) is 16 heads only");
  attention_tc_kernel<<<grid, kThreadsAttn, AttnSmem::kTotal, st>>>(mq, mk, me, b->row_off, b->valid_rows, b->qtile128_clip,
                                                                   b->qtile128_q0, (__nv_bfloat16*)out, hpc);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
