// Conformer convolution-module middle on the tensor cores (bf16 production path of b2t_dwconv_ln_swish):
//   causal depthwise conv1d (k = 31, per clip) -> LayerNorm(1024) -> swish        (HF modeling_wav2vec2_bert.py:213-221)
//
// Why: the CUDA-core tap loop of dwconv.cu is bound by register-file bandwidth, not by HBM — every FMA reads a weight and an
// accumulator that differ from the previous instruction's (only the input is served by the operand reuse cache) (measured on B200, tools/micro/fma_rate.cu and
// tools/dwconv_ab.py: scalar FFMA 2.2 clk, FFMA2 3.8 clk per warp instruction and SM sub-partition; the tap loop alone is
// 108 us of the 162 us launch, against an HBM floor of 41 us).  A depthwise conv has no shared operand to build a GEMM
// from — except along time: for ONE channel pair (2i, 2i+1) and one 64-row item starting at clip row T0
//
//   D[m][n] = sum_k A[m][k] * B[k][n]        m = 0..15,  n = 2 n' + p' (n' = 0..3 time phase, p' = channel parity),
//                                            k = 2 j + p (j = 0..35 window row, p = channel parity)
//   A[m][(j, p)]        = x[T0 - 30 + 4 m + j][2 i + p]            (a Hankel matrix of the input)
//   B[(j, p)][(n', p')] = (p == p') * w[2 i + p][j - n']           (block-diagonal Toeplitz matrix of the 31 taps)
//   D[m][(n', p')]      = y[T0 + 4 m + n'][2 i + p']
//
// maps onto mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with NO data rearrangement at all:
//   * the A fragment register of lane (g, q) for k-step ks is (row g, k = 2q, 2q+1) = both channels of the pair at window
//     row lane + 8 ks: one 32-bit word of the channels-last activation tile exactly as TMA delivers it.  One 128-bit
//     shared-memory load per lane feeds four channel pairs; with the 128-byte TMA swizzle the 32 lanes (32 consecutive
//     rows, same column) are conflict-free;
//   * the D fragment of lane (g, q) is (y[T0 + lane][2i], y[T0 + lane][2i+1]) and the same for row T0 + 32 + lane: the
//     packed bf16x2 OUTPUT word.  A thread owns two rows of the item for every channel pair its warp processes, so the
//     LayerNorm sums are plain per-thread accumulations (no shuffles) and the normalised row is stored straight from
//     registers with 256-bit stores (32 contiguous bytes per lane = one full sector);
//   * the B fragment is one weight per register: a word of the table w2[pair][tap] = (w[2i][tap], w[2i+1][tap]) masked to
//     the lane's channel parity; the table (with zero margins for the taps outside 0..30) lives in shared memory with the
//     four pairs of a warp's group interleaved, so one 128-bit load fetches a tap for all four MMAs of a step.
// K = 72 (36 window rows: 4 k16 steps + 1 k8 step); 61 % of the MMA work multiplies zeros, which is free: the whole launch
// needs 2.4 M warp MMAs.  Products of bf16 values are exact in fp32; the accumulation order differs from the CUDA-core
// kernels', i.e. a conv value can differ in its last fp32 bit before it is rounded to bf16.
//
// CTA = 512 threads = 4 quads of 4 warps, persistent over a contiguous run of 64-row items.  A chunk = 64 channels x 96
// window rows (12 KB, one TMA box); quad Q takes chunks Q, Q+4, Q+8, Q+12 of an item, each warp 8 channel pairs of a chunk.
// Every quad runs its own double-buffered TMA pipeline (full mbarrier per buffer, a 128-thread named barrier to release).
#include "tc_ptx.cuh"

namespace {

constexpr int kC = 1024, kTaps = 31, kItem = 64, kWin = 96, kThreads = 512;
constexpr int kNBQ = 2;                              // chunk buffers per quad
constexpr int kBufBytes = kWin * 128;                // 64 channels x 96 rows
constexpr int kTabE = 41;                            // table entries per pair: e = tap + 3, taps -3..36 (+1: odd pitch)
constexpr int kOffTab = 4 * kNBQ * kBufBytes;
constexpr int kOffLn = kOffTab + 512 * kTabE * 4;
constexpr int kOffPart = kOffLn + 512 * 16;
constexpr int kOffBar = kOffPart + 16 * kItem * 8;
constexpr int kSmemBytes = kOffBar + 4 * kNBQ * 8 + 1024;   // + slack for the 1024-byte alignment of the buffers

B2T_DEVICE void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
B2T_DEVICE void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
B2T_DEVICE void lds128(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
B2T_DEVICE uint32_t lds32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
B2T_DEVICE uint32_t pack_bf2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
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
B2T_DEVICE uint64_t add2p(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
B2T_DEVICE uint64_t mul2p(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <bool kK8>
__global__ void __launch_bounds__(kThreads, 1)
dwconv_mma_kernel(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ w_dw, const float* __restrict__ ln_w,
                  const float* __restrict__ ln_b, const int32_t* __restrict__ row_off, const int32_t* __restrict__ ctile_clip,
                  const int32_t* __restrict__ ctile_t0, int n_items, __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint32_t* wtab = reinterpret_cast<uint32_t*>(sm + kOffTab);
  float4* lnp = reinterpret_cast<float4*>(sm + kOffLn);
  float2* spart = reinterpret_cast<float2*>(sm + kOffPart);
  const uint32_t wtab_s = base + kOffTab, lnp_s = base + kOffLn, bar_s = base + kOffBar;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Q = warp >> 2, wq = warp & 3;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t bmask = (g & 1) ? 0xffff0000u : 0x0000ffffu;        // channel parity of this lane's B column
  const int ebase = q - (g >> 1) + 3;                                // table entry of tap q - n'

  // ---- one-time: weight table (bf16 pairs, zero margins), LayerNorm parameters (halved: the tail computes y / 2) ----
  for (int idx = tid; idx < 512 * kTabE; idx += kThreads) {
    const int e = idx >> 9, pair = idx & 511, tap = e - 3;
    uint32_t v = 0u;
    if (tap >= 0 && tap < kTaps) {
      const float2 w = __ldg(reinterpret_cast<const float2*>(w_dw + (size_t)tap * kC + 2 * pair));
      v = pack_bf2(w.x, w.y);
    }
    wtab[((pair >> 2) * kTabE + e) * 4 + (pair & 3)] = v;          // [group of 4 pairs][entry][pair in group]
  }
  lnp[tid] = make_float4(0.5f * __ldg(ln_w + 2 * tid), 0.5f * __ldg(ln_w + 2 * tid + 1), 0.5f * __ldg(ln_b + 2 * tid),
                         0.5f * __ldg(ln_b + 2 * tid + 1));
  if (tid == 0) {
    for (int i = 0; i < 4 * kNBQ; ++i) mbar_init(bar_s + 8 * i, 1);
    fence_barrier_init();
    tma_prefetch_desc(&xmap);
  }
  __syncthreads();

  const int per = (n_items + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(n_items, i0 + per);
  const int n_total = (i1 - i0) * 4;                                 // chunks this quad will process
  const bool leader = (wq == 0 && lane == 0);
  const uint32_t qbuf = base + (uint32_t)(Q * kNBQ) * kBufBytes, qbar = bar_s + 8u * (Q * kNBQ);

  auto issue = [&](int n, int first_row) {                           // chunk n of this quad -> buffer n % kNBQ
    const int slot = n % kNBQ;
    mbar_expect_tx(qbar + 8u * slot, kBufBytes);
    tma_load_2d(qbuf + (uint32_t)slot * kBufBytes, &xmap, qbar + 8u * slot, (Q + 4 * (n & 3)) * 64, first_row);
  };
  static_assert(kNBQ <= 4, "the first chunks in flight belong to the first item");

  int n = 0;
  // item descriptors are fetched one item ahead (two dependent global loads otherwise stall every warp at each item start)
  int nx_tile0 = 0, nx_r0 = 0, nx_rows = 0;
  if (i0 < i1) {
    const int clip = __ldg(ctile_clip + i0);
    nx_tile0 = __ldg(ctile_t0 + i0); nx_r0 = __ldg(row_off + clip); nx_rows = __ldg(row_off + clip + 1) - nx_r0;
  }
  if (leader)
    for (int k = 0; k < kNBQ && k < n_total; ++k) issue(k, nx_r0 + nx_tile0 - 30);
#pragma unroll 1
  for (int item = i0; item < i1; ++item) {
    const int tile0 = nx_tile0, r0 = nx_r0, rows = nx_rows;
    if (item + 1 < i1) {
      const int clip = __ldg(ctile_clip + item + 1);
      nx_tile0 = __ldg(ctile_t0 + item + 1); nx_r0 = __ldg(row_off + clip); nx_rows = __ldg(row_off + clip + 1) - nx_r0;
    }
    uint32_t areg[64];                                               // [cc][grp][pr][row half]: bf16x2 conv outputs
    float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;                // sum a, sum a^2 of rows tile0 + lane, tile0 + 32 + lane
#pragma unroll
    for (int cc = 0; cc < 4; ++cc, ++n) {
      const int slot = n % kNBQ;
      const uint32_t buf = qbuf + (uint32_t)slot * kBufBytes;
      mbar_wait(qbar + 8u * slot, (uint32_t)(n / kNBQ) & 1u);
      if (tile0 == 0) {
        // window rows 0..29 are the causal zero padding (the tile holds the previous clip's rows there)
        if (lane < 30) {
#pragma unroll
          for (int grp = 0; grp < 2; ++grp)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(buf + lane * 128 + (((2 * wq + grp) ^ (lane & 7)) << 4)), "r"(0u) : "memory");
        }
        __syncwarp();
      }
      const int chunk = Q + 4 * cc;
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
        const int c16 = 2 * wq + grp;                                // logical 16-byte column of this warp's four pairs
        const uint32_t a_base0 = buf + lane * 128 + ((c16 ^ (lane & 7)) << 4);          // window rows lane + 8 i
        const uint32_t a_base4 = buf + lane * 128 + ((c16 ^ ((lane + 4) & 7)) << 4);    // window rows lane + 8 i + 4
        const uint32_t w_base = wtab_s + 16u * (uint32_t)((chunk * 8 + 2 * wq + grp) * kTabE + ebase);
        float acc[4][4];
#pragma unroll
        for (int pr = 0; pr < 4; ++pr)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[pr][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (kK8) {                                                 // two m16n8k8 per step: half the live fragment registers
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t A0[4], A1[4], B0[4];
              lds128(A0, (hh ? a_base4 : a_base0) + (8 * ks + 4 * hh) * 128);
              lds128(A1, (hh ? a_base4 : a_base0) + (8 * ks + 4 * hh + 32) * 128);
              lds128(B0, w_base + 16u * (8 * ks + 4 * hh));
#pragma unroll
              for (int pr = 0; pr < 4; ++pr) mma1688(acc[pr], A0[pr], A1[pr], B0[pr] & bmask);
            }
            continue;
          }
          uint32_t A0[4], A1[4], A2[4], A3[4];
          lds128(A0, a_base0 + (8 * ks) * 128);
          lds128(A1, a_base0 + (8 * ks + 32) * 128);
          lds128(A2, a_base4 + (8 * ks + 4) * 128);
          lds128(A3, a_base4 + (8 * ks + 36) * 128);
          uint32_t B0[4], B1[4];                                     // one 128-bit load = the same tap of the four pairs
          lds128(B0, w_base + 16u * (8 * ks));
          lds128(B1, w_base + 16u * (8 * ks + 4));
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) mma16816(acc[pr], A0[pr], A1[pr], A2[pr], A3[pr], B0[pr] & bmask, B1[pr] & bmask);
        }
        {
          uint32_t A0[4], A1[4];
          lds128(A0, a_base0 + 32 * 128);
          lds128(A1, a_base0 + 64 * 128);
          uint32_t B0[4];
          lds128(B0, w_base + 16u * 32);
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) mma1688(acc[pr], A0[pr], A1[pr], B0[pr] & bmask);
        }
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {
          const uint32_t wa = pack_bf2(acc[pr][0], acc[pr][1]), wb = pack_bf2(acc[pr][2], acc[pr][3]);
          areg[((cc * 2 + grp) * 4 + pr) * 2 + 0] = wa;
          areg[((cc * 2 + grp) * 4 + pr) * 2 + 1] = wb;
          // (scalar on purpose: measured here, FFMA2 / FADD2 occupy the FMA pipe 1.7x as long as one scalar op and the
          // packed sums need four more registers in a kernel that is at its register limit)
          const float ax = __uint_as_float(wa << 16), ay = __uint_as_float(wa & 0xffff0000u);
          const float bx = __uint_as_float(wb << 16), by = __uint_as_float(wb & 0xffff0000u);
          s1a += ax + ay; s2a = fmaf(ax, ax, s2a); s2a = fmaf(ay, ay, s2a);
          s1b += bx + by; s2b = fmaf(bx, bx, s2b); s2b = fmaf(by, by, s2b);
        }
      }
      // release the buffer: every warp of the quad is done reading it; the leader refills it two chunks ahead
      if (tile0 == 0) fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + Q) : "memory");
      if (leader && n + kNBQ < n_total) issue(n + kNBQ, cc + kNBQ < 4 ? r0 + tile0 - 30 : nx_r0 + nx_tile0 - 30);
    }
    // ---- LayerNorm statistics of the 64 rows: 16 warps x (sum, sum of squares) per row ----
    spart[warp * kItem + lane] = make_float2(s1a, s2a);
    spart[warp * kItem + 32 + lane] = make_float2(s1b, s2b);
    __syncthreads();
    float ta = 0.f, ua = 0.f, tb = 0.f, ub = 0.f;
#pragma unroll
    for (int wi = 0; wi < 16; ++wi) {
      const float2 pa = spart[wi * kItem + lane], pb = spart[wi * kItem + 32 + lane];
      ta += pa.x; ua += pa.y; tb += pb.x; ub += pb.y;
    }
    __syncthreads();                                                 // spart is rewritten at the end of the next item
    const float mean_a = ta * (1.0f / kC), mean_b = tb * (1.0f / kC);
    const float rstd_a = rsqrtf(fmaxf(fmaf(-mean_a, mean_a, ua * (1.0f / kC)), 0.f) + 1e-5f);
    const float rstd_b = rsqrtf(fmaxf(fmaf(-mean_b, mean_b, ub * (1.0f / kC)), 0.f) + 1e-5f);
    const uint64_t ra = pk2(rstd_a, rstd_a), na = pk2(-mean_a * rstd_a, -mean_a * rstd_a);
    const uint64_t rb = pk2(rstd_b, rstd_b), nb = pk2(-mean_b * rstd_b, -mean_b * rstd_b);
    // ---- y / 2 = a * (rstd g / 2) + (b / 2 - mean rstd g / 2);  swish(y) = y/2 + y/2 tanh(y/2);  256-bit stores ----
    const bool va = tile0 + lane < rows, vb = tile0 + 32 + lane < rows;
    __nv_bfloat16* orow = out + (size_t)(r0 + tile0 + lane) * kC + 16 * wq;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int chunk = Q + 4 * cc;
#pragma unroll
      for (int j = 0; j < 8; ++j) {                                  // j = grp * 4 + pr: consecutive channel pairs
        float4 gb;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(gb.x), "=f"(gb.y), "=f"(gb.z), "=f"(gb.w)
                     : "r"(lnp_s + 16u * (uint32_t)(chunk * 32 + 8 * wq + j)));
        const uint64_t gh = pk2(gb.x, gb.y), bh = pk2(gb.z, gb.w);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t w = areg[(cc * 8 + j) * 2 + half];
          const uint64_t a = pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
          const uint64_t h = fma2p(a, mul2p(half ? rb : ra, gh), fma2p(half ? nb : na, gh, bh));
          const float2 hf = up2(h);
          float t0, t1;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(hf.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(hf.y));
          const float2 o = up2(fma2p(h, pk2(t0, t1), h));
          areg[(cc * 8 + j) * 2 + half] = pack_bf2(o.x, o.y);
        }
      }
      __nv_bfloat16* p = orow + chunk * 64;
      if (va)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(areg[(cc * 8 + 0) * 2]),
                     "r"(areg[(cc * 8 + 1) * 2]), "r"(areg[(cc * 8 + 2) * 2]), "r"(areg[(cc * 8 + 3) * 2]), "r"(areg[(cc * 8 + 4) * 2]),
                     "r"(areg[(cc * 8 + 5) * 2]), "r"(areg[(cc * 8 + 6) * 2]), "r"(areg[(cc * 8 + 7) * 2]) : "memory");
      if (vb)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p + (size_t)32 * kC), "r"(areg[(cc * 8 + 0) * 2 + 1]),
                     "r"(areg[(cc * 8 + 1) * 2 + 1]), "r"(areg[(cc * 8 + 2) * 2 + 1]), "r"(areg[(cc * 8 + 3) * 2 + 1]),
                     "r"(areg[(cc * 8 + 4) * 2 + 1]), "r"(areg[(cc * 8 + 5) * 2 + 1]), "r"(areg[(cc * 8 + 6) * 2 + 1]),
                     "r"(areg[(cc * 8 + 7) * 2 + 1]) : "memory");
    }
  }
}

// Measured and rejected in the same session (profiles/r02_dwconv_ab.txt): a 1024-thread version that stores the un-normalised
// rows to `out` right after the MMAs, keeps nothing in registers (64 per thread, 8 warps per scheduler) and normalises in a
// second phase from L2 — 177 us.  The formulation moves 1.2 MB of A fragments and 0.6 MB of B fragments through shared
// memory per 64-row item (every window word is read by nine k-steps), about 14 000 wavefronts per item whoever issues
// them; twice the warps only queue up behind the same shared-memory pipe.

}  // namespace

// bf16 tensor-core path of b2t_dwconv_ln_swish (dwconv.cu dispatches here for dwconv_ring = 7)
int b2t_dwconv_mma_launch(const void* x, const float* w_dw, const float* ln_w, const float* ln_b, const b2t_batch* b, void* out,
                          int variant, cudaStream_t st) {
  CUtensorMap map;
  int rc = make_map(&map, x, b->total_rows, kC, kC, kWin);
  if (rc != B2T_OK) return rc;
  int grid = b2t_num_sms();
  if (b->n_ctiles < grid) grid = b->n_ctiles;
  if (variant == 1) {
    B2T_SMEM_OPT_IN(kSmemBytes, dwconv_mma_kernel<true>);
    dwconv_mma_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(map, w_dw, ln_w, ln_b, b->row_off, b->ctile_clip, b->ctile_t0, b->n_ctiles,
                                                               (__nv_bfloat16*)out);
  } else {
    B2T_SMEM_OPT_IN(kSmemBytes, dwconv_mma_kernel<false>);
    dwconv_mma_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(map, w_dw, ln_w, ln_b, b->row_off, b->ctile_clip, b->ctile_t0, b->n_ctiles,
                                                                (__nv_bfloat16*)out);
  }
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
