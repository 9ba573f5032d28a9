// Relative-key self attention (reference audiotoken/modeling_wav2vec2_bert.py:37-77):
//   out_i = sum_j softmax_j( q_i.k_j / 8 + q_i.E[clamp(j-i,-64,8)+64] / 8 + padmask_j ) v_j
// The reference materialises E[T,T,64] and the bias [B,16,T,T]; here R_i = q_i.E^T (73 values per
// query) is computed once per query tile and gathered on the fly, keys >= valid_rows are skipped
// (their softmax weight is exactly 0 in the reference: finfo.min additive mask), and nothing of
// size T x T touches HBM.  Work item = (clip, 64-query tile, head); online softmax over 64-key tiles.
//
// Two implementations:
//   attention_simt_kernel : CUDA-core, fp32 maths, any activation type — B2T_PREC_FP32 path and the
//                           on-device cross-check of the tensor-core kernel.
//   attention_mma_kernel  : bf16 mma.sync m16n8k16 flash-style kernel (B2T_IMPL_MMA_SYNC; kept as a cross-check).
//   attention_tc_kernel   : tcgen05 / TMEM / TMA kernel in attention_tc.cu (B2T_PREC_BF16 default).
#include "common.cuh"

namespace {

constexpr int kHeads = 16, kHD = 64, kQT = 64, kKT = 64, kRel = 73, kLeft = 64, kRight = 8;
constexpr int kQKV = 3 * kHeads * kHD;   // 3072
constexpr int kH = kHeads * kHD;         // 1024

// ------------------------------------------------------------------------------------------------
// SIMT kernel: 256 threads = 8 warps, each warp owns 8 query rows of the tile.
// ------------------------------------------------------------------------------------------------
struct SimtSmem {
  float q[kQT][kHD];          // 16 KB
  float k[kKT][kHD + 1];      // 16.6 KB
  float v[kKT][kHD];          // 16 KB
  float e[kRel][kHD + 1];     // 19 KB
  float r[kQT][kRel + 1];     // 18.9 KB   R = q . E^T
  float p[8][kKT];            // 2 KB
};

template <typename T, bool kBF16>
__global__ void __launch_bounds__(256)
attention_simt_kernel(const T* __restrict__ qkv, const T* __restrict__ dist_emb,
                      const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                      const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0,
                      T* __restrict__ out, int H /* hidden width = 64 * heads; qkv rows are 3 H wide */) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SimtSmem& s = *reinterpret_cast<SimtSmem*>(smem_raw);
  const int QKV = 3 * H;
  const int clip = qtile_clip[blockIdx.x], q0 = qtile_q0[blockIdx.x], head = blockIdx.y;
  const int r0 = row_off[clip], rows = row_off[clip + 1] - r0, nkeys = valid_rows[clip];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kQT * kHD; i += 256) {
    int qi = i >> 6, d = i & 63;
    s.q[qi][d] = (q0 + qi < rows) ? ld_act(qkv + (size_t)(r0 + q0 + qi) * QKV + head * kHD + d) : 0.f;
  }
  for (int i = tid; i < kRel * kHD; i += 256) s.e[i >> 6][i & 63] = ld_act(dist_emb + i);
  __syncthreads();
  // R[qi][r] = q_i . E_r  (rounded to bf16 on the autocast path: einsum output dtype)
  for (int i = tid; i < kQT * kRel; i += 256) {
    int qi = i / kRel, r = i - qi * kRel;
    float acc = 0.f;
#pragma unroll 8
    for (int d = 0; d < kHD; ++d) acc = fmaf(s.q[qi][d], s.e[r][d], acc);
    s.r[qi][r] = r16<kBF16>(acc);
  }
  float m[8], l[8], o0[8], o1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -INFINITY; l[i] = 0.f; o0[i] = 0.f; o1[i] = 0.f; }

  for (int k0 = 0; k0 < nkeys; k0 += kKT) {
    __syncthreads();
    for (int i = tid; i < kKT * kHD; i += 256) {
      int kj = i >> 6, d = i & 63;
      bool ok = k0 + kj < nkeys;
      const T* base = qkv + (size_t)(r0 + k0 + kj) * QKV + head * kHD + d;
      s.k[kj][d] = ok ? ld_act(base + H) : 0.f;
      s.v[kj][d] = ok ? ld_act(base + 2 * H) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int qi = warp * 8 + i;
      const int qpos = q0 + qi;
      float sc[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int kj = lane + 32 * h;
        float acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < kHD; ++d) acc = fmaf(s.q[qi][d], s.k[kj][d], acc);
        int dist = (k0 + kj) - qpos;
        dist = max(-kLeft, min(kRight, dist)) + kLeft;
        acc = acc * 0.125f + s.r[qi][dist] * 0.125f;
        sc[h] = (k0 + kj < nkeys) ? acc : -INFINITY;
      }
      float mx = warp_max(fmaxf(sc[0], sc[1]));
      float mnew = fmaxf(m[i], mx);
      float corr = __expf(m[i] - mnew);
      float p0 = __expf(sc[0] - mnew), p1 = __expf(sc[1] - mnew);
      l[i] = l[i] * corr + warp_sum(p0 + p1);
      m[i] = mnew;
      s.p[warp][lane] = r16<kBF16>(p0);
      s.p[warp][lane + 32] = r16<kBF16>(p1);
      __syncwarp();
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 16
      for (int kj = 0; kj < kKT; ++kj) {
        float pj = s.p[warp][kj];
        a0 = fmaf(pj, s.v[kj][lane], a0);
        a1 = fmaf(pj, s.v[kj][lane + 32], a1);
      }
      o0[i] = o0[i] * corr + a0;
      o1[i] = o1[i] * corr + a1;
      __syncwarp();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int qpos = q0 + warp * 8 + i;
    if (qpos < rows) {
      float inv = 1.0f / l[i];
      T* o = out + (size_t)(r0 + qpos) * H + head * kHD;
      st_act(o + lane, o0[i] * inv);
      st_act(o + lane + 32, o1[i] * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core kernel (bf16 mma.sync.m16n8k16, fp32 accumulate): 128 threads = 4 warps x 16 query rows.
// ------------------------------------------------------------------------------------------------
B2T_DEVICE void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
B2T_DEVICE void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
B2T_DEVICE void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
B2T_DEVICE uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
B2T_DEVICE void cp_async16(uint32_t dst, const void* src, bool pred) {
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
B2T_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
B2T_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> B2T_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// smem tiles are [rows][64] bf16 = 128 B per row; 16-byte chunks XOR-swizzled by (row & 7)
B2T_DEVICE uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

struct MmaSmem {
  __nv_bfloat16 q[kQT * kHD];          // 8 KB
  __nv_bfloat16 kv[2][2][kKT * kHD];   // 32 KB: [buffer][K|V]; buffer 1 first stages the distance embedding E
                                       //        (80 x 64 bf16 = 10 KB, 73 rows + zero pad) until R is computed
  __nv_bfloat16 r[kQT][kRel + 3];      // 9.5 KB  bf16(q.E^T) — the einsum output the reference rounds to bf16
};
static_assert(sizeof(MmaSmem) <= 57 * 1024, "4 CTAs per SM need <= 56.75 KB each");

__global__ void __launch_bounds__(128, 4)
attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dist_emb,
                     const int32_t* __restrict__ row_off, const int32_t* __restrict__ valid_rows,
                     const int32_t* __restrict__ qtile_clip, const int32_t* __restrict__ qtile_q0,
                     __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  MmaSmem& s = *reinterpret_cast<MmaSmem*>(smem_raw);
  const int clip = qtile_clip[blockIdx.x], q0 = qtile_q0[blockIdx.x], head = blockIdx.y;
  const int r0 = row_off[clip], rows = row_off[clip + 1] - r0, nkeys = valid_rows[clip];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sq = (uint32_t)__cvta_generic_to_shared(s.q);
  const uint32_t se = (uint32_t)__cvta_generic_to_shared(s.kv[1][0]);
  const uint32_t sk[2] = {(uint32_t)__cvta_generic_to_shared(s.kv[0][0]), (uint32_t)__cvta_generic_to_shared(s.kv[1][0])};
  const uint32_t sv[2] = {(uint32_t)__cvta_generic_to_shared(s.kv[0][1]), (uint32_t)__cvta_generic_to_shared(s.kv[1][1])};

  const __nv_bfloat16* qbase = qkv + (size_t)r0 * kQKV + head * kHD;
  auto load_kv = [&](int buf, int k0) {
    // 64 rows x 8 chunks for K and for V: 1024 16-byte copies over 128 threads
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int i = tid + it * 128;          // 0..511
      int row = i >> 3, ch = i & 7;
      bool ok = k0 + row < nkeys;
      const __nv_bfloat16* src = qbase + (size_t)(ok ? k0 + row : 0) * kQKV + ch * 8;
      cp_async16(sk[buf] + swz(row, ch), src + kH, ok);
      cp_async16(sv[buf] + swz(row, ch), src + 2 * kH, ok);
    }
  };
  // Q tile + E
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    int i = tid + it * 128;
    int row = i >> 3, ch = i & 7;
    bool ok = q0 + row < rows;
    cp_async16(sq + swz(row, ch), qbase + (size_t)(ok ? q0 + row : 0) * kQKV + ch * 8, ok);
  }
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    int i = tid + it * 128;            // 0..639 = 80 rows x 8 chunks
    int row = i >> 3, ch = i & 7;
    bool ok = row < kRel;
    cp_async16(se + swz(row, ch), dist_emb + (size_t)(ok ? row : 0) * kHD + ch * 8, ok);
  }
  cp_async_commit();
  if (nkeys > 0) load_kv(0, 0);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();

  // Q fragments for this warp's 16 rows: 4 k-steps x 4 regs
  uint32_t qf[4][4];
  {
    const int row = warp * 16 + (lane & 15);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], sq + swz(row, ks * 2 + (lane >> 4)));
  }
  // R = Q . E^T for the warp's rows (10 n-tiles of 8), stored /8 and bf16-rounded in smem
  {
#pragma unroll
    for (int nt2 = 0; nt2 < 5; ++nt2) {
      float acc[2][4] = {};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bfr[4];
        // rows (n) nt2*16 + (lane&7) + 8*(lane>>4), k chunk ks*2 + ((lane>>3)&1)
        const int nrow = nt2 * 16 + (lane & 7) + ((lane >> 4) << 3);
        ldmatrix_x4(bfr, se + swz(nrow, ks * 2 + ((lane >> 3) & 1)));
        uint32_t b0[2] = {bfr[0], bfr[1]}, b1[2] = {bfr[2], bfr[3]};
        mma_bf16_16816(acc[0], qf[ks], b0);
        mma_bf16_16816(acc[1], qf[ks], b1);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = nt2 * 16 + h * 8 + 2 * (lane & 3);
        const int rlo = warp * 16 + (lane >> 2);
        if (col < kRel) { s.r[rlo][col] = __float2bfloat16_rn(acc[h][0]); s.r[rlo + 8][col] = __float2bfloat16_rn(acc[h][2]); }
        if (col + 1 < kRel) { s.r[rlo][col + 1] = __float2bfloat16_rn(acc[h][1]); s.r[rlo + 8][col + 1] = __float2bfloat16_rn(acc[h][3]); }
      }
    }
  }
  __syncthreads();   // every warp is done with E before key tile 1 overwrites its staging area
  const int rloc0 = warp * 16 + (lane >> 2), rloc1 = rloc0 + 8;   // the two query rows of this thread
  const int qp0 = q0 + rloc0, qp1 = q0 + rloc1;
  // everything below works in the log2 domain: t = (q.k + bf16(q.E_r)) * (log2e / 8), p = 2^(t - m)
  constexpr float kScale = 0.125f * 1.4426950408889634f;
  const float rl0 = __bfloat162float(s.r[rloc0][0]) * kScale, rr0 = __bfloat162float(s.r[rloc0][kRel - 1]) * kScale;
  const float rl1 = __bfloat162float(s.r[rloc1][0]) * kScale, rr1 = __bfloat162float(s.r[rloc1][kRel - 1]) * kScale;

  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[8][4] = {};
  const int nkt = (nkeys + kKT - 1) / kKT;

  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1, k0 = kt * kKT;
    if (kt + 1 < nkt) load_kv(buf ^ 1, k0 + kKT);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    // S = Q K^T : 8 n-tiles (keys) x 4 k-steps
    float sc[8][4] = {};
#pragma unroll
    for (int nt2 = 0; nt2 < 4; ++nt2) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bfr[4];
        const int nrow = nt2 * 16 + (lane & 7) + ((lane >> 4) << 3);
        ldmatrix_x4(bfr, sk[buf] + swz(nrow, ks * 2 + ((lane >> 3) & 1)));
        uint32_t b0[2] = {bfr[0], bfr[1]}, b1[2] = {bfr[2], bfr[3]};
        mma_bf16_16816(sc[nt2 * 2], qf[ks], b0);
        mma_bf16_16816(sc[nt2 * 2 + 1], qf[ks], b1);
      }
    }
    // scale + relative-key bias (constant per row outside the diagonal band) + key mask
    const int dmin = k0 - (q0 + warp * 16 + 15), dmax = (k0 + kKT - 1) - (q0 + warp * 16);
    if (dmax <= -kLeft) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[nt][0] = fmaf(sc[nt][0], kScale, rl0); sc[nt][1] = fmaf(sc[nt][1], kScale, rl0);
        sc[nt][2] = fmaf(sc[nt][2], kScale, rl1); sc[nt][3] = fmaf(sc[nt][3], kScale, rl1);
      }
    } else if (dmin >= kRight) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[nt][0] = fmaf(sc[nt][0], kScale, rr0); sc[nt][1] = fmaf(sc[nt][1], kScale, rr0);
        sc[nt][2] = fmaf(sc[nt][2], kScale, rr1); sc[nt][3] = fmaf(sc[nt][3], kScale, rr1);
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int kj = k0 + nt * 8 + 2 * (lane & 3) + c;
          const float b0v = __bfloat162float(s.r[rloc0][max(-kLeft, min(kRight, kj - qp0)) + kLeft]);
          const float b1v = __bfloat162float(s.r[rloc1][max(-kLeft, min(kRight, kj - qp1)) + kLeft]);
          sc[nt][c] = (sc[nt][c] + b0v) * kScale;
          sc[nt][2 + c] = (sc[nt][2 + c] + b1v) * kScale;
        }
      }
    }
    if (k0 + kKT > nkeys) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (k0 + nt * 8 + 2 * (lane & 3) + c >= nkeys) { sc[nt][c] = -INFINITY; sc[nt][2 + c] = -INFINITY; }
        }
      }
    }
    // online softmax (rows live in 4-lane groups)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = ex2_approx(m0 - mn0), c1 = ex2_approx(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
    uint32_t pf[4][4];   // P as A fragments: 4 k-steps (16 keys each)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p00 = ex2_approx(sc[nt][0] - mn0), p01 = ex2_approx(sc[nt][1] - mn0);
      const float p10 = ex2_approx(sc[nt][2] - mn1), p11 = ex2_approx(sc[nt][3] - mn1);
      ps0 += p00 + p01; ps1 += p10 + p11;
      const int ks = nt >> 1, hi = nt & 1;
      pf[ks][hi * 2 + 0] = pack_bf16(p00, p01);
      pf[ks][hi * 2 + 1] = pack_bf16(p10, p11);
    }
    l0 = l0 * c0 + ps0; l1 = l1 * c1 + ps1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
    // O += P V : 8 n-tiles (head dims) x 4 k-steps (keys); V^T fragments via ldmatrix.trans
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int dt2 = 0; dt2 < 4; ++dt2) {
        uint32_t bfr[4];
        const int krow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        ldmatrix_x4_trans(bfr, sv[buf] + swz(krow, dt2 * 2 + (lane >> 4)));
        uint32_t b0[2] = {bfr[0], bfr[1]}, b1[2] = {bfr[2], bfr[3]};
        mma_bf16_16816(o[dt2 * 2], pf[ks], b0);
        mma_bf16_16816(o[dt2 * 2 + 1], pf[ks], b1);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  // finalize: row sums across the 4-lane group
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = out + (size_t)r0 * kH + head * kHD;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = dt * 8 + 2 * (lane & 3);
    if (qp0 < rows) *reinterpret_cast<uint32_t*>(ob + (size_t)qp0 * kH + col) = pack_bf16(o[dt][0] * i0, o[dt][1] * i0);
    if (qp1 < rows) *reinterpret_cast<uint32_t*>(ob + (size_t)qp1 * kH + col) = pack_bf16(o[dt][2] * i1, o[dt][3] * i1);
  }
}

}  // namespace

int b2t_attention_tensor_tc(const void* qkv, const void* dist_emb, const b2t_batch* b, void* out, int heads, cudaStream_t st);  // attention_tc.cu

extern "C" int b2t_relkey_attention(const void* qkv, const void* dist_emb, const b2t_batch* b,
                                    void* out, int precision, int impl, void* stream) {
  return b2t_attention(qkv, dist_emb, b, out, kHeads, precision, impl, stream);
}

extern "C" int b2t_attention(const void* qkv, const void* dist_emb, const b2t_batch* b, void* out, int heads,
                             int precision, int impl, void* stream) {
  B2T_REQUIRE(qkv && dist_emb && b && out, B2T_ERR_ARG, "b2t_attention: null argument");
  B2T_REQUIRE(heads >= 1 && heads <= 64, B2T_ERR_ARG, "b2t_attention: heads must be in [1, 64] (got %d)", heads);
  B2T_REQUIRE(heads == kHeads || impl != B2T_IMPL_MMA_SYNC, B2T_ERR_ARG, "b2t_attention: the mma.sync cross-check kernel is 16 heads only");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (b->n_qtiles <= 0) return B2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int H = heads * kHD;
  dim3 grid(b->n_qtiles, heads);
  if (precision == B2T_PREC_FP32) {
    B2T_REQUIRE(impl != B2T_IMPL_TENSOR, B2T_ERR_ARG, "b2t_relkey_attention: tensor path is bf16 only");
    B2T_SMEM_OPT_IN(sizeof(SimtSmem), attention_simt_kernel<float, false>);
    attention_simt_kernel<float, false><<<grid, 256, sizeof(SimtSmem), st>>>(
        (const float*)qkv, (const float*)dist_emb, b->row_off, b->valid_rows, b->qtile_clip, b->qtile_q0, (float*)out, H);
  } else if (impl == B2T_IMPL_SIMT) {
    B2T_SMEM_OPT_IN(sizeof(SimtSmem), attention_simt_kernel<__nv_bfloat16, true>);
    attention_simt_kernel<__nv_bfloat16, true><<<grid, 256, sizeof(SimtSmem), st>>>(
        (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dist_emb, b->row_off, b->valid_rows, b->qtile_clip, b->qtile_q0,
        (__nv_bfloat16*)out, H);
  } else if (impl != B2T_IMPL_MMA_SYNC) {
    return b2t_attention_tensor_tc(qkv, dist_emb, b, out, heads, st);
  } else {
    B2T_SMEM_OPT_IN(sizeof(MmaSmem), attention_mma_kernel);
    attention_mma_kernel<<<grid, 128, sizeof(MmaSmem), st>>>(
        (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dist_emb, b->row_off, b->valid_rows, b->qtile_clip, b->qtile_q0,
        (__nv_bfloat16*)out);
  }
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}
