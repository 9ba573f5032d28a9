// Acoustic path: EnCodec 24 kHz SEANet encoder + 2-layer LSTM + residual VQ
// (reference audiotoken/encoder.py:44-57 -> third-party `encodec`; architecture as mirrored by
//  transformers models/encodec/modeling_encodec.py:82-176, 222-313, 364-438).
//
// Round-1 implementation: fp32 CUDA-core kernels (the reference's CPU numerics, BASELINE config 1),
// channels-last activations, ragged clips:
//   seanet_conv_kernel : causal strided conv as an implicit GEMM (64x64x16 tiles).  The A-operand gather
//                        applies the reflect left padding, the reflect "extra" right padding and the
//                        short-input rule of EncodecConv1d._pad1d on the fly (no padded copies), an
//                        optional ELU on the input, bias and an optional residual add.
//   lstm_step_kernel   : one time step of one LSTM layer for all still-active clips (clips sorted by
//                        length, so the active set is a prefix); gate GEMM h.W_hh^T + fused cell update.
//   rvq_kernel         : all n_q residual-VQ stages for 64 frames per block; the residual stays in shared
//                        memory, scores are never written out, every stage's winner is re-checked in fp64.
#include <map>
#include <string>
#include <vector>
#include "vq_cand.cuh"
#include "seanet_tc.h"

void b2t_reset_launch_count();
int b2t_rvq_tensor(const float* emb, int rows, const void* c2, int n_q_total, const float* codebooks,
                   const float* half_norm, const float* cmax_half, int n_q, int16_t* codes, unsigned int* stats,
                   cudaStream_t st);   // rvq_tc.cu

namespace {

constexpr int kLevels = 5;

B2T_DEVICE float eluf(float x) { return x > 0.f ? x : expm1f(x); }

// index into a length-`len` signal for padded position j in [-padL, len + padR)
// (F.pad(mode='reflect') with the zero-extension of short inputs); -1 = zero
B2T_DEVICE int reflect_index(int j, int len, int padL, int padR) {
  const int maxpad = padL > padR ? padL : padR;
  const int le = len <= maxpad ? maxpad + 1 : len;
  int i = j < 0 ? -j : (j < le ? j : 2 * (le - 1) - j);
  return (i >= 0 && i < len) ? i : -1;
}

struct ConvArgs {
  const float* in; int cin;
  const float* w; int kdim_pad;          // [cout][kdim_pad], k index = tap * cin + ci
  const float* bias; float* out; int cout;
  const float* resid;                    // [rows_out, cout] or null
  int k, s, elu_in;
  const int32_t* in_off; const int32_t* in_len; const int32_t* out_off; const int32_t* out_len;
  const int32_t* tile_clip; const int32_t* tile_t0;
  const int64_t* wave_off; const int32_t* true_len;   // layer 0 only (in = waveform)
};

__global__ void __launch_bounds__(256)
seanet_conv_kernel(ConvArgs a) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tid = threadIdx.x;
  const int clip = a.tile_clip[blockIdx.x], t0 = a.tile_t0[blockIdx.x];   // tiles on x: up to 2^31-1 blocks
  const int n0 = blockIdx.y * 64;
  const int in_len = a.in_len[clip], out_len = a.out_len[clip];
  const int padL = a.k - a.s;
  const int padR = out_len * a.s - in_len;           // "extra" padding: ceil(len/s)*s - len
  const long long in_base = a.wave_off ? (long long)a.wave_off[clip] : (long long)a.in_off[clip] * a.cin;
  const int valid_len = a.true_len ? a.true_len[clip] : in_len;   // samples >= true_len are zeros
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const int t = t0 + lr;
  float acc[4][4] = {};
  const int kdim = a.k * a.cin;
  for (int k0 = 0; k0 < a.kdim_pad; k0 += 16) {
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < out_len) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kidx = k0 + lk + i;
        if (kidx < kdim) {
          const int tap = kidx / a.cin, ci = kidx - tap * a.cin;
          const int src = reflect_index(t * a.s - padL + tap, in_len, padL, padR);
          if (src >= 0 && src < valid_len) {
            float v = a.in[in_base + (long long)src * a.cin + ci];
            av[i] = a.elu_in ? eluf(v) : v;
          }
        }
      }
    }
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + lr < a.cout) wv = *reinterpret_cast<const float4*>(a.w + (size_t)(n0 + lr) * a.kdim_pad + k0 + lk);
#pragma unroll
    for (int i = 0; i < 4; ++i) As[lk + i][lr] = av[i];
    Ws[lk][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 x4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float ar[4] = {x4.x, x4.y, x4.z, x4.w}, wr[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int out_base = a.out_off[clip];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tt = t0 + ty * 4 + i;
    if (tt >= out_len) continue;
    const size_t row = (size_t)(out_base + tt) * a.cout;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < a.cout) {
        float v = acc[i][j] + a.bias[n];
        if (a.resid) v += a.resid[row + n];
        a.out[row + n] = v;
      }
    }
  }
}

// One LSTM time step.  Block = 32 clips x 32 hidden units (x 4 gates); 256 threads.
// gates[b, g*512 + j] = xg[row_b, g*512 + j] + sum_d h_prev[b, d] * w_hh[g*512 + j, d]     (i, f, g, o)
__global__ void __launch_bounds__(256)
lstm_step_kernel(const float* __restrict__ xg, const float* __restrict__ w_hh, const float* __restrict__ h_prev,
                 float* __restrict__ h_next, float* __restrict__ c_state, float* __restrict__ out_seq,
                 const float* __restrict__ skip, const int32_t* __restrict__ order,
                 const int32_t* __restrict__ off4, int t, int n_active) {
  __shared__ float As[16][32 + 4];     // h_prev tile  [k][clip]
  __shared__ float Ws[16][128 + 4];    // w_hh tile    [k][gate-row]
  __shared__ float G[32][128 + 1];
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const int ty = tid >> 5, tx = tid & 31;          // 8 x 32: thread -> clips ty*4..+3, gate-rows tx*4..+3
  float acc[4][4] = {};
  for (int k0 = 0; k0 < 512; k0 += 16) {
    if (tid < 128) {                                 // 32 clips x 16 k = 512 floats = 128 float4
      const int r = tid >> 2, kq = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + r < n_active && t > 0) v = *reinterpret_cast<const float4*>(h_prev + (size_t)(b0 + r) * 512 + k0 + kq);
      As[kq][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {                 // 128 gate-rows x 16 k = 512 float4
      const int idx = tid + it * 256;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const int grow = (r >> 5) * 512 + j0 + (r & 31);     // local row r = gate*32 + jj
      const float4 v = *reinterpret_cast<const float4*>(w_hh + (size_t)grow * 512 + k0 + kq);
      Ws[kq][r] = v.x; Ws[kq + 1][r] = v.y; Ws[kq + 2][r] = v.z; Ws[kq + 3][r] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 x4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float ar[4] = {x4.x, x4.y, x4.z, x4.w}, wr[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) G[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();
  // cell update: 32 clips x 32 units = 1024 pairs, 4 per thread
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int p = tid + it * 256;
    const int bl = p >> 5, jj = p & 31;
    const int b = b0 + bl;
    if (b >= n_active) continue;
    const int clip = order[b];
    const size_t row = (size_t)(off4[clip] + t);
    const float* xr = xg + row * 2048 + j0 + jj;
    const float gi = G[bl][jj] + xr[0], gf = G[bl][32 + jj] + xr[512];
    const float gg = G[bl][64 + jj] + xr[1024], go = G[bl][96 + jj] + xr[1536];
    const float si = 1.0f / (1.0f + expf(-gi)), sf = 1.0f / (1.0f + expf(-gf)), so = 1.0f / (1.0f + expf(-go));
    const size_t sidx = (size_t)b * 512 + j0 + jj;
    const float cprev = t > 0 ? c_state[sidx] : 0.f;
    const float c = sf * cprev + si * tanhf(gg);
    const float h = so * tanhf(c);
    c_state[sidx] = c;
    h_next[sidx] = h;
    const size_t oidx = row * 512 + j0 + jj;
    out_seq[oidx] = skip ? h + skip[oidx] : h;
  }
}

// Residual VQ, all stages, 64 frames per block.
struct RvqSmem {
  float rT[128][64 + 4];      // residual, transposed [d][row]
  float eT[16][64 + 4];       // codebook chunk, transposed [k][code]
  Cand merge[64][16];
  int winner[64];
};

__global__ void __launch_bounds__(256)
rvq_kernel(const float* __restrict__ emb, int rows, const float* __restrict__ codebooks /*[n_q][1024][128]*/,
           const float* __restrict__ half_norm /*[n_q][1024]*/, const float* __restrict__ cmax_half /*[n_q]*/,
           int n_q, int16_t* __restrict__ codes /*[n_q][rows]*/) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  RvqSmem& s = *reinterpret_cast<RvqSmem*>(smem_raw);
  const int tid = threadIdx.x, m0 = blockIdx.x * 64;
  const int ty = tid >> 4, tx = tid & 15;
  for (int i = tid; i < 64 * 128; i += 256) {
    const int r = i >> 7, d = i & 127;
    s.rT[d][r] = (m0 + r < rows) ? emb[(size_t)(m0 + r) * 128 + d] : 0.f;
  }
  __syncthreads();
  for (int q = 0; q < n_q; ++q) {
    const float* E = codebooks + (size_t)q * 1024 * 128;
    const float* hn = half_norm + (size_t)q * 1024;
    Cand loc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) loc[i] = cand_empty();
    for (int n0 = 0; n0 < 1024; n0 += 64) {
      float acc[4][4] = {};
      for (int k0 = 0; k0 < 128; k0 += 16) {
        {
          const int r = tid >> 2, kq = (tid & 3) * 4;
          const float4 v = *reinterpret_cast<const float4*>(E + (size_t)(n0 + r) * 128 + k0 + kq);
          s.eT[kq][r] = v.x; s.eT[kq + 1][r] = v.y; s.eT[kq + 2][r] = v.z; s.eT[kq + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          float4 x4 = *reinterpret_cast<const float4*>(&s.rT[k0 + kk][ty * 4]);
          float4 w4 = *reinterpret_cast<const float4*>(&s.eT[kk][tx * 4]);
          const float ar[4] = {x4.x, x4.y, x4.z, x4.w}, wr[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int code = n0 + tx * 4 + j;
        const float h = __ldg(hn + code);
#pragma unroll
        for (int i = 0; i < 4; ++i) cand_insert_ordered(loc[i], acc[i][j] - h, code);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) s.merge[ty * 4 + i][tx] = loc[i];
    __syncthreads();
    if (tid < 64) {
      Cand c = cand_empty();
      for (int u = 0; u < 16; ++u) cand_merge(c, s.merge[tid][u]);
      // fp32 dot of length 128: |err| <= 128 * 2^-24 * |r||e| ; 2x margin
      double rr = 0.0;
      for (int d = 0; d < 128; ++d) rr += (double)s.rT[d][tid] * (double)s.rT[d][tid];
      const float ch = cmax_half[q];
      const float delta = 2.0f * 128.0f * 5.9604645e-8f * ((float)sqrt(rr) * sqrtf(2.0f * ch) + ch);
      int best;
      if (!(c.v4 >= c.v1 - 2.0f * delta)) {
        const int ci[3] = {c.i1, c.i2, c.i3};
        const float cv[3] = {c.v1, c.v2, c.v3};
        double bd = INFINITY;
        best = c.i1;
        for (int u = 0; u < 3; ++u) {
          if (u > 0 && !(cv[u] >= c.v1 - 2.0f * delta)) break;
          if (ci[u] >= 1024) break;
          const float* e = E + (size_t)ci[u] * 128;
          double dd = 0.0;
          for (int d = 0; d < 128; ++d) { const double a_ = (double)s.rT[d][tid] - (double)e[d]; dd += a_ * a_; }
          if (dd < bd || (dd == bd && ci[u] < best)) { bd = dd; best = ci[u]; }
        }
      } else {
        float run = -INFINITY;
        double bd = INFINITY;
        best = 0;
        for (int code = 0; code < 1024; ++code) {
          const float* e = E + (size_t)code * 128;
          float a = 0.f;
          for (int d = 0; d < 128; ++d) a = fmaf(s.rT[d][tid], e[d], a);
          a -= hn[code];
          if (a >= run - 2.0f * delta) {
            double dd = 0.0;
            for (int d = 0; d < 128; ++d) { const double t_ = (double)s.rT[d][tid] - (double)e[d]; dd += t_ * t_; }
            if (dd < bd) { bd = dd; best = code; }
          }
          run = fmaxf(run, a);
        }
      }
      s.winner[tid] = best;
      if (m0 + tid < rows) codes[(size_t)q * rows + m0 + tid] = (int16_t)best;
    }
    __syncthreads();
    // residual update in fp32, as the reference: r = r - E[idx]
    for (int i = tid; i < 64 * 128; i += 256) {
      const int r = i >> 7, d = i & 127;
      s.rT[d][r] -= E[(size_t)s.winner[r] * 128 + d];
    }
    __syncthreads();
  }
}

struct ConvSpec { int cin, cout, k, s; };
const ConvSpec kConvs[18] = {
    {1, 32, 7, 1},   {32, 16, 3, 1},   {16, 32, 1, 1},   {32, 32, 1, 1},   {32, 64, 4, 2},
    {64, 32, 3, 1},  {32, 64, 1, 1},   {64, 64, 1, 1},   {64, 128, 8, 4},  {128, 64, 3, 1},
    {64, 128, 1, 1}, {128, 128, 1, 1}, {128, 256, 10, 5}, {256, 128, 3, 1}, {128, 256, 1, 1},
    {256, 256, 1, 1}, {256, 512, 16, 8}, {512, 128, 7, 1}};

size_t align_up_a(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct b2t_acoustic_model {
  std::map<std::string, const void*> t;
};

extern "C" b2t_acoustic_model* b2t_acoustic_create(void) { return new b2t_acoustic_model(); }
extern "C" void b2t_acoustic_destroy(b2t_acoustic_model* m) { delete m; }
extern "C" int b2t_acoustic_set_tensor(b2t_acoustic_model* m, const char* name, const void* ptr) {
  B2T_REQUIRE(m && name && ptr, B2T_ERR_ARG, "b2t_acoustic_set_tensor: null argument");
  m->t[name] = ptr;
  return B2T_OK;
}

namespace {
struct AcWs {
  float* a[kLevels];   // level input / block output (C = 32,64,128,256,512)
  float* h[4];         // residual-block hidden (C/2)
  float* y[4];         // residual-block output (C)
  float* xg; float* s1; float* s2; float* emb; float* hA; float* hB; float* c;
  size_t total;
};
AcWs ac_carve(void* base, const b2t_acoustic_batch* b) {
  AcWs w;
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? (void*)(p + off) : nullptr; off += align_up_a(bytes, 256); return (float*)r; };
  const int ch[kLevels] = {32, 64, 128, 256, 512};
  for (int l = 0; l < kLevels; ++l) w.a[l] = take((size_t)b->total[l] * ch[l] * 4);
  for (int l = 0; l < 4; ++l) { w.h[l] = take((size_t)b->total[l] * (ch[l] / 2) * 4); w.y[l] = take((size_t)b->total[l] * ch[l] * 4); }
  const size_t t4 = b->total[4];
  w.xg = take(t4 * 2048 * 4); w.s1 = take(t4 * 512 * 4); w.s2 = take(t4 * 512 * 4); w.emb = take(t4 * 128 * 4);
  w.hA = take((size_t)b->n_clips * 512 * 4); w.hB = take((size_t)b->n_clips * 512 * 4); w.c = take((size_t)b->n_clips * 512 * 4);
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t b2t_acoustic_workspace_bytes(const b2t_acoustic_batch* b, int precision) {
  if (!b) return 0;
  if (precision == B2T_PREC_BF16) return b2t_seanet_tc_workspace_bytes(b);
  return ac_carve(nullptr, b).total;
}

bool g_rvq_tensor = true;   // b2t_set_option("rvq_tensor", 0/1)
bool b2t_profile_on();      // pipeline.cu

// ---- optional CUDA-event phase timing of b2t_acoustic_encode (b2t_profile_enable) -------------------
namespace {
struct AcProf {
  std::vector<cudaEvent_t> pool; size_t used = 0;
  struct Span { int cls; size_t e0, e1; };
  std::vector<Span> spans;
  cudaEvent_t get() {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
  }
};
thread_local AcProf g_acprof;
}  // namespace
void b2t_acoustic_mark(int cls, int begin, cudaStream_t st) {
  static thread_local size_t open_e0[4];
  if (!b2t_profile_on() || cls < 0 || cls >= 4) return;
  if (begin) { open_e0[cls] = g_acprof.used; cudaEventRecord(g_acprof.get(), st); }
  else { const size_t e1 = g_acprof.used; cudaEventRecord(g_acprof.get(), st); g_acprof.spans.push_back({cls, open_e0[cls], e1}); }
}
extern "C" int b2t_acoustic_profile_read(float* ms4) {
  B2T_REQUIRE(ms4, B2T_ERR_ARG, "b2t_acoustic_profile_read: null argument");
  for (int i = 0; i < 4; ++i) ms4[i] = 0.f;
  for (auto& sp : g_acprof.spans) {
    B2T_CUDA(cudaEventSynchronize(g_acprof.pool[sp.e1]));
    float ms = 0.f;
    B2T_CUDA(cudaEventElapsedTime(&ms, g_acprof.pool[sp.e0], g_acprof.pool[sp.e1]));
    ms4[sp.cls] += ms;
  }
  g_acprof.spans.clear(); g_acprof.used = 0;
  return B2T_OK;
}

namespace {
int rvq_launch(const b2t_acoustic_model* m, const float* emb, int rows, int n_q, int16_t* codes, cudaStream_t st,
               int impl = B2T_IMPL_AUTO) {
  const char* names[3] = {"rvq.codebooks", "rvq.half_norm", "rvq.cmax_half"};
  const float* t[3];
  for (int i = 0; i < 3; ++i) {
    auto it = m->t.find(names[i]);
    B2T_REQUIRE(it != m->t.end(), B2T_ERR_STATE, "b2t_acoustic_encode: tensor '%s' not set", names[i]);
    t[i] = (const float*)it->second;
  }
  auto c2 = m->t.find("rvq.c2");
  B2T_REQUIRE(impl != B2T_IMPL_TENSOR || c2 != m->t.end(), B2T_ERR_STATE, "b2t_rvq_encode: tensor 'rvq.c2' not set");
  if ((impl == B2T_IMPL_TENSOR || (impl == B2T_IMPL_AUTO && g_rvq_tensor)) && c2 != m->t.end())
    return b2t_rvq_tensor(emb, rows, c2->second, n_q, t[0], t[1], t[2], n_q, codes,
                          (unsigned int*)(m->t.count("rvq.stats") ? m->t.at("rvq.stats") : nullptr), st);
  B2T_SMEM_OPT_IN(sizeof(RvqSmem), rvq_kernel);
  rvq_kernel<<<(rows + 63) / 64, 256, sizeof(RvqSmem), st>>>(emb, rows, t[0], t[1], t[2], n_q, codes);
  B2T_LAUNCH_CHECK();
  return B2T_OK;
}

int encode_tc(const b2t_acoustic_model* m, const float* wave, const b2t_acoustic_batch* b, int n_q, void* workspace,
              size_t workspace_bytes, int16_t* codes, float* emb_out, const int32_t* active_host, cudaStream_t st) {
  SeanetTcWeights wt{};
  bool missing = false;
  std::string miss;
  auto T = [&](const std::string& n) -> const void* {
    auto it = m->t.find(n);
    if (it == m->t.end()) { if (!missing) miss = n; missing = true; return nullptr; }
    return it->second;
  };
  wt.conv0_w = (const float*)T("conv0.w"); wt.conv0_b = (const float*)T("conv0.b");
  const int ch[4] = {32, 64, 128, 256}, st_[4] = {2, 4, 5, 8};
  for (int l = 0; l < 4; ++l) {
    const std::string s = std::to_string(l);
    wt.k3_w[l] = (const __nv_bfloat16*)T("tc.k3" + s + ".w"); wt.k3_b[l] = (const float*)T("tc.k3" + s + ".b");
    wt.res_w[l] = (const __nv_bfloat16*)T("tc.res" + s + ".w"); wt.res_b[l] = (const float*)T("tc.res" + s + ".b");
    wt.down_w[l] = (const __nv_bfloat16*)T("tc.down" + s + ".w"); wt.down_b[l] = (const float*)T("tc.down" + s + ".b");
    wt.k3_kpad[l] = (3 * ch[l] + 63) / 64 * 64;
    wt.res_kpad[l] = (ch[l] * 3 / 2 + 63) / 64 * 64;
    wt.down_kpad[l] = 2 * st_[l] * ch[l];
  }
  for (int j = 0; j < 2; ++j) {
    wt.lstm_w[j] = (const __nv_bfloat16*)T("tc.lstm" + std::to_string(j) + ".w");
    wt.lstm_b[j] = (const float*)T("tc.lstm" + std::to_string(j) + ".b");
  }
  wt.final_w = (const __nv_bfloat16*)T("tc.final.w"); wt.final_b = (const float*)T("tc.final.b");
  B2T_REQUIRE(!missing, B2T_ERR_STATE, "b2t_acoustic_encode(bf16): tensor '%s' not set", miss.c_str());
  float* emb = emb_out ? emb_out : b2t_seanet_tc_emb(workspace, b);
  int rc = b2t_seanet_tc_encode(wt, wave, b, workspace, workspace_bytes, emb, active_host, st);
  if (rc != B2T_OK) return rc;
  b2t_acoustic_mark(3, 1, st);
  rc = rvq_launch(m, emb, b->total[4], n_q, codes, st);
  b2t_acoustic_mark(3, 0, st);
  return rc;
}
}  // namespace

extern "C" int b2t_rvq_encode(const b2t_acoustic_model* m, const float* emb, int rows, int n_q, int impl,
                              int16_t* codes, void* stream) {
  B2T_REQUIRE(m && emb && codes, B2T_ERR_ARG, "b2t_rvq_encode: null argument");
  B2T_REQUIRE(n_q >= 1 && n_q <= 32 && rows >= 0, B2T_ERR_ARG, "b2t_rvq_encode: bad n_q / rows");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  if (rows == 0) return B2T_OK;
  return rvq_launch(m, emb, rows, n_q, codes, (cudaStream_t)stream, impl);
}

extern "C" int b2t_acoustic_encode(const b2t_acoustic_model* m, const float* wave, const b2t_acoustic_batch* b,
                                   int n_q, int precision, void* workspace, size_t workspace_bytes, int16_t* codes,
                                   float* emb_out, const int32_t* active_host, void* stream) {
  B2T_REQUIRE(m && wave && b && workspace && codes && active_host, B2T_ERR_ARG, "b2t_acoustic_encode: null argument");
  B2T_REQUIRE(n_q >= 1 && n_q <= 32, B2T_ERR_ARG, "b2t_acoustic_encode: n_q must be in [1, 32]");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  b2t_reset_launch_count();
  if (b->n_clips <= 0 || b->total[4] <= 0) return B2T_OK;
  if (precision == B2T_PREC_BF16)
    return encode_tc(m, wave, b, n_q, workspace, workspace_bytes, codes, emb_out, active_host, (cudaStream_t)stream);
  AcWs w = ac_carve(workspace, b);
  B2T_REQUIRE(workspace_bytes >= w.total, B2T_ERR_WORKSPACE, "b2t_acoustic_encode: workspace %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  bool missing = false;
  std::string miss;
  auto T = [&](const std::string& n) -> const float* {
    auto it = m->t.find(n);
    if (it == m->t.end()) { if (!missing) miss = n; missing = true; return nullptr; }
    return (const float*)it->second;
  };
  auto conv = [&](int ci, int lin, int lout, const float* in, float* out, const float* resid, int elu) -> int {
    const ConvSpec& cs = kConvs[ci];
    ConvArgs a{};
    a.in = in; a.cin = cs.cin; a.w = T("conv" + std::to_string(ci) + ".w"); a.bias = T("conv" + std::to_string(ci) + ".b");
    B2T_REQUIRE(!missing, B2T_ERR_STATE, "b2t_acoustic_encode: tensor '%s' not set", miss.c_str());
    a.kdim_pad = (cs.k * cs.cin + 15) / 16 * 16;
    a.out = out; a.cout = cs.cout; a.resid = resid; a.k = cs.k; a.s = cs.s; a.elu_in = elu;
    a.in_off = b->off[lin]; a.in_len = b->len[lin]; a.out_off = b->off[lout]; a.out_len = b->len[lout];
    a.tile_clip = b->tile_clip[lout]; a.tile_t0 = b->tile_t0[lout];
    a.wave_off = (ci == 0) ? b->wave_off : nullptr;
    a.true_len = (ci == 0) ? b->true_len : nullptr;
    if (b->n_tiles[lout] <= 0) return B2T_OK;
    dim3 grid(b->n_tiles[lout], (cs.cout + 63) / 64);
    seanet_conv_kernel<<<grid, 256, 0, st>>>(a);
    B2T_LAUNCH_CHECK();
    return B2T_OK;
  };
#define RUN(call) do { int rc__ = (call); if (rc__ != B2T_OK) return rc__; } while (0)
  // conv0, then 4 x (residual block, ELU + strided conv)
  RUN(conv(0, 0, 0, wave, w.a[0], nullptr, 0));
  for (int l = 0; l < 4; ++l) {
    const int c0 = 1 + 4 * l;
    RUN(conv(c0, l, l, w.a[l], w.h[l], nullptr, 1));            // ELU, k3, C -> C/2
    RUN(conv(c0 + 2, l, l, w.a[l], w.y[l], nullptr, 0));        // 1x1 shortcut on the block input
    RUN(conv(c0 + 1, l, l, w.h[l], w.y[l], w.y[l], 1));         // ELU, k1, C/2 -> C, + shortcut
    RUN(conv(c0 + 3, l, l + 1, w.y[l], w.a[l + 1], nullptr, 1)); // ELU, strided conv k=2s
  }
  // LSTM (2 layers) + skip
  const int t4 = b->total[4];
  for (int layer = 0; layer < 2; ++layer) {
    const std::string L = "lstm" + std::to_string(layer) + ".";
    const float* wih = T(L + "w_ih"); const float* whh = T(L + "w_hh"); const float* bias = T(L + "b");
    B2T_REQUIRE(!missing, B2T_ERR_STATE, "b2t_acoustic_encode: tensor '%s' not set", miss.c_str());
    {
      ConvArgs a{};
      a.in = layer == 0 ? w.a[4] : w.s1; a.cin = 512; a.w = wih; a.kdim_pad = 512; a.bias = bias; a.out = w.xg; a.cout = 2048;
      a.k = 1; a.s = 1; a.elu_in = 0;
      a.in_off = b->off[4]; a.in_len = b->len[4]; a.out_off = b->off[4]; a.out_len = b->len[4];
      a.tile_clip = b->tile_clip[4]; a.tile_t0 = b->tile_t0[4];
      dim3 grid(b->n_tiles[4], 2048 / 64);
      seanet_conv_kernel<<<grid, 256, 0, st>>>(a);
      B2T_LAUNCH_CHECK();
    }
    float* hp = w.hA; float* hn = w.hB;
    for (int t = 0; t < b->t_max; ++t) {
      const int na = active_host[t];
      if (na <= 0) break;
      dim3 grid(512 / 32, (na + 31) / 32);
      lstm_step_kernel<<<grid, 256, 0, st>>>(w.xg, whh, hp, hn, w.c, layer == 0 ? w.s1 : w.s2,
                                             layer == 1 ? w.a[4] : nullptr, b->order, b->off[4], t, na);
      B2T_LAUNCH_CHECK();
      float* tmp = hp; hp = hn; hn = tmp;
    }
  }
  // ELU + final conv k7 512 -> 128
  float* emb = emb_out ? emb_out : w.emb;
  RUN(conv(17, 4, 4, w.s2, emb, nullptr, 1));
  // residual VQ
  return rvq_launch(m, emb, t4, n_q, codes, st);
#undef RUN
}

// =====================================================================================================
// Decode half (reference audiotoken/decoder.py:62-76: `quantizer.decode` + `model.decoder`,
// architecture as mirrored by transformers modeling_encodec.py:179-219, 316-353, 440-447).  fp32 CUDA-core kernels
// (the reference's CPU numerics), same channels-last ragged layout and level tables as the fp32 encoder: level 4 is
// the 75 Hz frame rate, level 0 the waveform; every level length is R_l * frames.
// =====================================================================================================
namespace {

// codes int16 [n_q, rows] -> emb fp32 [rows, 128] = sum over stages of the selected codewords, in stage order
__global__ void __launch_bounds__(256)
rvq_decode_kernel(const int16_t* __restrict__ codes, int rows, const float* __restrict__ codebooks, int n_q,
                  float* __restrict__ emb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // (row, 4-float group)
  const int row = i >> 5, g = i & 31;
  if (row >= rows) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int q = 0; q < n_q; ++q) {
    const int c = codes[(size_t)q * rows + row];
    const float4 e = __ldg(reinterpret_cast<const float4*>(codebooks + ((size_t)q * 1024 + c) * 128 + 4 * g));
    acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
  }
  reinterpret_cast<float4*>(emb + (size_t)row * 128)[g] = acc;
}

// Causal transposed conv (k = 2s, stride s) with ELU on its input, as one implicit GEMM per output phase j = u mod s:
//   out[t*s + j, co] = b[co] + sum_ci ELU(x[t, ci]) W[ci, co, j] + ELU(x[t-1, ci]) W[ci, co, j + s]
// (the k - s trailing samples of the full transposed convolution are trimmed, so output length = s * input length).
// w: [s][cout][kdim_pad], k index = half * cin + ci (half 0: x[t], half 1: x[t-1]).  grid = (tiles of 64 input rows,
// ceil(cout / 64), s).
struct ConvTArgs {
  const float* in; int cin; const float* w; int kdim_pad; const float* bias; float* out; int cout; int s;
  const int32_t* in_off; const int32_t* in_len; const int32_t* out_off;
  const int32_t* tile_clip; const int32_t* tile_t0;
};

__global__ void __launch_bounds__(256)
seanet_convt_kernel(ConvTArgs a) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tid = threadIdx.x;
  const int clip = a.tile_clip[blockIdx.x], t0 = a.tile_t0[blockIdx.x];
  const int n0 = blockIdx.y * 64, phase = blockIdx.z;
  const int in_len = a.in_len[clip];
  const long long in_base = (long long)a.in_off[clip] * a.cin;
  const float* W = a.w + (size_t)phase * a.cout * a.kdim_pad;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const int t = t0 + lr;
  float acc[4][4] = {};
  const int kdim = 2 * a.cin;
  for (int k0 = 0; k0 < a.kdim_pad; k0 += 16) {
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < in_len) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kidx = k0 + lk + i;
        if (kidx < kdim) {
          const int half = kidx >= a.cin, ci = kidx - half * a.cin;
          const int src = t - half;
          if (src >= 0) av[i] = eluf(a.in[in_base + (long long)src * a.cin + ci]);
        }
      }
    }
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + lr < a.cout) wv = *reinterpret_cast<const float4*>(W + (size_t)(n0 + lr) * a.kdim_pad + k0 + lk);
#pragma unroll
    for (int i = 0; i < 4; ++i) As[lk + i][lr] = av[i];
    Ws[lk][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 x4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float ar[4] = {x4.x, x4.y, x4.z, x4.w}, wr[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
  const long long out_base = a.out_off[clip];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tt = t0 + ty * 4 + i;
    if (tt >= in_len) continue;
    const size_t row = (size_t)(out_base + (long long)tt * a.s + phase) * a.cout;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < a.cout) a.out[row + n] = acc[i][j] + a.bias[n];
    }
  }
}

// decoder convs in forward order (transposed ones are handled by seanet_convt_kernel): index into "dec.conv<i>"
const ConvSpec kDecConvs[14] = {
    {128, 512, 7, 1},                                             // 0: layers.0
    {256, 128, 3, 1}, {128, 256, 1, 1}, {256, 256, 1, 1},         // 1-3: block 4 (k3, k1, shortcut)
    {128, 64, 3, 1},  {64, 128, 1, 1},  {128, 128, 1, 1},         // 4-6: block 7
    {64, 32, 3, 1},   {32, 64, 1, 1},   {64, 64, 1, 1},           // 7-9: block 10
    {32, 16, 3, 1},   {16, 32, 1, 1},   {32, 32, 1, 1},           // 10-12: block 13
    {32, 1, 7, 1}};                                               // 13: layers.15
const int kDecUpC[4] = {512, 256, 128, 64};                       // input channels of the 4 transposed convs
const int kDecUpS[4] = {8, 5, 4, 2};

struct DecWs {
  float* emb; float* a4; float* xg; float* s1; float* s2; float* hA; float* hB; float* c;
  float* u[4]; float* h[4]; float* y[4];      // per finer level (3, 2, 1, 0): upsampled, block hidden, block output
  size_t total;
};
DecWs dec_carve(void* base, const b2t_acoustic_batch* b) {
  DecWs w;
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? (void*)(p + off) : nullptr; off += align_up_a(bytes, 256); return (float*)r; };
  const size_t t4 = b->total[4];
  w.emb = take(t4 * 128 * 4); w.a4 = take(t4 * 512 * 4); w.xg = take(t4 * 2048 * 4); w.s1 = take(t4 * 512 * 4); w.s2 = take(t4 * 512 * 4);
  w.hA = take((size_t)b->n_clips * 512 * 4); w.hB = take((size_t)b->n_clips * 512 * 4); w.c = take((size_t)b->n_clips * 512 * 4);
  for (int i = 0; i < 4; ++i) {
    const int lvl = 3 - i, C = kDecUpC[i] / 2;
    w.u[i] = take((size_t)b->total[lvl] * C * 4); w.h[i] = take((size_t)b->total[lvl] * (C / 2) * 4); w.y[i] = take((size_t)b->total[lvl] * C * 4);
  }
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t b2t_acoustic_decode_workspace_bytes(const b2t_acoustic_batch* b) {
  if (!b) return 0;
  return dec_carve(nullptr, b).total;
}

extern "C" int b2t_acoustic_decode(const b2t_acoustic_model* m, const int16_t* codes, const b2t_acoustic_batch* b, int n_q,
                                   void* workspace, size_t workspace_bytes, float* wave_out, const int32_t* active_host,
                                   void* stream) {
  B2T_REQUIRE(m && codes && b && workspace && wave_out && active_host, B2T_ERR_ARG, "b2t_acoustic_decode: null argument");
  B2T_REQUIRE(n_q >= 1 && n_q <= 32, B2T_ERR_ARG, "b2t_acoustic_decode: n_q must be in [1, 32]");
  B2T_REQUIRE(b->aligned320, B2T_ERR_ARG, "b2t_acoustic_decode: every clip must be frames * 320 samples long");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  b2t_reset_launch_count();
  if (b->n_clips <= 0 || b->total[4] <= 0) return B2T_OK;
  DecWs w = dec_carve(workspace, b);
  B2T_REQUIRE(workspace_bytes >= w.total, B2T_ERR_WORKSPACE, "b2t_acoustic_decode: workspace %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  bool missing = false;
  std::string miss;
  auto T = [&](const std::string& n) -> const float* {
    auto it = m->t.find(n);
    if (it == m->t.end()) { if (!missing) miss = n; missing = true; return nullptr; }
    return (const float*)it->second;
  };
  auto conv = [&](int ci, int lvl, const float* in, float* out, const float* resid, int elu) -> int {
    const ConvSpec& cs = kDecConvs[ci];
    ConvArgs a{};
    a.in = in; a.cin = cs.cin; a.w = T("dec.conv" + std::to_string(ci) + ".w"); a.bias = T("dec.conv" + std::to_string(ci) + ".b");
    B2T_REQUIRE(!missing, B2T_ERR_STATE, "b2t_acoustic_decode: tensor '%s' not set", miss.c_str());
    a.kdim_pad = (cs.k * cs.cin + 15) / 16 * 16;
    a.out = out; a.cout = cs.cout; a.resid = resid; a.k = cs.k; a.s = 1; a.elu_in = elu;
    a.in_off = b->off[lvl]; a.in_len = b->len[lvl]; a.out_off = b->off[lvl]; a.out_len = b->len[lvl];
    a.tile_clip = b->tile_clip[lvl]; a.tile_t0 = b->tile_t0[lvl];
    if (b->n_tiles[lvl] <= 0) return B2T_OK;
    dim3 grid(b->n_tiles[lvl], (cs.cout + 63) / 64);
    seanet_conv_kernel<<<grid, 256, 0, st>>>(a);
    B2T_LAUNCH_CHECK();
    return B2T_OK;
  };
#define RUN(call) do { int rc__ = (call); if (rc__ != B2T_OK) return rc__; } while (0)
  const int t4 = b->total[4];
  const float* cbs = T("rvq.codebooks");
  B2T_REQUIRE(!missing, B2T_ERR_STATE, "b2t_acoustic_decode: tensor '%s' not set", miss.c_str());
  rvq_decode_kernel<<<(t4 * 32 + 255) / 256, 256, 0, st>>>(codes, t4, cbs, n_q, w.emb);
  B2T_LAUNCH_CHECK();
  RUN(conv(0, 4, w.emb, w.a4, nullptr, 0));                       // Conv(128 -> 512, k7)
  for (int layer = 0; layer < 2; ++layer) {                       // LSTM x2 + skip
    const std::string L = "dec.lstm" + std::to_string(layer) + ".";
    const float* wih = T(L + "w_ih"); const float* whh = T(L + "w_hh"); const float* bias = T(L + "b");
    B2T_REQUIRE(!missing, B2T_ERR_STATE, "b2t_acoustic_decode: tensor '%s' not set", miss.c_str());
    {
      ConvArgs a{};
      a.in = layer == 0 ? w.a4 : w.s1; a.cin = 512; a.w = wih; a.kdim_pad = 512; a.bias = bias; a.out = w.xg; a.cout = 2048;
      a.k = 1; a.s = 1; a.elu_in = 0;
      a.in_off = b->off[4]; a.in_len = b->len[4]; a.out_off = b->off[4]; a.out_len = b->len[4];
      a.tile_clip = b->tile_clip[4]; a.tile_t0 = b->tile_t0[4];
      dim3 grid(b->n_tiles[4], 2048 / 64);
      seanet_conv_kernel<<<grid, 256, 0, st>>>(a);
      B2T_LAUNCH_CHECK();
    }
    float* hp = w.hA; float* hn = w.hB;
    for (int t = 0; t < b->t_max; ++t) {
      const int na = active_host[t];
      if (na <= 0) break;
      dim3 grid(512 / 32, (na + 31) / 32);
      lstm_step_kernel<<<grid, 256, 0, st>>>(w.xg, whh, hp, hn, w.c, layer == 0 ? w.s1 : w.s2,
                                             layer == 1 ? w.a4 : nullptr, b->order, b->off[4], t, na);
      B2T_LAUNCH_CHECK();
      float* tmp = hp; hp = hn; hn = tmp;
    }
  }
  const float* x = w.s2;
  for (int i = 0; i < 4; ++i) {                                   // ELU, transposed conv, residual block
    const int lin = 4 - i, lout = 3 - i, cin = kDecUpC[i], cout = cin / 2, s = kDecUpS[i];
    ConvTArgs a{};
    a.in = x; a.cin = cin; a.w = T("dec.convt" + std::to_string(i) + ".w"); a.bias = T("dec.convt" + std::to_string(i) + ".b");
    B2T_REQUIRE(!missing, B2T_ERR_STATE, "b2t_acoustic_decode: tensor '%s' not set", miss.c_str());
    a.kdim_pad = (2 * cin + 15) / 16 * 16; a.out = w.u[i]; a.cout = cout; a.s = s;
    a.in_off = b->off[lin]; a.in_len = b->len[lin]; a.out_off = b->off[lout];
    a.tile_clip = b->tile_clip[lin]; a.tile_t0 = b->tile_t0[lin];
    dim3 grid(b->n_tiles[lin], (cout + 63) / 64, s);
    seanet_convt_kernel<<<grid, 256, 0, st>>>(a);
    B2T_LAUNCH_CHECK();
    const int c0 = 1 + 3 * i;
    RUN(conv(c0, lout, w.u[i], w.h[i], nullptr, 1));              // ELU, k3, C -> C/2
    RUN(conv(c0 + 2, lout, w.u[i], w.y[i], nullptr, 0));          // 1x1 shortcut on the block input
    RUN(conv(c0 + 1, lout, w.h[i], w.y[i], w.y[i], 1));           // ELU, k1, C/2 -> C, + shortcut
    x = w.y[i];
  }
  RUN(conv(13, 0, x, wave_out, nullptr, 1));                      // ELU, Conv(32 -> 1, k7)
  return B2T_OK;
#undef RUN
}
