// Whole semantic encoder: wave -> log-mel -> stacked features -> conformer layers -> LN -> VQ tokens
// (reference Wav2VecBertEncoder.forward, audiotoken/encoder.py:163-186).  Host-side orchestration of
// the kernels in this directory on one stream; no allocation, no synchronisation.
//
// Tensor names accepted by b2t_semantic_set_tensor (prepared by audiotoken_b200/encoder.py from the
// HF state dict; "act" = bf16 in B2T_PREC_BF16, fp32 in B2T_PREC_FP32; everything else fp32):
//   fp.ln.w fp.ln.b [160] | fp.proj.w [1024,160] act | fp.proj.b [1024]
//   codebook [K,1024] | codebook.half_norm [K] (optional)
//   L<i>.ffn1.ln.w/.b | L<i>.ffn1.w1 [4096,1024] act | .b1 [4096] | .w2 [1024,4096] act | .b2 [1024]
//   L<i>.attn.ln.w/.b | L<i>.attn.wqkv [3072,1024] act (q|k|v rows) | .bqkv [3072] | .wo [1024,1024] act
//   | .bo [1024] | .dist [73,64] act
//   L<i>.conv.ln.w/.b | L<i>.conv.pw1 [2048,1024] act, rows interleaved (a0,g0,a1,g1,...) | .dw [31,1024] (tap-major)
//   | .dwln.w/.b | .pw2 [1024,1024] act
//   L<i>.ffn2.* (as ffn1) | L<i>.final.ln.w/.b
#include <map>
#include <string>
#include <vector>
#include "common.cuh"

void b2t_reset_launch_count();

// ---- optional per-kernel-class CUDA-event profiling (bench.py roofline numbers) ------------------
namespace {
enum { PC_FBANK = 0, PC_LN, PC_GEMM, PC_ATTN, PC_DWCONV, PC_VQ, PC_COUNT };
struct Prof {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Span { int cls; size_t e0, e1; };
  std::vector<Span> spans;
  double flops_gemm = 0.0;
  cudaEvent_t get() {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
  }
};
thread_local Prof g_prof;
struct Scope {
  int cls; cudaStream_t st; size_t e0 = 0; bool on;
  Scope(int c, cudaStream_t s) : cls(c), st(s), on(g_prof.on) {
    if (on) { e0 = g_prof.used; cudaEventRecord(g_prof.get(), st); }
  }
  ~Scope() {
    if (on) { size_t e1 = g_prof.used; cudaEventRecord(g_prof.get(), st); g_prof.spans.push_back({cls, e0, e1}); }
  }
};
}  // namespace

extern "C" int b2t_profile_enable(int on) { g_prof.on = on != 0; return B2T_OK; }
bool b2t_profile_on() { return g_prof.on; }

// Synchronises, adds up the event spans recorded since the last read into ms_per_class[6]
// (fbank, layernorm, gemm, attention, dwconv, vq) and returns the GEMM FLOPs issued in *gemm_flops.
extern "C" int b2t_profile_read(float* ms_per_class, double* gemm_flops) {
  B2T_REQUIRE(ms_per_class, B2T_ERR_ARG, "b2t_profile_read: null argument");
  for (int i = 0; i < PC_COUNT; ++i) ms_per_class[i] = 0.f;
  for (auto& sp : g_prof.spans) {
    B2T_CUDA(cudaEventSynchronize(g_prof.pool[sp.e1]));
    float ms = 0.f;
    B2T_CUDA(cudaEventElapsedTime(&ms, g_prof.pool[sp.e0], g_prof.pool[sp.e1]));
    ms_per_class[sp.cls] += ms;
  }
  if (gemm_flops) *gemm_flops = g_prof.flops_gemm;
  g_prof.spans.clear(); g_prof.used = 0; g_prof.flops_gemm = 0.0;
  return B2T_OK;
}

// ---- developer hook: in-situ determinism check -------------------------------------------------------------------
// b2t_debug_stage_sums(buf, capacity): after every stage of b2t_semantic_encode a 64-bit position-weighted checksum
// of the WHOLE workspace is written to buf[stage] (integer atomics: order independent).  Two runs of the same batch
// from the same initial workspace contents must produce identical vectors; the first differing entry names the stage
// whose kernel is not deterministic (tools/stage_sums.py).
namespace {
unsigned long long* g_sums = nullptr;
int g_sums_cap = 0, g_sums_n = 0;
__global__ void __launch_bounds__(256) workspace_sum_kernel(const uint4* __restrict__ p, size_t n16, unsigned long long* out) {
  unsigned long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    const unsigned long long a = ((unsigned long long)v.x << 32) | v.y, b = ((unsigned long long)v.z << 32) | v.w;
    acc += (a ^ (b * 0x9E3779B97F4A7C15ull)) * (2 * (unsigned long long)i + 1);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
}  // namespace
extern "C" int b2t_debug_stage_sums(unsigned long long* buf, int capacity) {
  g_sums = buf; g_sums_cap = capacity; g_sums_n = 0;
  return B2T_OK;
}
extern "C" int b2t_debug_stage_count(void) { return g_sums_n; }

// b2t_set_option("ffn_resid_epilogue", 0/1): the two K = 4096 feed-forward GEMMs of a layer update the fp32 stream in their
// epilogue (read-modify-write hidden under 32 k-steps per tile) and the LayerNorm that follows reads x alone: same
// bytes in total, but 6 KB (resp. 2 KB) per row leave the stand-alone HBM-bound LayerNorm kernel.  Same arithmetic.
bool g_ffn_resid_epilogue = false;  // measured: LayerNorm -20 ms / step, GEMM +26 ms (the K = 4096 epilogues are not free): off

struct b2t_semantic_model {
  int n_layers;
  int codebook_size;
  int precision;
  int gemm_impl = B2T_IMPL_AUTO;
  int attn_impl = B2T_IMPL_AUTO;
  int mel_bf16 = -1;   // -1: follow precision
  std::map<std::string, const void*> t;
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  float* logmel; float* mean; float* std_; uint8_t* row_valid; void* a160; float* x; void* ln_out;
  void* big; void* att; void* d; void* vq; size_t vq_bytes; size_t total;
};

Workspace carve(void* base, int M, int F, int n_clips, int K, int precision) {
  const size_t act = precision == B2T_PREC_BF16 ? 2 : 4;
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? (void*)(p + off) : nullptr; off += align_up(bytes, 256); return r; };
  Workspace w;
  w.logmel = (float*)take((size_t)F * 80 * 4);
  w.mean = (float*)take((size_t)n_clips * 80 * 4);
  w.std_ = (float*)take((size_t)n_clips * 80 * 4);
  w.row_valid = (uint8_t*)take((size_t)M);
  w.a160 = take((size_t)M * 160 * act);
  w.x = (float*)take((size_t)M * 1024 * 4);
  w.ln_out = take((size_t)M * 1024 * act);
  w.big = take((size_t)M * 4096 * act);
  w.att = take((size_t)M * 1024 * act);
  w.d = take((size_t)M * 1024 * act);
  w.vq_bytes = b2t_vq_workspace_bytes(M, 1024, K);
  w.vq = take(w.vq_bytes);
  w.total = off;
  return w;
}

}  // namespace

extern "C" b2t_semantic_model* b2t_semantic_create(int n_layers, int codebook_size, int precision) {
  if (n_layers < 0 || n_layers > 64 || codebook_size < 1 || codebook_size > 32768 ||
      (precision != B2T_PREC_BF16 && precision != B2T_PREC_FP32)) {
    b2t_set_error("b2t_semantic_create: bad arguments (n_layers=%d K=%d precision=%d)", n_layers, codebook_size, precision);
    return nullptr;
  }
  auto* m = new b2t_semantic_model();
  m->n_layers = n_layers; m->codebook_size = codebook_size; m->precision = precision;
  return m;
}

extern "C" void b2t_semantic_destroy(b2t_semantic_model* m) { delete m; }

extern "C" int b2t_semantic_set_tensor(b2t_semantic_model* m, const char* name, const void* ptr) {
  B2T_REQUIRE(m && name, B2T_ERR_ARG, "b2t_semantic_set_tensor: null argument");
  std::string n(name);
  if (n == "opt.gemm_impl") { m->gemm_impl = (int)(intptr_t)ptr; return B2T_OK; }
  if (n == "opt.attn_impl") { m->attn_impl = (int)(intptr_t)ptr; return B2T_OK; }
  if (n == "opt.mel_bf16") { m->mel_bf16 = (int)(intptr_t)ptr; return B2T_OK; }
  B2T_REQUIRE(ptr, B2T_ERR_ARG, "b2t_semantic_set_tensor: null pointer for %s", name);
  m->t[n] = ptr;
  return B2T_OK;
}

extern "C" size_t b2t_semantic_workspace_bytes(const b2t_semantic_model* m, int total_rows, int total_frames,
                                               int n_clips) {
  if (!m) return 0;
  return carve(nullptr, total_rows, total_frames, n_clips, m->codebook_size, m->precision).total;
}

extern "C" int b2t_semantic_encode(const b2t_semantic_model* m, const float* wave, const b2t_batch* b,
                                   const b2t_fbank_tables* tables, void* workspace, size_t workspace_bytes,
                                   int16_t* tokens, int tap_layer, float* tap_out, void* stream) {
  B2T_REQUIRE(m && wave && b && tables && workspace && tokens, B2T_ERR_ARG, "b2t_semantic_encode: null argument");
  int rc = b2t_arch_ok();
  if (rc != B2T_OK) return rc;
  b2t_reset_launch_count();
  const int M = b->total_rows;
  if (M <= 0) return B2T_OK;
  Workspace w = carve(workspace, M, b->total_frames, b->n_clips, m->codebook_size, m->precision);
  B2T_REQUIRE(workspace_bytes >= w.total, B2T_ERR_WORKSPACE, "b2t_semantic_encode: workspace %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  const int prec = m->precision;
  const bool bf = prec == B2T_PREC_BF16;
  const size_t act = bf ? 2 : 4;

  bool missing = false;
  std::string miss_name;
  auto T = [&](const std::string& n) -> const void* {
    auto it = m->t.find(n);
    if (it == m->t.end()) { if (!missing) miss_name = n; missing = true; return nullptr; }
    return it->second;
  };
  g_sums_n = 0;
  auto stage_sum = [&]() -> int {
    if (g_sums == nullptr || g_sums_n >= g_sums_cap) return B2T_OK;
    B2T_CUDA(cudaMemsetAsync(g_sums + g_sums_n, 0, 8, st));
    workspace_sum_kernel<<<4 * b2t_num_sms(), 256, 0, st>>>((const uint4*)workspace, w.total / 16, g_sums + g_sums_n);
    B2T_LAUNCH_CHECK();
    ++g_sums_n;
    return B2T_OK;
  };
#define RUN(call) do { int rc__ = (call); if (rc__ != B2T_OK) return rc__; rc__ = stage_sum(); if (rc__ != B2T_OK) return rc__; } while (0)
#define NEED() B2T_REQUIRE(!missing, B2T_ERR_STATE, "b2t_semantic_encode: tensor '%s' not set", miss_name.c_str())

  auto gemm = [&](const void* A, int lda, const void* W, const void* bias, void* out, int ldo, float* resid,
                  int N, int K, int epi, float alpha, int round_resid) -> int {
    b2t_gemm_args g{};
    g.A = A; g.lda = lda; g.W = W; g.bias = (const float*)bias; g.out = out; g.ldo = ldo; g.resid = resid;
    g.row_valid = w.row_valid; g.M = M; g.N = N; g.K = K; g.epilogue = epi; g.alpha = alpha;
    g.round_resid_bf16 = round_resid; g.precision = prec; g.impl = bf ? m->gemm_impl : B2T_IMPL_SIMT;
    Scope sc(PC_GEMM, st);
    if (g_prof.on) g_prof.flops_gemm += 2.0 * M * (double)N * K;
    return b2t_gemm(&g, stream);
  };

  auto ln = [&](const float* x, const void* lw, const void* lb, const uint8_t* rv, void* out, int oprec) -> int {
    Scope sc(PC_LN, st);
    return b2t_layernorm(x, (const float*)lw, (const float*)lb, rv, out, M, 1024, oprec, stream);
  };
  // ---- front end
  const int mel_bf16 = m->mel_bf16 >= 0 ? m->mel_bf16 : (bf ? 1 : 0);
  {
    Scope sc(PC_FBANK, st);
    RUN(b2t_fbank_logmel(wave, b, tables, w.logmel, mel_bf16, stream));
    RUN(b2t_fbank_stats(w.logmel, b, w.mean, w.std_, stream));
  }
  const void* fplw = T("fp.ln.w"); const void* fplb = T("fp.ln.b");
  const void* fpw = T("fp.proj.w"); const void* fpb = T("fp.proj.b");
  NEED();
  {
    Scope sc(PC_FBANK, st);
    RUN(b2t_fbank_stack_ln(w.logmel, w.mean, w.std_, b, (const float*)fplw, (const float*)fplb, w.a160, nullptr,
                           w.row_valid, prec, stream));
  }
  // feature projection; padded rows zeroed (HF :493); residual stream x starts here
  RUN(gemm(w.a160, 160, fpw, fpb, nullptr, 0, w.x, 1024, 160, B2T_EPI_BIAS_MASK, 1.f, 0));
  if (tap_layer == 0 && tap_out) B2T_CUDA(cudaMemcpyAsync(tap_out, w.x, (size_t)M * 4096, cudaMemcpyDeviceToDevice, st));

  // Every residual add of the conformer layer is followed by a LayerNorm, so the Linear that produces the
  // branch output writes it as a plain activation `d` and one fused kernel does x += alpha*d and the LayerNorm
  // (b2t_add_layernorm) — no read-modify-write epilogue on the fp32 stream.
  auto add_ln = [&](float alpha, int rr, const void* w1, const void* b1, const void* w2, const void* b2,
                    const uint8_t* rv, void* out) -> int {
    Scope sc(PC_LN, st);
    return b2t_add_layernorm(w.x, w.d, alpha, rr, (const float*)w1, (const float*)b1, (const float*)w2, (const float*)b2, rv,
                             out, M, prec, stream);
  };
  // the same with the residual update already done by the producing GEMM's epilogue (delta = NULL: x is only read)
  auto add_ln_x = [&](const void* w1, const void* b1, const void* w2, const void* b2, const uint8_t* rv, void* out) -> int {
    Scope sc(PC_LN, st);
    return b2t_add_layernorm(w.x, nullptr, 0.f, 0, (const float*)w1, (const float*)b1, (const float*)w2, (const float*)b2, rv,
                             out, M, prec, stream);
  };
  if (m->n_layers > 0) {
    const void* lw = T("L0.ffn1.ln.w"); const void* lb = T("L0.ffn1.ln.b");
    NEED();
    RUN(ln(w.x, lw, lb, nullptr, w.ln_out, prec));
  }
  for (int i = 0; i < m->n_layers; ++i) {
    const std::string L = "L" + std::to_string(i) + ".";
    const int rr = (bf && i == 0) ? 1 : 0;   // layer 0 of the autocast path keeps a bf16 residual stream
    // ---- half-step feed forward 1 (ln_out already holds LN_ffn1(x))
    {
      const void* w1 = T(L + "ffn1.w1"); const void* b1 = T(L + "ffn1.b1"); const void* w2 = T(L + "ffn1.w2"); const void* b2 = T(L + "ffn1.b2");
      const void* nlw = T(L + "attn.ln.w"); const void* nlb = T(L + "attn.ln.b");
      NEED();
      RUN(gemm(w.ln_out, 1024, w1, b1, w.big, 4096, nullptr, 4096, 1024, B2T_EPI_BIAS_SWISH, 1.f, 0));
      if (bf && g_ffn_resid_epilogue) {
        RUN(gemm(w.big, 4096, w2, b2, nullptr, 0, w.x, 1024, 4096, B2T_EPI_RESID, 0.5f, rr));
        RUN(add_ln_x(nlw, nlb, nullptr, nullptr, nullptr, w.ln_out));
      } else {
        RUN(gemm(w.big, 4096, w2, b2, w.d, 1024, nullptr, 1024, 4096, B2T_EPI_BIAS, 1.f, 0));
        RUN(add_ln(0.5f, rr, nlw, nlb, nullptr, nullptr, nullptr, w.ln_out));
      }
    }
    // ---- self attention
    {
      const void* wqkv = T(L + "attn.wqkv"); const void* bqkv = T(L + "attn.bqkv");
      const void* wo = T(L + "attn.wo"); const void* bo = T(L + "attn.bo"); const void* dist = T(L + "attn.dist");
      const void* nlw = T(L + "conv.ln.w"); const void* nlb = T(L + "conv.ln.b");
      NEED();
      RUN(gemm(w.ln_out, 1024, wqkv, bqkv, w.big, 3072, nullptr, 3072, 1024, B2T_EPI_BIAS, 1.f, 0));
      { Scope sc(PC_ATTN, st); RUN(b2t_relkey_attention(w.big, dist, b, w.att, prec, bf ? m->attn_impl : B2T_IMPL_SIMT, stream)); }
      RUN(gemm(w.att, 1024, wo, bo, w.d, 1024, nullptr, 1024, 1024, B2T_EPI_BIAS, 1.f, 0));
      RUN(add_ln(1.f, rr, nlw, nlb, nullptr, nullptr, w.row_valid, w.ln_out));    // conv-module LN, padded rows zeroed
    }
    // ---- convolution module
    {
      const void* pw1 = T(L + "conv.pw1"); const void* dw = T(L + "conv.dw");
      const void* dlw = T(L + "conv.dwln.w"); const void* dlb = T(L + "conv.dwln.b"); const void* pw2 = T(L + "conv.pw2");
      const void* nlw = T(L + "ffn2.ln.w"); const void* nlb = T(L + "ffn2.ln.b");
      NEED();
      RUN(gemm(w.ln_out, 1024, pw1, nullptr, w.big, 1024, nullptr, 2048, 1024, B2T_EPI_GLU, 1.f, 0));
      { Scope sc(PC_DWCONV, st); RUN(b2t_dwconv_ln_swish(w.big, (const float*)dw, (const float*)dlw, (const float*)dlb, b, w.att, prec, stream)); }
      RUN(gemm(w.att, 1024, pw2, nullptr, w.d, 1024, nullptr, 1024, 1024, B2T_EPI_BIAS, 1.f, 0));
      RUN(add_ln(1.f, rr, nlw, nlb, nullptr, nullptr, nullptr, w.ln_out));
    }
    // ---- half-step feed forward 2, final LayerNorm, next layer's first LayerNorm
    {
      const void* w1 = T(L + "ffn2.w1"); const void* b1 = T(L + "ffn2.b1"); const void* w2 = T(L + "ffn2.w2"); const void* b2 = T(L + "ffn2.b2");
      const void* flw = T(L + "final.ln.w"); const void* flb = T(L + "final.ln.b");
      const bool last = i + 1 == m->n_layers;
      const std::string N = "L" + std::to_string(i + 1) + ".";
      const void* nlw = last ? nullptr : T(N + "ffn1.ln.w"); const void* nlb = last ? nullptr : T(N + "ffn1.ln.b");
      NEED();
      RUN(gemm(w.ln_out, 1024, w1, b1, w.big, 4096, nullptr, 4096, 1024, B2T_EPI_BIAS_SWISH, 1.f, 0));
      if (bf && g_ffn_resid_epilogue) {
        RUN(gemm(w.big, 4096, w2, b2, nullptr, 0, w.x, 1024, 4096, B2T_EPI_RESID, 0.5f, rr));
        RUN(add_ln_x(flw, flb, nlw, nlb, nullptr, last ? nullptr : w.ln_out));
      } else {
        RUN(gemm(w.big, 4096, w2, b2, w.d, 1024, nullptr, 1024, 4096, B2T_EPI_BIAS, 1.f, 0));
        RUN(add_ln(0.5f, rr, flw, flb, nlw, nlb, nullptr, last ? nullptr : w.ln_out));
      }
    }
    if (tap_layer == i + 1 && tap_out)
      B2T_CUDA(cudaMemcpyAsync(tap_out, w.x, (size_t)M * 4096, cudaMemcpyDeviceToDevice, st));
  }
  // ---- tail: affine-free LN + nearest codeword (reference encoder.py:175-181)
  const void* cb = T("codebook");
  NEED();
  auto it = m->t.find("codebook.half_norm");
  const float* hn = it == m->t.end() ? nullptr : (const float*)it->second;
  { Scope sc(PC_VQ, st); RUN(b2t_vq_argmin(w.x, 1024, M, 1024, (const float*)cb, hn, m->codebook_size, 1, bf ? B2T_IMPL_AUTO : B2T_IMPL_SIMT, tokens, nullptr, w.vq, w.vq_bytes, stream)); }
  (void)act;
  return B2T_OK;
#undef RUN
#undef NEED
}
