// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and the host-side tensor-map encoder shared by the
// sm_100a tensor-core kernels (gemm_tc.cu, attention_tc.cu).  Bit layouts follow
// cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor) of the CUTLASS headers in this image.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace {

constexpr int kBK = 64;   // bf16 elements per 128-byte swizzle row

// ---- PTX wrappers -------------------------------------------------------------------------------
B2T_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

B2T_DEVICE void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
B2T_DEVICE void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
B2T_DEVICE void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
B2T_DEVICE bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time before it reports failure)
B2T_DEVICE bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch failure reported by the host) instead of hanging
// the GPU box.  2^26 probes of a hardware-sleeping try_wait is seconds, far beyond any real wait.
B2T_DEVICE void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) b2t_trap_report("mbar_wait ran out of its bound (bar smem address, parity)", bar, parity);
  }
}
// One lane of a CONVERGED warp.  Code guarded by this predicate inside warp-uniform control flow lets ptxas issue
// tcgen05.mma / tcgen05.commit / TMA loads straight from the uniform datapath; `if (lane == 0)` around a whole loop
// makes the region divergent and every such instruction is wrapped in an elect-and-retry loop with R2UR copies.
B2T_DEVICE bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
B2T_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
B2T_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
B2T_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
B2T_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

B2T_DEVICE void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// pair mode: the load fills THIS CTA's shared memory, the mbarrier may live in the peer (leader) CTA
B2T_DEVICE void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
// multicast: the box lands at the same shared-memory offset in every CTA of `mask` and signals each one's mbarrier
B2T_DEVICE void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank 0
B2T_DEVICE uint32_t mapa_rank0(uint32_t addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
  return r;
}
B2T_DEVICE void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// The same without memory ordering.  A release at cluster scope compiles to MEMBAR.ALL.GPU + ERRBAR: the arriving warp
// waits until every global store it has in flight is acknowledged.  Where the hand-over is about TENSOR memory whose
// reads have already completed (tcgen05.wait::ld, then tcgen05.fence::before_thread_sync) no memory needs publishing.
B2T_DEVICE void mbar_arrive_cluster_relaxed(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
B2T_DEVICE void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

B2T_DEVICE void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
B2T_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, single CTA, bf16 inputs, fp32 accumulate
B2T_DEVICE void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
B2T_DEVICE void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// single-CTA MMAs, arrival multicast to the same barrier offset in every CTA of `mask`
B2T_DEVICE void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
B2T_DEVICE void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
B2T_DEVICE void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
B2T_DEVICE void tmem_alloc_pair(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
B2T_DEVICE void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// 2^x on the FMA / ALU pipes (Cody-Waite: x = n + f, |f| <= 1/2, 2^f by a degree-3 minimax polynomial, n added to the
// exponent field): relative error 1e-4, far below the bf16 rounding of P.  x <= 127; anything below -126 flushes to 0.
B2T_DEVICE float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float xr = x + 12582912.0f;                 // 1.5 * 2^23: the integer nearest to x lands in the low mantissa bits
  const float f = x - (xr - 12582912.0f);
  float p = fmaf(f, 0.05550410866f, 0.24022650696f);
  p = fmaf(p, f, 0.69314718056f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(xr) << 23));
}
// packed fp32 pairs (sm_100 FFMA2 / FADD2): two results per issue slot
B2T_DEVICE float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
B2T_DEVICE float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// 2^x for a pair on the FMA / ALU pipes (see ex2_poly): 2 FMNMX + 3 FADD2 + 3 FFMA2 + 2 LEA per pair.  x <= 127;
// anything below -126 (masked keys: -inf) gives 2^-126 = 1e-38 instead of 0 — below every quantity it is added to.
B2T_DEVICE float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f); x.y = fmaxf(x.y, -126.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
  const float2 xr = fadd2(x, magic);                // the integer nearest to x lands in the low mantissa bits
  const float2 n = fadd2(xr, nmagic);
  const float2 f = fadd2(x, make_float2(-n.x, -n.y));
  float2 p = ffma2(f, make_float2(0.05550410866f, 0.05550410866f), make_float2(0.24022650696f, 0.24022650696f));
  p = ffma2(p, f, make_float2(0.69314718056f, 0.69314718056f));
  p = ffma2(p, f, make_float2(1.0f, 1.0f));
  return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(xr.x) << 23)),
                     __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(xr.y) << 23)));
}
// swish(v) = v / (1 + e^-v) for a pair with ONE special-function op per element: e^-v = 2^(-v log2 e) by the packed
// polynomial above (argument clamped to +-126: v < -87 gives v * 2^-126, i.e. -0 after the bf16 rounding), then MUFU.RCP.
// The MUFU.EX2 + MUFU.RCP form made the N = 4096 swish GEMM's epilogue XU-bound (XU pipe 49 % over the whole kernel with
// the F2F / F2FP conversions on the same pipe).  Relative error 1e-4, far below the bf16 rounding of the result.
B2T_DEVICE float2 swish2(float2 v) {
  float2 x = make_float2(v.x * -1.4426950408889634f, v.y * -1.4426950408889634f);
  x.x = fminf(x.x, 126.0f); x.y = fminf(x.y, 126.0f);
  const float2 e = ex2_poly2(x);
  float ra, rb;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(e.x + 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(e.y + 1.0f));
  return make_float2(v.x * ra, v.y * rb);
}

B2T_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
B2T_DEVICE void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64).
B2T_DEVICE uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1, a_format BF16 [7,10)=1, b_format BF16 [10,13)=1,
// a/b K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29).
constexpr uint32_t make_idesc(int m, int n, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// ---- host: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// [rows, K] bf16 row-major with leading dimension ld (elements); box = 64 (K) x box_rows, 128B swizzle
int make_map(CUtensorMap* map, const void* ptr, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  B2T_REQUIRE(fn != nullptr, B2T_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2T_REQUIRE(r == CUDA_SUCCESS, B2T_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld);
  return B2T_OK;
}


}  // namespace
