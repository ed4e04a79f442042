// Shared device/host helpers for the procyon_b200 sm_100a kernels.
// PTX wrappers (mbarrier, TMA, tcgen05, TMEM) are written by hand; encodings were
// checked against the CUTLASS 4.5 sm100 headers (cute/arch/mma_sm100_desc.hpp).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/procyon_b200.h"

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember, per call site and per
// device ordinal, the largest size already opted into (one process may drive several GPUs; racing threads at worst set
// the attribute twice, which is harmless).
struct SmemOptIn {
  size_t set[64] = {};
  bool need(size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& cur = set[dev & 63];
    if (bytes <= cur) return false;
    cur = bytes;
    return true;
  }
};

namespace pcy {

// ---------------------------------------------------------------------------------------------
// Host-side error plumbing (never throw across the C ABI).
// ---------------------------------------------------------------------------------------------
// status codes: PCY_OK / PCY_ERR_* macros from include/procyon_b200.h

int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what, const char* file, int line);
int num_sms();
// Counts kernels launched by this library (bench.py reports it as gpu_launches).
void count_launch(int n = 1);
extern bool g_fused_rope;
extern bool g_gemm_cluster;
extern int g_gemm_pair_mma;
extern int g_gemm_force_tile;
extern int g_esm_attention_kernel;
extern int g_esm_attention_tail_rows;
extern bool g_esm_attention_q_rope;
extern bool g_skinny_mma;
extern bool g_pdl;  // pcy_set_pdl: programmatic dependent launch for the one-launch-per-op decode chain (default on)

#define PCY_CUDA(expr)                                                     \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) return ::pcy::cuda_error(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define PCY_LAUNCH_CHECK()                                                 \
  do {                                                                     \
    ::pcy::count_launch();                                                 \
    cudaError_t _e = cudaGetLastError();                                   \
    if (_e != cudaSuccess) return ::pcy::cuda_error(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define PCY_REQUIRE(cond, ...)                                             \
  do {                                                                     \
    if (!(cond)) return ::pcy::set_error(PCY_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define PCY_TRY(expr)                                                      \
  do {                                                                     \
    int _rc = (expr);                                                      \
    if (_rc != 0) return _rc;                                              \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int ceil_div(int64_t x, int64_t m) { return (int)((x + m - 1) / m); }

#ifdef __CUDACC__
// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// The one-launch-per-op decode step is a chain of ~260 short kernels; a plain stream (or graph) edge lets kernel n + 1
// start only when kernel n has drained, so every link costs a launch ramp during which HBM idles.  With the
// programmatic-serialization launch attribute kernel n + 1 may become resident as soon as every CTA of kernel n has
// executed pdl_launch_dependents() (first statement of every kernel in the chain) and runs until its own pdl_wait():
// the weight-streaming kernels put the prefetch of their first weight stages BEFORE the wait (weights depend on
// nothing), i.e. the next op's HBM stream starts under the previous op's tail - what the persistent kernel's producer
// warp does inside one launch.  pdl_wait() returns when the preceding grid has completed and flushed, so everything
// that is read or written after it is ordered exactly as with a normal edge; only immutable data (weights, RoPE /
// norm tables) may be touched before it.  Both instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = ::pcy::g_pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// ---------------------------------------------------------------------------------------------
// Small device utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// two floats -> one register of two bf16 (round to nearest even), `lo` in bits 0..15.  As one cvt: through
// __floats2bfloat162_rn + a reinterpret_cast the compiler rebuilt the word from its halves with two PRMTs per pair
// (ncu of the fc1 epilogue: 64 PRMT per 64 columns).
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---- packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 process two fp32 lanes per issued instruction) ----
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t r, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_bcast(float v) { return f2_pack(v, v); }

// erf GELU, as torch.nn.GELU() / fair-esm `gelu`:  gelu(x) = x * Phi(x) = max(x, 0) - |x| * erfc(t)/2 with
// t = |x|/sqrt(2) (for x >= 0 that is x - x * erfc/2, for x < 0 it is x * erfc/2), and erfc(t) = 2^(t*Q(t)) where Q is
// the degree-6 near-minimax fit of log2(erfc(t))/t on [0, 4]; beyond 4 the exponent t*Q(t) - 1 stays below -26.9 and
// falls monotonically (checked up to the fp32 overflow of t^7, where ex2(-inf) = 0), so no clamp is needed.
// Measured against the fp64 erf form over [-8, 8] in fp32 arithmetic: |error| <= 1.3e-6 absolute — 170x below one
// bf16 ulp, the precision every caller stores the result in.  12 instructions with one MUFU.EX2, where erff() costs
// 27 (two coefficient sets picked by 9 FSELs).  The fc1 epilogue of the ESM2 encoder is bound by the FMA pipe on this
// function (round 2: one more FFMA per element cost 4.5 ms of a 51 ms GEMM), hence gelu_erf2: the same arithmetic on
// PAIRS of values with FFMA2 / FMUL2 — bit-identical results, half the FMA-pipe instructions (fc1 51.3 -> 47.5 ms).  (An Abramowitz-Stegun
// erf with rcp + ex2 was measured SLOWER than erff: two quarter-rate MUFU ops.)
#define PCY_GELU_C0 -1.2784041246050037e-05f
#define PCY_GELU_C1 0.00038682681042701006f
#define PCY_GELU_C2 -0.004717740695923567f
#define PCY_GELU_C3 0.03266161307692528f
#define PCY_GELU_C4 -0.1509079933166504f
#define PCY_GELU_C5 -0.9179017543792725f
#define PCY_GELU_C6 -1.6279263496398926f
__device__ __forceinline__ float gelu_erf(float x) {
  const float na = -fabsf(x);
  const float t = na * -0.70710678118654752440f;
  float q = PCY_GELU_C0;
  q = fmaf(q, t, PCY_GELU_C1);
  q = fmaf(q, t, PCY_GELU_C2);
  q = fmaf(q, t, PCY_GELU_C3);
  q = fmaf(q, t, PCY_GELU_C4);
  q = fmaf(q, t, PCY_GELU_C5);
  q = fmaf(q, t, PCY_GELU_C6);
  float h;  // erfc(t) / 2
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(fmaf(t, q, -1.0f)));
  return fmaf(na, h, fmaxf(x, 0.f));
}
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x) {
  const uint64_t na = x | 0x8000000080000000ull;  // -|x| in both lanes
  const uint64_t t = f2_mul(na, f2_bcast(-0.70710678118654752440f));
  uint64_t q = f2_fma(f2_bcast(PCY_GELU_C0), t, f2_bcast(PCY_GELU_C1));
  q = f2_fma(q, t, f2_bcast(PCY_GELU_C2));
  q = f2_fma(q, t, f2_bcast(PCY_GELU_C3));
  q = f2_fma(q, t, f2_bcast(PCY_GELU_C4));
  q = f2_fma(q, t, f2_bcast(PCY_GELU_C5));
  q = f2_fma(q, t, f2_bcast(PCY_GELU_C6));
  const uint64_t e = f2_fma(t, q, f2_bcast(-1.0f));
  float e0, e1, h0, h1, x0, x1;
  f2_unpack(e, e0, e1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h0) : "f"(e0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h1) : "f"(e1));
  f2_unpack(x, x0, x1);
  return f2_fma(na, f2_pack(h0, h1), f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// 16-byte streaming loads / stores
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a pipeline bug must trap (kills the context) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint64_t t0 = 0;
  for (uint32_t it = 0;; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if ((it & 0xfffu) == 0xfffu) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("pcy: mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x,
               bar, parity);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* desc, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const void* desc, uint32_t bar, int32_t c0, int32_t c1,
                                            int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// multicast variant: the box lands at the same CTA-relative offset in every CTA of `cta_mask`, and completes
// `bytes` on the mbarrier at the same CTA-relative offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const void* desc, uint32_t bar, int32_t c0, int32_t c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// thread-block clusters
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tcgen05.commit: arrive on an mbarrier once all previously issued MMAs of this thread retire.
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// same, arriving on the mbarrier at this CTA-relative offset in every CTA of `cta_mask`
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}

// A operand from TMEM (lane = row, two consecutive 16-bit K elements per 32-bit column), B from shared memory
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: thread i of the warp writes 16 consecutive 32-bit columns of lane base + i
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- cta_group::2: the two CTAs of a cluster pair run ONE MMA of M = 256; each holds its 128 rows of A, half of B's
// rows and its 128 rows of the accumulator in its own shared memory / TMEM, and only the even-ranked CTA issues ----
constexpr uint32_t PAIR_LEADER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address (pair of 2)
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {  // one warp of EACH CTA, same offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on the LEADER CTA's mbarrier at the same offset
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const void* desc, uint32_t bar, int32_t c0,
                                                 int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(desc), "r"(bar & PAIR_LEADER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this offset in every CTA of `cta_mask` once the pair's MMAs issued so far retire
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}
// plain arrive on the LEADER CTA's mbarrier at this offset (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PAIR_LEADER_MASK) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle:
// rows are 128 B apart, 8-row groups 1024 B apart (SBO), LBO unused (=1).
// bits [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand (the contraction index is the slow one in smem), 128-byte swizzle:
// 64 MN-elements (128 B) contiguous per k-row; 8 k-rows form a 1024 B atom (SBO between k-groups);
// LBO = distance between successive 64-element MN blocks.
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
// [4,6) D fmt (1=f32) | [7,10) A fmt (1=bf16) | [10,13) B fmt | 15 A major | 16 B major (0 = K-major)
// [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (thread i of the warp gets lane base+i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
#endif  // __CUDACC__

}  // namespace pcy
