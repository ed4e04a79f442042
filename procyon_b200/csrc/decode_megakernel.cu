// One persistent kernel per Llama decode step (rows = n_inputs * beams <= 4): one CTA per SM streams its balanced
// slice of every weight matrix exactly once, phases are separated by a grid-wide barrier, activations (a few KB)
// bounce through L2.  Per layer:  P1 qkv = Wqkv . rms(x)  |  P2 RoPE + KV append + split-KV attention (the last CTA
// of a kv head merges the splits)  |  P3 x += Wo . attn  |  P4 act = silu(Wg . rms(x)) * (Wu . rms(x))
// |  P5 x += Wdown . act, then logits = Wlm . rms(x).
//
// Batch-1 decode is pure weight streaming (15 GB per token for Llama-3-8B), so the kernel is built around keeping
// HBM busy: a dedicated producer warp walks the CTA's weight slices of ALL phases in order and copies them with
// cp.async.bulk (TMA, 8 KB per copy) into a 192 KB shared-memory ring guarded by full/empty mbarriers.  The weights
// do not depend on activations, so the producer runs ahead across grid barriers, activation staging, epilogues and
// the attention phase: up to a ring's worth of the next phase is already on chip when its consumers start.  The
// number of slots is a multiple of 12 and slot s always belongs to consumer warp s % 12, which takes its chunks in
// order (one consumer per mbarrier: parity waits cannot alias), FMAs them against the activations staged in shared
// memory and reduces.  12 consumer warps + 1 producer warp = 13 warps keeps 128 registers per thread.
//
// Replaces the per-token HF LlamaForCausalLM forward of the reference's generate loops
// (procyon/model/model_unified.py:769, :887 -> procyon/model/pmc_llama.py:581).
#include <algorithm>
#include <cstdlib>
#include <vector>
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace pcy {

namespace {

constexpr int MK_WARPS = 12;                 // consumer warps
constexpr int MK_THREADS = MK_WARPS * 32;    // 384 consumer threads
constexpr int MK_BLOCK = MK_THREADS + 32;    // + one producer warp (warp 12)
constexpr int CH = 4096;                     // weight elements per ring slot (one bulk copy)
constexpr int SLOT_BYTES = CH * 2;
constexpr int HD = 128;
constexpr int KPW = 8;                       // keys per warp and attention work item
constexpr int ATT_CHUNK = KPW * MK_WARPS;    // 96 keys per attention work item
constexpr int PSTR = HD + 4;    // floats per (split, head) attention partial: 128 outputs, max, sum (16-byte rows)
constexpr int MERGE_B = 12;     // splits merged per batch of independent loads

// The poll is a relaxed load and no fence follows it: an acquire load / __threadfence() compiles to ... + CCTL.IVALL,
// which flushes the SM's whole L1 including the local-memory lines (parameters, spills) every warp keeps there, and
// every phase would then start with a chain of L2 round trips.  Nothing needs the invalidation: data written by
// other CTAs is only ever read with ld.cg (L2), after the CTA barrier that follows the poll.
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// barrier over the consumer threads only (the producer warp never joins a CTA-wide barrier)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MK_THREADS) : "memory"); }
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy (TMA, no tensor map), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
  return u;
}
// activations written by other CTAs earlier in this kernel must be read through L2 (L1 is not coherent)
__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

// Grid-wide barrier over the consumer threads of all CTAs (cooperative launch: all CTAs are resident).
// Thread 0 publishes the CTA's writes with a release reduction (cumulative over the CTA barrier before it) and polls
// until everybody has arrived; the second CTA barrier hands that view to the other threads, which read remote data
// through L2 (ld.cg) only.
struct GridBarrier {
  unsigned int* counter;
  unsigned int target;
  unsigned int nblocks;
  __device__ __forceinline__ void sync() {
    consumer_sync();
    if (threadIdx.x == 0) {
      target += nblocks;
      red_release_add(counter, 1u);
      uint64_t t0 = 0;
      for (uint32_t it = 0; ld_relaxed_u32(counter) < target; ++it) {
        if ((it & 0x3fffu) == 0x3fffu) {  // bounded: a scheduling bug must trap, not hang the GPU
          const uint64_t now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000ull) __trap();
        }
      }
    }
    consumer_sync();
  }
};

// ---- legacy tensor-core path for the ring consumers -------------------------------------------------------------
// A weight chunk is one row: 256 consecutive weights (32 segments of 16 B) are read as a 16x16 A fragment by one
// ldmatrix.x4 (lane l supplies segment l: conflict-free), the matching 256 activations as two 16x8 B fragments, and
// two mma.sync m16n8k16 put the 16 partial dot products of the segments pairs on the diagonals D1[n][n] and
// D2[n+8][n].  Everything off the diagonals is discarded - the point is not the FLOPs but the issue slots: 4
// instructions per 512 B of weights instead of ~27 with scalar FMAs (unpacking bf16 costs more than the math), which
// lets the consumers drain a prefetched ring ~6x faster than HBM refills it.
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

enum : int { EPI_BF16 = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_FP32 = 3 };
enum : int { STAGE_PLAIN = 0, STAGE_RMS = 1 };

// slot / parity of ring chunk g (g mod ns and g / ns by multiply-high: exact for g < 2^32 / ns)
struct RingGeom {
  int ns;
  uint32_t magic;  // ceil(2^32 / ns)
  __device__ __forceinline__ void locate(uint32_t g, uint32_t& slot, uint32_t& par) const {
    const uint32_t q = __umulhi(g, magic);
    slot = g - q * (uint32_t)ns;
    par = q & 1u;
  }
};

// This CTA's slice of every phase, as a pair of cumulative fractions scaled to 2^32.  Equal slices by default; after
// pcy_set_decode_sm_shares() proportional to the streaming rate measured for the SM the CTA runs on: with all 148
// SMs pulling flat out some get up to ~20 % less HBM bandwidth than others (systematic by SM id, different from GPU
// to GPU - scripts/profile_decode_skew.py), and with equal slices every grid barrier waits for the slowest.
struct Share {
  uint64_t lo, hi;
};
__device__ __forceinline__ Share cta_share(const uint64_t* table) {
  if (table != nullptr) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    return Share{table[smid], table[smid + 1]};
  }
  return Share{((uint64_t)blockIdx.x << 32) / gridDim.x, ((uint64_t)(blockIdx.x + 1) << 32) / gridDim.x};
}
// contiguous range of `n` items for this CTA
__device__ __forceinline__ void cta_range(const Share& sh, int n, int& lo, int& hi) {
  lo = (int)(((uint64_t)n * sh.lo) >> 32);
  hi = (int)(((uint64_t)n * sh.hi) >> 32);
}
struct Smem {
  bf16* a;       // [MT][K] staged activations
  float* out;    // [out_rows][MT] partial sums
  float* red;    // [MK_WARPS * 4] scratch
  // weight ring: slot i at ring + i * SLOT_BYTES, full barrier at bars + 8 i, empty barrier at bars + 8 (ns + i)
  uint32_t ring, bars;
  RingGeom rg;
  uint32_t chunk0;  // ring chunks consumed by the phases before the current one (same count in the producer)
  Share share;      // this CTA's slice of every phase
  // self-refill mode (no producer warp): the next chunk THIS warp will request - always ring_slots positions ahead of
  // the chunk it consumes, whatever phase that falls into
  int f_ph, f_c, f_n, f_cpr, f_lo, f_K, f_epi;
  uint32_t f_pos;
  const bf16* f_W;
  int64_t f_ldw;
  unsigned long long* tbuf;  // profiling stamps (CTA 0, thread 0) or null
  int tix;
  int phase_ix;  // weight phases finished so far (profiling: per-CTA stream-end times at tbuf[4096 + ...])
  bool skew;     // the timing buffer is large enough for them (the caller wrote the "SKEW" tag at tbuf[4095])
  __device__ __forceinline__ void stamp() {
    if (tbuf != nullptr) {
      if (blockIdx.x == 0 && threadIdx.x == 0) tbuf[tix] = globaltimer_ns();
      ++tix;
    }
  }
};

// weight row of local row r (within the CTA's range starting at logical output o_lo)
__device__ __forceinline__ int weight_row(int epi, int o_lo, int r) {
  if (epi != EPI_SWIGLU) return o_lo + r;
  const int j = o_lo + (r >> 1);  // logical output; even local rows = gate, odd = up
  return (j >> 4) * 32 + (j & 15) + ((r & 1) ? 16 : 0);
}

constexpr int PL = 8;  // producer lanes: chunks g .. g+7 are issued side by side (needs ns >= 12 > PL)

// Producer side of one phase: this CTA's rows of W, cut into chunks of <= CH elements of one row, each one bulk copy
// into the next ring slot.  `g` is the running chunk counter shared by convention with the consumers, which walk the
// identical chunk list in gemv_phase.  One thread cannot issue the copies fast enough for 44 GB/s per SM (wait +
// address arithmetic ~0.25 us per chunk), so PL lanes of the producer warp each take every PL-th chunk.  The lanes
// stay converged: nobody issues before every lane's slot is free.  (Chunk g reuses the slot of chunk g - ns, issued in
// an earlier iteration because ns > PL, so no lane ever waits on a lane of its own iteration.)
//
// `window` bounds the copies IN FLIGHT: chunk g is not issued before chunk g - window has landed.  Consumers drain
// slots faster than HBM fills them, so without the bound every slot of every SM would sit in the DRAM queues (28 MB,
// ~4 us) and the latency-critical DRAM reads of the attention phase (K / V rows) would wait behind them.  The rest
// of the ring fills up only while the consumers are stalled, which is what it is for.
__device__ __noinline__ void produce_phase(uint32_t ring, uint32_t bars, RingGeom rg, uint32_t& g, int window,
                                           Share sh, const bf16* W, int64_t ldw, int n_out, int K, int epi,
                                           uint64_t pol) {
  int lo, hi;
  cta_range(sh, n_out, lo, hi);
  const int rpo = (epi == EPI_SWIGLU) ? 2 : 1;
  const int n_rows = (hi - lo) * rpo;
  const int cpr = (K + CH - 1) / CH;
  const int n_chunks = n_rows * cpr;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t LANES = (1u << PL) - 1u;
  for (int c0 = 0; c0 < n_chunks; c0 += PL) {
    const int c = c0 + lane;
    const bool active = c < n_chunks;
    uint32_t slot = 0, par = 0, bytes = 0, wslot = 0, wpar = 0;
    const bf16* src = W;
    bool ok = !active, landed = true;
    if (active) {
      const int r = c / cpr, kc = c - r * cpr;
      const uint32_t gc = g + (uint32_t)c;
      rg.locate(gc, slot, par);
      src = W + (int64_t)weight_row(epi, lo, r) * ldw + kc * CH;
      bytes = (uint32_t)min(CH, K - kc * CH) * 2u;
      if (gc >= (uint32_t)window) {
        rg.locate(gc - (uint32_t)window, wslot, wpar);
        landed = false;
      }
    }
    uint64_t t0 = 0;
    for (uint32_t it = 0;; ++it) {
      if (!landed) landed = mbar_try_wait(bars + 8u * wslot, wpar);               // chunk g - window has arrived
      if (!ok) ok = mbar_try_wait(bars + 8u * (rg.ns + slot), par ^ 1u);          // slot drained by its consumer warp
      if (__all_sync(LANES, ok && landed)) break;
      if ((it & 0xfffu) == 0xfffu) {  // bounded: a pipeline bug must trap, not hang the GPU
        const uint64_t now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ull) __trap();
      }
    }
    if (active) {
      mbar_arrive_expect_tx(bars + 8u * slot, bytes);
      bulk_g2s(ring + slot * SLOT_BYTES, src, bytes, bars + 8u * slot, pol);
    }
  }
  g += (uint32_t)n_chunks;
}

// ---- self-refill mode ------------------------------------------------------------------------------------------------
// Without the producer warp the block is 12 warps = 3 per scheduler partition = 168 registers per thread instead of 128
// (13 warps round up to 4 per partition), which removes the kernel's local-memory frame: spilled values come back
// through L2 at phase boundaries, ~28 KB of L1 being all that is left beside the ring.  Ring position g belongs to warp
// g % 12 as before; after consuming the chunk at position g a warp requests the chunk at position g + ring_slots
// itself (same slot, and again its own), so every warp always has ring_slots / 12 copies in flight and the empty
// barriers disappear.
struct MegaParams;
struct PhaseW {
  const bf16* W;
  int64_t ldw;
  int n_out, K, epi;
};
__device__ __forceinline__ PhaseW phase_weights(const MegaParams& p, const LlamaLayerPtrs* s_layers, int ph);

__device__ __forceinline__ void fill_load_phase(Smem& sm, const MegaParams& p, const LlamaLayerPtrs* s_layers) {
  const PhaseW w = phase_weights(p, s_layers, sm.f_ph);
  int lo, hi;
  cta_range(sm.share, w.n_out, lo, hi);
  sm.f_cpr = (w.K + CH - 1) / CH;
  sm.f_n = (hi - lo) * (w.epi == EPI_SWIGLU ? 2 : 1) * sm.f_cpr;
  sm.f_lo = lo; sm.f_K = w.K; sm.f_epi = w.epi; sm.f_W = w.W; sm.f_ldw = w.ldw;
}
__device__ __forceinline__ void fill_normalise(Smem& sm, const MegaParams& p, const LlamaLayerPtrs* s_layers,
                                               int n_phases) {
  while (sm.f_ph < n_phases && sm.f_c >= sm.f_n) {
    sm.f_c -= sm.f_n;
    ++sm.f_ph;
    if (sm.f_ph < n_phases) fill_load_phase(sm, p, s_layers);
  }
}
// request this warp's next chunk (it lands in the slot the warp has just drained, or in a still untouched one)
__device__ __forceinline__ void fill_next(Smem& sm, const MegaParams& p, const LlamaLayerPtrs* s_layers, int n_phases,
                                          uint64_t pol) {
  if (sm.f_ph >= n_phases) return;
  if ((threadIdx.x & 31) == 0) {
    const int r = sm.f_c / sm.f_cpr, kc = sm.f_c - r * sm.f_cpr;
    uint32_t slot, par;
    sm.rg.locate(sm.f_pos, slot, par);
    const bf16* src = sm.f_W + (int64_t)weight_row(sm.f_epi, sm.f_lo, r) * sm.f_ldw + kc * CH;
    const uint32_t bytes = (uint32_t)min(CH, sm.f_K - kc * CH) * 2u;
    fence_proxy_async_smem();  // the warp's reads of the slot are ordered before the bulk write
    mbar_arrive_expect_tx(sm.bars + 8u * slot, bytes);
    bulk_g2s(sm.ring + slot * SLOT_BYTES, src, bytes, sm.bars + 8u * slot, pol);
  }
  sm.f_c += MK_WARPS;
  sm.f_pos += MK_WARPS;
  fill_normalise(sm, p, s_layers, n_phases);
}

// out = epi(W[n_out(x2), K] . A[MT, K]).  The CTA's slice of W arrives through the ring in chunks of <= CH elements
// of one row.
// `hook` runs once per warp right after its first chunk (or after the loop if it has none): work that only has to be
// ISSUED during the phase (the attention K / V requests) goes there, where the warp would otherwise wait for HBM.
template <int MT, bool SELF, class Hook>
__device__ __forceinline__ void gemv_phase(Smem& sm, const MegaParams& p, const LlamaLayerPtrs* s_layers, int n_phases,
                                        uint64_t pol, int n_out, int K, int stage, const bf16* A, int64_t lda, int rows,
                                        const bf16* __restrict__ rms_w, float eps, int epi, void* out, int64_t ldo,
                                        Hook&& hook) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int o_lo, o_hi;
  cta_range(sm.share, n_out, o_lo, o_hi);
  const int rpo = (epi == EPI_SWIGLU) ? 2 : 1;
  const int n_rows = (o_hi - o_lo) * rpo;
  const int cpr = (K + CH - 1) / CH;
  const int n_chunks = n_rows * cpr;

  // ---- stage A (plain | RMS-normalised with HF rounding) ----
  constexpr int HOLD = 2;  // 16-byte pieces of a row a thread keeps in registers between the two RMS passes
  const bool rms_in_regs = K <= MK_THREADS * 8 * HOLD;
  for (int m = 0; m < MT; ++m) {
    bf16* dst = sm.a + (int64_t)m * K;
    if (m >= rows) {
      for (int k = tid * 8; k < K; k += MK_THREADS * 8) *reinterpret_cast<uint4*>(dst + k) = make_uint4(0, 0, 0, 0);
      continue;
    }
    const bf16* src = A + (int64_t)m * lda;
    if (stage == STAGE_PLAIN) {
      // all loads of a batch are issued before the first store (one L2 round trip per batch, not per piece)
      constexpr int PB = 5;
      for (int k0 = tid * 8; k0 < K; k0 += MK_THREADS * 8 * PB) {
        uint4 u[PB];
#pragma unroll
        for (int i = 0; i < PB; ++i) {
          const int k = k0 + i * MK_THREADS * 8;
          if (k < K) u[i] = __ldcg(reinterpret_cast<const uint4*>(src + k));
        }
#pragma unroll
        for (int i = 0; i < PB; ++i) {
          const int k = k0 + i * MK_THREADS * 8;
          if (k < K) *reinterpret_cast<uint4*>(dst + k) = u[i];
        }
      }
      continue;
    }
    uint4 held[HOLD], held_w[HOLD];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < HOLD; ++i) {
      const int k = (tid + i * MK_THREADS) * 8;
      const bool in = rms_in_regs && k < K;
      held[i] = in ? __ldcg(reinterpret_cast<const uint4*>(src + k)) : make_uint4(0, 0, 0, 0);
      held_w[i] = in ? *reinterpret_cast<const uint4*>(rms_w + k) : make_uint4(0, 0, 0, 0);  // off the critical path
    }
    if (rms_in_regs) {
#pragma unroll
      for (int i = 0; i < HOLD; ++i) {
        const float2 a = unpack_bf16x2(held[i].x), b = unpack_bf16x2(held[i].y), c = unpack_bf16x2(held[i].z),
                     d = unpack_bf16x2(held[i].w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
      }
    } else {
      for (int k = tid * 8; k < K; k += MK_THREADS * 8) {
        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(src + k));
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) sm.red[warp] = ss;
    consumer_sync();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < MK_WARPS; ++w) tot += sm.red[w];
    const float rstd = rsqrtf(tot / (float)K + eps);
    consumer_sync();
    auto normalise = [&](const uint4& u, const uint4& g, int k) {
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
      const uint32_t gg[4] = {g.x, g.y, g.z, g.w};
      uint32_t oo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = unpack_bf16x2(uu[i]);
        const float2 w = unpack_bf16x2(gg[i]);
        oo[i] = pack_bf16x2(w.x * bf16_round(x.x * rstd), w.y * bf16_round(x.y * rstd));  // HF LlamaRMSNorm
      }
      *reinterpret_cast<uint4*>(dst + k) = make_uint4(oo[0], oo[1], oo[2], oo[3]);
    };
    if (rms_in_regs) {
#pragma unroll
      for (int i = 0; i < HOLD; ++i) {
        const int k = (tid + i * MK_THREADS) * 8;
        if (k < K) normalise(held[i], held_w[i], k);
      }
    } else {
      for (int k = tid * 8; k < K; k += MK_THREADS * 8)
        normalise(__ldcg(reinterpret_cast<const uint4*>(src + k)), *reinterpret_cast<const uint4*>(rms_w + k), k);
    }
  }
  if (cpr > 1)
    for (int i = tid; i < n_rows * MT; i += MK_THREADS) sm.out[i] = 0.f;
  // residual values are final since two phases ago: fetch them now, off the critical path
  const int n_o = o_hi - o_lo;
  float res_pre = 0.f;
  if (epi == EPI_RESIDUAL && tid < n_o * MT) {
    const int o = tid / MT, m = tid % MT;
    if (m < rows) res_pre = ldcg_bf16(reinterpret_cast<const bf16*>(out) + (int64_t)m * ldo + o_lo + o);
  }
  consumer_sync();
  sm.stamp();

  // ---- consume the ring: chunk g of the step lives in slot g % ns and belongs to warp g % 12 ----
  // (Taking a warp's two slots together to share the activation unpack was measured slower: with 24 slots a warp
  // owns exactly two, and holding both leaves the producer nothing to refill while the warp computes.)
  const uint32_t a_base = smem_u32(sm.a);
  bool hooked = false;
  for (int c = (int)((warp + MK_WARPS - sm.chunk0 % MK_WARPS) % MK_WARPS); c < n_chunks; c += MK_WARPS) {
    uint32_t slot, par;
    sm.rg.locate(sm.chunk0 + (uint32_t)c, slot, par);
    const int r = c / cpr, kc = c - r * cpr;
    const int len = min(CH, K - kc * CH);
    const uint32_t wsm = sm.ring + slot * SLOT_BYTES + lane * 16;
    const uint32_t asm0 = a_base + (uint32_t)(kc * CH) * 2u + lane * 16;
    mbar_wait(sm.bars + 8u * slot, par);
    float d1[MT][4], d2[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int i = 0; i < 4; ++i) d1[m][i] = d2[m][i] = 0.f;
    }
    // activation segment of this lane for the B fragments: [0-7 | 16-23 | 8-15 | 24-31] -> (B1 k-lo, B1 k-hi, B2 k-lo,
    // B2 k-hi), pairing segment n with A rows n and segment 8 + n with A rows 8 + n
    const uint32_t bseg = (lane < 8 || lane >= 24) ? lane : (lane < 16 ? lane + 8 : lane - 8);
    const uint32_t asmB = asm0 - lane * 16 + bseg * 16;
#pragma unroll 4
    for (int j = 0; j < len / 256; ++j) {
      uint32_t af[4];
      ldmatrix_x4(wsm + j * 512, af);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        uint32_t bf[4];
        ldmatrix_x4(asmB + (uint32_t)m * (uint32_t)K * 2u + j * 512, bf);
        mma_bf16_16816(d1[m], af, bf[0], bf[1]);
        mma_bf16_16816(d2[m], af, bf[2], bf[3]);
      }
    }
    __syncwarp();
    if (SELF) fill_next(sm, p, s_layers, n_phases, pol);  // slot free: request the chunk this warp needs a ring later
    else if (lane == 0) mbar_arrive(sm.bars + 8u * (sm.rg.ns + slot));  // slot free: the producer may refill it
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      // diagonals: D1[n][n] sits in c0 / c1 and D2[n+8][n] in c2 / c3 of the lanes with lane/4 == 2 (lane%4) (+1)
      const int gq = lane >> 2, q2 = (lane & 3) * 2;
      const float v = warp_sum(gq == q2 ? d1[m][0] + d2[m][2] : (gq == q2 + 1 ? d1[m][1] + d2[m][3] : 0.f));
      if (lane == 0) {
        if (cpr > 1) atomicAdd(&sm.out[r * MT + m], v);
        else sm.out[r * MT + m] = v;
      }
    }
    if (!hooked) {
      hook();
      hooked = true;
    }
  }
  if (!hooked) hook();
  sm.chunk0 += (uint32_t)n_chunks;
  consumer_sync();
  if (sm.skew && threadIdx.x == 0) {  // profiling: when did each CTA finish streaming this phase?
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* sk = sm.tbuf + 4096 + ((size_t)sm.phase_ix * gridDim.x + blockIdx.x) * 2;
    sk[0] = globaltimer_ns();
    sk[1] = smid;
  }
  ++sm.phase_ix;
  sm.stamp();

  // ---- epilogue ----
  for (int i = tid; i < n_o * MT; i += MK_THREADS) {
    const int o = i / MT, m = i % MT;
    if (m >= rows) continue;
    const int col = o_lo + o;
    if (epi == EPI_SWIGLU) {
      const float g = sm.out[(2 * o) * MT + m], up = sm.out[(2 * o + 1) * MT + m];
      reinterpret_cast<bf16*>(out)[(int64_t)m * ldo + col] = __float2bfloat16_rn(silu(g) * up);
    } else {
      float v = sm.out[o * MT + m];
      if (epi == EPI_FP32) {
        reinterpret_cast<float*>(out)[(int64_t)m * ldo + col] = v;
      } else {
        bf16* op = reinterpret_cast<bf16*>(out) + (int64_t)m * ldo + col;
        if (epi == EPI_RESIDUAL) v += (i == tid) ? res_pre : ldcg_bf16(op);
        *op = __float2bfloat16_rn(v);
      }
    }
  }
}

struct MegaParams {
  pcy_llama_config cfg;
  const bf16* embed;
  const bf16* lm_head;
  const bf16* norm;
  const LlamaLayerPtrs* layers;
  const float* rope;
  // session
  int rows, beams, S, max_gen;
  const bf16* kv_prompt;
  const uint8_t* prompt_valid;
  bf16* kv_gen;
  const int32_t* tokens;
  const int32_t* slots;
  const int32_t* state;
  float* logits;
  // scratch (global)
  bf16* x;      // [rows][d]
  bf16* qkv;    // [rows][qkv_dim]
  bf16* attn;   // [rows][H * HD] merged attention output
  bf16* act;    // [rows][ffn]
  float* part;  // [rows][KVH][max_splits][GQ][PSTR]
  int max_splits;
  int ring_slots;
  int out_rows;         // capacity (rows) of the per-CTA partial-sum buffer
  int layers_off;       // byte offset of the shared-memory copy of the layer pointer table
  int window;           // bulk copies in flight per SM (chunks)
  uint32_t ring_magic;  // ceil(2^32 / ring_slots)
  const uint64_t* shares;  // [n_sms + 1] cumulative slice boundaries by SM id (x 2^32), or null = equal slices
  unsigned int* barrier;  // [0] grid barrier counter, [32 + row * KVH + kvh] attention tickets
  unsigned long long* timing;  // optional: globaltimer at every phase boundary (CTA 0), for profiling
};

// The K and V rows one thread needs for an attention work item (see attention_phase for the thread mapping), held in
// registers.  For a CTA's first item they are requested at the START of the qkv phase of the layer: requested just
// before the attention phase (even before the grid barrier in front of it) they arrived ~5 us late - the SM can only
// keep so many missing lines in flight - while seven microseconds of weight streaming hide them completely.
struct AttnLoads {
  uint4 k[KPW / 4][2];
  uint2 v[KPW];
  uint32_t vmask;  // bit j: V row of key j was requested from global memory
  bool khave[KPW / 4], kvalid[KPW / 4];

  __device__ __forceinline__ void request(const MegaParams& p, int layer, int t, int item) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane >> 3, l8 = lane & 7;
    const int KVH = p.cfg.n_kv_heads, kvd = KVH * HD;
    const int pos_cur = p.S + t - 1, ctx = pos_cur + 1;
    const int n_splits = (ctx + ATT_CHUNK - 1) / ATT_CHUNK;
    const int64_t n_prompt = (int64_t)(p.rows / p.beams) * p.S, n_gen = (int64_t)p.rows * p.max_gen;
    const bf16* kp = p.kv_prompt + ((int64_t)layer * 2 + 0) * n_prompt * kvd;
    const bf16* vp = p.kv_prompt + ((int64_t)layer * 2 + 1) * n_prompt * kvd;
    const bf16* kg = p.kv_gen + ((int64_t)layer * 2 + 0) * n_gen * kvd;
    const bf16* vg = p.kv_gen + ((int64_t)layer * 2 + 1) * n_gen * kvd;
    const int split = item % n_splits, kvh = (item / n_splits) % KVH, row = item / (n_splits * KVH);
    const int input = row / p.beams;
    const int k0 = split * ATT_CHUNK;
    const int n_keys = min(ctx, k0 + ATT_CHUNK) - k0;
    // element offset of the K / V row of key position `pos`
    auto row_off = [&](int pos, bool& in_prompt) -> int64_t {
      in_prompt = pos < p.S;
      if (in_prompt) return ((int64_t)input * p.S + pos) * kvd + kvh * HD;
      const int g = pos - p.S;
      const int prow = p.slots[(int64_t)row * p.max_gen + g];
      return ((int64_t)prow * p.max_gen + g) * kvd + kvh * HD;
    };
#pragma unroll
    for (int it = 0; it < KPW / 4; ++it) {
      const int kk = warp * KPW + it * 4 + sub;
      const int pos = k0 + kk;
      kvalid[it] = kk < n_keys;
      khave[it] = kvalid[it] && pos != pos_cur;
      if (khave[it]) {
        bool in_prompt;
        const int64_t off = row_off(pos, in_prompt) + l8 * 16;
        const bf16* kptr = (in_prompt ? kp : kg) + off;
        if (in_prompt && p.prompt_valid) kvalid[it] = p.prompt_valid[(int64_t)input * p.S + pos] != 0;
        k[it][0] = __ldcg(reinterpret_cast<const uint4*>(kptr));
        k[it][1] = __ldcg(reinterpret_cast<const uint4*>(kptr + 8));
      }
    }
    vmask = 0;
    const int pos0 = k0 + warp * KPW;
    if (pos0 + KPW <= p.S) {
      // the warp's 8 keys are all prompt positions (the common case): one base pointer, constant stride
      const bf16* vb = vp + ((int64_t)input * p.S + pos0) * kvd + kvh * HD + lane * 4;
#pragma unroll
      for (int j = 0; j < KPW; ++j) v[j] = __ldcg(reinterpret_cast<const uint2*>(vb + (int64_t)j * kvd));
      vmask = (1u << KPW) - 1u;
    } else {
#pragma unroll
      for (int j = 0; j < KPW; ++j) {
        const int kk = warp * KPW + j;
        const int pos = k0 + kk;
        v[j] = make_uint2(0, 0);
        if (kk < n_keys && pos != pos_cur) {
          bool in_prompt;
          const int64_t off = row_off(pos, in_prompt) + lane * 4;
          v[j] = __ldcg(reinterpret_cast<const uint2*>((in_prompt ? vp : vg) + off));
          vmask |= 1u << j;
        }
      }
    }
  }
};

// Merge of the split-KV partials of one (row, kv head): out[h][dim] = sum_s w_s acc_s / sum_s w_s l_s, w_s = 2^(m_s - max).
// All loads of a batch of MERGE_B splits are independent (one L2 round trip per batch).
template <int GQ>
__device__ __noinline__ void merge_splits(const float* base, int n_splits, bf16* out) {
  const int64_t stride = (int64_t)GQ * PSTR;
  for (int o = threadIdx.x; o < GQ * HD; o += MK_THREADS) {
    const int hh = o / HD, dim = o % HD;
    const float* ph = base + hh * PSTR;
    float mx = -INFINITY;
    if (n_splits > MERGE_B) {  // otherwise the one batch below carries its own maxima
      for (int s0 = 0; s0 < n_splits; s0 += MERGE_B) {
        float mv[MERGE_B];
#pragma unroll
        for (int i = 0; i < MERGE_B; ++i) mv[i] = (s0 + i < n_splits) ? __ldcg(ph + (s0 + i) * stride + HD) : -INFINITY;
#pragma unroll
        for (int i = 0; i < MERGE_B; ++i) mx = fmaxf(mx, mv[i]);
      }
    }
    float l = 0.f, acc = 0.f;
    for (int s0 = 0; s0 < n_splits; s0 += MERGE_B) {
      float mv[MERGE_B], lv[MERGE_B], vv[MERGE_B];
#pragma unroll
      for (int i = 0; i < MERGE_B; ++i) {
        const bool in = s0 + i < n_splits;
        mv[i] = in ? __ldcg(ph + (s0 + i) * stride + HD) : -INFINITY;
        lv[i] = in ? __ldcg(ph + (s0 + i) * stride + HD + 1) : 0.f;
        vv[i] = in ? __ldcg(ph + (s0 + i) * stride + dim) : 0.f;
      }
      if (n_splits <= MERGE_B) {
#pragma unroll
        for (int i = 0; i < MERGE_B; ++i) mx = fmaxf(mx, mv[i]);
      }
#pragma unroll
      for (int i = 0; i < MERGE_B; ++i) {
        const float w = (mv[i] == -INFINITY) ? 0.f : exp2f(mv[i] - mx);
        l = fmaf(w, lv[i], l);
        acc = fmaf(w, vv[i], acc);
      }
    }
    out[hh * HD + dim] = __float2bfloat16_rn(l > 0.f ? acc / l : 0.f);
  }
}

// P2: one work item = (row, kv head, split of ATT_CHUNK keys); warp w owns keys 8w .. 8w+7 of the split.
//   scores : lane (sub, l8) dots 16 dims of key 8w + 4 it + sub with the 4 query heads, 3 shuffles finish the dot
//   P.V    : lane owns dims 4 lane .. 4 lane + 3 of the 4 heads and loops over the warp's 8 keys (no shuffles);
//            the 12 per-warp partial outputs meet in shared memory
// The K / V rows of an item (everything but the current token, which is still being produced by P1) are requested
// BEFORE the grid barrier that ends P1, so their HBM latency overlaps the barrier.  The CTA that finishes the last
// split of a (row, kv head) merges the partials (ticket counter) and writes the bf16 attention output, so that the
// o_proj phase stages 8 KB per row instead of every CTA re-reading every partial.
template <int GQ>
__device__ __forceinline__ void attention_phase(const MegaParams& p, uint8_t* smem_raw, int layer, int t,
                                                GridBarrier& bar, Smem& sm, AttnLoads& L) {
  static_assert(GQ == 4, "the P.V loop reads the 4 head probabilities of a key as one float4");
  sm.stamp();
  // [GQ][HD], pre-scaled by log2(e) / sqrt(HD), dims permuted: dim d = 16 l8 + 4 j + e sits at 32 j + 4 l8 + e, so that
  // the 8 lanes sharing a key read 128 contiguous bytes per float4 (the natural order is a 4-way bank conflict on
  // every one of the 16 loads per head: measured as ~3 us per layer)
  float* s_q = reinterpret_cast<float*>(smem_raw);
  float* s_knew = s_q + GQ * HD;                    // [HD]
  float* s_vnew = s_knew + HD;                      // [HD]
  float* s_sc = s_vnew + HD;                        // [ATT_CHUNK][GQ]
  float* s_ml = s_sc + GQ * ATT_CHUNK;              // [GQ][2]
  int* s_flag = reinterpret_cast<int*>(s_ml + GQ * 2);  // [4]
  float* s_po = s_ml + GQ * 2 + 4;                  // [MK_WARPS][GQ][HD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.cfg.n_heads, KVH = p.cfg.n_kv_heads;
  const int kvd = KVH * HD, qkv_dim = (H + 2 * KVH) * HD;
  const int g_cur = t - 1, pos_cur = p.S + g_cur, ctx = pos_cur + 1;
  const int n_splits = (ctx + ATT_CHUNK - 1) / ATT_CHUNK;
  const int n_items = p.rows * KVH * n_splits;
  const float scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
  const int64_t n_prompt = (int64_t)(p.rows / p.beams) * p.S, n_gen = (int64_t)p.rows * p.max_gen;
  const bf16* kp = p.kv_prompt + ((int64_t)layer * 2 + 0) * n_prompt * kvd;
  const bf16* vp = p.kv_prompt + ((int64_t)layer * 2 + 1) * n_prompt * kvd;
  bf16* kg = p.kv_gen + ((int64_t)layer * 2 + 0) * n_gen * kvd;
  bf16* vg = p.kv_gen + ((int64_t)layer * 2 + 1) * n_gen * kvd;
  const int sub = lane >> 3, l8 = lane & 7;  // scores: 8 lanes per key, 16 dims per lane

  // (the first item's rows were requested at the start of the qkv phase, see the phase loop)
  for (int item = blockIdx.x;; item += gridDim.x) {
    const bool first_item = item == (int)blockIdx.x;  // profiling stamps cover the first item only
    if (!first_item) {  // (only with rows > 1 or very long contexts: more items than CTAs)
      if (item >= n_items) break;
      consumer_sync();  // the previous item is done with the shared buffers
    }
    if (!first_item) L.request(p, layer, t, item);
    if (first_item) {
      sm.stamp();
      bar.sync();  // qkv of this step is complete
      sm.stamp();
      if (item >= n_items) break;
    }
    const int split = item % n_splits, kvh = (item / n_splits) % KVH, row = item / (n_splits * KVH);
    const bf16* qkv_row = p.qkv + (int64_t)row * qkv_dim;
    const int k0 = split * ATT_CHUNK;
    {
      const float2* cs = reinterpret_cast<const float2*>(p.rope) + (int64_t)pos_cur * (HD / 2);
      for (int i = tid; i < (GQ + 1) * (HD / 2); i += MK_THREADS) {
        const int hh = i / (HD / 2), j = i % (HD / 2);
        const bf16* src = (hh < GQ) ? qkv_row + (kvh * GQ + hh) * HD : qkv_row + (H + kvh) * HD;
        const float lo = ldcg_bf16(src + j), hi = ldcg_bf16(src + j + HD / 2);
        const float2 c = cs[j];
        const float o_lo = bf16_round(lo * c.x - hi * c.y), o_hi = bf16_round(hi * c.x + lo * c.y);
        if (hh < GQ) {
          const int d0 = j, d1 = j + HD / 2;
          s_q[hh * HD + ((d0 >> 2) & 3) * 32 + (d0 >> 4) * 4 + (d0 & 3)] = o_lo * scale_log2;
          s_q[hh * HD + ((d1 >> 2) & 3) * 32 + (d1 >> 4) * 4 + (d1 & 3)] = o_hi * scale_log2;
        } else {
          s_knew[j] = o_lo;
          s_knew[j + HD / 2] = o_hi;
        }
      }
      if (tid < HD) s_vnew[tid] = ldcg_bf16(qkv_row + (H + KVH + kvh) * HD + tid);
    }
    consumer_sync();
    if (first_item) sm.stamp();
    if (pos_cur >= k0 && pos_cur < k0 + ATT_CHUNK && tid < HD) {
      const int64_t off = ((int64_t)row * p.max_gen + g_cur) * kvd + kvh * HD + tid;
      kg[off] = __float2bfloat16_rn(s_knew[tid]);
      vg[off] = __float2bfloat16_rn(s_vnew[tid]);
    }
    // ---- scores ----
#pragma unroll
    for (int it = 0; it < KPW / 4; ++it) {
      const int kk = warp * KPW + it * 4 + sub;
      float kf[16];
      if (L.khave[it]) {
        const uint32_t w[8] = {L.k[it][0].x, L.k[it][0].y, L.k[it][0].z, L.k[it][0].w,
                               L.k[it][1].x, L.k[it][1].y, L.k[it][1].z, L.k[it][1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          kf[2 * j] = f.x;
          kf[2 * j + 1] = f.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) kf[j] = s_knew[l8 * 16 + j];  // current token (or unused)
      }
      // one head at a time (not unrolled: 64 query values in flight would push the K / V rows out of registers)
#pragma unroll 1
      for (int h = 0; h < GQ; ++h) {
        const float4* q4 = reinterpret_cast<const float4*>(s_q + h * HD + l8 * 4);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 q = q4[j * 8];
          s0 = fmaf(kf[4 * j], q.x, s0);
          s1 = fmaf(kf[4 * j + 1], q.y, s1);
          s0 = fmaf(kf[4 * j + 2], q.z, s0);
          s1 = fmaf(kf[4 * j + 3], q.w, s1);
        }
        float a = s0 + s1;
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (l8 == 0) s_sc[kk * GQ + h] = L.kvalid[it] ? a : -INFINITY;
      }
    }
    consumer_sync();
    if (first_item) sm.stamp();
    if (warp < GQ) {
      const int h = warp;
      float m = -INFINITY;
      for (int k = lane; k < ATT_CHUNK; k += 32) m = fmaxf(m, s_sc[k * GQ + h]);
      m = warp_max(m);
      float l = 0.f;
      for (int k = lane; k < ATT_CHUNK; k += 32) {
        const float pr = (m == -INFINITY) ? 0.f : exp2f(s_sc[k * GQ + h] - m);
        s_sc[k * GQ + h] = pr;
        l += pr;
      }
      l = warp_sum(l);
      if (lane == 0) { s_ml[h * 2] = m; s_ml[h * 2 + 1] = l; }
    }
    consumer_sync();
    if (first_item) sm.stamp();
    // ---- P.V ----
    {
      float acc[GQ][4];
#pragma unroll
      for (int h = 0; h < GQ; ++h) acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;
#pragma unroll
      for (int j = 0; j < KPW; ++j) {
        const int kk = warp * KPW + j;
        float v0, v1, v2, v3;
        if (L.vmask & (1u << j)) {
          const float2 f0 = unpack_bf16x2(L.v[j].x), f1 = unpack_bf16x2(L.v[j].y);
          v0 = f0.x; v1 = f0.y; v2 = f1.x; v3 = f1.y;
        } else {  // the current token (its probability is 0 for keys past the context)
          const float4 f = *reinterpret_cast<const float4*>(s_vnew + lane * 4);
          v0 = bf16_round(f.x); v1 = bf16_round(f.y); v2 = bf16_round(f.z); v3 = bf16_round(f.w);
        }
        const float4 pr = *reinterpret_cast<const float4*>(s_sc + kk * GQ);
        const float ph[GQ] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
        for (int h = 0; h < GQ; ++h) {
          acc[h][0] = fmaf(ph[h], v0, acc[h][0]);
          acc[h][1] = fmaf(ph[h], v1, acc[h][1]);
          acc[h][2] = fmaf(ph[h], v2, acc[h][2]);
          acc[h][3] = fmaf(ph[h], v3, acc[h][3]);
        }
      }
#pragma unroll
      for (int h = 0; h < GQ; ++h)
        *reinterpret_cast<float4*>(s_po + (warp * GQ + h) * HD + lane * 4) =
            make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
    }
    consumer_sync();
    if (first_item) sm.stamp();
    float* part = p.part + (((int64_t)row * KVH + kvh) * p.max_splits + split) * GQ * PSTR;
    for (int o = tid; o < GQ * HD; o += MK_THREADS) {
      const int hh = o / HD, dim = o % HD;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < MK_WARPS; ++w) s += s_po[(w * GQ + hh) * HD + dim];
      part[hh * PSTR + dim] = s;
    }
    if (tid < GQ) {
      part[tid * PSTR + HD] = s_ml[tid * 2];
      part[tid * PSTR + HD + 1] = s_ml[tid * 2 + 1];
    }
    // ---- ticket: the CTA that completes the last split of (row, kv head) merges them ----
    consumer_sync();
    if (tid == 0) {
      unsigned int tk;  // release: the CTA's partials are visible before the ticket is; partials are read with ld.cg
      asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;"
                   : "=r"(tk) : "l"(p.barrier + 32 + row * KVH + kvh) : "memory");
      s_flag[0] = tk == (unsigned int)(n_splits * (layer + 1) - 1);
    }
    consumer_sync();
    if (s_flag[0])
      merge_splits<GQ>(p.part + (((int64_t)row * KVH + kvh) * p.max_splits) * GQ * PSTR, n_splits,
                       p.attn + (int64_t)row * (H * HD) + kvh * GQ * HD);
    if (first_item) sm.stamp();
  }
}

// Touches the K / V pages this CTA's first attention item of `layer` will read (one word each, result unused), long
// before the attention phase: each layer's K and V live on 2 MB pages last used a whole step (15 GB of weight
// traffic) ago, so the first access pays a full page-table walk under load (~4 us, measured as a stall of the load
// ISSUE in the attention phase).  Issued from one thread while the CTA streams the MLP weights, the walk is free.
__device__ __forceinline__ void touch_kv_pages(const MegaParams& p, int layer, int t) {
  const int KVH = p.cfg.n_kv_heads, kvd = KVH * HD;
  const int ctx = p.S + t;
  const int n_splits = (ctx + ATT_CHUNK - 1) / ATT_CHUNK;
  const int item = blockIdx.x;
  if (layer >= p.cfg.n_layers || item >= p.rows * KVH * n_splits) return;
  const int split = item % n_splits, kvh = (item / n_splits) % KVH, row = item / (n_splits * KVH);
  const int input = row / p.beams;
  const int64_t n_prompt = (int64_t)(p.rows / p.beams) * p.S, n_gen = (int64_t)p.rows * p.max_gen;
  const int pos0 = split * ATT_CHUNK, pos1 = min(ctx - 1, pos0 + ATT_CHUNK - 1);
#pragma unroll
  for (int kv = 0; kv < 2; ++kv) {
    const bf16* pp = p.kv_prompt + ((int64_t)layer * 2 + kv) * n_prompt * kvd;
    const bf16* pg = p.kv_gen + ((int64_t)layer * 2 + kv) * n_gen * kvd;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int pos = e ? pos1 : pos0;
      const bf16* a = pos < p.S ? pp + ((int64_t)input * p.S + pos) * kvd + kvh * HD
                                : pg + ((int64_t)row * p.max_gen + (pos - p.S)) * kvd + kvh * HD;
      unsigned int dummy;
      asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(dummy) : "l"(a) : "memory");
    }
  }
}

__device__ __forceinline__ PhaseW phase_weights(const MegaParams& p, const LlamaLayerPtrs* s_layers, int ph) {
  const pcy_llama_config& c = p.cfg;
  const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, KVH = c.n_kv_heads;
  if (ph == 4 * c.n_layers) return PhaseW{p.lm_head, d, c.vocab, d, EPI_FP32};
  const LlamaLayerPtrs& y = s_layers[ph >> 2];
  switch (ph & 3) {
    case 0: return PhaseW{y.wqkv, d, (H + 2 * KVH) * HD, d, EPI_BF16};
    case 1: return PhaseW{y.wo, H * HD, d, H * HD, EPI_RESIDUAL};
    case 2: return PhaseW{y.wgu, d, f, d, EPI_SWIGLU};
    default: return PhaseW{y.wdown, f, d, f, EPI_RESIDUAL};
  }
}

template <int MT, int GQ, bool SELF>
__global__ void __launch_bounds__(SELF ? MK_THREADS : MK_BLOCK, 1)
llama_decode_megakernel(const MegaParams p) {
  extern __shared__ __align__(128) uint8_t mk_smem[];
  const pcy_llama_config& c = p.cfg;
  const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, KVH = c.n_kv_heads;
  const int qkv_dim = (H + 2 * KVH) * HD;
  const int kmax = f > d ? f : d;
  const int ns = p.ring_slots;
  // layout: [ring: ns slots][barriers: 2 ns x 8 B, padded to 1 KB][activations | attention scratch][out][red]
  const uint32_t ring = smem_u32(mk_smem);
  const uint32_t bars = ring + (uint32_t)ns * SLOT_BYTES;
  uint8_t* work = mk_smem + (size_t)ns * SLOT_BYTES + 1024;
  // the per-layer weight pointers, copied once so that no phase starts with a dependent global load
  LlamaLayerPtrs* s_layers = reinterpret_cast<LlamaLayerPtrs*>(mk_smem + p.layers_off);
  for (int i = threadIdx.x; i < c.n_layers * (int)(sizeof(LlamaLayerPtrs) / 8); i += blockDim.x)
    reinterpret_cast<uint64_t*>(s_layers)[i] = reinterpret_cast<const uint64_t*>(p.layers)[i];

  if (threadIdx.x == 0) {
    for (int i = 0; i < ns; ++i) {
      mbar_init(bars + 8u * i, 1);         // full: the producer's arrive.expect_tx (+ the copy's bytes)
      mbar_init(bars + 8u * (ns + i), 1);  // empty: lane 0 of the consuming warp
    }
    fence_barrier_init();
  }
  __syncthreads();  // the only CTA-wide barrier: after it the producer warp goes its own way

  if (!SELF && threadIdx.x >= MK_THREADS) {
    // ===================== producer warp: every weight byte of this CTA, in phase order =====================
    if (threadIdx.x < MK_THREADS + PL) {
      const RingGeom rg{ns, p.ring_magic};
      const uint64_t pol = l2_evict_first_policy();  // weights are read once per step: keep L2 for KV / activations
      const Share sh = cta_share(p.shares);
      uint32_t g = 0;
      for (int l = 0; l < c.n_layers; ++l) {
        const LlamaLayerPtrs& y = s_layers[l];
        produce_phase(ring, bars, rg, g, p.window, sh, y.wqkv, d, qkv_dim, d, EPI_BF16, pol);
        produce_phase(ring, bars, rg, g, p.window, sh, y.wo, H * HD, d, H * HD, EPI_RESIDUAL, pol);
        produce_phase(ring, bars, rg, g, p.window, sh, y.wgu, d, f, d, EPI_SWIGLU, pol);
        produce_phase(ring, bars, rg, g, p.window, sh, y.wdown, f, d, f, EPI_RESIDUAL, pol);
      }
      produce_phase(ring, bars, rg, g, p.window, sh, p.lm_head, d, c.vocab, d, EPI_FP32, pol);
    }
    return;
  }

  Smem sm;
  sm.a = reinterpret_cast<bf16*>(work);
  sm.out = reinterpret_cast<float*>(work + (size_t)MT * kmax * 2);
  sm.red = sm.out + (size_t)p.out_rows * MT;
  sm.share = cta_share(p.shares);
  sm.ring = ring; sm.bars = bars; sm.rg = RingGeom{ns, p.ring_magic}; sm.chunk0 = 0;
  sm.tbuf = p.timing;
  sm.tix = 0;
  sm.phase_ix = 0;
  sm.skew = p.timing != nullptr && p.timing[4095] == 0x534B4557ull;
  uint8_t* att_smem = work;  // the attention phase reuses the activation staging area
  const int n_phases = 4 * c.n_layers + 1;
  const uint64_t fill_pol = l2_evict_first_policy();  // weights are read once per step: keep L2 for KV / activations
  if (SELF) {
    // this warp's first chunks: ring positions warp, warp + 12, ... < ring_slots
    sm.f_ph = 0;
    sm.f_c = (int)(threadIdx.x >> 5);
    sm.f_pos = threadIdx.x >> 5;
    fill_load_phase(sm, p, s_layers);
    fill_normalise(sm, p, s_layers, n_phases);
    for (int i = 0; i < ns / MK_WARPS; ++i) fill_next(sm, p, s_layers, n_phases, fill_pol);
  }

  GridBarrier bar{p.barrier, 0u, gridDim.x};
  const int t = p.state[0];
  if (threadIdx.x == 32) touch_kv_pages(p, 0, t);
  sm.stamp();

  // One flat loop over the 4 L + 1 weight phases, so that gemv_phase and attention_phase each have a single call site
  // and can be inlined: as out-of-line functions they took the kernel parameters and `sm` by reference, i.e. through
  // per-thread copies in local memory (~500 B x 416 threads against the ~28 KB of L1 left beside the ring), and every
  // phase began with chains of local loads served by L2.
  const int n_att_items = p.rows * KVH * ((p.S + t + ATT_CHUNK - 1) / ATT_CHUNK);
  AttnLoads att_loads;
  att_loads.vmask = 0;
  for (int ph = 0; ph < n_phases; ++ph) {
    const int l = ph >> 2;
    const int kind = ph == n_phases - 1 ? 4 : (ph & 3);  // 0 qkv | 1 o_proj | 2 gate/up | 3 down | 4 lm head
    const LlamaLayerPtrs& y = s_layers[kind == 4 ? 0 : l];
    int n_out, K, stage, epi, rows = p.rows;
    const bf16 *A, *rms_w = nullptr;
    int64_t lda, ldo;
    void* out;
    if (kind == 0) {         // qkv = Wqkv . rms(x)
      n_out = qkv_dim; K = d; stage = STAGE_RMS; A = p.x; lda = d; rms_w = y.ln1; epi = EPI_BF16; out = p.qkv; ldo = qkv_dim;
      if (l == 0) {
        // the residual stream starts as the embedding of the last token of every row: each CTA mirrors its column
        // slice of the rows into x
        for (int m = 0; m < p.rows; ++m) {
          const int tok = p.tokens[(int64_t)m * p.max_gen + (t - 1)];
          const bf16* src = p.embed + (int64_t)tok * d;
          int clo, chi;
          cta_range(sm.share, d / 8, clo, chi);
          for (int k8 = clo + threadIdx.x; k8 < chi; k8 += MK_THREADS)
            *reinterpret_cast<uint4*>(p.x + (int64_t)m * d + k8 * 8) = *reinterpret_cast<const uint4*>(src + k8 * 8);
        }
        if (p.rows == 1) A = p.embed + (int64_t)p.tokens[t - 1] * d;  // straight from the table (x is not visible yet)
        else bar.sync();
      }
    } else if (kind == 1) {  // attention, then x += Wo . attn
      attention_phase<GQ>(p, att_smem, l, t, bar, sm, att_loads);
      sm.stamp();
      bar.sync();
      sm.stamp();
      n_out = d; K = H * HD; stage = STAGE_PLAIN; A = p.attn; lda = H * HD; epi = EPI_RESIDUAL; out = p.x; ldo = d;
    } else if (kind == 2) {  // act = silu(Wg . rms(x)) * (Wu . rms(x))
      if (threadIdx.x == 32) touch_kv_pages(p, l + 1, t);
      n_out = f; K = d; stage = STAGE_RMS; A = p.x; lda = d; rms_w = y.ln2; epi = EPI_SWIGLU; out = p.act; ldo = f;
    } else if (kind == 3) {  // x += Wdown . act
      n_out = d; K = f; stage = STAGE_PLAIN; A = p.act; lda = f; epi = EPI_RESIDUAL; out = p.x; ldo = d;
    } else {                 // logits = Wlm . rms(x)
      n_out = c.vocab; K = d; stage = STAGE_RMS; A = p.x; lda = d; rms_w = p.norm; epi = EPI_FP32; out = p.logits;
      ldo = c.vocab;
    }
    gemv_phase<MT, SELF>(sm, p, s_layers, n_phases, fill_pol, n_out, K, stage, A, lda, rows, rms_w, c.rms_eps, epi, out, ldo, [&]() {
      // the K / V rows of this layer's attention item: requested while the qkv weights stream
      if (kind == 0 && (int)blockIdx.x < n_att_items) att_loads.request(p, l, t, blockIdx.x);
    });
    if (kind == 4) {
      consumer_sync();
      sm.stamp();
    } else if (kind != 0) {  // (the barrier after the qkv phase is inside attention_phase)
      sm.stamp();
      bar.sync();
      sm.stamp();
    }
  }
}

unsigned long long* g_timing = nullptr;
bool g_self_refill = [] {
  const char* e = getenv("PCY_DECODE_SELF_REFILL");
  return e && atoi(e) != 0;
}();

constexpr size_t SMEM_LIMIT = 227 * 1024;
constexpr double MAX_SHARE = 1.15;
uint64_t* g_shares = nullptr;  // device table, see decode_megakernel_set_shares

// rows of the widest per-CTA slice (SwiGLU phases hold a gate and an up row per output)
int out_rows_for(const pcy_llama_config& c) {
  const int sms = num_sms();
  const int qkv = (c.n_heads + 2 * c.n_kv_heads) * HD;
  int r = ceil_div(c.vocab, sms) + 1;
  r = std::max(r, 2 * (ceil_div(c.ffn_dim, sms) + 1));
  r = std::max(r, ceil_div(qkv, sms) + 1);
  r = std::max(r, ceil_div(c.d_model, sms) + 1);
  // head-room for measured (unequal) slices: shares are capped at MAX_SHARE x the equal slice
  return (int)round_up((int)(r * MAX_SHARE) + 2, 64);
}
// bytes of the non-ring part of shared memory: barriers + max(gemv staging, attention scratch)
size_t work_smem_bytes(const pcy_llama_config& c, int mt) {
  const int kmax = std::max(std::max(c.ffn_dim, c.d_model), c.n_heads * HD);
  const size_t smem_gemv = (size_t)mt * kmax * 2 + (size_t)out_rows_for(c) * mt * 4 + MK_WARPS * 4 * 4;
  const size_t smem_att = (size_t)(4 * HD + 2 * HD + 4 * ATT_CHUNK + 8 + 4 + MK_WARPS * 4 * HD) * 4 + 64;
  return 1024 + (smem_gemv > smem_att ? smem_gemv : smem_att);
}
size_t layer_table_bytes(const pcy_llama_config& c) { return round_up((size_t)c.n_layers * sizeof(LlamaLayerPtrs), 16); }
// slots: a multiple of 12 (slot s is owned by consumer warp s % 12), at most 60 (2 x 60 barriers fit the 1 KB block)
int ring_slots_for(const pcy_llama_config& c, int mt) {
  const size_t w = work_smem_bytes(c, mt) + layer_table_bytes(c);
  if (w + MK_WARPS * SLOT_BYTES > SMEM_LIMIT) return 0;
  const int ns = (int)((SMEM_LIMIT - w) / SLOT_BYTES) / MK_WARPS * MK_WARPS;
  return ns > 60 ? 60 : ns;
}

}  // namespace

void decode_megakernel_set_timing(unsigned long long* dev_buf) { g_timing = dev_buf; }
void decode_megakernel_set_self_refill(int enabled) { g_self_refill = enabled != 0; }

int decode_megakernel_set_shares(const float* shares, int n) {
  if (shares == nullptr || n == 0) {  // back to equal slices
    g_shares = nullptr;               // (the table stays allocated: a launch in flight may still read it)
    return 0;
  }
  PCY_REQUIRE(n == num_sms(), "decode shares: %d entries for %d SMs", n, num_sms());
  std::vector<double> f(n);
  double tot = 0.0;
  for (int i = 0; i < n; ++i) {
    PCY_REQUIRE(shares[i] > 0.f, "decode shares: entry %d is not positive", i);
    tot += shares[i];
  }
  // clamp to [2 - MAX_SHARE, MAX_SHARE] x the equal slice (the per-CTA partial-sum buffer is sized for MAX_SHARE)
  double used = 0.0;
  for (int i = 0; i < n; ++i) {
    f[i] = std::min(std::max(shares[i] / tot * n, 2.0 - MAX_SHARE), MAX_SHARE - 0.01);
    used += f[i];
  }
  std::vector<uint64_t> cum(n + 1);
  double run = 0.0;
  for (int i = 0; i < n; ++i) {
    cum[i] = (uint64_t)(run / used * 4294967296.0);
    run += f[i];
  }
  cum[0] = 0;
  cum[n] = 1ull << 32;
  for (int i = 0; i < n; ++i)  // renormalisation must not push a slice over the cap
    PCY_REQUIRE((double)(cum[i + 1] - cum[i]) / 4294967296.0 * n < MAX_SHARE, "decode shares: slice %d too large", i);
  static uint64_t* dev = nullptr;
  if (dev == nullptr) PCY_CUDA(cudaMalloc(&dev, (size_t)(n + 1) * sizeof(uint64_t)));
  PCY_CUDA(cudaMemcpy(dev, cum.data(), (size_t)(n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
  g_shares = dev;
  return 0;
}

int64_t decode_megakernel_scratch_bytes(const pcy_llama_config& c, int rows, int S, int max_gen) {
  const int64_t d = c.d_model, qkv = (int64_t)(c.n_heads + 2 * c.n_kv_heads) * HD;
  const int max_splits = ceil_div(S + max_gen, ATT_CHUNK);
  int64_t b = 512 + round_up(rows * d * 2, 256) * 2 + round_up(rows * qkv * 2, 256) +
              round_up((int64_t)rows * c.n_heads * HD * 2, 256) + round_up((int64_t)rows * c.ffn_dim * 2, 256);
  b += round_up((int64_t)rows * c.n_kv_heads * max_splits * (c.n_heads / c.n_kv_heads) * PSTR * 4, 256);
  b += 1024;
  return b;
}

bool decode_megakernel_supported(const pcy_llama_config& c, int rows) {
  const int gq = c.n_heads / c.n_kv_heads;
  if (rows < 1 || rows > 4 || c.head_dim != HD || gq != 4) return false;
  if (c.d_model % 256 != 0 || c.ffn_dim % 256 != 0) return false;
  if (rows * c.n_kv_heads > 96) return false;  // attention ticket counters live in the 512-byte barrier block
  const int mt = rows <= 1 ? 1 : rows <= 2 ? 2 : 4;
  return ring_slots_for(c, mt) >= MK_WARPS;
}

int decode_megakernel(const pcy_llama_config& c, const LlamaLayerPtrs* layers_dev, const bf16* embed,
                      const bf16* lm_head, const bf16* norm, const float* rope, const pcy_decode_buffers* b,
                      void* scratch, cudaStream_t stream) {
  const int rows = b->n_inputs * b->beams;
  PCY_REQUIRE(decode_megakernel_supported(c, rows), "decode megakernel: unsupported configuration");
  MegaParams p;
  p.cfg = c; p.embed = embed; p.lm_head = lm_head; p.norm = norm; p.layers = layers_dev; p.rope = rope;
  p.rows = rows; p.beams = b->beams; p.S = b->S; p.max_gen = b->max_gen;
  p.kv_prompt = reinterpret_cast<const bf16*>(b->kv_prompt); p.prompt_valid = b->prompt_valid;
  p.kv_gen = reinterpret_cast<bf16*>(b->kv_gen); p.tokens = b->tokens; p.slots = b->slots; p.state = b->state;
  p.logits = b->logits_cur;
  const int64_t d = c.d_model, qkv = (int64_t)(c.n_heads + 2 * c.n_kv_heads) * HD;
  uint8_t* s = reinterpret_cast<uint8_t*>(round_up(reinterpret_cast<int64_t>(scratch), 256));
  auto carve = [&](int64_t bytes) { uint8_t* r = s; s += round_up(bytes, 256); return r; };
  p.barrier = reinterpret_cast<unsigned int*>(carve(512));
  p.x = reinterpret_cast<bf16*>(carve(rows * d * 2));
  carve(rows * d * 2);
  p.qkv = reinterpret_cast<bf16*>(carve(rows * qkv * 2));
  p.attn = reinterpret_cast<bf16*>(carve((int64_t)rows * c.n_heads * HD * 2));
  p.act = reinterpret_cast<bf16*>(carve((int64_t)rows * c.ffn_dim * 2));
  p.max_splits = ceil_div(b->S + b->max_gen, ATT_CHUNK);
  p.part = reinterpret_cast<float*>(s);
  p.timing = g_timing;
  p.shares = g_shares;
  PCY_CUDA(cudaMemsetAsync(p.barrier, 0, 512, stream));  // barrier counter + tickets

  const int mt = rows <= 1 ? 1 : rows <= 2 ? 2 : 4;
  p.ring_slots = ring_slots_for(c, mt);
  p.out_rows = out_rows_for(c);
  p.ring_magic = (uint32_t)(((1ull << 32) + p.ring_slots - 1) / p.ring_slots);
  static const int window_env = [] {
    const char* e = getenv("PCY_DECODE_WINDOW");  // tuning knob: weight chunks (8 KB) in flight per SM
    return e ? atoi(e) : 24;
  }();
  p.window = std::min(std::max(window_env, PL), p.ring_slots);
  p.layers_off = (int)((size_t)p.ring_slots * SLOT_BYTES + work_smem_bytes(c, mt));
  const size_t smem = (size_t)p.layers_off + layer_table_bytes(c);
  PCY_REQUIRE(smem <= SMEM_LIMIT, "decode megakernel: needs %zu bytes of shared memory", smem);
  // pcy_set_decode_self_refill(1) / PCY_DECODE_SELF_REFILL=1: no producer warp (12 warps, 168 registers, no spills; every warp refills its own ring
  // slots).  Measured on B200: 3.049 ms per token against 3.047 ms with the producer warp - what this kernel loses is
  // arrival skew at the grid barriers (scripts/profile_decode_phases.py), not local-memory traffic, unlike the beam
  // kernel, where the same change was worth 9 %.  Kept as an option, parity-tested (tests/test_gpu_llama.py).
  const bool self_refill = g_self_refill;
  void* fn = nullptr;
  if (self_refill) {
    if (mt == 1) fn = (void*)llama_decode_megakernel<1, 4, true>;
    else if (mt == 2) fn = (void*)llama_decode_megakernel<2, 4, true>;
    else fn = (void*)llama_decode_megakernel<4, 4, true>;
  } else {
    if (mt == 1) fn = (void*)llama_decode_megakernel<1, 4, false>;
    else if (mt == 2) fn = (void*)llama_decode_megakernel<2, 4, false>;
    else fn = (void*)llama_decode_megakernel<4, 4, false>;
  }
  static SmemOptIn opt[6];
  const int slot = (mt == 1 ? 0 : mt == 2 ? 1 : 2) + (self_refill ? 3 : 0);
  if (opt[slot].need(smem)) PCY_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&p};
  // cooperative launch: guarantees that all CTAs are co-resident (the grid barrier needs it)
  PCY_CUDA(cudaLaunchCooperativeKernel(fn, dim3(num_sms()), dim3(self_refill ? MK_THREADS : MK_BLOCK), args, smem,
                                       stream));
  count_launch();
  return 0;
}

}  // namespace pcy
