// One persistent kernel per Llama decode step (rows = n_inputs * beams <= 4): one CTA per SM streams its balanced
// slice of every weight matrix exactly once, phases are separated by a grid-wide barrier, activations (a few KB)
// bounce through L2.  Per layer:  P1 qkv = Wqkv . rms(x)  |  P2 RoPE + KV append + split-KV attention partials
// |  P2c combine  |  P3 x += Wo . attn  |  P4 act = silu(Wg . rms(x)) * (Wu . rms(x))  |  P5 x += Wdown . act,
// then logits = Wlm . rms(x).  Batch-1 decode is pure weight streaming (15 GB per token for Llama-3-8B): what this
// design buys over one launch per op is no launch / ramp-up / tail per op (160 of them per token) and L2 prefetch
// of the next phase's first weight rows while CTAs wait at the barrier.
//
// Replaces the per-token HF LlamaForCausalLM forward of the reference's generate loops
// (procyon/model/model_unified.py:769, :887 -> procyon/model/pmc_llama.py:581).
#include "common.cuh"
#include "ops.h"

namespace pcy {

namespace {

constexpr int MK_THREADS = 512;
constexpr int MK_WARPS = MK_THREADS / 32;
constexpr int UL = 8;          // 512-byte weight segments in flight per warp (one 16-byte load per lane each)
constexpr int HD = 128;
constexpr int ATT_CHUNK = 128;  // keys per attention work item (8 per warp)
constexpr int PSTR = HD + 4;    // floats per (split, head) attention partial: 128 outputs, max, sum (16-byte rows)
constexpr int MAX_OUT_PER_CTA = 1024;

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// activations written by other CTAs earlier in this kernel must be read through L2 (L1 is not coherent)
__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

struct GridBarrier {
  unsigned int* counter;
  unsigned int target;
  unsigned int nblocks;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += nblocks;
      __threadfence();
      atomicAdd(counter, 1u);
      uint64_t t0 = 0;
      for (uint32_t it = 0; ld_acquire_u32(counter) < target; ++it) {
        if ((it & 0x3fffu) == 0x3fffu) {  // bounded: a scheduling bug must trap, not hang the GPU
          const uint64_t now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000ull) __trap();
        }
      }
      __threadfence();
    }
    __syncthreads();
  }
};

__device__ __forceinline__ float dot8f(const uint4& a, const uint4& w, float s) {
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  const float2 w0 = unpack_bf16x2(w.x), w1 = unpack_bf16x2(w.y), w2 = unpack_bf16x2(w.z), w3 = unpack_bf16x2(w.w);
  s = fmaf(a0.x, w0.x, s); s = fmaf(a0.y, w0.y, s);
  s = fmaf(a1.x, w1.x, s); s = fmaf(a1.y, w1.y, s);
  s = fmaf(a2.x, w2.x, s); s = fmaf(a2.y, w2.y, s);
  s = fmaf(a3.x, w3.x, s); s = fmaf(a3.y, w3.y, s);
  return s;
}

enum : int { EPI_BF16 = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_FP32 = 3 };
enum : int { STAGE_PLAIN = 0, STAGE_RMS = 1, STAGE_ATTN = 2 };

struct Smem {
  bf16* a;       // [MT][K] staged activations
  float* out;    // [MAX_OUT_PER_CTA * 2][MT] partial sums
  float* red;    // [MK_WARPS * 4] scratch
  unsigned long long* tbuf;  // profiling stamps (CTA 0, thread 0) or null
  int* tix;
  __device__ __forceinline__ void stamp() const {
    if (tbuf != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tbuf[*tix] = globaltimer_ns();
    if (tbuf != nullptr) ++*tix;
  }
};

// balanced contiguous range of `n` items for this CTA
__device__ __forceinline__ void cta_range(int n, int& lo, int& hi) {
  lo = (int)(((int64_t)n * blockIdx.x) / gridDim.x);
  hi = (int)(((int64_t)n * (blockIdx.x + 1)) / gridDim.x);
}

// weight row of local row r (within the CTA's range starting at logical output o_lo)
__device__ __forceinline__ int weight_row(int epi, int o_lo, int r) {
  if (epi != EPI_SWIGLU) return o_lo + r;
  const int j = o_lo + (r >> 1);  // logical output; even local rows = gate, odd = up
  return (j >> 4) * 32 + (j & 15) + ((r & 1) ? 16 : 0);
}

// L2 prefetch of this CTA's slice of a coming phase: lines [line0, line0 + n_lines) of the slice, issued before the
// grid barrier so that HBM keeps streaming while CTAs wait and stage activations.
__device__ void prefetch_phase(const bf16* W, int64_t ldw, int n_out, int K, int epi, int line0, int n_lines) {
  int lo, hi;
  cta_range(n_out, lo, hi);
  const int rpo = (epi == EPI_SWIGLU) ? 2 : 1;
  const int n_rows = (hi - lo) * rpo;
  const int lines_per_row = (K * 2) / 128;
  const int total = n_rows * lines_per_row;
  const int end = min(total, line0 + n_lines);
  for (int i = line0 + threadIdx.x; i < end; i += MK_THREADS) {
    const int r = i / lines_per_row, l = i % lines_per_row;
    prefetch_l2(W + (int64_t)weight_row(epi, lo, r) * ldw + l * 64);
  }
}

struct AttnSrc {  // STAGE_ATTN: A[m][head*128 + dim] = merge over splits of the attention partials
  const float* part;
  int n_splits, max_splits, KVH, GQ;
};

// out = epi(W[n_out(x2), K] . A[MT, K]).  The CTA's slice of W is treated as a flat list of 512-byte segments that is
// divided evenly among the 16 warps (a warp's range may start and end inside a row); the first batch of weight loads
// is issued before the activations are staged, so its latency overlaps the staging.
template <int MT>
__device__ void gemv_phase(const Smem& sm, const bf16* __restrict__ W, int64_t ldw, int n_out, int K, int stage,
                           const bf16* A, int64_t lda, const AttnSrc& asrc, int rows, const bf16* __restrict__ rms_w,
                           float eps, int epi, void* out, int64_t ldo) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int o_lo, o_hi;
  cta_range(n_out, o_lo, o_hi);
  const int rpo = (epi == EPI_SWIGLU) ? 2 : 1;
  const int n_rows = (o_hi - o_lo) * rpo;
  // work unit = (local row, K chunk of 2048 elements): 8 x 16-byte loads per lane; units go round-robin to warps
  constexpr int KCH = UL * 256;
  const int kc_per_row = (K + KCH - 1) / KCH;
  const int n_units = n_rows * kc_per_row;
  uint4 wb[UL];
#define PCY_LOAD_UNIT(U)                                                                     \
  {                                                                                          \
    const int r_ = (U) / kc_per_row, kc_ = (U) - r_ * kc_per_row;                            \
    const bf16* wrow_ = W + (int64_t)weight_row(epi, o_lo, r_) * ldw + kc_ * KCH;            \
    const int k_rem_ = K - kc_ * KCH;                                                        \
    _Pragma("unroll") for (int j = 0; j < UL; ++j) {                                         \
      const int k_ = j * 256 + lane * 8;                                                     \
      wb[j] = (k_ < k_rem_) ? ldg_nc_v4(wrow_ + k_) : make_uint4(0, 0, 0, 0);                \
    }                                                                                        \
  }

  // ---- stage A (plain | RMS-normalised with HF rounding | merged attention partials) ----
  for (int m = 0; m < MT; ++m) {
    bf16* dst = sm.a + (int64_t)m * K;
    if (m >= rows) {
      for (int k = tid * 8; k < K; k += MK_THREADS * 8) *reinterpret_cast<uint4*>(dst + k) = make_uint4(0, 0, 0, 0);
      continue;
    }
    if (stage == STAGE_ATTN) {
      for (int k = tid * 8; k < K; k += MK_THREADS * 8) {
        const int head = k / HD, dim = k % HD;
        const int kvh = head / asrc.GQ, hq = head % asrc.GQ;
        const float* ps = asrc.part + (((int64_t)m * asrc.KVH + kvh) * asrc.max_splits) * asrc.GQ * PSTR + hq * PSTR;
        const int64_t stride = (int64_t)asrc.GQ * PSTR;
        float mx = -INFINITY;
        for (int sp = 0; sp < asrc.n_splits; ++sp) mx = fmaxf(mx, __ldcg(ps + sp * stride + HD));
        float l = 0.f, acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int sp = 0; sp < asrc.n_splits; ++sp) {
          const float ms = __ldcg(ps + sp * stride + HD);
          if (ms == -INFINITY) continue;
          const float w = exp2f(ms - mx);
          l += w * __ldcg(ps + sp * stride + HD + 1);
          const float4 v0 = __ldcg(reinterpret_cast<const float4*>(ps + sp * stride + dim));
          const float4 v1 = __ldcg(reinterpret_cast<const float4*>(ps + sp * stride + dim + 4));
          acc[0] += w * v0.x; acc[1] += w * v0.y; acc[2] += w * v0.z; acc[3] += w * v0.w;
          acc[4] += w * v1.x; acc[5] += w * v1.y; acc[6] += w * v1.z; acc[7] += w * v1.w;
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        *reinterpret_cast<uint4*>(dst + k) =
            make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                       pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
      }
      continue;
    }
    const bf16* src = A + (int64_t)m * lda;
    float rstd = 1.f;
    if (stage == STAGE_RMS) {
      float ss = 0.f;
      for (int k = tid * 8; k < K; k += MK_THREADS * 8) {
        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(src + k));
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
      }
      ss = warp_sum(ss);
      if (lane == 0) sm.red[warp] = ss;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < MK_WARPS; ++w) tot += sm.red[w];
      rstd = rsqrtf(tot / (float)K + eps);
      __syncthreads();
    }
    for (int k = tid * 8; k < K; k += MK_THREADS * 8) {
      const uint4 u = __ldcg(reinterpret_cast<const uint4*>(src + k));
      if (stage == STAGE_RMS) {
        const uint4 g = *reinterpret_cast<const uint4*>(rms_w + k);
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
      } else {
        *reinterpret_cast<uint4*>(dst + k) = u;
      }
    }
  }
  for (int i = tid; i < n_rows * MT; i += MK_THREADS) sm.out[i] = 0.f;
  // residual values are final since two phases ago: fetch them now, off the critical path
  const int n_o = o_hi - o_lo;
  float res_pre = 0.f;
  if (epi == EPI_RESIDUAL && tid < n_o * MT) {
    const int o = tid / MT, m = tid % MT;
    if (m < rows) res_pre = ldcg_bf16(reinterpret_cast<const bf16*>(out) + (int64_t)m * ldo + o_lo + o);
  }
  __syncthreads();
  sm.stamp();

  // ---- stream the weights ----
  for (int u = warp; u < n_units; u += MK_WARPS) {
    PCY_LOAD_UNIT(u)
    const int r = u / kc_per_row, kc = u - r * kc_per_row;
    const int k_rem = K - kc * KCH;
    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.f;
#pragma unroll
    for (int j = 0; j < UL; ++j) {
      const int k = j * 256 + lane * 8;
      if (k < k_rem) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint4 a = *reinterpret_cast<const uint4*>(sm.a + (int64_t)m * K + kc * KCH + k);
          acc[m] = dot8f(a, wb[j], acc[m]);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const float v = warp_sum(acc[m]);
      if (lane == 0) atomicAdd(&sm.out[r * MT + m], v);
    }
  }
#undef PCY_LOAD_UNIT
  __syncthreads();
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
  bf16* act;    // [rows][ffn]
  float* part;  // [rows][KVH][max_splits][GQ][PSTR]
  int max_splits;
  unsigned int* barrier;
  unsigned long long* timing;  // optional: globaltimer at every phase boundary (CTA 0), for profiling
};

// P2: one work item = (row, kv head, split of 256 keys): RoPE(q, k_new), KV append, scores, softmax, P.V -> partial.
// The partials are merged by the next phase while it stages its activations (no combine pass, no extra barrier).
template <int GQ>
__device__ void attention_items(const MegaParams& p, uint8_t* smem_raw, int layer) {
  float* s_q = reinterpret_cast<float*>(smem_raw);  // [GQ][HD]
  float* s_knew = s_q + GQ * HD;                    // [HD]
  float* s_vnew = s_knew + HD;                      // [HD]
  float* s_sc = s_vnew + HD;                        // [GQ][ATT_CHUNK]
  float* s_ml = s_sc + GQ * ATT_CHUNK;              // [GQ][2]
  float* s_po = s_ml + GQ * 2;                      // [MK_WARPS][GQ][HD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.cfg.n_heads, KVH = p.cfg.n_kv_heads;
  const int kvd = KVH * HD, qkv_dim = (H + 2 * KVH) * HD;
  const int t = p.state[0];
  const int g_cur = t - 1, pos_cur = p.S + g_cur, ctx = pos_cur + 1;
  const int n_splits = (ctx + ATT_CHUNK - 1) / ATT_CHUNK;
  const int n_items = p.rows * KVH * n_splits;
  const float scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
  const int64_t n_prompt = (int64_t)(p.rows / p.beams) * p.S, n_gen = (int64_t)p.rows * p.max_gen;
  const bf16* kp = p.kv_prompt + ((int64_t)layer * 2 + 0) * n_prompt * kvd;
  const bf16* vp = p.kv_prompt + ((int64_t)layer * 2 + 1) * n_prompt * kvd;
  bf16* kg = p.kv_gen + ((int64_t)layer * 2 + 0) * n_gen * kvd;
  bf16* vg = p.kv_gen + ((int64_t)layer * 2 + 1) * n_gen * kvd;
  const int sub = lane >> 3, l8 = lane & 7;  // 8 lanes per key, 16 dims per lane
  constexpr int KPW = ATT_CHUNK / MK_WARPS / 4;  // key iterations per warp (4 keys each)

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int split = item % n_splits, kvh = (item / n_splits) % KVH, row = item / (n_splits * KVH);
    const int input = row / p.beams;
    const bf16* qkv_row = p.qkv + (int64_t)row * qkv_dim;
    const int k0 = split * ATT_CHUNK;
    const int n_keys = min(ctx, k0 + ATT_CHUNK) - k0;
    // K/V row pointers of this lane's keys (nullptr = the current token, held in smem; or out of range)
    const bf16* kptr[KPW];
    const bf16* vptr[KPW];
    bool valid[KPW];
    uint4 kreg[KPW][2];
#pragma unroll
    for (int it = 0; it < KPW; ++it) {
      const int kk = (it * MK_WARPS + warp) * 4 + sub;
      const int pos = k0 + kk;
      kptr[it] = vptr[it] = nullptr;
      valid[it] = kk < n_keys;
      if (valid[it] && pos != pos_cur) {
        if (pos < p.S) {
          const int64_t base = ((int64_t)input * p.S + pos) * kvd + kvh * HD + l8 * 16;
          kptr[it] = kp + base;
          vptr[it] = vp + base;
          if (p.prompt_valid) valid[it] = p.prompt_valid[(int64_t)input * p.S + pos] != 0;
        } else {
          const int g = pos - p.S;
          const int prow = p.slots[(int64_t)row * p.max_gen + g];
          const int64_t base = ((int64_t)prow * p.max_gen + g) * kvd + kvh * HD + l8 * 16;
          kptr[it] = kg + base;
          vptr[it] = vg + base;
        }
      }
      if (kptr[it] != nullptr) {  // issue all K loads before touching shared memory
        kreg[it][0] = __ldcg(reinterpret_cast<const uint4*>(kptr[it]));
        kreg[it][1] = __ldcg(reinterpret_cast<const uint4*>(kptr[it] + 8));
      }
    }
    __syncthreads();  // previous item done with the shared buffers
    {
      const float2* cs = reinterpret_cast<const float2*>(p.rope) + (int64_t)pos_cur * (HD / 2);
      for (int i = tid; i < (GQ + 1) * (HD / 2); i += MK_THREADS) {
        const int hh = i / (HD / 2), j = i % (HD / 2);
        const bf16* src = (hh < GQ) ? qkv_row + (kvh * GQ + hh) * HD : qkv_row + (H + kvh) * HD;
        const float lo = ldcg_bf16(src + j), hi = ldcg_bf16(src + j + HD / 2);
        const float2 c = cs[j];
        const float o_lo = bf16_round(lo * c.x - hi * c.y), o_hi = bf16_round(hi * c.x + lo * c.y);
        if (hh < GQ) {
          s_q[hh * HD + j] = o_lo * scale_log2;
          s_q[hh * HD + j + HD / 2] = o_hi * scale_log2;
        } else {
          s_knew[j] = o_lo;
          s_knew[j + HD / 2] = o_hi;
        }
      }
      if (tid < HD) s_vnew[tid] = ldcg_bf16(qkv_row + (H + KVH + kvh) * HD + tid);
    }
    __syncthreads();
    if (pos_cur >= k0 && pos_cur < k0 + ATT_CHUNK && tid < HD) {
      const int64_t off = ((int64_t)row * p.max_gen + g_cur) * kvd + kvh * HD + tid;
      kg[off] = __float2bfloat16_rn(s_knew[tid]);
      vg[off] = __float2bfloat16_rn(s_vnew[tid]);
    }
    // ---- scores ----
#pragma unroll
    for (int it = 0; it < KPW; ++it) {
      const int kk = (it * MK_WARPS + warp) * 4 + sub;
      float kf[16];
      if (kptr[it] != nullptr) {
        const uint32_t w[8] = {kreg[it][0].x, kreg[it][0].y, kreg[it][0].z, kreg[it][0].w,
                               kreg[it][1].x, kreg[it][1].y, kreg[it][1].z, kreg[it][1].w};
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
#pragma unroll
      for (int h = 0; h < GQ; ++h) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) a = fmaf(kf[j], s_q[h * HD + l8 * 16 + j], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (l8 == 0) s_sc[h * ATT_CHUNK + kk] = valid[it] ? a : -INFINITY;
      }
    }
    // V loads in flight while the softmax statistics are computed
    uint4 vreg[KPW][2];
#pragma unroll
    for (int it = 0; it < KPW; ++it) {
      if (vptr[it] != nullptr) {
        vreg[it][0] = __ldcg(reinterpret_cast<const uint4*>(vptr[it]));
        vreg[it][1] = __ldcg(reinterpret_cast<const uint4*>(vptr[it] + 8));
      }
    }
    __syncthreads();
    if (warp < GQ) {
      const int h = warp;
      float m = -INFINITY;
      for (int k = lane; k < ATT_CHUNK; k += 32) m = fmaxf(m, s_sc[h * ATT_CHUNK + k]);
      m = warp_max(m);
      float l = 0.f;
      for (int k = lane; k < ATT_CHUNK; k += 32) {
        const float pr = (m == -INFINITY) ? 0.f : exp2f(s_sc[h * ATT_CHUNK + k] - m);
        s_sc[h * ATT_CHUNK + k] = pr;
        l += pr;
      }
      l = warp_sum(l);
      if (lane == 0) { s_ml[h * 2] = m; s_ml[h * 2 + 1] = l; }
    }
    __syncthreads();
    // ---- P.V: two heads at a time to bound the accumulator registers ----
#pragma unroll
    for (int hp = 0; hp < GQ; hp += 2) {
      float acc[2][16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[0][j] = acc[1][j] = 0.f;
#pragma unroll
      for (int it = 0; it < KPW; ++it) {
        const int kk = (it * MK_WARPS + warp) * 4 + sub;
        float vf[16];
        if (vptr[it] != nullptr) {
          const uint32_t w[8] = {vreg[it][0].x, vreg[it][0].y, vreg[it][0].z, vreg[it][0].w,
                                 vreg[it][1].x, vreg[it][1].y, vreg[it][1].z, vreg[it][1].w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            vf[2 * j] = f.x;
            vf[2 * j + 1] = f.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) vf[j] = bf16_round(s_vnew[l8 * 16 + j]);
        }
        const float p0 = s_sc[hp * ATT_CHUNK + kk];
        const float p1 = (hp + 1 < GQ) ? s_sc[(hp + 1) * ATT_CHUNK + kk] : 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          acc[0][j] = fmaf(p0, vf[j], acc[0][j]);
          acc[1][j] = fmaf(p1, vf[j], acc[1][j]);
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (hp + q < GQ) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = acc[q][j];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (sub == 0) s_po[(warp * GQ + hp + q) * HD + l8 * 16 + j] = v;
          }
        }
      }
    }
    __syncthreads();
    float* part = p.part + (((int64_t)row * KVH + kvh) * p.max_splits + split) * GQ * PSTR;
    {
      const int h = tid / HD, dim = tid % HD;  // 512 threads = 4 heads x 128 dims
      for (int hh = h; hh < GQ; hh += MK_THREADS / HD) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < MK_WARPS; ++w) s += s_po[(w * GQ + hh) * HD + dim];
        part[hh * PSTR + dim] = s;
      }
      if (tid < GQ) {
        part[tid * PSTR + HD] = s_ml[tid * 2];
        part[tid * PSTR + HD + 1] = s_ml[tid * 2 + 1];
      }
    }
  }
}

template <int MT, int GQ>
__global__ void __launch_bounds__(MK_THREADS, 1)
llama_decode_megakernel(const MegaParams p) {
  extern __shared__ __align__(16) uint8_t mk_smem[];
  const pcy_llama_config& c = p.cfg;
  const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, KVH = c.n_kv_heads;
  const int qkv_dim = (H + 2 * KVH) * HD;
  const int kmax = f > d ? f : d;
  Smem sm;
  sm.a = reinterpret_cast<bf16*>(mk_smem);
  sm.out = reinterpret_cast<float*>(mk_smem + (size_t)MT * kmax * 2);
  sm.red = sm.out + MAX_OUT_PER_CTA * 2 * MT;
  uint8_t* att_smem = mk_smem;  // the attention phase reuses the activation staging area

  GridBarrier bar{p.barrier, 0u, gridDim.x};
  const int t = p.state[0];
  const int n_splits = (p.S + t + ATT_CHUNK - 1) / ATT_CHUNK;
  const AttnSrc no_attn{nullptr, 0, 0, 0, 0};
  const AttnSrc attn_src{p.part, n_splits, p.max_splits, KVH, GQ};
  int tix = 0;
  sm.tbuf = p.timing;
  sm.tix = &tix;
  auto stamp = [&]() { sm.stamp(); };
  stamp();

  for (int l = 0; l < c.n_layers; ++l) {
    const LlamaLayerPtrs& y = p.layers[l];
    // ---- P1: qkv = Wqkv . rms(x) ----
    if (l == 0) {
      // the residual stream starts as the embedding of the last token of every row: each CTA mirrors its column
      // slice of the rows into x
      for (int m = 0; m < p.rows; ++m) {
        const int tok = p.tokens[(int64_t)m * p.max_gen + (t - 1)];
        const bf16* src = p.embed + (int64_t)tok * d;
        int clo, chi;
        cta_range(d / 8, clo, chi);
        for (int k8 = clo + threadIdx.x; k8 < chi; k8 += MK_THREADS)
          *reinterpret_cast<uint4*>(p.x + (int64_t)m * d + k8 * 8) = *reinterpret_cast<const uint4*>(src + k8 * 8);
      }
      if (p.rows == 1) {  // read the row straight from the table (x is not globally visible yet)
        const int tok = p.tokens[t - 1];
        gemv_phase<MT>(sm, y.wqkv, d, qkv_dim, d, STAGE_RMS, p.embed + (int64_t)tok * d, d, no_attn, 1, y.ln1,
                       c.rms_eps, EPI_BF16, p.qkv, qkv_dim);
      } else {
        bar.sync();
        gemv_phase<MT>(sm, y.wqkv, d, qkv_dim, d, STAGE_RMS, p.x, d, no_attn, p.rows, y.ln1, c.rms_eps, EPI_BF16,
                       p.qkv, qkv_dim);
      }
    } else {
      gemv_phase<MT>(sm, y.wqkv, d, qkv_dim, d, STAGE_RMS, p.x, d, no_attn, p.rows, y.ln1, c.rms_eps, EPI_BF16, p.qkv,
                     qkv_dim);
    }
    prefetch_phase(y.wo, H * HD, d, H * HD, EPI_RESIDUAL, 0, 512);
    stamp();
    bar.sync();
    stamp();
    // ---- P2: attention partials ----
    attention_items<GQ>(p, att_smem, l);
    stamp();
    bar.sync();
    stamp();
    // ---- P3: x += Wo . attn (the attention partials are merged while staging) ----
    gemv_phase<MT>(sm, y.wo, H * HD, d, H * HD, STAGE_ATTN, nullptr, 0, attn_src, p.rows, nullptr, 0.f, EPI_RESIDUAL,
                   p.x, d);
    stamp();
    bar.sync();
    stamp();
    // ---- P4: act = silu(Wg . rms(x)) * (Wu . rms(x)) ----
    gemv_phase<MT>(sm, y.wgu, d, f, d, STAGE_RMS, p.x, d, no_attn, p.rows, y.ln2, c.rms_eps, EPI_SWIGLU, p.act, f);
    stamp();
    bar.sync();
    stamp();
    // ---- P5: x += Wdown . act ----
    gemv_phase<MT>(sm, y.wdown, f, d, f, STAGE_PLAIN, p.act, f, no_attn, p.rows, nullptr, 0.f, EPI_RESIDUAL, p.x, d);
    stamp();
    bar.sync();
    stamp();
  }
  // ---- logits = Wlm . rms(x) ----
  gemv_phase<MT>(sm, p.lm_head, d, c.vocab, d, STAGE_RMS, p.x, d, no_attn, p.rows, p.norm, c.rms_eps, EPI_FP32,
                 p.logits, c.vocab);
  __syncthreads();
  stamp();
}

unsigned long long* g_timing = nullptr;

}  // namespace

void decode_megakernel_set_timing(unsigned long long* dev_buf) { g_timing = dev_buf; }

int64_t decode_megakernel_scratch_bytes(const pcy_llama_config& c, int rows, int S, int max_gen) {
  const int64_t d = c.d_model, qkv = (int64_t)(c.n_heads + 2 * c.n_kv_heads) * HD;
  const int max_splits = ceil_div(S + max_gen, ATT_CHUNK);
  int64_t b = round_up(rows * d * 2, 256) * 2 + round_up(rows * qkv * 2, 256) + round_up((int64_t)rows * c.ffn_dim * 2, 256);
  b += round_up((int64_t)rows * c.n_kv_heads * max_splits * (c.n_heads / c.n_kv_heads) * PSTR * 4, 256);
  b += 1024;
  return b;
}

bool decode_megakernel_supported(const pcy_llama_config& c, int rows) {
  const int gq = c.n_heads / c.n_kv_heads;
  if (rows < 1 || rows > 4 || c.head_dim != HD || gq != 4) return false;
  if (c.d_model % 256 != 0 || c.ffn_dim % 256 != 0) return false;
  const int sms = num_sms();
  // every CTA's slice of the widest phase must fit the shared-memory partial-sum buffer
  if (ceil_div(c.vocab, sms) + 1 > MAX_OUT_PER_CTA || ceil_div(c.ffn_dim, sms) + 1 > MAX_OUT_PER_CTA) return false;
  return true;
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
  p.barrier = reinterpret_cast<unsigned int*>(carve(256));
  carve(256);
  p.x = reinterpret_cast<bf16*>(carve(rows * d * 2));
  carve(rows * d * 2);
  p.qkv = reinterpret_cast<bf16*>(carve(rows * qkv * 2));
  p.act = reinterpret_cast<bf16*>(carve((int64_t)rows * c.ffn_dim * 2));
  p.max_splits = ceil_div(b->S + b->max_gen, ATT_CHUNK);
  p.part = reinterpret_cast<float*>(s);
  p.timing = g_timing;
  PCY_CUDA(cudaMemsetAsync(p.barrier, 0, 512, stream));  // barrier counter + tickets

  const int mt = rows <= 1 ? 1 : rows <= 2 ? 2 : 4;
  const int kmax = c.ffn_dim > c.d_model ? c.ffn_dim : c.d_model;
  const size_t smem_gemv = (size_t)mt * kmax * 2 + (size_t)MAX_OUT_PER_CTA * 2 * mt * 4 + MK_WARPS * 4 * 4;
  const size_t smem_att = (size_t)(4 * HD + 2 * HD + 4 * ATT_CHUNK + 8 + MK_WARPS * 4 * HD) * 4 + 64;
  const size_t smem = smem_gemv > smem_att ? smem_gemv : smem_att;
  PCY_REQUIRE(smem <= 220 * 1024, "decode megakernel: needs %zu bytes of shared memory", smem);
  void* fn = nullptr;
  if (mt == 1) fn = (void*)llama_decode_megakernel<1, 4>;
  else if (mt == 2) fn = (void*)llama_decode_megakernel<2, 4>;
  else fn = (void*)llama_decode_megakernel<4, 4>;
  static size_t smem_set[3] = {0, 0, 0};
  const int slot = mt == 1 ? 0 : mt == 2 ? 1 : 2;
  if (smem > smem_set[slot]) {
    PCY_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[slot] = smem;
  }
  void* args[] = {(void*)&p};
  // cooperative launch: guarantees that all CTAs are co-resident (the grid barrier needs it)
  PCY_CUDA(cudaLaunchCooperativeKernel(fn, dim3(num_sms()), dim3(MK_THREADS), args, smem, stream));
  count_launch();
  return 0;
}

}  // namespace pcy
