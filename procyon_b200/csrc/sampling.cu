// Decode-step token selection on the device: per-row log-sum-exp + top-K of the fp32 logits, then greedy
// or diverse-beam selection with the beam bookkeeping (token history, log-probs, KV slot ancestry) done by one
// CTA per input — no logits leave the GPU and nothing syncs with the host.
//
// Replaces the host loop of UnifiedProCyon._generate_beam_search (procyon/model/model_unified.py:782-833:
// LogSoftmax, per-group bincount Hamming penalty, ravel().topk, beam reorder of out / log-probs / per-step
// logits / every layer's K and V) and the greedy branch of _generate_sampling (:891-911).
#include "common.cuh"
#include "ops.h"

namespace pcy {

namespace {

constexpr int TK_THREADS = 512;
constexpr int TK_SEG = 8;      // segments per row (CTAs per row)
constexpr int TK_MAXK = 16;    // max candidates kept per segment
constexpr int SEL_MAX_CAND = 16 * TK_SEG * TK_MAXK;  // beam_group_size = beam_size = 16 (vanilla beam search)

struct Cand {
  float v;
  int idx;
};
// order: larger value first; ties -> smaller index first (deterministic)
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ Cand block_argmax(Cand c, Cand* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, c.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, c.idx, o);
    if (better(ov, oi, c.v, c.idx)) { c.v = ov; c.idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) s_red[warp] = c;
  __syncthreads();
  if (warp == 0) {
    Cand d = (lane < (int)(blockDim.x >> 5)) ? s_red[lane] : Cand{-INFINITY, 0x7fffffff};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, d.v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, d.idx, o);
      if (better(ov, oi, d.v, d.idx)) { d.v = ov; d.idx = oi; }
    }
    if (lane == 0) s_red[0] = d;
  }
  __syncthreads();
  return s_red[0];
}

// grid (TK_SEG, rows). Per segment: max, sum exp(x - max), top-K (value, index), optional copy to history.
__global__ void __launch_bounds__(TK_THREADS)
topk_segment_kernel(const float* __restrict__ logits, int V, int K, float* __restrict__ seg_max,
                    float* __restrict__ seg_sum, float* __restrict__ cand_val, int32_t* __restrict__ cand_idx,
                    float* __restrict__ hist, const int32_t* __restrict__ state, int rows) {
  __shared__ Cand s_red[TK_THREADS / 32];
  __shared__ float s_f[TK_THREADS / 32];
  const int seg = blockIdx.x, row = blockIdx.y;
  const int per = (V + TK_SEG - 1) / TK_SEG;
  const int lo = seg * per, hi = min(V, lo + per);
  const float* x = logits + (int64_t)row * V;
  float* h = nullptr;
  if (hist != nullptr) h = hist + ((int64_t)state[0] * rows + row) * V;

  float m = -INFINITY;
  for (int i = lo + threadIdx.x; i < hi; i += TK_THREADS) {
    const float v = x[i];
    if (h) h[i] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_f[threadIdx.x >> 5] = m;
  __syncthreads();
  m = s_f[0];
  for (int w = 1; w < TK_THREADS / 32; ++w) m = fmaxf(m, s_f[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += TK_THREADS) s += expf(x[i] - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_f[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < TK_THREADS / 32; ++w) t += s_f[w];
    seg_max[row * TK_SEG + seg] = m;
    seg_sum[row * TK_SEG + seg] = t;
  }
  // K rounds of block arg-max over the elements strictly after the previous pick in (value desc, index asc) order
  float pv = INFINITY;
  int pi = -1;
  for (int k = 0; k < K; ++k) {
    Cand c{-INFINITY, 0x7fffffff};
    for (int i = lo + threadIdx.x; i < hi; i += TK_THREADS) {
      const float v = x[i];
      const bool after = (v < pv) || (v == pv && i > pi);
      if (after && better(v, i, c.v, c.idx)) { c.v = v; c.idx = i; }
    }
    c = block_argmax(c, s_red);
    if (threadIdx.x == 0) {
      cand_val[((int64_t)row * TK_SEG + seg) * TK_MAXK + k] = c.v;
      cand_idx[((int64_t)row * TK_SEG + seg) * TK_MAXK + k] = (c.idx == 0x7fffffff) ? -1 : c.idx;
    }
    pv = c.v;
    pi = c.idx;
    __syncthreads();
  }
}

struct SelectParams {
  const float* seg_max;
  const float* seg_sum;
  const float* cand_val;
  const int32_t* cand_idx;
  int32_t* tokens;  // [rows][max_gen]
  int32_t* slots;   // [rows][max_gen]
  float* logprobs;  // [rows]
  int32_t* state;   // [0]=t, [2]=finished, [3]=finish step
  int32_t* gstate;  // lock-step session group (or null): [0]=inputs whose beams all hold EOS this step, [1]=finished
  int beams, group, max_gen, K, eos_id;
  float penalty;
  int greedy;
};

// One CTA per input. Dynamic smem: old token/slot rows of all beams.
__global__ void __launch_bounds__(256)
select_kernel(const SelectParams p) {
  extern __shared__ int32_t s_hist[];  // [beams][t] tokens then [beams][t] slots
  __shared__ float s_lse[32];
  __shared__ float s_cv[SEL_MAX_CAND];  // candidate values of the rows of one group (<= 16 rows x 8 segments x 16)
  __shared__ int s_ct[SEL_MAX_CAND];
  __shared__ int s_cr[SEL_MAX_CAND];
  __shared__ Cand s_red[8];
  __shared__ int s_sel_tok[32], s_sel_par[32];
  __shared__ float s_sel_val[32];
  __shared__ int s_all_eos;

  const int input = blockIdx.x;
  const int t = p.state[0];
  // generation already finished (all beams of all inputs hit EOS): later steps are no-ops
  if ((p.gstate ? p.gstate[1] : p.state[2]) != 0) return;
  const int beams = p.beams, G = p.group;
  const int r0 = input * beams;
  const int tid = threadIdx.x;

  // log-sum-exp per row from the segment statistics
  if (tid < beams) {
    const int row = r0 + tid;
    float m = -INFINITY;
    for (int s = 0; s < TK_SEG; ++s) m = fmaxf(m, p.seg_max[row * TK_SEG + s]);
    float tot = 0.f;
    for (int s = 0; s < TK_SEG; ++s) tot += p.seg_sum[row * TK_SEG + s] * expf(p.seg_max[row * TK_SEG + s] - m);
    s_lse[tid] = m + logf(tot);
  }
  // stage old histories
  for (int i = tid; i < beams * t; i += blockDim.x) {
    const int b = i / t, j = i % t;
    s_hist[i] = p.tokens[(int64_t)(r0 + b) * p.max_gen + j];
    s_hist[beams * t + i] = p.slots[(int64_t)(r0 + b) * p.max_gen + j];
  }
  __syncthreads();

  const int n_groups = beams / G;
  const int per_row = TK_SEG * p.K;
  for (int g = 0; g < n_groups; ++g) {
    const int gs = g * G;                       // first beam of the group (relative)
    const int n_check = (t == 0) ? 1 : G;       // identical beams at step 0: look at one (model_unified.py:788-795)
    const int n_cand = n_check * per_row;
    // gather candidates: value = logit - lse + running log-prob - penalty * (#earlier-group picks of that token)
    for (int i = tid; i < n_cand; i += blockDim.x) {
      const int rr = i / per_row, c = i % per_row;
      const int seg = c / p.K, k = c % p.K;
      const int row = r0 + gs + rr;
      const int64_t off = ((int64_t)row * TK_SEG + seg) * TK_MAXK + k;
      const int tok = p.cand_idx[off];
      float v = -INFINITY;
      if (tok >= 0) {
        v = p.cand_val[off] - s_lse[gs + rr] + (p.greedy ? 0.f : p.logprobs[row]);
        if (!p.greedy && g > 0) {
          int cnt = 0;
          for (int q = 0; q < gs; ++q) cnt += (s_sel_tok[q] == tok);
          v -= p.penalty * (float)cnt;
        }
      }
      s_cv[i] = v;
      s_ct[i] = tok;
      s_cr[i] = gs + rr;
    }
    __syncthreads();
    // pick the top-G (value desc; ties: lower flat index (row*V + tok) first, as a stable ravel().topk would)
    for (int j = 0; j < G; ++j) {
      Cand c{-INFINITY, 0x7fffffff};
      for (int i = tid; i < n_cand; i += blockDim.x)
        if (s_ct[i] >= 0 && better(s_cv[i], i, c.v, c.idx)) { c.v = s_cv[i]; c.idx = i; }
      c = block_argmax(c, s_red);
      if (tid == 0) {
        s_sel_tok[gs + j] = s_ct[c.idx];
        s_sel_par[gs + j] = s_cr[c.idx];
        s_sel_val[gs + j] = c.v;
        s_ct[c.idx] = -1;  // consumed
      }
      __syncthreads();
    }
  }

  // write the new histories
  for (int i = tid; i < beams * t; i += blockDim.x) {
    const int b = i / t, j = i % t;
    const int par = s_sel_par[b];
    p.tokens[(int64_t)(r0 + b) * p.max_gen + j] = s_hist[par * t + j];
    if (j < t - 1) p.slots[(int64_t)(r0 + b) * p.max_gen + j] = s_hist[beams * t + par * t + j];
  }
  if (tid == 0) s_all_eos = 1;
  __syncthreads();
  if (tid < beams) {
    const int b = tid;
    const int par = s_sel_par[b];
    p.tokens[(int64_t)(r0 + b) * p.max_gen + t] = s_sel_tok[b];
    // KV of generation slot t-1 (written by this step's forward) lives in the parent's physical row
    if (t >= 1) p.slots[(int64_t)(r0 + b) * p.max_gen + (t - 1)] = r0 + par;
    if (p.greedy) p.logprobs[r0 + b] += s_sel_val[b];
    else p.logprobs[r0 + b] = s_sel_val[b];
    bool has_eos = (s_sel_tok[b] == p.eos_id);
    for (int j = 0; j < t && !has_eos; ++j) has_eos = (s_hist[par * t + j] == p.eos_id);
    if (!has_eos) s_all_eos = 0;
  }
  __syncthreads();
  if (tid == 0 && p.eos_id >= 0) {
    // counts the inputs whose beams all contain EOS at this step (across every session of a lock-step group)
    if (s_all_eos) atomicAdd(p.gstate ? &p.gstate[0] : &p.state[4], 1);
  }
}

// gstate (nullable): shared by the sessions that split one batch (<= 16 beam rows each) and step in lock-step on one
// stream — the reference stops the WHOLE batch at the first step where every beam of every input holds an EOS
// (procyon/model/model_unified.py:833), so the count is accumulated over the group and evaluated by its last session.
__global__ void advance_kernel(int32_t* state, int32_t* gstate, int n_inputs, int check_eos, int group_last) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (gstate != nullptr) {
      if (gstate[1] == 0) {
        if (group_last) {
          if (check_eos && gstate[0] == gstate[3]) {
            gstate[1] = 1;
            gstate[2] = state[0];
          }
          gstate[0] = 0;
        }
        state[0] += 1;
      }
      return;
    }
    if (state[2] == 0) {
      if (check_eos && state[4] == n_inputs) {
        state[2] = 1;
        state[3] = state[0];  // step index at which every beam had produced EOS
      }
      state[0] += 1;
    }
    state[4] = 0;
  }
}

}  // namespace

int topk_workspace_floats(int rows) { return rows * TK_SEG * (2 + 2 * TK_MAXK); }

int decode_select(const DecodeSelectArgs& a, cudaStream_t stream) {
  PCY_REQUIRE(a.beams >= 1 && a.beams <= 32, "select: beams=%d out of range [1,32]", a.beams);
  PCY_REQUIRE(a.group >= 1 && a.beams % a.group == 0, "beam_group_size must evenly divide beam_size, got: %d %% %d != 0",
              a.beams, a.group);
  const int K = a.greedy ? 1 : a.beams;
  PCY_REQUIRE(K <= TK_MAXK, "select: beams=%d exceeds the per-segment candidate limit %d", a.beams, TK_MAXK);
  PCY_REQUIRE(a.group * TK_SEG * K <= SEL_MAX_CAND, "select: group too large");
  const int rows = a.n_inputs * a.beams;
  float* seg_max = a.workspace;
  float* seg_sum = seg_max + rows * TK_SEG;
  float* cand_val = seg_sum + rows * TK_SEG;
  int32_t* cand_idx = reinterpret_cast<int32_t*>(cand_val + (int64_t)rows * TK_SEG * TK_MAXK);
  dim3 grid(TK_SEG, rows);
  topk_segment_kernel<<<grid, TK_THREADS, 0, stream>>>(a.logits, a.vocab, K, seg_max, seg_sum, cand_val, cand_idx,
                                                       a.logits_hist, a.state, rows);
  PCY_LAUNCH_CHECK();
  SelectParams p;
  p.seg_max = seg_max; p.seg_sum = seg_sum; p.cand_val = cand_val; p.cand_idx = cand_idx;
  p.tokens = a.tokens; p.slots = a.slots; p.logprobs = a.logprobs; p.state = a.state; p.gstate = a.group_state;
  p.beams = a.beams; p.group = a.greedy ? 1 : a.group; p.max_gen = a.max_gen; p.K = K; p.eos_id = a.eos_id;
  p.penalty = a.diversity_penalty; p.greedy = a.greedy;
  const size_t smem = (size_t)2 * a.beams * a.max_gen * sizeof(int32_t);
  PCY_REQUIRE(smem <= 160 * 1024, "select: beams*max_gen too large for shared memory");
  static SmemOptIn opt;
  if (smem > 16 * 1024 && opt.need(smem))
    PCY_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  select_kernel<<<a.n_inputs, 256, smem, stream>>>(p);
  PCY_LAUNCH_CHECK();
  advance_kernel<<<1, 32, 0, stream>>>(a.state, a.group_state, a.n_inputs, a.stop_on_all_eos, a.group_last);
  PCY_LAUNCH_CHECK();
  return 0;
}

}  // namespace pcy
