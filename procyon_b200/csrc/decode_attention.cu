// Single-token (decode) attention for the Llama KV-cache step, fused with RoPE of the new q/k and the append of
// the new k/v row to the cache.  One CTA per (key split, kv head, beam row); the 4 (= H/KVH) query heads of a
// kv head share every K/V read.  Beam search never copies the cache: the prompt part is stored once per input
// and shared by all its beams, the generated part is indexed through `slots[row][g]` = the physical row that
// holds generation slot g of this beam's history (see pcy_decode_select_beam).
//
// Replaces, per layer and step, HF LlamaAttention on a (beams,1) query with `torch.cat` KV growth
// (procyon/model/model_unified.py:769 -> pmc_llama.py:581) and the per-layer fancy-index cache reorder
// (procyon/model/model_unified.py:830-832).
#include "common.cuh"
#include "ops.h"

namespace pcy {

namespace {

constexpr int DA_THREADS = 128;
constexpr int DA_CHUNK = 128;  // keys per CTA
constexpr int HD = 128;        // Llama head_dim (checked on the host)
constexpr int MAX_GQ = 8;      // max query heads per kv head

struct DecAttnParams {
  const bf16* qkv;  // [rows, (H + 2*KVH) * HD], un-roped
  int64_t qkv_ld;
  const float* cos_sin;  // [P][HD/2][2]
  const bf16* kp;        // prompt K  [n_inputs][S][KVH*HD]
  const bf16* vp;
  bf16* kg;  // generated K [rows][max_gen][KVH*HD]
  bf16* vg;
  const int32_t* slots;         // [rows][max_gen]
  const uint8_t* prompt_valid;  // [n_inputs][S] or null
  const int32_t* state;         // state[0] = t (this step consumes token t-1 at position S + t - 1)
  float* part;                  // [rows][KVH][n_splits][GQ][HD + 2]
  int32_t* tickets;             // [rows][KVH]
  bf16* out;                    // [rows, H*HD]
  int H, KVH, S, max_gen, beams, n_splits;
  float scale_log2;
};

template <int GQ>
__global__ void __launch_bounds__(DA_THREADS)
decode_attn_kernel(const DecAttnParams p) {
  __shared__ float s_q[GQ][HD];       // roped q (bf16-rounded), pre-scaled by softmax_scale*log2e
  __shared__ float s_knew[HD], s_vnew[HD];
  __shared__ float s_sc[GQ][DA_CHUNK];  // scores -> probabilities
  __shared__ float s_red[GQ][4];
  __shared__ float s_m[GQ], s_l[GQ];
  __shared__ int s_last;

  const int split = blockIdx.x, kvh = blockIdx.y, row = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = p.state[0];
  const int g_cur = t - 1;           // generation slot written by this step
  const int pos_cur = p.S + g_cur;   // its position
  const int ctx = pos_cur + 1;       // keys visible: positions [0, ctx)
  const int input = row / p.beams;
  const int kvdim = p.KVH * HD;
  const bf16* qkv_row = p.qkv + (int64_t)row * p.qkv_ld;

  // ---- RoPE of q (GQ heads) and of the new k; stage v ----
  {
    const float2* cs = reinterpret_cast<const float2*>(p.cos_sin) + (int64_t)pos_cur * (HD / 2);
    for (int i = tid; i < (GQ + 1) * (HD / 2); i += DA_THREADS) {
      const int hh = i / (HD / 2), j = i % (HD / 2);
      const bf16* src = (hh < GQ) ? qkv_row + (kvh * GQ + hh) * HD : qkv_row + (p.H + kvh) * HD;
      const float lo = __bfloat162float(src[j]), hi = __bfloat162float(src[j + HD / 2]);
      const float2 c = cs[j];
      const float o_lo = bf16_round(lo * c.x - hi * c.y), o_hi = bf16_round(hi * c.x + lo * c.y);
      if (hh < GQ) {
        s_q[hh][j] = o_lo * p.scale_log2;
        s_q[hh][j + HD / 2] = o_hi * p.scale_log2;
      } else {
        s_knew[j] = o_lo;
        s_knew[j + HD / 2] = o_hi;
      }
    }
    if (tid < HD) s_vnew[tid] = __bfloat162float(qkv_row[(p.H + p.KVH + kvh) * HD + tid]);
  }
  __syncthreads();

  const int k0 = split * DA_CHUNK;
  const int k1 = min(ctx, k0 + DA_CHUNK);
  const int n_keys = max(0, k1 - k0);

  // the CTA whose range holds the current position appends the new k/v row to the cache
  if (pos_cur >= k0 && pos_cur < k0 + DA_CHUNK && tid < HD) {
    const int64_t off = ((int64_t)row * p.max_gen + g_cur) * kvdim + kvh * HD + tid;
    p.kg[off] = __float2bfloat16_rn(s_knew[tid]);
    p.vg[off] = __float2bfloat16_rn(s_vnew[tid]);
  }

  // ---- phase 1: scores. 8 lanes per key (16 dims each), 4 keys per warp iteration ----
  const int sub = lane >> 3;   // key within the warp's group of 4
  const int l8 = lane & 7;     // 16-dim slice
  for (int kk = warp * 4 + sub; kk < DA_CHUNK; kk += 16) {
    const int pos = k0 + kk;
    float acc[GQ];
#pragma unroll
    for (int h = 0; h < GQ; ++h) acc[h] = 0.f;
    bool valid = pos < ctx;
    if (valid) {
      float kf[16];
      if (pos == pos_cur) {
#pragma unroll
        for (int j = 0; j < 16; ++j) kf[j] = s_knew[l8 * 16 + j];
      } else {
        const bf16* kptr;
        if (pos < p.S) {
          kptr = p.kp + ((int64_t)input * p.S + pos) * kvdim + kvh * HD;
          if (p.prompt_valid) valid = p.prompt_valid[(int64_t)input * p.S + pos] != 0;
        } else {
          const int g = pos - p.S;
          const int prow = p.slots[(int64_t)row * p.max_gen + g];
          kptr = p.kg + ((int64_t)prow * p.max_gen + g) * kvdim + kvh * HD;
        }
        const uint4 u0 = *reinterpret_cast<const uint4*>(kptr + l8 * 16);
        const uint4 u1 = *reinterpret_cast<const uint4*>(kptr + l8 * 16 + 8);
        const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          kf[2 * j] = f.x;
          kf[2 * j + 1] = f.y;
        }
      }
#pragma unroll
      for (int h = 0; h < GQ; ++h) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) a = fmaf(kf[j], s_q[h][l8 * 16 + j], a);
        acc[h] = a;
      }
    }
#pragma unroll
    for (int h = 0; h < GQ; ++h) {
      float a = acc[h];
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      if (l8 == 0) s_sc[h][kk] = valid ? a : -INFINITY;
    }
  }
  __syncthreads();

  // ---- softmax statistics per head over the chunk (warp w handles heads w, w+4, ...) ----
  for (int h = warp; h < GQ; h += 4) {
    float m = -INFINITY;
    for (int k = lane; k < DA_CHUNK; k += 32) m = fmaxf(m, s_sc[h][k]);
    m = warp_max(m);
    float l = 0.f;
    for (int k = lane; k < DA_CHUNK; k += 32) {
      const float pr = (m == -INFINITY) ? 0.f : exp2f(s_sc[h][k] - m);
      s_sc[h][k] = pr;
      l += pr;
    }
    l = warp_sum(l);
    if (lane == 0) { s_m[h] = m; s_l[h] = l; }
  }
  __syncthreads();

  // ---- phase 2: o[h][dim] = sum_k p[h][k] v[k][dim]; thread = dim ----
  float o[GQ];
#pragma unroll
  for (int h = 0; h < GQ; ++h) o[h] = 0.f;
  if (tid < HD) {
    for (int kk = 0; kk < n_keys; ++kk) {
      const int pos = k0 + kk;
      float v;
      if (pos == pos_cur) {
        v = bf16_round(s_vnew[tid]);
      } else if (pos < p.S) {
        v = __bfloat162float(p.vp[((int64_t)input * p.S + pos) * kvdim + kvh * HD + tid]);
      } else {
        const int g = pos - p.S;
        const int prow = p.slots[(int64_t)row * p.max_gen + g];
        v = __bfloat162float(p.vg[((int64_t)prow * p.max_gen + g) * kvdim + kvh * HD + tid]);
      }
#pragma unroll
      for (int h = 0; h < GQ; ++h) o[h] = fmaf(s_sc[h][kk], v, o[h]);
    }
  }

  // ---- write the partial, take a ticket; the last CTA of this (row, kvh) combines ----
  float* part = p.part + (((int64_t)row * p.KVH + kvh) * p.n_splits + split) * GQ * (HD + 2);
  if (tid < HD) {
#pragma unroll
    for (int h = 0; h < GQ; ++h) part[h * (HD + 2) + tid] = o[h];
  }
  if (tid < GQ) {
    part[tid * (HD + 2) + HD] = s_m[tid];
    part[tid * (HD + 2) + HD + 1] = s_l[tid];
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int ticket = atomicAdd(&p.tickets[row * p.KVH + kvh], 1);
    s_last = (ticket == p.n_splits - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* base = p.part + ((int64_t)row * p.KVH + kvh) * p.n_splits * GQ * (HD + 2);
  if (tid < HD) {
#pragma unroll
    for (int h = 0; h < GQ; ++h) {
      float m = -INFINITY;
      for (int s = 0; s < p.n_splits; ++s) m = fmaxf(m, __ldcg(base + (s * GQ + h) * (HD + 2) + HD));
      float l = 0.f, acc = 0.f;
      for (int s = 0; s < p.n_splits; ++s) {
        const float* ps = base + (s * GQ + h) * (HD + 2);
        const float ms = __ldcg(ps + HD);
        if (ms == -INFINITY) continue;
        const float w = exp2f(ms - m);
        l += w * __ldcg(ps + HD + 1);
        acc += w * __ldcg(ps + tid);
      }
      const float r = (l > 0.f) ? acc / l : 0.f;
      p.out[(int64_t)row * (p.H * HD) + (kvh * GQ + h) * HD + tid] = __float2bfloat16_rn(r);
    }
  }
  if (tid == 0) p.tickets[row * p.KVH + kvh] = 0;  // ready for the next launch (graph replay)
}

}  // namespace

int decode_attention(const DecodeAttnArgs& a, cudaStream_t stream) {
  PCY_REQUIRE(a.head_dim == HD, "decode attention: head_dim must be 128 (got %d)", a.head_dim);
  PCY_REQUIRE(a.H % a.KVH == 0, "decode attention: H %% KVH != 0");
  const int gq = a.H / a.KVH;
  DecAttnParams p;
  p.qkv = a.qkv; p.qkv_ld = a.qkv_ld; p.cos_sin = a.cos_sin; p.kp = a.k_prompt; p.vp = a.v_prompt;
  p.kg = a.k_gen; p.vg = a.v_gen; p.slots = a.slots; p.prompt_valid = a.prompt_valid; p.state = a.state;
  p.part = a.partials; p.tickets = a.tickets; p.out = a.out; p.H = a.H; p.KVH = a.KVH; p.S = a.S;
  p.max_gen = a.max_gen; p.beams = a.beams; p.n_splits = decode_attention_splits(a.S, a.max_gen);
  p.scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
  dim3 grid(p.n_splits, a.KVH, a.rows);
  if (gq == 4) decode_attn_kernel<4><<<grid, DA_THREADS, 0, stream>>>(p);
  else if (gq == 1) decode_attn_kernel<1><<<grid, DA_THREADS, 0, stream>>>(p);
  else if (gq == 2) decode_attn_kernel<2><<<grid, DA_THREADS, 0, stream>>>(p);
  else if (gq == 8) decode_attn_kernel<8><<<grid, DA_THREADS, 0, stream>>>(p);
  else return set_error(PCY_ERR_UNSUPPORTED, "decode attention: H/KVH=%d unsupported (1,2,4,8)", gq);
  PCY_LAUNCH_CHECK();
  return 0;
}

int decode_attention_splits(int S, int max_gen) { return ceil_div(S + max_gen, DA_CHUNK); }
int64_t decode_attention_partial_floats(int rows, int H, int KVH, int S, int max_gen) {
  return (int64_t)rows * KVH * decode_attention_splits(S, max_gen) * (H / KVH) * (HD + 2);
}

}  // namespace pcy
