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
  int split0;  // first key split handled by the per-row kernel (the splits before it: shared-prompt kernel)
  float scale_log2;
};

// The CTA that delivers the last split of (row, kv head) merges all of them: out[h][dim] = sum_s w_s acc_s / sum_s w_s l_s
template <int GQ>
__device__ __forceinline__ void combine_splits(const DecAttnParams& p, int row, int kvh, int tid) {
  const float* base = p.part + ((int64_t)row * p.KVH + kvh) * p.n_splits * GQ * (HD + 2);
  constexpr int CB = 8;  // splits per batch of independent loads (one L2 round trip per batch, not per split)
  if (tid < HD) {
#pragma unroll
    for (int h = 0; h < GQ; ++h) {
      const float* ph = base + h * (HD + 2);
      const int64_t stride = (int64_t)GQ * (HD + 2);
      float m = -INFINITY;
      if (p.n_splits > CB) {  // otherwise the single batch below carries its own maxima
        for (int s0 = 0; s0 < p.n_splits; s0 += CB) {
          float mv[CB];
#pragma unroll
          for (int i = 0; i < CB; ++i) mv[i] = (s0 + i < p.n_splits) ? __ldcg(ph + (s0 + i) * stride + HD) : -INFINITY;
#pragma unroll
          for (int i = 0; i < CB; ++i) m = fmaxf(m, mv[i]);
        }
      }
      float l = 0.f, acc = 0.f;
      for (int s0 = 0; s0 < p.n_splits; s0 += CB) {
        float mv[CB], lv[CB], vv[CB];
#pragma unroll
        for (int i = 0; i < CB; ++i) {
          const bool in = s0 + i < p.n_splits;
          mv[i] = in ? __ldcg(ph + (s0 + i) * stride + HD) : -INFINITY;
          lv[i] = in ? __ldcg(ph + (s0 + i) * stride + HD + 1) : 0.f;
          vv[i] = in ? __ldcg(ph + (s0 + i) * stride + tid) : 0.f;
        }
        if (p.n_splits <= CB) {
#pragma unroll
          for (int i = 0; i < CB; ++i) m = fmaxf(m, mv[i]);
        }
#pragma unroll
        for (int i = 0; i < CB; ++i) {
          const float w = (mv[i] == -INFINITY) ? 0.f : exp2f(mv[i] - m);
          l = fmaf(w, lv[i], l);
          acc = fmaf(w, vv[i], acc);
        }
      }
      const float r = (l > 0.f) ? acc / l : 0.f;
      p.out[(int64_t)row * (p.H * HD) + (kvh * GQ + h) * HD + tid] = __float2bfloat16_rn(r);
    }
  }
}

template <int GQ>
__global__ void __launch_bounds__(DA_THREADS)
decode_attn_kernel(const DecAttnParams p) {
  __shared__ float s_q[GQ][HD];       // roped q (bf16-rounded), pre-scaled by softmax_scale*log2e
  __shared__ float s_knew[HD], s_vnew[HD];
  __shared__ float s_sc[GQ][DA_CHUNK];  // scores -> probabilities
  __shared__ float s_red[GQ][4];
  __shared__ float s_m[GQ], s_l[GQ];
  __shared__ int s_last;

  const int split = blockIdx.x + p.split0, kvh = blockIdx.y, row = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  pdl_wait();  // qkv, the step state and the partial buffers belong to the preceding kernels
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
  combine_splits<GQ>(p, row, kvh, tid);
  if (tid == 0) p.tickets[row * p.KVH + kvh] = 0;  // ready for the next launch (graph replay)
}

// ---------------------------------------------------------------------------------------------------------------
// Beam search: all beams of an input attend to the SAME prompt K/V, so the key splits that lie entirely inside the
// prompt are handled once per (input, kv head, split) for all beams x GQ query heads together, on the tensor cores:
// S = Q K^T and O = P V with mma.sync.m16n8k16 (Q [<= 64 rows] x K tile [128 keys] x 128 dims).  At 10 beams the
// per-row kernel above read the prompt K/V ten times and took 98 us per layer; this one reads it once.
// The generated tail (per-beam rows, slot indirection) stays with the per-row kernel; both write the same partial
// buffer and share the ticket counters, whoever delivers the last split of a row merges.
constexpr int SP_THREADS = 256;
constexpr int SP_KEYS = DA_CHUNK;   // 128 keys per CTA
constexpr int SP_MAXR = 64;         // beams * GQ query rows per input
constexpr int SP_SPITCH = SP_KEYS + 4;   // fp32 score row pitch (floats)
constexpr int SP_PPITCH = SP_KEYS + 8;   // bf16 probability row pitch (elements): 272 B rows, conflict-free ldmatrix
constexpr int SP_SMEM = SP_MAXR * HD * 2 + 2 * SP_KEYS * HD * 2 + SP_MAXR * SP_SPITCH * 4 + SP_MAXR * SP_PPITCH * 2 +
                        2 * SP_MAXR * 4 + 64;

__device__ __forceinline__ uint32_t sw_addr(uint32_t base, int row, int chunk) {  // 256-byte rows, 16-byte chunks
  return base + row * 256 + (((chunk & 8) | ((chunk ^ row) & 7)) << 4);
}
__device__ __forceinline__ void sp_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void sp_ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void sp_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int GQ>
__global__ void __launch_bounds__(SP_THREADS)
decode_attn_shared_prompt_kernel(const DecAttnParams p) {
  extern __shared__ __align__(128) uint8_t sp_smem[];
  const uint32_t sQ = smem_u32(sp_smem);
  const uint32_t sK = sQ + SP_MAXR * HD * 2;
  const uint32_t sV = sK + SP_KEYS * HD * 2;
  float* sS = reinterpret_cast<float*>(sp_smem + SP_MAXR * HD * 2 + 2 * SP_KEYS * HD * 2);
  bf16* sP = reinterpret_cast<bf16*>(sS + SP_MAXR * SP_SPITCH);
  float* sM = reinterpret_cast<float*>(sP + SP_MAXR * SP_PPITCH);
  float* sL = sM + SP_MAXR;
  int* s_last = reinterpret_cast<int*>(sL + SP_MAXR);
  const uint32_t sP_u = smem_u32(sP);

  const int split = blockIdx.x, kvh = blockIdx.y, input = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  const int R = p.beams * GQ;             // live query rows
  const int m_tiles = (R + 15) / 16;
  const int t = p.state[0];
  const int pos_cur = p.S + t - 1;
  const int kvdim = p.KVH * HD;
  const int k0 = split * SP_KEYS;

  // ---- K / V tiles: 128 keys x 256 B, swizzled 16-byte chunks ----
  {
    const bf16* kb = p.kp + ((int64_t)input * p.S + k0) * kvdim + kvh * HD;
    const bf16* vb = p.vp + ((int64_t)input * p.S + k0) * kvdim + kvh * HD;
    for (int c = tid; c < SP_KEYS * 16; c += SP_THREADS) {
      const int row = c >> 4, ch = c & 15;
      const uint32_t dk = sw_addr(sK, row, ch), dv = sw_addr(sV, row, ch);
      const bf16* srck = kb + (int64_t)row * kvdim + ch * 8;
      const bf16* srcv = vb + (int64_t)row * kvdim + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dk), "l"(srck) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dv), "l"(srcv) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // (the prompt's K / V rows were written by the prefill and never change during decoding: they are requested above,
  // before the dependency wait; qkv and the step state come from the preceding kernels)
  pdl_wait();
  // ---- Q: RoPE (rotate-half, rounded to bf16 like the stored k) of the beams' GQ heads; rows >= R are zero ----
  {
    const float2* cs = reinterpret_cast<const float2*>(p.cos_sin) + (int64_t)pos_cur * (HD / 2);
    for (int i = tid; i < SP_MAXR * (HD / 2); i += SP_THREADS) {
      const int r = i / (HD / 2), j = i % (HD / 2);
      float o_lo = 0.f, o_hi = 0.f;
      if (r < R) {
        const int row = input * p.beams + r / GQ, hh = r % GQ;
        const bf16* src = p.qkv + (int64_t)row * p.qkv_ld + (kvh * GQ + hh) * HD;
        const float lo = __bfloat162float(src[j]), hi = __bfloat162float(src[j + HD / 2]);
        const float2 c = cs[j];
        o_lo = lo * c.x - hi * c.y;
        o_hi = hi * c.x + lo * c.y;
      }
      const uint32_t a0 = sw_addr(sQ, r, j >> 3) + (j & 7) * 2, a1 = sw_addr(sQ, r, (j + HD / 2) >> 3) + (j & 7) * 2;
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(a0), "h"(__bfloat16_as_ushort(__float2bfloat16_rn(o_lo))) : "memory");
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(a1), "h"(__bfloat16_as_ushort(__float2bfloat16_rn(o_hi))) : "memory");
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- S = Q K^T: warp w owns keys 16 w .. 16 w + 15 ----
  {
    float acc[SP_MAXR / 16][2][4];
#pragma unroll
    for (int mt = 0; mt < SP_MAXR / 16; ++mt)
#pragma unroll
      for (int n = 0; n < 2; ++n) acc[mt][n][0] = acc[mt][n][1] = acc[mt][n][2] = acc[mt][n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      uint32_t bf[4];  // K as B: (keys 0-7, k lo) (keys 0-7, k hi) (keys 8-15, k lo) (keys 8-15, k hi)
      {
        const int r = warp * 16 + (lane & 7) + (lane >> 4) * 8, ch = ks * 2 + ((lane >> 3) & 1);
        sp_ldsm_x4(sw_addr(sK, r, ch), bf);
      }
#pragma unroll
      for (int mt = 0; mt < SP_MAXR / 16; ++mt) {
        if (mt < m_tiles) {
          uint32_t af[4];  // Q as A: (rows 0-7, k lo) (rows 8-15, k lo) (rows 0-7, k hi) (rows 8-15, k hi)
          const int r = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ch = ks * 2 + (lane >> 4);
          sp_ldsm_x4(sw_addr(sQ, r, ch), af);
          sp_mma(acc[mt][0], af, bf[0], bf[1]);
          sp_mma(acc[mt][1], af, bf[2], bf[3]);
        }
      }
    }
    const int g = lane >> 2, q2 = (lane & 3) * 2;
#pragma unroll
    for (int mt = 0; mt < SP_MAXR / 16; ++mt) {
      if (mt < m_tiles) {
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int key = warp * 16 + n * 8 + q2;
          bool v0 = true, v1 = true;
          if (p.prompt_valid) {
            v0 = p.prompt_valid[(int64_t)input * p.S + k0 + key] != 0;
            v1 = p.prompt_valid[(int64_t)input * p.S + k0 + key + 1] != 0;
          }
          float* r0 = sS + (mt * 16 + g) * SP_SPITCH + key;
          float* r1 = sS + (mt * 16 + g + 8) * SP_SPITCH + key;
          r0[0] = v0 ? acc[mt][n][0] * p.scale_log2 : -INFINITY;
          r0[1] = v1 ? acc[mt][n][1] * p.scale_log2 : -INFINITY;
          r1[0] = v0 ? acc[mt][n][2] * p.scale_log2 : -INFINITY;
          r1[1] = v1 ? acc[mt][n][3] * p.scale_log2 : -INFINITY;
        }
      }
    }
  }
  __syncthreads();
  // ---- softmax statistics and probabilities, one warp per query row ----
  for (int r = warp; r < m_tiles * 16; r += SP_THREADS / 32) {
    float x[4], m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[i] = (r < R) ? sS[r * SP_SPITCH + lane + 32 * i] : -INFINITY;
      m = fmaxf(m, x[i]);
    }
    m = warp_max(m);
    float l = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float pr = (m == -INFINITY) ? 0.f : exp2f(x[i] - m);
      l += pr;
      sP[r * SP_PPITCH + lane + 32 * i] = __float2bfloat16_rn(pr);
    }
    l = warp_sum(l);
    if (lane == 0) { sM[r] = m; sL[r] = l; }
  }
  __syncthreads();
  // ---- O = P V: warp w owns output dims 16 w .. 16 w + 15 ----
  {
    float acc[SP_MAXR / 16][2][4];
#pragma unroll
    for (int mt = 0; mt < SP_MAXR / 16; ++mt)
#pragma unroll
      for (int n = 0; n < 2; ++n) acc[mt][n][0] = acc[mt][n][1] = acc[mt][n][2] = acc[mt][n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < SP_KEYS / 16; ++ks) {
      uint32_t bf[4];  // V [key][dim] as B through transposing loads: (keys lo, dims lo) (keys hi, dims lo) (keys lo, dims hi) ...
      {
        const int kr = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ch = warp * 2 + (lane >> 4);
        sp_ldsm_x4_t(sw_addr(sV, kr, ch), bf);
      }
#pragma unroll
      for (int mt = 0; mt < SP_MAXR / 16; ++mt) {
        if (mt < m_tiles) {
          uint32_t af[4];  // P as A
          const int r = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kc = ks * 16 + (lane >> 4) * 8;
          sp_ldsm_x4(sP_u + (r * SP_PPITCH + kc) * 2, af);
          sp_mma(acc[mt][0], af, bf[0], bf[1]);
          sp_mma(acc[mt][1], af, bf[2], bf[3]);
        }
      }
    }
    // partials: [row][KVH][n_splits][GQ][HD + 2]
    const int g = lane >> 2, q2 = (lane & 3) * 2;
#pragma unroll
    for (int mt = 0; mt < SP_MAXR / 16; ++mt) {
      if (mt < m_tiles) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int r = mt * 16 + g + hf * 8;
          if (r < R) {
            const int row = input * p.beams + r / GQ, hh = r % GQ;
            float* part = p.part + ((((int64_t)row * p.KVH + kvh) * p.n_splits + split) * GQ + hh) * (HD + 2);
#pragma unroll
            for (int n = 0; n < 2; ++n) {
              const int dim = warp * 16 + n * 8 + q2;
              part[dim] = acc[mt][n][hf * 2];
              part[dim + 1] = acc[mt][n][hf * 2 + 1];
            }
          }
        }
      }
    }
  }
  for (int r = tid; r < R; r += SP_THREADS) {
    const int row = input * p.beams + r / GQ, hh = r % GQ;
    float* part = p.part + ((((int64_t)row * p.KVH + kvh) * p.n_splits + split) * GQ + hh) * (HD + 2);
    part[HD] = sM[r];
    part[HD + 1] = sL[r];
  }
  // ---- tickets: one per beam row ----
  __threadfence();
  __syncthreads();
  for (int b = 0; b < p.beams; ++b) {
    const int row = input * p.beams + b;
    if (tid == 0) {
      const int ticket = atomicAdd(&p.tickets[row * p.KVH + kvh], 1);
      *s_last = (ticket == p.n_splits - 1);
    }
    __syncthreads();
    if (*s_last) {
      __threadfence();
      combine_splits<GQ>(p, row, kvh, tid);
      if (tid == 0) p.tickets[row * p.KVH + kvh] = 0;
    }
    __syncthreads();
  }
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
  p.split0 = 0;
  // beam search: the key splits that lie entirely inside the (shared) prompt go through the tensor-core kernel once
  // for all beams of an input
  if (a.beams > 1 && gq == 4 && a.beams * gq <= SP_MAXR && g_skinny_mma) {
    const int n_shared = a.S / SP_KEYS;
    if (n_shared > 0) {
      static SmemOptIn opt;
      if (opt.need(SP_SMEM))
        PCY_CUDA(cudaFuncSetAttribute(decode_attn_shared_prompt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SP_SMEM));
      dim3 sgrid(n_shared, a.KVH, a.rows / a.beams);
      PCY_CUDA(launch_pdl(decode_attn_shared_prompt_kernel<4>, sgrid, dim3(SP_THREADS), (size_t)SP_SMEM, stream, p));
      PCY_LAUNCH_CHECK();
      p.split0 = n_shared;
    }
  }
  dim3 grid(p.n_splits - p.split0, a.KVH, a.rows);
  if (gq == 4) PCY_CUDA(launch_pdl(decode_attn_kernel<4>, grid, dim3(DA_THREADS), 0, stream, p));
  else if (gq == 1) PCY_CUDA(launch_pdl(decode_attn_kernel<1>, grid, dim3(DA_THREADS), 0, stream, p));
  else if (gq == 2) PCY_CUDA(launch_pdl(decode_attn_kernel<2>, grid, dim3(DA_THREADS), 0, stream, p));
  else if (gq == 8) PCY_CUDA(launch_pdl(decode_attn_kernel<8>, grid, dim3(DA_THREADS), 0, stream, p));
  else return set_error(PCY_ERR_UNSUPPORTED, "decode attention: H/KVH=%d unsupported (1,2,4,8)", gq);
  PCY_LAUNCH_CHECK();
  return 0;
}

int decode_attention_splits(int S, int max_gen) { return ceil_div(S + max_gen, DA_CHUNK); }
int64_t decode_attention_partial_floats(int rows, int H, int KVH, int S, int max_gen) {
  return (int64_t)rows * KVH * decode_attention_splits(S, max_gen) * (H / KVH) * (HD + 2);
}

}  // namespace pcy
