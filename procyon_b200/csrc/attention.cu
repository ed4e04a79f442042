// Flash-style multi-head attention for the ESM2 encoder (bidirectional, key-padding mask) and the Llama
// prefill (causal, left-pad key mask, grouped-query heads).  Scores never touch HBM: 64-query x 64-key tiles,
// K/V double-buffered in shared memory with cp.async, QK^T and PV on mma.sync bf16 tensor-core fragments with
// fp32 online softmax (exp2, running max / sum per row, warp-quad shuffles).
//
// Replaces fair-esm MultiheadAttention's bmm + fp32 softmax + bmm (reached via procyon/model/esm.py:536) and
// HF LlamaAttention's eager matmul/softmax (math shown in procyon/model/pmc_llama.py:221-247).
// Round-1 note: this kernel is 2-6 % of the path's FLOPs; a tcgen05/TMEM variant is the planned next step.
#include "common.cuh"
#include "ops.h"

namespace pcy {

namespace {

constexpr int ATT_BM = 64;
constexpr int ATT_BN = 64;
constexpr int ATT_THREADS = 128;

struct AttnParams {
  const bf16* q;
  const bf16* k;
  const bf16* v;
  bf16* o;
  int64_t q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;
  int q_hs, k_hs, v_hs, o_hs;
  int B, H, KVH, Tq, Tk, head_dim;
  const uint8_t* key_valid;  // [B, Tk] (1 = attend) or null
  int64_t kv_bs;
  float scale_log2;
  int causal_offset;  // key j is visible to query i iff j <= i + causal_offset
};

template <int HD>
__device__ __forceinline__ int swz_chunk(int row, int chunk) {
  if (HD == 32) return chunk ^ ((row >> 1) & 3);
  if (HD == 64) return chunk ^ (row & 7);
  return (chunk & ~7) | ((chunk & 7) ^ (row & 7));
}
template <int HD>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return (uint32_t)(row * (HD * 2) + swz_chunk<HD>(row, chunk) * 16);
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// copy a [64 x HD] tile (rows r0.., at most n_rows valid, hd valid columns) global -> swizzled smem
template <int HD>
__device__ __forceinline__ void load_tile(uint32_t smem_tile, const bf16* base, int64_t row_stride, int r0,
                                          int n_rows, int hd) {
  constexpr int CH = HD / 8;
  for (int idx = threadIdx.x; idx < 64 * CH; idx += ATT_THREADS) {
    const int row = idx / CH, chunk = idx % CH;
    const int grow = r0 + row;
    const bool ok = (grow < n_rows) && (chunk * 8 < hd);
    const bf16* src = base + (ok ? ((int64_t)grow * row_stride + chunk * 8) : 0);
    cp_async_16(smem_tile + tile_off<HD>(row, chunk), src, ok ? 16 : 0);
  }
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS)
flash_attn_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  constexpr int TILE_BYTES = 64 * HD * 2;
  const uint32_t sQ = smem_u32(att_smem);
  const uint32_t sK = sQ + TILE_BYTES;
  const uint32_t sV = sK + 2 * TILE_BYTES;

  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KVH);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int q0 = q_tile * ATT_BM;

  const bf16* qb = p.q + b * p.q_bs + (int64_t)h * p.q_hs;
  const bf16* kb = p.k + b * p.k_bs + (int64_t)kvh * p.k_hs;
  const bf16* vb = p.v + b * p.v_bs + (int64_t)kvh * p.v_hs;
  const uint8_t* valid = p.key_valid ? p.key_valid + b * p.kv_bs : nullptr;

  int k_end = p.Tk;
  if (CAUSAL) k_end = min(p.Tk, q0 + ATT_BM + p.causal_offset);  // keys >= this are masked for every row here
  const int n_tiles = k_end > 0 ? (k_end + ATT_BN - 1) / ATT_BN : 0;

  load_tile<HD>(sQ, qb, p.q_rs, q0, p.Tq, p.head_dim);
  cp_async_commit();
  if (n_tiles > 0) {
    load_tile<HD>(sK, kb, p.k_rs, 0, p.Tk, p.head_dim);
    load_tile<HD>(sV, vb, p.v_rs, 0, p.Tk, p.head_dim);
  }
  cp_async_commit();

  uint32_t qf[HD / 16][4];
  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};

  for (int j = 0; j < n_tiles; ++j) {
    const int stage = j & 1;
    if (j + 1 < n_tiles) {
      load_tile<HD>(sK + (stage ^ 1) * TILE_BYTES, kb, p.k_rs, (j + 1) * ATT_BN, p.Tk, p.head_dim);
      load_tile<HD>(sV + (stage ^ 1) * TILE_BYTES, vb, p.v_rs, (j + 1) * ATT_BN, p.Tk, p.head_dim);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = ks * 2 + (lane >> 4);
        ldmatrix_x4(sQ + tile_off<HD>(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t sKs = sK + stage * TILE_BYTES;
    const uint32_t sVs = sV + stage * TILE_BYTES;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[i][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int mid = lane >> 3;
        const int row = np * 16 + (lane & 7) + (mid >> 1) * 8;
        const int chunk = ks * 2 + (mid & 1);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(sKs + tile_off<HD>(row, chunk), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
        mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
      }
    }

    // ---- mask ----
    const int kbase = j * ATT_BN;
    const bool need_mask = (kbase + ATT_BN > p.Tk) || (valid != nullptr) ||
                           (CAUSAL && (kbase + ATT_BN - 1 > q0 + p.causal_offset));
    if (need_mask) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int kidx = kbase + nt * 8 + tig * 2 + c;
          bool ok = kidx < p.Tk;
          if (ok && valid) ok = valid[kidx] != 0;
          bool ok0 = ok, ok1 = ok;
          if (CAUSAL) {
            ok0 = ok0 && (kidx <= qrow[0] + p.causal_offset);
            ok1 = ok1 && (kidx <= qrow[1] + p.causal_offset);
          }
          if (!ok0) s[nt][c] = -INFINITY;
          if (!ok1) s[nt][2 + c] = -INFINITY;
        }
      }
    }

    // ---- online softmax ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float corr[2], moff[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      if (m_new == -INFINITY) { corr[r] = 1.f; moff[r] = 0.f; }
      else { corr[r] = exp2f((m_run[r] - m_new) * p.scale_log2); moff[r] = m_new * p.scale_log2; }
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] * p.scale_log2 - moff[0]);
      const float p1 = exp2f(s[nt][1] * p.scale_log2 - moff[0]);
      const float p2 = exp2f(s[nt][2] * p.scale_log2 - moff[1]);
      const float p3 = exp2f(s[nt][3] * p.scale_log2 - moff[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
    }

    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {  // 16 keys per step
#pragma unroll
      for (int np = 0; np < HD / 16; ++np) {  // pairs of 8-wide output column tiles
        const int mid = lane >> 3;
        const int row = ks * 16 + (lane & 7) + (mid & 1) * 8;
        const int chunk = np * 2 + (mid >> 1);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(sVs + tile_off<HD>(row, chunk), b0, b1, b2, b3);
        mma_bf16_16816(o_acc[2 * np], pf[ks], b0, b1);
        mma_bf16_16816(o_acc[2 * np + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();
  }

  // ---- finalize: O /= l, stage through smem (Q region), coalesced 16-byte stores ----
  if (n_tiles == 0) {
    cp_async_wait<0>();
    __syncthreads();
  }
  float inv_l[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    inv_l[r] = (l > 0.f) ? 1.f / l : 0.f;
  }
#pragma unroll
  for (int nt = 0; nt < HD / 8; ++nt) {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const uint32_t a0 = sQ + tile_off<HD>(r0, nt) + tig * 4;
    const uint32_t a1 = sQ + tile_off<HD>(r1, nt) + tig * 4;
    const uint32_t v0 = pack_bf16x2(o_acc[nt][0] * inv_l[0], o_acc[nt][1] * inv_l[0]);
    const uint32_t v1 = pack_bf16x2(o_acc[nt][2] * inv_l[1], o_acc[nt][3] * inv_l[1]);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a0), "r"(v0) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a1), "r"(v1) : "memory");
  }
  __syncthreads();
  bf16* ob = p.o + b * p.o_bs + (int64_t)h * p.o_hs;
  constexpr int CH = HD / 8;
  for (int idx = threadIdx.x; idx < 64 * CH; idx += ATT_THREADS) {
    const int row = idx / CH, chunk = idx % CH;
    const int grow = q0 + row;
    if (grow < p.Tq && chunk * 8 < p.head_dim) {
      uint4 u;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                   : "r"(sQ + tile_off<HD>(row, chunk)));
      *reinterpret_cast<uint4*>(ob + (int64_t)grow * p.o_rs + chunk * 8) = u;
    }
  }
}

template <int HD, bool CAUSAL>
int launch_attn(const AttnParams& p, cudaStream_t stream) {
  constexpr int SMEM = 5 * 64 * HD * 2;
  static SmemOptIn opt;
  if (SMEM > 48 * 1024 && opt.need(SMEM))
    PCY_CUDA(cudaFuncSetAttribute(flash_attn_kernel<HD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  dim3 grid(ceil_div(p.Tq, ATT_BM), p.H, p.B);
  flash_attn_kernel<HD, CAUSAL><<<grid, ATT_THREADS, SMEM, stream>>>(p);
  PCY_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int flash_attention(const AttnArgs& a, cudaStream_t stream) {
  PCY_REQUIRE(a.head_dim % 8 == 0 && a.head_dim >= 8 && a.head_dim <= 128, "attention: head_dim=%d unsupported",
              a.head_dim);
  PCY_REQUIRE(a.H % a.KVH == 0, "attention: H=%d not a multiple of KVH=%d", a.H, a.KVH);
  PCY_REQUIRE(a.q_rs % 8 == 0 && a.k_rs % 8 == 0 && a.v_rs % 8 == 0 && a.o_rs % 8 == 0 && a.q_hs % 8 == 0 &&
                  a.k_hs % 8 == 0 && a.v_hs % 8 == 0 && a.o_hs % 8 == 0,
              "attention: strides must be multiples of 8 elements");
  if (a.B == 0 || a.Tq == 0) return 0;
  AttnParams p;
  p.q = a.q; p.k = a.k; p.v = a.v; p.o = a.o;
  p.q_bs = a.q_bs; p.q_rs = a.q_rs; p.k_bs = a.k_bs; p.k_rs = a.k_rs; p.v_bs = a.v_bs; p.v_rs = a.v_rs;
  p.o_bs = a.o_bs; p.o_rs = a.o_rs; p.q_hs = a.q_hs; p.k_hs = a.k_hs; p.v_hs = a.v_hs; p.o_hs = a.o_hs;
  p.B = a.B; p.H = a.H; p.KVH = a.KVH; p.Tq = a.Tq; p.Tk = a.Tk; p.head_dim = a.head_dim;
  p.key_valid = a.key_valid; p.kv_bs = a.key_valid_bs;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.causal_offset = a.Tk - a.Tq;
  const int hd = a.head_dim <= 32 ? 32 : a.head_dim <= 64 ? 64 : 128;
  if (a.causal) {
    if (hd == 32) return launch_attn<32, true>(p, stream);
    if (hd == 64) return launch_attn<64, true>(p, stream);
    return launch_attn<128, true>(p, stream);
  }
  if (hd == 32) return launch_attn<32, false>(p, stream);
  if (hd == 64) return launch_attn<64, false>(p, stream);
  return launch_attn<128, false>(p, stream);
}

}  // namespace pcy
