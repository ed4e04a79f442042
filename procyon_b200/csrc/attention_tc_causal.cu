// Causal (Llama prefill) attention on the sm_100a tensor cores, head_dim 128, grouped-query heads, left-pad key mask.
//
// Replaces HF LlamaAttention's eager  softmax(q k^T / sqrt(d) + causal + padding mask) v  over the whole prompt
// (procyon/model/pmc_llama.py:221-247, reached from LlamaPostTokenization.forward :571-584) for prompts of >= 128
// positions; shorter prompts keep the mma.sync kernel in attention.cu.
//
// One CTA = 128 queries of one (sequence, query head); 64 keys per step; two CTAs per SM (112 KB of shared memory and
// 256 TMEM columns each: two S buffers + O), so that one CTA's softmax runs under the other's MMAs.
//   warp 8, one thread   TMA loads (Q once: two 64-dim blocks of [128 x 128 B], 128B swizzle; K / V per step: two blocks of
//                        [64 x 128 B] each, 2 stages) and tcgen05.mma issue:  S = Q K^T  (M 128, N 64, K 128 = 8 UMMA_K
//                        over the two dim blocks)  and  O += P V  (M 128, N 2 x 64, K 64 keys; V consumed in place as an
//                        MN-major operand, one N = 64 MMA chain per dim block)
//   warps 0-7            softmax, two threads per query row (32 key columns each) straight out of TMEM; P goes back
//                        through shared memory as the K-major A operand of the second product
// O is NOT carried in registers: it accumulates in TMEM across all steps (use_acc) and is rescaled there, by the row's
// two threads, only in steps where the running row maximum grows (tcgen05.ld -> scale -> tcgen05.st between the
// completion of P V (j-1) and the release of P(j)) - with 64 output columns per thread a register copy would not fit
// the 113-register budget of two CTAs per SM.
// Causality: a tile of queries [q0, q0 + 128) needs the key steps up to ceil((q0 + 128) / 64); inside the last two the
// mask is per element (key <= query), folded into the same 32-bit validity word as the padding mask; heavy tiles
// (late queries) are scheduled first.  The last query tile is shifted back to end at S (rows it repeats are written
// twice with identical values), so every S >= 128 is covered without a tail kernel.
#include "common.cuh"
#include "ops.h"

namespace pcy {

bool g_llama_tc_attention = true;  // pcy_set_llama_tc_attention(0): mma.sync kernel for every prefill (tests, A/B)

namespace {

constexpr int CBM = 128;   // queries per CTA
constexpr int CBN = 64;    // keys per step
constexpr int CHD = 128;   // head dim
constexpr int C_SM_WARPS = 8;
constexpr int C_THREADS = (C_SM_WARPS + 1) * 32;
constexpr int QBLK = CBM * 128;          // one 64-dim block of the Q tile: 128 rows x 128 B
constexpr int KVBLK = CBN * 128;         // one 64-dim block of a K / V tile: 64 rows x 128 B
constexpr int Q_BYTES_C = 2 * QBLK;      // 32 KB
constexpr int KV_BYTES_C = 2 * KVBLK;    // 16 KB per tensor and stage
constexpr int P_BYTES_C = CBM * 128;     // 16 KB: [128 rows][64 keys] bf16
constexpr int C_STAGES = 2;
constexpr int C_SMEM = Q_BYTES_C + 2 * C_STAGES * KV_BYTES_C + P_BYTES_C + 2 * CBM * 2 /*row max exchange (bf16)*/ +
                       128 /*barriers*/;
static_assert(2 * CBM * 4 <= P_BYTES_C, "the final row-sum exchange reuses the P region");
static_assert(2 * (C_SMEM + 1024) <= 228 * 1024, "two CTAs per SM must fit");
constexpr int C_TMEM_COLS = 256;  // S (two buffers): [0, 64) and [64, 128), O: [128, 256)

struct CausalParams {
  bf16* o;
  int64_t o_rs;
  const uint8_t* key_valid;  // [B][S] or null
  int B, H, KVH, S, n_q_tiles;
  float scale_log2;
};

__global__ void __launch_bounds__(C_THREADS, 2)
llama_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                          const CausalParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();  // 128B-swizzled tiles need 1024-byte alignment
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Q_BYTES_C;                   // C_STAGES x [2 blocks]
  const uint32_t sV = sK + C_STAGES * KV_BYTES_C;
  const uint32_t sP = sV + C_STAGES * KV_BYTES_C;
  const uint32_t sX = sP + P_BYTES_C;                   // bf16 [2 halves][128 rows]
  const uint32_t bars = sX + 2 * CBM * 2;
  const uint32_t q_full = bars, kv_full0 = bars + 8, kv_empty0 = bars + 24, s_full0 = bars + 40, p_ready = bars + 56,
                 o_full = bars + 64, tmem_slot = bars + 72;
  __nv_bfloat16* xchg = reinterpret_cast<__nv_bfloat16*>(smem_raw + (sX - base));
  float* xchg_f = reinterpret_cast<float*>(smem_raw + (sP - base));  // P region, free after the last P V

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = p.n_q_tiles - 1 - (int)blockIdx.x;  // late (long) tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KVH);
  const int q0 = min(q_tile * CBM, p.S - CBM);
  const int row_base = b * p.S;
  const int n_kv = min((q0 + CBM + CBN - 1) / CBN, (p.S + CBN - 1) / CBN);
  const int col_q = h * CHD, col_k = (p.H + kvh) * CHD, col_v = (p.H + p.KVH + kvh) * CHD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(q_full, 1);
    mbar_init(kv_full0, 1);
    mbar_init(kv_full0 + 8, 1);
    mbar_init(kv_empty0, 1);
    mbar_init(kv_empty0 + 8, 1);
    mbar_init(s_full0, 1);
    mbar_init(s_full0 + 8, 1);
    mbar_init(p_ready, C_SM_WARPS * 32);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == C_SM_WARPS) {
    tmem_alloc(tmem_slot, C_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == C_SM_WARPS) {
    if (lane == 0) {
      // ---------------- TMA producer + MMA issuer (one thread) ----------------
      mbar_arrive_expect_tx(q_full, Q_BYTES_C);
      tma_load_2d(sQ, &tmap_q, q_full, col_q, row_base + q0);
      tma_load_2d(sQ + QBLK, &tmap_q, q_full, col_q + 64, row_base + q0);
      auto load_kv = [&](int j) {
        const int st = j & 1;
        const uint32_t bar = kv_full0 + 8 * st;
        mbar_arrive_expect_tx(bar, 2 * KV_BYTES_C);
        const int r0 = row_base + j * CBN;
        tma_load_2d(sK + st * KV_BYTES_C, &tmap_kv, bar, col_k, r0);
        tma_load_2d(sK + st * KV_BYTES_C + KVBLK, &tmap_kv, bar, col_k + 64, r0);
        tma_load_2d(sV + st * KV_BYTES_C, &tmap_kv, bar, col_v, r0);
        tma_load_2d(sV + st * KV_BYTES_C + KVBLK, &tmap_kv, bar, col_v + 64, r0);
      };
      // S is double-buffered in TMEM: S(j + 1) is issued BEFORE the wait for softmax(j), so the softmax of consecutive
      // steps runs back to back and the MMAs / TMA loads hide under it (with one S buffer a step was the serial chain
      // S -> softmax -> P V, 2.9 us per step; the longest tile of S = 1024 has 16 steps)
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(kv_full0 + 8 * st, (j >> 1) & 1);
        tc_fence_after();
        const int keys = min(CBN, p.S - j * CBN);
        const int n_mma = (keys + 15) & ~15;
        const uint32_t idesc_s = make_idesc_bf16(CBM, n_mma);
#pragma unroll
        for (int k = 0; k < CHD / 16; ++k) {
          const uint64_t qd = make_desc_kmajor_sw128(sQ + (k >> 2) * QBLK) + 2 * (k & 3);
          const uint64_t kd = make_desc_kmajor_sw128(sK + st * KV_BYTES_C + (k >> 2) * KVBLK) + 2 * (k & 3);
          tc_mma_bf16(tmem_base + st * CBN, qd, kd, idesc_s, k > 0 ? 1u : 0u);
        }
        tc_commit(s_full0 + 8 * st);
      };
      load_kv(0);
      if (n_kv > 1) load_kv(1);
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        // (S buffer (j + 1) & 1 was last read by softmax(j - 1), whose p_ready this thread has already waited for)
        if (j + 1 < n_kv) issue_s(j + 1);
        // O += P V : M = 128, N = 64 per dim block, K = n_mma keys; A = P (K-major), B = V (MN-major)
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const int keys = min(CBN, p.S - j * CBN);
        const int n_mma = (keys + 15) & ~15;
        const uint32_t idesc_o = make_idesc_bf16(CBM, 64, 0, 1);
        for (int k = 0; k < n_mma / 16; ++k) {
          const uint64_t pd = make_desc_kmajor_sw128(sP) + 2 * k;
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            const uint64_t vd = make_desc_mnmajor_sw128(sV + st * KV_BYTES_C + nb * KVBLK + k * 2048, 1024);
            tc_mma_bf16(tmem_base + 2 * CBN + nb * 64, pd, vd, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          }
        }
        tc_commit(kv_empty0 + 8 * st);
        tc_commit(o_full);
        if (j + 2 < n_kv) {  // K / V of step j + 2 into this stage as soon as P V (j) has retired
          mbar_wait(kv_empty0 + 8 * st, (j >> 1) & 1);
          load_kv(j + 2);
        }
      }
    }
  } else {
    // ---------------- softmax: two threads per query row ----------------
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;  // query row within the tile = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int q_pos = q0 + r;
    float m_run = -INFINITY, l_run = 0.f;
    const uint8_t* valid_g = p.key_valid ? p.key_valid + (int64_t)b * p.S : nullptr;
    for (int j = 0; j < n_kv; ++j) {
      const int k_lo = j * CBN + half * 32;  // first key of this thread's 32 columns
      // validity of the warp's 32 keys (inside the sequence, not padding): one ballot; then the row's causal limit
      bool ok = k_lo + lane < p.S;
      if (ok && valid_g) ok = valid_g[k_lo + lane] != 0;
      uint32_t mw = __ballot_sync(0xffffffffu, ok);
      const int n_vis = q_pos - k_lo + 1;  // keys k_lo .. k_lo + n_vis - 1 are not in this query's future
      mw &= n_vis >= 32 ? 0xffffffffu : (n_vis <= 0 ? 0u : ((1u << n_vis) - 1u));
      mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_lane + (j & 1) * CBN + half * 32, v);
      tc_wait_ld();
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (mw == 0xffffffffu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], ((mw >> i) & 1u) ? __uint_as_float(v[i]) : -INFINITY);
      }
      // the two threads of a row must agree on the offset exactly; any value >= the true maximum works, so the maxima
      // are exchanged rounded UP to bf16 (the two-CTA shared-memory budget leaves 1 KB beside the tiles)
      const __nv_bfloat16 mx_own_b = __float2bfloat16_ru(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])));
      asm volatile("bar.sync 1, 256;" ::: "memory");  // everybody has read the previous step's slots
      xchg[half * CBM + r] = mx_own_b;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float m_new = fmaxf(m_run, fmaxf(__bfloat162float(mx_own_b), __bfloat162float(xchg[(half ^ 1) * CBM + r])));
      const float corr = (m_new == -INFINITY) ? 1.f : exp2f((m_run - m_new) * p.scale_log2);
      const float moff = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
      if (j > 0) {
        // P V (j - 1) has finished: P and this K / V stage are free, O may be rescaled
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, corr != 1.f)) {
#pragma unroll
          for (int c = 0; c < 64; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(t_lane + 2 * CBN + half * 64 + c, o);
            tc_wait_ld();
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              lo[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
              hi[i] = __float_as_uint(__uint_as_float(o[16 + i]) * corr);
            }
            tmem_st_32x32b_x16(t_lane + 2 * CBN + half * 64 + c, lo);
            tmem_st_32x32b_x16(t_lane + 2 * CBN + half * 64 + c + 16, hi);
          }
          tc_wait_st();
        }
      }
      // p = exp2(s * scale - moff) -> bf16 -> swizzled row r of the P block (chunks 4 half .. 4 half + 3)
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0, p1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v[i]), p.scale_log2, -moff)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -moff)));
        p0 = ((mw >> i) & 1u) ? p0 : 0.f;
        p1 = ((mw >> (i + 1)) & 1u) ? p1 : 0.f;
        ls4[(i >> 1) & 3] += p0 + p1;
        packed[i >> 1] = pack_bf16x2(p0, p1);
      }
      const uint32_t prow = sP + (uint32_t)r * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = half * 4 + q;
        const uint32_t addr = prow + (uint32_t)((chunk ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(packed[4 * q]), "r"(packed[4 * q + 1]),
                     "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                     : "memory");
      }
      l_run = l_run * corr + ((ls4[0] + ls4[1]) + (ls4[2] + ls4[3]));
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // ---- finalize: O / (row sum of both halves) ----
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    xchg_f[half * CBM + r] = l_run;  // (every thread is past the last P V: the P region is free)
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_tot = l_run + xchg_f[(half ^ 1) * CBM + r];
    const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
    const bool write = q_pos < p.S && q_pos >= q_tile * CBM;  // rows below q_tile * CBM belong to the previous tile
    bf16* op = p.o + (int64_t)(row_base + q_pos) * p.o_rs + h * CHD + half * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_lane + 2 * CBN + half * 64 + c, o);
      tc_wait_ld();
      if (write) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c + i) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == C_SM_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C_TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D map over the packed qkv matrix [rows, cols] (bf16, row stride ld): boxes of 64 columns (128 B) x box_rows rows
int make_map(const bf16* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return set_error(PCY_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    enc = reinterpret_cast<EncodeTiledFn>(fp);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PCY_ERR_CUDA, "cuTensorMapEncodeTiled(llama qkv) failed (%d)", (int)r);
  return 0;
}

}  // namespace

// qkv bf16 [B*S, (H + 2 KVH) * 128] with RoPE applied to q and k; out bf16 [B*S, H*128] (row stride o_rs).
// Handles the whole prefill when S >= 128, head_dim 128 and H % KVH == 0 (*done = 1), otherwise leaves it to the
// caller's fallback (*done = 0).
int llama_attention_tc(const bf16* qkv, int64_t qkv_ld, const uint8_t* key_valid, bf16* out, int64_t o_rs, int B, int S,
                       int H, int KVH, int head_dim, float scale, int* done, cudaStream_t stream) {
  *done = 0;
  if (!g_llama_tc_attention || head_dim != CHD || S < CBM || KVH <= 0 || H % KVH != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) != 0 || qkv_ld % 8 != 0) return 0;
  static SmemOptIn opt;
  if (opt.need(C_SMEM))
    PCY_CUDA(cudaFuncSetAttribute(llama_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
  CUtensorMap tq, tkv;
  const int64_t cols = (int64_t)(H + 2 * KVH) * CHD;
  PCY_TRY(make_map(qkv, (int64_t)B * S, cols, qkv_ld, CBM, &tq));
  PCY_TRY(make_map(qkv, (int64_t)B * S, cols, qkv_ld, CBN, &tkv));
  CausalParams p;
  p.o = out; p.o_rs = o_rs; p.key_valid = key_valid; p.B = B; p.H = H; p.KVH = KVH; p.S = S;
  p.n_q_tiles = (S + CBM - 1) / CBM;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(p.n_q_tiles, H, B);
  llama_attention_tc_kernel<<<grid, C_THREADS, C_SMEM, stream>>>(tq, tkv, p);
  PCY_LAUNCH_CHECK();
  *done = 1;
  return 0;
}

}  // namespace pcy

extern "C" int pcy_set_llama_tc_attention(int enabled) {
  pcy::g_llama_tc_attention = enabled != 0;
  return 0;
}
