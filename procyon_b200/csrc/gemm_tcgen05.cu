// bf16 GEMM on the sm_100a 5th-gen tensor cores: TMA-staged 128B-swizzled tiles -> tcgen05.mma
// (accumulator in TMEM, double buffered) -> tcgen05.ld epilogue with fused bias / scale / GELU /
// SwiGLU / residual.  Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer
// (+ TMEM allocator), warps 2..9 = epilogue (two per TMEM lane quadrant).
//
// CL = 2: CTAs run as clusters of two that work on vertically adjacent output tiles (same columns).  The W tile they
// share is fetched once from L2: each CTA loads half of it with TMA multicast into both CTAs' shared memory.  A
// 128x256 tile per SM needs (128 + 256) x 64 x 2 B per 64-deep k-block, i.e. ~16 TB/s of L2 -> SM traffic at full
// tensor rate, which is what bounds the one-CTA kernel on the ESM2 / Llama shapes; sharing W cuts it by a third.
//
// U2 (with CL = 2): the pair runs ONE tcgen05.mma.cta_group::2 of M = 256 per k-step.  Each CTA stages only its 128
// rows of A and HALF of the W tile (32 KB instead of 48 KB per k-block for BN = 256, so the ring holds 6 stages
// instead of 4), both CTAs' TMA loads count on the leader's full barrier, the leader's single thread issues the MMAs
// for both SMs and its commits arrive on both CTAs' empty / accumulator-full barriers; each CTA's epilogue warps drain
// their own 128 accumulator rows and release the accumulator on the leader's barrier.
//
// Replaces, on the hot path, every nn.Linear of fair-esm ESM2 (q/k/v/out_proj, fc1, fc2 — reached via
// procyon/model/esm.py:536), of HF LlamaDecoderLayer (q/k/v/o_proj, gate/up/down_proj — reached via
// procyon/model/pmc_llama.py:571) and of create_mlp (procyon/model/model_utils.py:13-41).
#include <algorithm>
#include <mutex>
#include <cstdlib>
#include <unordered_map>

#include "common.cuh"
#include "ops.h"

namespace pcy {

int g_gemm_force_tile = [] { const char* e = getenv("PCY_GEMM_FORCE_TILE"); return e ? atoi(e) : 0; }();
bool g_gemm_cluster = false;  // pcy_set_gemm_cluster(1): 2-CTA clusters sharing the W tile by TMA multicast (measured: no gain)
// pcy_set_gemm_pair_mma: 2-CTA clusters issuing cta_group::2 MMAs (M = 256 per pair).  0 = never, 1 = when the problem
// has at least three waves of tiles (default), 2 = whenever there are two row-blocks (tests)
int g_gemm_pair_mma = 1;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle span
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;  // two per TMEM lane quadrant: each takes half of the tile's columns
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;

template <int BN, bool U2 = false>
struct TileCfg {
  static constexpr int STAGES = U2 ? ((BN == 256) ? 6 : 8) : ((BN >= 192) ? 4 : 6);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (U2 ? BN / 2 : BN) * BK * 2;  // U2: this CTA's half of the W tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // two accumulator buffers at columns 0 and BN; the allocation must be a power of two (BN = 192: 384 -> 512)
  static constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * 32 * 128;  // per warp: 32 rows x 64 bf16 columns
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct EpiParams {
  void* C;
  int64_t ldc;
  const float* bias;
  const bf16* residual;
  int64_t ldr;
  int M, N, K;
  int c_fp32;
  int act;
  float scale;
  int scale_ncols;
  const float* rope;  // fp32 [P][rope_hd/2][2] or null: rotate-half RoPE on columns < rope_ncols, position = row % rope_T
  int rope_hd, rope_T, rope_ncols;
};

// v += bias; v *= scale on the leading columns
__device__ __forceinline__ void epi_bias_scale(float (&v)[32], int col0, const EpiParams& p) {
  const bool full_cols = (col0 + 32 <= p.N);
  if (p.bias != nullptr) {
    if (full_cols) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  if (col0 < p.scale_ncols) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < p.scale_ncols) v[j] *= p.scale;
  }
}

// The same step on PAIRS of columns (FADD2 / FMUL2), for the coalesced store path, whose 32-column chunks are always
// complete (col0 + 32 <= N).  Bit-identical to epi_bias_scale lane by lane.
__device__ __forceinline__ void epi_bias_scale2(uint64_t (&v)[16], int col0, const EpiParams& p) {
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 2 * j));
      v[j] = f2_add(v[j], f2_pack(b.x, b.y));
      v[j + 1] = f2_add(v[j + 1], f2_pack(b.z, b.w));
    }
  }
  if (col0 + 32 <= p.scale_ncols) {
    const uint64_t sc = f2_bcast(p.scale);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = f2_mul(v[j], sc);
  } else if (col0 < p.scale_ncols) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float a, b;
      f2_unpack(v[j], a, b);
      if (col0 + 2 * j < p.scale_ncols) a *= p.scale;
      if (col0 + 2 * j + 1 < p.scale_ncols) b *= p.scale;
      v[j] = f2_pack(a, b);
    }
  }
}

__device__ __forceinline__ void epi_store_bf16(const float (&v)[32], int row, int col0, const EpiParams& p) {
  bf16* cp = reinterpret_cast<bf16*>(p.C) + (int64_t)row * p.ldc + col0;
  if ((col0 + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 u;
      u.x = pack_bf16x2(v[j], v[j + 1]);
      u.y = pack_bf16x2(v[j + 2], v[j + 3]);
      u.z = pack_bf16x2(v[j + 4], v[j + 5]);
      u.w = pack_bf16x2(v[j + 6], v[j + 7]);
      *reinterpret_cast<uint4*>(cp + j) = u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < p.N) cp[j] = __float2bfloat16_rn(v[j]);
  }
}

// Tile order: bands of GROUP_M row-blocks, row-block fastest inside a band. The ~148 tiles in flight then cover a
// compact (GROUP_M x ~18) patch of the output, so every A and B tile fetched from HBM is reused out of L2 by the
// other tiles of the patch (a plain row-major order re-reads A once per column block when A does not fit in L2).
constexpr int GROUP_M = 8;
__device__ __forceinline__ void tile_coords(int tile, int m_blocks, int n_blocks, int& m_blk, int& n_blk) {
  const int per_group = GROUP_M * n_blocks;
  const int g = tile / per_group;
  const int first_m = g * GROUP_M;
  const int rows = min(GROUP_M, m_blocks - first_m);
  const int local = tile - g * per_group;
  m_blk = first_m + local % rows;
  n_blk = local / rows;
}

// pair order for CL = 2: the same banding over PAIRS of row-blocks
__device__ __forceinline__ void pair_coords(int pair, int m_pairs, int n_blocks, int& m_pair, int& n_blk) {
  constexpr int GROUP_P = GROUP_M / 2;
  const int per_group = GROUP_P * n_blocks;
  const int g = pair / per_group;
  const int first = g * GROUP_P;
  const int rows = min(GROUP_P, m_pairs - first);
  const int local = pair - g * per_group;
  m_pair = first + local % rows;
  n_blk = local / rows;
}

template <int BN, bool ROPE, int CL, bool U2 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const EpiParams p) {
  static_assert(!U2 || CL == 2, "cta_group::2 needs the 2-CTA cluster");
  using Cfg = TileCfg<BN, U2>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = staging_base + Cfg::STAGING_BYTES;
  // barrier layout (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (p.M + BM - 1) / BM;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int k_blocks = (p.K + BK - 1) / BK;
  // work items: tiles (CL = 1) or vertical tile pairs shared by the two CTAs of a cluster (CL = 2; an odd last
  // row-block is paired with an all-padding one, whose rows the epilogue masks like any M tail)
  const int m_units = (m_blocks + CL - 1) / CL;
  const int num_tiles = m_units * n_blocks;
  const int unit0 = blockIdx.x / CL, unit_step = gridDim.x / CL;
  const uint32_t cta_rank = (CL == 2) ? cluster_ctarank() : 0u;
  auto coords = [&](int unit, int& m_blk, int& n_blk) {
    if (CL == 1) {
      tile_coords(unit, m_blocks, n_blocks, m_blk, n_blk);
    } else {
      int m_pair;
      pair_coords(unit, m_units, n_blocks, m_pair, n_blk);
      m_blk = 2 * m_pair + (int)cta_rank;
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      // a slot is refilled (in both CTAs) once both CTAs' MMAs have read it; U2: one pair-wide commit per use
      mbar_init(empty_bar(s), U2 ? 1 : CL);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      // U2: the leader's MMAs write both CTAs' accumulators, so both CTAs' epilogue warps release on the leader's
      mbar_init(tempty_bar(s), U2 ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (U2) {
    __syncthreads();
    cluster_sync_all();  // both CTAs are resident (and their barriers exist) before the pair-wide allocation
  }
  if (warp == 1) {
    if (U2) {
      tmem_alloc_pair(tmem_ptr_smem, Cfg::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // the peer's barriers exist before anything can arrive on them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        int m_blk, n_blk;
        coords(tile, m_blk, n_blk);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          if (U2) {
            // both CTAs' boxes are counted on the leader's barrier (the only one the MMA thread waits on)
            if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
            tma_load_2d_pair(sa, &tmap_a, full_bar(stage), kb * BK, m_blk * BM);
            tma_load_2d_pair(sb, &tmap_b, full_bar(stage), kb * BK, n_blk * BN + (int)cta_rank * (BN / 2));
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BK, m_blk * BM);
          if (CL == 1) {
            tma_load_2d(sb, &tmap_b, full_bar(stage), kb * BK, n_blk * BN);
          } else {
            // my half of the shared W tile, delivered to both CTAs (the other half arrives from the peer)
            tma_load_2d_mc(sb + cta_rank * (Cfg::B_BYTES / 2), &tmap_b, full_bar(stage), kb * BK,
                           n_blk * BN + (int)cta_rank * (BN / 2), (uint16_t)0x3);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && !(U2 && cta_rank != 0)) {  // U2: the leader issues for the pair
      constexpr uint32_t idesc = make_idesc_bf16(U2 ? 2 * BM : BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t a_desc = make_desc_kmajor_sw128(sa);
          const uint64_t b_desc = make_desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in the (addr >> 4) field
            if (U2)
              tc_mma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                               (kb > 0 || k > 0) ? 1u : 0u);
            else
              tc_mma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot once these MMAs retire (in both CTAs: either one's next load writes into both)
          if (CL == 1) tc_commit(empty_bar(stage));
          else if (U2) tc_commit_pair(empty_bar(stage), (uint16_t)0x3);
          else tc_commit_mc(empty_bar(stage), (uint16_t)0x3);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (U2: of both CTAs)
        if (U2) tc_commit_pair(tfull_bar(acc), (uint16_t)0x3);
        else tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int chalf = (warp - 2) >> 2;  // which half of the tile's columns this warp drains
    constexpr int C_PER = BN / 32 / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool swiglu = (p.act == ACT_SWIGLU);
    // bf16 output with 16-byte aligned rows (and residual rows): eligible for the coalesced smem-staged path
    const bool fast_store = !swiglu && !p.c_fp32 && (p.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            (p.residual == nullptr ||
                             ((p.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)));
    for (int tile = unit0; tile < num_tiles; tile += unit_step) {
      int m_blk, n_blk;
      coords(tile, m_blk, n_blk);
      const int row = m_blk * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = chalf * C_PER; c < (chalf + 1) * C_PER; ++c) {
        const int col0 = n_blk * BN + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        if (ROPE && p.rope != nullptr && col0 < p.rope_ncols) {
          // fused rotate-half RoPE: columns (i, i + hd/2) of a head are in chunks c and c + hd/64
          const int half = p.rope_hd >> 1;
          const int in_head = col0 % p.rope_hd;
          if (in_head >= half) continue;  // second-half chunk: written together with its partner
          uint32_t ra[32], rb[32];
          tmem_ld_32x32b_x32(t_addr + (uint32_t)(c * 32), ra);
          tmem_ld_32x32b_x32(t_addr + (uint32_t)(c * 32 + half), rb);
          tc_wait_ld();
          float lo[32], hi[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { lo[j] = __uint_as_float(ra[j]); hi[j] = __uint_as_float(rb[j]); }
          epi_bias_scale(lo, col0, p);
          epi_bias_scale(hi, col0 + half, p);
          if (row_ok) {
            const int pos = row % p.rope_T;
            const float2* cs = reinterpret_cast<const float2*>(p.rope) + (int64_t)pos * half + in_head;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float2 t = __ldg(cs + j);
              // the un-fused path stores q/k in bf16 before rotating them: keep that rounding point
              const float a = bf16_round(lo[j]), b = bf16_round(hi[j]);
              lo[j] = a * t.x - b * t.y;
              hi[j] = b * t.x + a * t.y;
            }
            epi_store_bf16(lo, row, col0, p);
            epi_store_bf16(hi, row, col0 + half, p);
          }
          continue;
        }
        if (fast_store && (c & 1) == 0 && c + 1 < (chalf + 1) * C_PER && col0 + 64 <= p.N) {
          // ---- coalesced path: 64 columns (two chunks) of this warp's 32 rows go through a swizzled smem block so
          // that every global load / store instruction moves four full 128-byte row segments ----
          const uint32_t stg = staging_base + (uint32_t)(warp - 2) * (32 * 128);
          const int row0 = m_blk * BM + quad * 32;
          if (p.residual != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = (lane >> 3) + 4 * i, ch = lane & 7;
              uint4 u = make_uint4(0, 0, 0, 0);
              if (row0 + rr < p.M)
                u = *reinterpret_cast<const uint4*>(p.residual + (int64_t)(row0 + rr) * p.ldr + col0 + ch * 8);
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg + rr * 128 + ((ch ^ (rr & 7)) << 4)),
                           "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
            }
            __syncwarp();
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r2[32];
            tmem_ld_32x32b_x32(t_addr + (uint32_t)((c + hf) * 32), r2);
            tc_wait_ld();
            // bias, scale and GELU on 16 PAIRS of columns: two fp32 lanes per FMA-pipe instruction (FFMA2), on which this
            // epilogue is bound under the fc1 GEMM of the ESM2 encoder
            uint64_t vp[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) vp[j] = f2_pack(__uint_as_float(r2[2 * j]), __uint_as_float(r2[2 * j + 1]));
            epi_bias_scale2(vp, col0 + hf * 32, p);
            if (p.act == ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) vp[j] = gelu_erf2(vp[j]);
            }
            float v2[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) f2_unpack(vp[j], v2[2 * j], v2[2 * j + 1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t addr = stg + lane * 128 + (((hf * 4 + q) ^ (lane & 7)) << 4);
              if (p.residual != nullptr) {
                uint4 u;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                             : "r"(addr));
                const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
                             f3 = unpack_bf16x2(u.w);
                v2[8 * q] += f0.x; v2[8 * q + 1] += f0.y; v2[8 * q + 2] += f1.x; v2[8 * q + 3] += f1.y;
                v2[8 * q + 4] += f2.x; v2[8 * q + 5] += f2.y; v2[8 * q + 6] += f3.x; v2[8 * q + 7] += f3.y;
              }
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pack_bf16x2(v2[8 * q], v2[8 * q + 1])),
                           "r"(pack_bf16x2(v2[8 * q + 2], v2[8 * q + 3])), "r"(pack_bf16x2(v2[8 * q + 4], v2[8 * q + 5])),
                           "r"(pack_bf16x2(v2[8 * q + 6], v2[8 * q + 7])) : "memory");
            }
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = (lane >> 3) + 4 * i, ch = lane & 7;
            if (row0 + rr < p.M) {
              uint4 u;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                           : "r"(stg + rr * 128 + ((ch ^ (rr & 7)) << 4)));
              *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.C) + (int64_t)(row0 + rr) * p.ldc + col0 + ch * 8) = u;
            }
          }
          __syncwarp();
          ++c;  // the partner chunk is done
          continue;
        }
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_addr + (uint32_t)(c * 32), r);
        tc_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const bool full_cols = (col0 + 32 <= p.N);
        if (p.bias != nullptr) {
          if (full_cols) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
          }
        }
        if (col0 < p.scale_ncols) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.scale_ncols) v[j] *= p.scale;
        }
        if (!swiglu) {
          if (p.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          }
          if (row_ok) {
            if (p.residual != nullptr) {
              const bf16* rp = p.residual + (int64_t)row * p.ldr + col0;
              if (full_cols && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  const uint4 u = *reinterpret_cast<const uint4*>(rp + j);
                  const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
                               f3 = unpack_bf16x2(u.w);
                  v[j] += f0.x; v[j + 1] += f0.y; v[j + 2] += f1.x; v[j + 3] += f1.y;
                  v[j + 4] += f2.x; v[j + 5] += f2.y; v[j + 6] += f3.x; v[j + 7] += f3.y;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) v[j] += __bfloat162float(rp[j]);
              }
            }
            if (p.c_fp32) {
              float* cp = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col0;
              if (full_cols && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) cp[j] = v[j];
              }
            } else {
              bf16* cp = reinterpret_cast<bf16*>(p.C) + (int64_t)row * p.ldc + col0;
              if (full_cols && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 u;
                  u.x = pack_bf16x2(v[j], v[j + 1]);
                  u.y = pack_bf16x2(v[j + 2], v[j + 3]);
                  u.z = pack_bf16x2(v[j + 4], v[j + 5]);
                  u.w = pack_bf16x2(v[j + 6], v[j + 7]);
                  *reinterpret_cast<uint4*>(cp + j) = u;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) cp[j] = __float2bfloat16_rn(v[j]);
              }
            }
          }
        } else {
          // SwiGLU: columns [0,16) of the chunk are gate, [16,32) are up -> 16 outputs at col0/2.
          // N is a multiple of 32 in this mode (checked on the host).
          if (row_ok) {
            const int ocol0 = col0 >> 1;
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = silu(v[j]) * v[16 + j];
            if (p.residual != nullptr) {
              const bf16* rp = p.residual + (int64_t)row * p.ldr + ocol0;
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] += __bfloat162float(rp[j]);
            }
            bf16* cp = reinterpret_cast<bf16*>(p.C) + (int64_t)row * p.ldc + ocol0;
            if ((reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                uint4 u;
                u.x = pack_bf16x2(o[j], o[j + 1]);
                u.y = pack_bf16x2(o[j + 2], o[j + 3]);
                u.z = pack_bf16x2(o[j + 4], o[j + 5]);
                u.w = pack_bf16x2(o[j + 6], o[j + 7]);
                *reinterpret_cast<uint4*>(cp + j) = u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) cp[j] = __float2bfloat16_rn(o[j]);
            }
          }
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above) -> release the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (U2) mbar_arrive_leader(tempty_bar(acc));
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // nobody leaves while the peer can still multicast into it / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    if (U2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor-map cache + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  int64_t rows, cols, ld;
  int box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ (size_t)k.rows;
    h = h * 1000003u ^ (size_t)k.cols;
    h = h * 1000003u ^ (size_t)k.ld;
    h = h * 1000003u ^ (size_t)k.box_rows;
    return h;
  }
};

// 2-D bf16 tensor map: inner dim = cols (K), outer = rows; box = {64, box_rows}; 128B swizzle; OOB -> 0.
int get_tensor_map(const bf16* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(PCY_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(PCY_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%lld cols=%lld ld=%lld", (int)r,
                     (const void*)ptr, (long long)rows, (long long)cols, (long long)ld);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

template <int BN, bool ROPE, int CL, bool U2 = false>
int launch(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = TileCfg<BN, U2>;
  static SmemOptIn opt;
  if (opt.need(Cfg::SMEM_BYTES))
    PCY_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, ROPE, CL, U2>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  CUtensorMap ta, tb;
  PCY_TRY(get_tensor_map(a.A, a.M, a.K, a.lda, BM, &ta));
  PCY_TRY(get_tensor_map(a.W, a.N, a.K, a.ldw, BN / CL, &tb));  // CL = 2: each CTA loads half of the W tile
  EpiParams p;
  p.C = a.C; p.ldc = a.ldc; p.bias = a.bias; p.residual = a.residual; p.ldr = a.ldr;
  p.M = a.M; p.N = a.N; p.K = a.K; p.c_fp32 = a.c_fp32; p.act = a.act; p.scale = a.scale;
  p.scale_ncols = a.scale_ncols;
  p.rope = a.rope; p.rope_hd = a.rope_hd; p.rope_T = a.rope_T; p.rope_ncols = a.rope ? a.rope_ncols : 0;
  if (CL == 1) {
    const int tiles = ceil_div(a.M, BM) * ceil_div(a.N, BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    gemm_bf16_tcgen05_kernel<BN, ROPE, CL, U2><<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  } else {
    const int pairs = ceil_div(ceil_div(a.M, BM), 2) * ceil_div(a.N, BN);
    const int clusters = std::min(pairs, num_sms() / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PCY_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, ROPE, CL, U2>, ta, tb, p));
  }
  PCY_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int gemm_bf16_tc(const GemmArgs& a, cudaStream_t stream) {
  PCY_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  PCY_REQUIRE(a.K % 8 == 0 && a.lda % 8 == 0 && a.ldw % 8 == 0, "gemm: K/lda/ldw must be multiples of 8");
  PCY_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0,
              "gemm: A and W must be 16-byte aligned");
  if (a.act == ACT_SWIGLU) {
    PCY_REQUIRE(a.N % 32 == 0 && !a.c_fp32, "gemm: SwiGLU needs N %% 32 == 0 and bf16 output");
  }
  if (a.rope != nullptr) {
    PCY_REQUIRE((a.rope_hd == 64 || a.rope_hd == 128) && a.rope_ncols % a.rope_hd == 0 && a.rope_ncols <= a.N &&
                    a.rope_T > 0 && a.act == ACT_NONE && !a.c_fp32 && a.residual == nullptr,
                "gemm: fused RoPE needs head_dim 64/128, bf16 output, no activation/residual");
  }
  // Tile choice: 128x256 unless that leaves most SMs idle / wastes a wide tail.
  const int sms = num_sms();
  const int tiles256 = ceil_div(a.M, BM) * ceil_div(a.N, 256);
  const int tiles128 = ceil_div(a.M, BM) * ceil_div(a.N, 128);
  // (a single, nearly full wave of 128x256 tiles beats two waves of 128x128 tiles: measured on B200 at M = 1024,
  // N = 4096: K = 14336 1188 vs 964 TFLOP/s, K = 4096 854 vs 794 - the wider tile halves the shared-memory traffic per
  // flop, which is what limits the narrow one)
  bool use128 = (a.N <= 128) || (tiles256 * 10 < sms * 7 && tiles128 > tiles256);
  if (!use128) {
    // wave quantisation: prefer the tile size with the better SM-time efficiency
    const double w256 = (double)tiles256 / (double)(ceil_div(tiles256, sms) * sms);
    const double w128 = (double)tiles128 / (double)(ceil_div(tiles128, sms) * sms);
    if (w128 > w256 * 1.15) use128 = true;
  }
  const int force = g_gemm_force_tile;  // pcy_set_gemm_tile: 0 = heuristic, 128 / 192 / 256 = that tile width (tests)
  if (force == 128) use128 = true;
  if ((force == 256 || force == 192) && a.N > 128) use128 = false;
  // two or more row-blocks: cluster pairs share the W tile — as one cta_group::2 MMA, or through TMA multicast
  // Measured on B200 (profiles/r01_gemm_shapes_pair.log): +5..12 % on the ESM2 shapes (M = 32 896) and on the Llama
  // gate/up GEMM, i.e. at or above cuBLAS, but -5 % on problems of one or two waves (Llama q/k/v/o at M = 1024), where
  // pairing halves the number of schedulable units.
  const int tiles = use128 ? tiles128 : tiles256;
  // ... and on single-wave problems with an even number of row blocks and a long K loop, where a pair occupies the two
  // SMs two single tiles would have occupied anyway (M = 1024, N = 4096, K = 14336: 1355 vs 1188 TFLOP/s, cuBLAS 1341)
  const bool single_wave_pair = !use128 && tiles256 <= sms && ceil_div(a.M, BM) % 2 == 0 && a.K >= 8192;
  if (ceil_div(a.M, BM) >= 2 && a.N > 128 && a.rope == nullptr &&
      (g_gemm_pair_mma == 2 || (g_gemm_pair_mma == 1 && (tiles >= 3 * sms || single_wave_pair))))
    return use128 ? launch<128, false, 2, true>(a, stream) : launch<256, false, 2, true>(a, stream);
  const bool pair = g_gemm_cluster && ceil_div(a.M, BM) >= 2;
  if (a.rope != nullptr) {
    if (pair) return use128 ? launch<128, true, 2>(a, stream) : launch<256, true, 2>(a, stream);
    return use128 ? launch<128, true, 1>(a, stream) : launch<256, true, 1>(a, stream);
  }
  if (pair) return use128 ? launch<128, false, 2>(a, stream) : launch<256, false, 2>(a, stream);
  {
    // 128x192 tiles when they fill the waves better than both other sizes (Llama q/k/v at M = 1024: N = 6144 is 192
    // tiles of 128x256 = 1.3 waves, 384 of 128x128 = 2.6 waves, but 256 of 128x192 = 1.73 waves with 93 % of the wide
    // tile's per-tile rate); relative per-tile rates measured on B200: 1.0 / 0.93 / 0.75
    const int tiles192 = ceil_div(a.M, BM) * ceil_div(a.N, 192);
    auto eff = [&](int tiles, double rate, int bn) {
      const double waste = (double)a.N / ((double)ceil_div(a.N, bn) * bn);  // columns of the last tile past N
      return (double)tiles / (double)(ceil_div(tiles, sms) * sms) * rate * waste;
    };
    const double e192 = eff(tiles192, 0.93, 192), e256 = eff(tiles256, 1.0, 256), e128 = eff(tiles128, 0.75, 128);
    if (force == 192 || (force == 0 && a.N >= 192 && e192 > 1.08 * (use128 ? e128 : e256)))
      return launch<192, false, 1>(a, stream);
  }
  return use128 ? launch<128, false, 1>(a, stream) : launch<256, false, 1>(a, stream);
}

}  // namespace pcy
