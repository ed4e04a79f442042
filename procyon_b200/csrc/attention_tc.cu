// Bidirectional (ESM2) attention on the sm_100a tensor cores: S = Q K^T and O_tile = P V are tcgen05.mma with
// accumulators in TMEM, Q/K/V tiles arrive by TMA (128B swizzle), the softmax runs one query row per thread
// straight out of TMEM (no shuffles), P goes back through shared memory as the A operand of the second MMA and
// V is consumed in place as an MN-major B operand (no transpose).  128 queries x 128 keys per step, head_dim 64,
// two CTAs per SM so one CTA's softmax overlaps the other's MMAs.
//
// Replaces fair-esm MultiheadAttention's bmm -> fp32 softmax -> bmm (reached via procyon/model/esm.py:536) for
// the full 128-row query tiles; the ragged tail rows use the mma.sync kernel in attention.cu.
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ops.h"

namespace pcy {

// pcy_set_esm_attention_kernel: 0 = 128-key steps (esm_attention_tc_kernel), 1 = 64-key steps with double-buffered
// S / P / P.V (esm_attention_tc64_kernel), 2 = 64-key steps with Q and P in TMEM (esm_attention_ts_kernel), 3 = the
// same with P packed on the ALU pipe, 4 = kernel 2 with pair barriers instead of CTA-wide bar.sync, 5 (default) =
// kernel 4 with the output tile accumulated in TMEM
int g_esm_attention_kernel = 5;
// pcy_set_esm_attention_tail_rows(n): when T leaves at most n query rows beyond the last full 128-row tile (ESM2 adds
// BOS + EOS: 512 residues are T = 514 = 4 tiles + 2 rows), those rows go to the mma.sync kernel of attention.cu instead
// of a fifth tcgen05 CTA that streams all of K / V for two live rows.  0 = every row on the tcgen05 kernel.
int g_esm_attention_tail_rows = 0;
// pcy_set_esm_attention_q_rope(1): kernels 2 / 3 / 4 rotate Q on its way into TMEM and the RoPE pass only covers K.
// Measured on B200 (profiles/r02_esm_breakdown.log): the RoPE pass drops 8.3 -> 4.3 ms per 256-protein step, but the
// attention kernel's prologue (cos / sin loads + 4 FMAs per element in front of the first MMA, on the critical path of
// every CTA) costs 51.5 -> 65.4 ms: off by default.
bool g_esm_attention_q_rope = false;

namespace {

constexpr int TBM = 128;  // queries per CTA
constexpr int TBN = 128;  // keys per step
constexpr int THD = 64;   // head dim
constexpr int SM_WARPS = 8;       // softmax warps: two threads per query row (64 of the 128 key columns each)
constexpr int TC_THREADS = (SM_WARPS + 1) * 32;  // + warp 8: TMA + MMA issue
constexpr int Q_BYTES = TBM * THD * 2;
constexpr int KV_BYTES = TBN * THD * 2;
constexpr int P_BYTES = TBM * TBN * 2;
constexpr int TC_SMEM = Q_BYTES + 2 * 2 * KV_BYTES + P_BYTES + 2 * TBM * 2 /*row max exchange (bf16)*/ + 32 /*valid*/ + 80 /*barriers*/;
static_assert(2 * (TC_SMEM + 1024) <= 228 * 1024, "two CTAs per SM must fit");
constexpr int TMEM_COLS_ATT = 256;  // S: [0,128), O_tile: [128,192)

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

struct TcAttnParams {
  bf16* o;
  int64_t o_rs;  // output row stride (elements)
  const uint8_t* key_valid;  // [B][T] or null
  // the same as bits: word w of a sequence covers keys [32 w, 32 w + 32) (bits past T are 0), 2 * ceil(T / 64) words per
  // sequence; null: kernels that want the words of every step up front build them from the bytes with ballots
  const uint32_t* valid_words;
  int B, H, T, d;
  int n_q_tiles;
  float scale_log2;
  const float* q_rope;  // fp32 [T][32][2] (cos, sin) or null: rotate-half RoPE applied to Q on its way into TMEM
};

__global__ void __launch_bounds__(TC_THREADS, 2)
esm_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment; the dynamic window starts 1024-aligned when the kernel has no
  // static shared memory (checked: a misaligned base would overrun the allocation)
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Q_BYTES;            // 2 stages
  const uint32_t sV = sK + 2 * KV_BYTES;       // 2 stages
  const uint32_t sP = sV + 2 * KV_BYTES;       // [2 blocks of 64 keys][128 rows][128 B]
  const uint32_t sXchg = sP + P_BYTES;         // bf16 [2 halves][128 rows]
  const uint32_t sValid = sXchg + 2 * TBM * 2; // 2 parities x 4 words
  const uint32_t bars = sValid + 32;
  const uint32_t q_full = bars, kv_full0 = bars + 8, kv_empty0 = bars + 24, s_full = bars + 40, p_ready = bars + 48,
                 o_full = bars + 56, tmem_slot = bars + 64;  // 68 bytes used of 80
  uint8_t* valid_smem = smem_raw + (sValid - smem_u32(smem_raw));
  __nv_bfloat16* xchg = reinterpret_cast<__nv_bfloat16*>(smem_raw + (sXchg - smem_u32(smem_raw)));
  float* xchg_f = reinterpret_cast<float*>(smem_raw + (sP - smem_u32(smem_raw)));  // P region, free after the last tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  // the last tile is shifted back to end exactly at T (it recomputes some rows of the previous tile; identical
  // values are written twice) so that no ragged tail is left for another kernel
  const int q0 = min(q_tile * TBM, p.T - TBM);
  const int row_base = b * p.T;  // first row of this sequence in the [B*T, 3d] matrix
  const int n_kv = (p.T + TBN - 1) / TBN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    mbar_init(kv_full0, 1);
    mbar_init(kv_full0 + 8, 1);
    mbar_init(kv_empty0, 1);
    mbar_init(kv_empty0 + 8, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, SM_WARPS * 32);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == SM_WARPS) {
    tmem_alloc(tmem_slot, TMEM_COLS_ATT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == SM_WARPS) {
    if (lane == 0) {
      // ---------------- TMA producer + MMA issuer (one thread) ----------------
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(sQ, &tmap, q_full, h * THD, row_base + q0);
      // prologue: first K/V stage
      auto load_kv = [&](int j) {
        const int st = j & 1;
        mbar_arrive_expect_tx(kv_full0 + 8 * st, 2 * KV_BYTES);
        tma_load_2d(sK + st * KV_BYTES, &tmap, kv_full0 + 8 * st, p.d + h * THD, row_base + j * TBN);
        tma_load_2d(sV + st * KV_BYTES, &tmap, kv_full0 + 8 * st, 2 * p.d + h * THD, row_base + j * TBN);
      };
      load_kv(0);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(kv_full0 + 8 * st, (j >> 1) & 1);
        tc_fence_after();
        // S = Q K^T : M=128, N=keys in this tile rounded up to 16, K=64 (4 x UMMA_K)
        const int keys = min(TBN, p.T - j * TBN);
        const int n_mma = (keys + 15) & ~15;
        const uint32_t idesc_s = make_idesc_bf16(TBM, n_mma);
        const uint64_t qd = make_desc_kmajor_sw128(sQ);
        const uint64_t kd = make_desc_kmajor_sw128(sK + st * KV_BYTES);
#pragma unroll
        for (int k = 0; k < THD / 16; ++k) tc_mma_bf16(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        tc_commit(s_full);
        if (j + 1 < n_kv) {  // prefetch the next K/V tile into the other stage once its previous user (PV of j-1) is done
          if (j + 1 >= 2) mbar_wait(kv_empty0 + 8 * ((j + 1) & 1), (((j + 1) >> 1) - 1) & 1);
          load_kv(j + 1);
        }
        // O_tile = P V : M=128, N=64, K=n_mma keys; A = P (K-major, two 64-key blocks), B = V (MN-major)
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const uint32_t idesc_o = make_idesc_bf16(TBM, THD, 0, 1);
        for (int k = 0; k < n_mma / 16; ++k) {
          const uint64_t pd = make_desc_kmajor_sw128(sP + (k >> 2) * (TBM * 128) + (k & 3) * 32);
          const uint64_t vd = make_desc_mnmajor_sw128(sV + st * KV_BYTES + k * 2048, 1024);
          tc_mma_bf16(tmem_base + TBN, pd, vd, idesc_o, k > 0 ? 1u : 0u);
        }
        tc_commit(kv_empty0 + 8 * st);
        tc_commit(o_full);
      }
    }
  } else {
    // ---------------- softmax / accumulate: two threads per query row ----------------
    // warp w may touch TMEM lanes 32*(w%4)..+31: warps w and w+4 share a row quadrant and split the key columns
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;  // query row within the tile = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    constexpr int OH = THD / 2;  // output columns per thread
    float o[OH];
#pragma unroll
    for (int i = 0; i < OH; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const uint8_t* valid_g = p.key_valid ? p.key_valid + (int64_t)b * p.T : nullptr;
    uint32_t* vwords = reinterpret_cast<uint32_t*>(valid_smem);  // [2 parities][4 words]: validity bit per key
    // In the shifted last tile the rows below q_tile*TBM were already produced by the previous tile.  A warp whose 32
    // rows are all such repeats keeps every barrier / mbarrier appointment (the MMAs are issued for the whole tile
    // anyway, rows are independent in both products, and its rows of P / O are never read) but does none of the
    // TMEM reads, exponentials or stores: at T = 514 that is 6 of the 8 softmax warps of every fifth tile.
    const bool live = q0 + quad * 32 + 31 >= q_tile * TBM;
    auto publish_valid = [&](int j) {  // validity bit of key j*TBN + r, one word per 32 keys (needed by every warp)
      const int kidx = j * TBN + r;
      bool ok = kidx < p.T;
      if (ok && valid_g) ok = valid_g[kidx] != 0;
      const uint32_t word = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) vwords[(j & 1) * 4 + quad] = word;
    };
    if (!live) {
      for (int j = 0; j < n_kv; ++j) {
        if (half == 0) publish_valid(j);
        mbar_wait(s_full, j & 1);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mbar_arrive(p_ready);
        mbar_wait(o_full, j & 1);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
    } else {
    for (int j = 0; j < n_kv; ++j) {
      const int kbase = j * TBN;
      if (half == 0) publish_valid(j);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int keys = min(TBN, p.T - kbase);
      const int n_mma = (keys + 15) & ~15;
      const int c_lo = half * 64, c_hi = min(n_mma, c_lo + 64);  // this thread's key columns
      // pass 1: row max over this thread's columns (validity is applied in pass 2: masked keys only raise the max,
      // which is harmless for the softmax value... but not for -inf rows, so mask here too)
      asm volatile("bar.sync 1, 256;" ::: "memory");  // validity words visible
      const uint32_t mwa = vwords[(j & 1) * 4 + half * 2], mwb = vwords[(j & 1) * 4 + half * 2 + 1];
      const bool all_valid = (mwa & mwb) == 0xffffffffu;
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + c, v);
        tc_wait_ld();
        if (all_valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
        } else {
          const uint32_t mw = (c == c_lo) ? mwa : mwb;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            mx4[i & 3] = fmaxf(mx4[i & 3], ((mw >> i) & 1u) ? __uint_as_float(v[i]) : -INFINITY);
        }
      }
      // the two threads of a row must agree on the offset exactly; any value >= the true max works, so exchange the
      // max rounded UP to bf16 (halves the exchange buffer: the kernel sits 100 bytes under the 2-CTA/SM smem limit)
      const __nv_bfloat16 mx_own_b = __float2bfloat16_ru(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])));
      xchg[half * TBM + r] = mx_own_b;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float mx = fmaxf(__bfloat162float(mx_own_b), __bfloat162float(xchg[(half ^ 1) * TBM + r]));
      const float m_new = fmaxf(m_run, mx);
      const float corr = (m_new == -INFINITY) ? 1.f : exp2f((m_run - m_new) * p.scale_log2);
      const float moff = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
      // pass 2: p = exp2(s*scale - moff) -> bf16 -> swizzled smem block `half` (A operand of the second MMA)
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + c, v);
        tc_wait_ld();
        uint32_t packed[16];
        const uint32_t mw = all_valid ? 0xffffffffu : ((c == c_lo) ? mwa : mwb);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0, p1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v[i]), p.scale_log2, -moff)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -moff)));
          if (!all_valid) {
            p0 = ((mw >> i) & 1u) ? p0 : 0.f;
            p1 = ((mw >> (i + 1)) & 1u) ? p1 : 0.f;
          }
          ls4[(i >> 1) & 3] += p0 + p1;
          packed[i >> 1] = pack_bf16x2(p0, p1);
        }
        const uint32_t blk = sP + (uint32_t)half * (TBM * 128) + (uint32_t)r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 63) >> 3) + q;
          const uint32_t addr = blk + (uint32_t)((chunk ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(packed[4 * q]), "r"(packed[4 * q + 1]),
                       "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                       : "memory");
        }
      }
      l_run = l_run * corr + ((ls4[0] + ls4[1]) + (ls4[2] + ls4[3]));  // partial sum over this thread's columns
      m_run = m_new;
      // publish P (the previous O_tile was consumed at the end of the previous iteration)
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      // O = O * corr + P V of this tile (this thread's 32 output columns)
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + TBN + half * OH, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < OH; ++i) o[i] = fmaf(o[i], corr, __uint_as_float(v[i]));
      }
      tc_fence_before();
    }
    // ---- finalize: row sum = both halves ----
    asm volatile("bar.sync 1, 256;" ::: "memory");
    xchg_f[half * TBM + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_tot = l_run + xchg_f[(half ^ 1) * TBM + r];
    const int qrow = q0 + r;
    if (qrow < p.T && qrow >= q_tile * TBM) {  // rows below q_tile*TBM belong to the previous tile
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      bf16* op = p.o + (int64_t)(row_base + qrow) * p.o_rs + h * THD + half * OH;
#pragma unroll
      for (int c = 0; c < OH; c += 8) {
        uint4 u;
        u.x = pack_bf16x2(o[c] * inv, o[c + 1] * inv);
        u.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
        u.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv);
        u.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
        *reinterpret_cast<uint4*>(op + c) = u;
      }
    }
    }  // live
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SM_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS_ATT);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// 64-key steps, everything the tensor core touches double-buffered.
//
// In the 128-key kernel above a CTA's softmax warps and its MMAs take turns: S(j) -> softmax(j) -> P.V(j) -> O read ->
// S(j+1) ... (ncu: 34 % of the warp samples sit in the waits for s_full and o_full; the second CTA on the SM hides only
// part of it).  Here a step is 64 keys, so TMEM (256 columns) holds TWO S tiles and TWO P.V tiles and shared memory
// two P tiles in the same 112 KB:
//   * S(j+1) is issued at the START of the MMA thread's iteration j, before it waits for softmax(j);
//   * softmax(j) keeps its 32 scores per thread in registers (one tcgen05.ld per step instead of two) and, once P(j) is
//     published, folds in the P.V tile of step j-1, which has been complete for a whole step;
// so in steady state neither side waits for the other.  Barriers come in pairs indexed by j & 1 (phase (j >> 1) & 1):
// a barrier's next phase needs work that every waiter of the current phase has already passed.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TBN2 = 64;                      // keys per step
constexpr int KV2_STAGES = 4;
constexpr int KV2_BYTES = TBN2 * THD * 2;     // 8 KB
constexpr int P2_BYTES = TBM * TBN2 * 2;      // 16 KB: [128 rows][128 B], 128B swizzle
constexpr int TC2_SMEM = Q_BYTES + 2 * KV2_STAGES * KV2_BYTES + 2 * P2_BYTES + 2 * TBM * 2 /*row max exchange*/ +
                         16 /*valid words*/ + 144 /*barriers*/;
static_assert(2 * (TC2_SMEM + 1024) <= 228 * 1024, "two CTAs per SM must fit");
static_assert(2 * TBM * 4 <= 2 * P2_BYTES, "row-sum exchange reuses the P region");

__global__ void __launch_bounds__(TC_THREADS, 2)
esm_attention_tc64_kernel(const __grid_constant__ CUtensorMap tmap, const TcAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Q_BYTES;                     // KV2_STAGES stages
  const uint32_t sV = sK + KV2_STAGES * KV2_BYTES;      // KV2_STAGES stages
  const uint32_t sP = sV + KV2_STAGES * KV2_BYTES;      // 2 buffers
  const uint32_t sXchg = sP + 2 * P2_BYTES;             // bf16 [2 halves][128 rows]
  const uint32_t sValid = sXchg + 2 * TBM * 2;          // [2 parities][2 words]
  const uint32_t bars = sValid + 16;
  const uint32_t q_full = bars, kv_full0 = bars + 8, kv_empty0 = kv_full0 + 8 * KV2_STAGES,
                 s_full0 = kv_empty0 + 8 * KV2_STAGES, p_ready0 = s_full0 + 16, o_full0 = p_ready0 + 16,
                 tmem_slot = o_full0 + 16;  // 124 bytes used of 144
  uint8_t* valid_smem = smem_raw + (sValid - base);
  __nv_bfloat16* xchg = reinterpret_cast<__nv_bfloat16*>(smem_raw + (sXchg - base));
  float* xchg_f = reinterpret_cast<float*>(smem_raw + (sP - base));  // P region, free after the last P.V

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = min(q_tile * TBM, p.T - TBM);  // shifted last tile, see the kernel above
  const int row_base = b * p.T;
  const int n_kv = (p.T + TBN2 - 1) / TBN2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV2_STAGES; ++s) {
      mbar_init(kv_full0 + 8 * s, 1);
      mbar_init(kv_empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full0 + 8 * s, 1);
      mbar_init(p_ready0 + 8 * s, SM_WARPS * 32);
      mbar_init(o_full0 + 8 * s, 1);
    }
    fence_barrier_init();
  }
  if (warp == SM_WARPS) {
    tmem_alloc(tmem_slot, TMEM_COLS_ATT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // TMEM columns: S buffers at 0 and 64, P.V tiles at 128 and 192

  if (warp == SM_WARPS) {
    if (lane == 0) {
      // ---------------- TMA producer + MMA issuer (one thread) ----------------
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(sQ, &tmap, q_full, h * THD, row_base + q0);
      tma_load_2d(sQ + Q_BYTES / 2, &tmap, q_full, h * THD, row_base + q0 + TBN2);
      auto load_kv = [&](int t) {
        const int st = t & (KV2_STAGES - 1);
        mbar_arrive_expect_tx(kv_full0 + 8 * st, 2 * KV2_BYTES);
        tma_load_2d(sK + st * KV2_BYTES, &tmap, kv_full0 + 8 * st, p.d + h * THD, row_base + t * TBN2);
        tma_load_2d(sV + st * KV2_BYTES, &tmap, kv_full0 + 8 * st, 2 * p.d + h * THD, row_base + t * TBN2);
      };
      auto issue_s = [&](int t) {  // S(t) = Q K(t)^T into S buffer t & 1
        const int st = t & (KV2_STAGES - 1);
        mbar_wait(kv_full0 + 8 * st, (t / KV2_STAGES) & 1);
        tc_fence_after();
        const int keys = min(TBN2, p.T - t * TBN2);
        const uint32_t idesc_s = make_idesc_bf16(TBM, (keys + 15) & ~15);
        const uint64_t qd = make_desc_kmajor_sw128(sQ);
        const uint64_t kd = make_desc_kmajor_sw128(sK + st * KV2_BYTES);
#pragma unroll
        for (int k = 0; k < THD / 16; ++k)
          tc_mma_bf16(tmem_base + (uint32_t)((t & 1) * TBN2), qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        tc_commit(s_full0 + 8 * (t & 1));
      };
      for (int t = 0; t < min(KV2_STAGES - 1, n_kv); ++t) load_kv(t);
      mbar_wait(q_full, 0);
      issue_s(0);
      const uint32_t idesc_o = make_idesc_bf16(TBM, THD, 0, 1);
      for (int j = 0; j < n_kv; ++j) {
        // S(j+1) goes into the buffer S(j-1) was read from; its readers arrived on p_ready(j-1), waited for below in
        // the previous iteration
        if (j + 1 < n_kv) issue_s(j + 1);
        const int t = j + KV2_STAGES - 1;  // refill the stage tile j-1 used, once P.V(j-1) has retired
        if (t < n_kv) {
          if (t >= KV2_STAGES) mbar_wait(kv_empty0 + 8 * (t & (KV2_STAGES - 1)), ((t / KV2_STAGES) - 1) & 1);
          load_kv(t);
        }
        // O_tile(j) = P(j) V(j) : M=128, N=64, K = keys of this step rounded up to 16; A = P (K-major), B = V (MN-major)
        mbar_wait(p_ready0 + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        const int st = j & (KV2_STAGES - 1);
        const int n_mma = (min(TBN2, p.T - j * TBN2) + 15) & ~15;
        for (int k = 0; k < n_mma / 16; ++k) {
          const uint64_t pd = make_desc_kmajor_sw128(sP + (j & 1) * P2_BYTES + k * 32);
          const uint64_t vd = make_desc_mnmajor_sw128(sV + st * KV2_BYTES + k * 2048, 1024);
          tc_mma_bf16(tmem_base + (uint32_t)(2 * TBN2 + (j & 1) * THD), pd, vd, idesc_o, k > 0 ? 1u : 0u);
        }
        tc_commit(kv_empty0 + 8 * st);
        tc_commit(o_full0 + 8 * (j & 1));
      }
    }
  } else {
    // ---------------- softmax / accumulate: two threads per query row, 32 of the step's 64 keys each ----------------
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;  // query row within the tile = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    constexpr int OH = THD / 2;  // output columns per thread
    const uint8_t* valid_g = p.key_valid ? p.key_valid + (int64_t)b * p.T : nullptr;
    uint32_t* vwords = reinterpret_cast<uint32_t*>(valid_smem);  // [2 parities][2 words]: validity bit per key
    const bool live = q0 + quad * 32 + 31 >= q_tile * TBM;  // see the kernel above
    auto publish_valid = [&](int j) {  // threads r < 64 of half 0: validity bit of key j*64 + r
      if (half == 0 && quad < 2) {
        const int kidx = j * TBN2 + r;
        bool ok = kidx < p.T;
        if (ok && valid_g) ok = valid_g[kidx] != 0;
        const uint32_t word = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) vwords[(j & 1) * 2 + quad] = word;
      }
    };
    if (!live) {
      for (int j = 0; j < n_kv; ++j) {
        publish_valid(j);
        mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mbar_arrive(p_ready0 + 8 * (j & 1));
        if (j > 0) mbar_wait(o_full0 + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
      }
      mbar_wait(o_full0 + 8 * ((n_kv - 1) & 1), ((n_kv - 1) >> 1) & 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
    } else {
      float o[OH];
#pragma unroll
      for (int i = 0; i < OH; ++i) o[i] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, corr_prev = 1.f;
      auto fold_o = [&](int jj, float corr) {  // o = o * corr + P.V tile of step jj (this thread's 32 columns)
        mbar_wait(o_full0 + 8 * (jj & 1), (jj >> 1) & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + (uint32_t)(2 * TBN2 + (jj & 1) * THD + half * OH), v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < OH; ++i) o[i] = fmaf(o[i], corr, __uint_as_float(v[i]));
      };
      for (int j = 0; j < n_kv; ++j) {
        publish_valid(j);
        mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        uint32_t v[32];  // this thread's 32 scores of the step stay in registers for both passes
        tmem_ld_32x32b_x32(t_lane + (uint32_t)((j & 1) * TBN2 + half * 32), v);
        tc_wait_ld();
        asm volatile("bar.sync 1, 256;" ::: "memory");  // validity words visible
        const uint32_t mw = vwords[(j & 1) * 2 + half];
        // pass 1: row max over the valid columns (columns past the step's keys hold stale scores: their bits are 0)
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (mw == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            mx4[i & 3] = fmaxf(mx4[i & 3], ((mw >> i) & 1u) ? __uint_as_float(v[i]) : -INFINITY);
        }
        // the two threads of a row must agree on the offset exactly; any value >= the true max works: exchange the
        // max rounded UP to bf16
        const __nv_bfloat16 mx_own_b = __float2bfloat16_ru(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])));
        xchg[half * TBM + r] = mx_own_b;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float mx = fmaxf(__bfloat162float(mx_own_b), __bfloat162float(xchg[(half ^ 1) * TBM + r]));
        const float m_new = fmaxf(m_run, mx);
        const float corr = (m_new == -INFINITY) ? 1.f : exp2f((m_run - m_new) * p.scale_log2);
        const float moff = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
        // pass 2: p = exp2(s*scale - moff) -> bf16 -> swizzled P buffer j & 1 (A operand of the second MMA)
        float ls4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0, p1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v[i]), p.scale_log2, -moff)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -moff)));
          if (mw != 0xffffffffu) {
            p0 = ((mw >> i) & 1u) ? p0 : 0.f;
            p1 = ((mw >> (i + 1)) & 1u) ? p1 : 0.f;
          }
          ls4[(i >> 1) & 3] += p0 + p1;
          packed[i >> 1] = pack_bf16x2(p0, p1);
        }
        const uint32_t blk = sP + (uint32_t)(j & 1) * P2_BYTES + (uint32_t)r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t addr = blk + (uint32_t)(((half * 4 + q) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(packed[4 * q]), "r"(packed[4 * q + 1]),
                       "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                       : "memory");
        }
        l_run = l_run * corr + ((ls4[0] + ls4[1]) + (ls4[2] + ls4[3]));
        m_run = m_new;
        // publish P(j); this arrive also tells the MMA thread that S buffer j & 1 and P.V tile (j-1) & 1's
        // predecessor (folded at the end of the previous step) are free
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_ready0 + 8 * (j & 1));
        // fold in the P.V tile of the previous step (complete for about a step by now) with ITS correction factor
        if (j > 0) fold_o(j - 1, corr_prev);
        corr_prev = corr;
      }
      fold_o(n_kv - 1, corr_prev);
      // ---- finalize: row sum = both halves ----
      asm volatile("bar.sync 1, 256;" ::: "memory");
      xchg_f[half * TBM + r] = l_run;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float l_tot = l_run + xchg_f[(half ^ 1) * TBM + r];
      const int qrow = q0 + r;
      if (qrow < p.T && qrow >= q_tile * TBM) {  // rows below q_tile*TBM belong to the previous tile
        const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
        bf16* op = p.o + (int64_t)(row_base + qrow) * p.o_rs + h * THD + half * OH;
#pragma unroll
        for (int c = 0; c < OH; c += 8) {
          uint4 u;
          u.x = pack_bf16x2(o[c] * inv, o[c + 1] * inv);
          u.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
          u.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv);
          u.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
          *reinterpret_cast<uint4*>(op + c) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SM_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS_ATT);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Same schedule as the 64-key kernel above, but both A operands live in TMEM.
//
// A tcgen05.mma whose A operand is a shared-memory descriptor reads 128 rows x 32 B of it per K = 16 step whatever N
// is, so with N = 64 (this head dim, these 64-key steps) an instruction costs ~170 cycles instead of the N/2 = 32 the
// tensor core needs (ncu of the 128-key kernel: 12 MMAs per 128 keys keep the tensor pipe busy for ~2050 cycles).
// Here Q is written to TMEM once per CTA by the softmax threads (each its own row; it is the A operand of every
// S = Q K^T), and P(j) is written over the first 32 columns of the S buffer it was computed from (bf16 pairs, one
// tcgen05.st per thread) and consumed from there by P.V — no P tile in shared memory, no async-proxy fence.
// TMEM columns: S/P buffers at 0 and 64, the P.V tile at 128, Q at 192 (32 columns).
// ---------------------------------------------------------------------------------------------------------------
// K / V tiles requested ahead of the S = Q K^T that consumes them, for kernels 2-7.  With KV2_STAGES - 1 = 3 the tile
// requested in iteration j goes into the stage of tile j - 1, whose P.V was issued a moment ago: the MMA thread sits
// ~280 cycles at kv_empty in every step (clock64 stamps of the thread, profiles/r02_attention_step_stamps.log: issue of
// S 430 + this wait 280 + wait for the softmax threads 500-680 + issue of P.V 525 cycles per step).  With 2 it goes into
// the stage of tile j - 2, free for a whole step, and the tile still has a step (~2 000 cycles) to arrive: parity-green
// on B200 and measured at the end of round 2 — no faster (47.4 / 46.2 ms for kernels 5 / 6 against 46.4 / 45.6-46.1:
// the thread then waits that much longer for the softmax threads), so the value every validation run of the round used
// stays.  1 would dead-lock (S(j+1) is issued before the iteration's load); scripts/sim_attention_phases.py replays all
// three.
constexpr int KV2_AHEAD = KV2_STAGES - 1;
constexpr int TS_SMEM = 2 * KV2_STAGES * KV2_BYTES + 4 * TBM * 2 /*row max exchange, 2 parities*/ + 2 * TBM * 4 /*row sums*/ +
                        16 /*valid words*/ + 144 /*barriers*/;

// ALU_PACK: the bf16 pairs of P are built with two integer adds (round half up) and one byte permute instead of
// F2FP.BF16.PACK_AB, which shares the quarter-rate XU pipe with MUFU.EX2 (ncu: XU 43 % busy, 1.5 XU ops per score).
// PAIR: the only data two warps ever exchange is the row maximum between the two threads of a row, i.e. between warps w
// and w + 4 (same TMEM lane quadrant, same SM sub-partition).  Instead of two 256-thread bar.sync per step (18 % of the
// warp samples of the plain kernel sit there: eight warps on four sub-partitions drift apart) every warp takes the
// validity bits of its own 32 keys with one ballot, the pair meets at ONE 64-thread named barrier per step (row-max
// slots double-buffered by step parity), and the masked / unmasked forms of the exponential pass are separate loops
// (as one loop the compiler predicates 66 LOP3 / FSEL into every step, masked or not).
// OACC (kernel 5): the output tile is not carried in registers and folded in step by step (a TMEM read of the P.V
// tile + a wait for its MMAs in EVERY step of the softmax threads' serial chain) but accumulates in TMEM across all
// steps (use_acc), like the causal prefill kernel's.  The running maximum is then only raised when a step's maximum
// exceeds it by more than 2^8 (the exponentials stay <= 256, exact in the fp32 row sum and well inside bf16's range for
// P; both threads of a row take the same decision from the same two numbers), and only such steps rescale O in TMEM
// (tcgen05.ld -> scale -> tcgen05.st between the completion of P.V(j-1) and the release of P(j)).
// 256-protein step of ESM2-650M on B200: 51.4 -> 49.7 ms.  (Requesting the scores of step j + 1 from TMEM at the end of
// step j — they are complete a step ahead — was measured slower, 49.7 -> 53.8 ms: with v[] live across the loop edge
// the 96-register budget of two CTAs per SM spills.)
template <bool ALU_PACK, bool PAIR = false, bool OACC = false>
__global__ void __launch_bounds__(TC_THREADS, 2)
esm_attention_ts_kernel(const __grid_constant__ CUtensorMap tmap, const TcAttnParams p, const bf16* __restrict__ qkv) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t sK = base;                             // KV2_STAGES stages
  const uint32_t sV = sK + KV2_STAGES * KV2_BYTES;      // KV2_STAGES stages
  const uint32_t sXchg = sV + KV2_STAGES * KV2_BYTES;   // bf16 [2 halves][128 rows]
  const uint32_t sSum = sXchg + 4 * TBM * 2;            // float [2 halves][128 rows]
  const uint32_t sValid = sSum + 2 * TBM * 4;           // [2 parities][2 words]
  const uint32_t bars = sValid + 16;
  const uint32_t q_full = bars, kv_full0 = bars + 8, kv_empty0 = kv_full0 + 8 * KV2_STAGES,
                 s_full0 = kv_empty0 + 8 * KV2_STAGES, p_ready0 = s_full0 + 16, o_full = p_ready0 + 16,
                 tmem_slot = o_full + 8;
  uint8_t* valid_smem = smem_raw + (sValid - base);
  __nv_bfloat16* xchg = reinterpret_cast<__nv_bfloat16*>(smem_raw + (sXchg - base));
  float* xchg_f = reinterpret_cast<float*>(smem_raw + (sSum - base));
  constexpr uint32_t COL_O = 2 * TBN2, COL_Q = 2 * TBN2 + THD;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = min(q_tile * TBM, p.T - TBM);  // shifted last tile, see the first kernel
  const int row_base = b * p.T;
  const int n_kv = (p.T + TBN2 - 1) / TBN2;

  // The softmax threads' first global loads — their half of the query row and the validity words — are issued BEFORE
  // the CTA's setup (barrier init, TMEM allocation, __syncthreads): a CTA lives for nine key steps and these loads head
  // the critical path of its first one.
  uint32_t qr[16];
  uint32_t step_masks = 0;  // PAIR: lane j holds the validity word of this warp's keys of step j
  const bool early_q = warp < SM_WARPS && p.q_rope == nullptr;
  if (early_q) {
    const int e_quad = warp & 3, e_half = warp >> 2;
    const uint4* qp = reinterpret_cast<const uint4*>(
        qkv + (int64_t)(row_base + q0 + e_quad * 32 + lane) * (3 * p.d) + h * THD + e_half * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = __ldg(qp + i);
      qr[4 * i] = u.x; qr[4 * i + 1] = u.y; qr[4 * i + 2] = u.z; qr[4 * i + 3] = u.w;
    }
    if (PAIR && n_kv <= 32 && p.valid_words != nullptr && lane < n_kv)
      step_masks = __ldg(p.valid_words + (int64_t)b * (2 * n_kv) + 2 * lane + e_half);
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, SM_WARPS * 32);
    for (int s = 0; s < KV2_STAGES; ++s) {
      mbar_init(kv_full0 + 8 * s, 1);
      mbar_init(kv_empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full0 + 8 * s, 1);
      mbar_init(p_ready0 + 8 * s, SM_WARPS * 32);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == SM_WARPS) {
    tmem_alloc(tmem_slot, TMEM_COLS_ATT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == SM_WARPS) {
    if (lane == 0) {
      // ---------------- TMA producer + MMA issuer (one thread) ----------------
      auto load_kv = [&](int t) {
        const int st = t & (KV2_STAGES - 1);
        mbar_arrive_expect_tx(kv_full0 + 8 * st, 2 * KV2_BYTES);
        tma_load_2d(sK + st * KV2_BYTES, &tmap, kv_full0 + 8 * st, p.d + h * THD, row_base + t * TBN2);
        tma_load_2d(sV + st * KV2_BYTES, &tmap, kv_full0 + 8 * st, 2 * p.d + h * THD, row_base + t * TBN2);
      };
      auto issue_s = [&](int t) {  // S(t) = Q K(t)^T into S buffer t & 1; A = Q in TMEM
        const int st = t & (KV2_STAGES - 1);
        mbar_wait(kv_full0 + 8 * st, (t / KV2_STAGES) & 1);
        tc_fence_after();
        const int keys = min(TBN2, p.T - t * TBN2);
        const uint32_t idesc_s = make_idesc_bf16(TBM, (keys + 15) & ~15);
        const uint64_t kd = make_desc_kmajor_sw128(sK + st * KV2_BYTES);
#pragma unroll
        for (int k = 0; k < THD / 16; ++k)
          tc_mma_bf16_ts(tmem_base + (uint32_t)((t & 1) * TBN2), tmem_base + COL_Q + 8 * k, kd + 2 * k, idesc_s,
                         k > 0 ? 1u : 0u);
        tc_commit(s_full0 + 8 * (t & 1));
      };
      for (int t = 0; t < min(KV2_AHEAD, n_kv); ++t) load_kv(t);
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      const uint32_t idesc_o = make_idesc_bf16(TBM, THD, 0, 1);
      for (int j = 0; j < n_kv; ++j) {
        // S(j+1) overwrites the buffer that held S(j-1) and P(j-1): the tensor core runs this thread's MMAs in issue
        // order, so it follows P.V(j-1), and the readers of S(j-1) arrived on p_ready(j-1) before that was issued
        if (j + 1 < n_kv) issue_s(j + 1);
        const int t = j + KV2_AHEAD;  // refill the stage tile j-1 used, once P.V(j-1) has retired (see KV2_AHEAD)
        if (t < n_kv) {
          if (t >= KV2_STAGES) mbar_wait(kv_empty0 + 8 * (t & (KV2_STAGES - 1)), ((t / KV2_STAGES) - 1) & 1);
          load_kv(t);
        }
        // O_tile(j) = P(j) V(j) : M=128, N=64, K = keys of this step rounded up to 16; A = P in TMEM, B = V (MN-major)
        mbar_wait(p_ready0 + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        const int st = j & (KV2_STAGES - 1);
        const int n_mma = (min(TBN2, p.T - j * TBN2) + 15) & ~15;
        for (int k = 0; k < n_mma / 16; ++k) {
          const uint64_t vd = make_desc_mnmajor_sw128(sV + st * KV2_BYTES + k * 2048, 1024);
          tc_mma_bf16_ts(tmem_base + COL_O, tmem_base + (uint32_t)((j & 1) * TBN2 + 8 * k), vd, idesc_o,
                         (k > 0 || (OACC && j > 0)) ? 1u : 0u);
        }
        tc_commit(kv_empty0 + 8 * st);
        tc_commit(o_full);
      }
    }
  } else {
    // ---------------- softmax / accumulate: two threads per query row, 32 of the step's 64 keys each ----------------
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;  // query row within the tile = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    constexpr int OH = THD / 2;  // output columns per thread
    {
      // this thread's half of its query row (32 bf16 = 16 packed columns) -> TMEM: the A operand of every S = Q K^T
      const bf16* qrow_p = qkv + (int64_t)(row_base + q0 + r) * (3 * p.d) + h * THD;
      if (p.q_rope == nullptr) {
        // (loaded before the CTA's setup)
      } else {
        // Q arrives un-rotated (the RoPE pass then only touches K: a third less HBM traffic for it); same arithmetic
        // as rope_kernel: out[i] = x[i] cos_i - x[i+32] sin_i, out[i+32] = x[i+32] cos_i + x[i] sin_i, one rounding
        const float4* cs = reinterpret_cast<const float4*>(p.q_rope + (int64_t)(q0 + r) * THD);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float lo[8], hi[8], out[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(qrow_p + 8 * c)), lo);
          unpack8(__ldg(reinterpret_cast<const uint4*>(qrow_p + 32 + 8 * c)), hi);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float4 t = __ldg(cs + 4 * c + (j >> 1));  // (cos_j, sin_j, cos_j+1, sin_j+1)
            if (half == 0) {
              out[j] = lo[j] * t.x - hi[j] * t.y;
              out[j + 1] = lo[j + 1] * t.z - hi[j + 1] * t.w;
            } else {
              out[j] = hi[j] * t.x + lo[j] * t.y;
              out[j + 1] = hi[j + 1] * t.z + lo[j + 1] * t.w;
            }
          }
          const uint4 u = pack8(out);
          qr[4 * c] = u.x; qr[4 * c + 1] = u.y; qr[4 * c + 2] = u.z; qr[4 * c + 3] = u.w;
        }
      }
      if (PAIR && n_kv <= 32 && p.valid_words != nullptr) {
        // one load per lane, issued with the Q loads: a single global round trip in the CTA's prologue (as bytes +
        // ballots the loop below makes one round trip per four steps, ncu: 11 % of the warp samples of kernel 6)
        if (!early_q && lane < n_kv) step_masks = __ldg(p.valid_words + (int64_t)b * (2 * n_kv) + 2 * lane + half);
      } else if (PAIR && n_kv <= 32) {
        // Validity bits of this warp's 32 keys of EVERY step, taken here, under the latency of the Q loads above: lane j
        // keeps the word of step j.  (ncu: as one byte load + ballot per step, the compare behind the load was the
        // single hottest stall of the step loop, 8 % of the kernel's warp samples.)
        const uint8_t* vg = p.key_valid ? p.key_valid + (int64_t)b * p.T : nullptr;
#pragma unroll 4
        for (int j = 0; j < n_kv; ++j) {
          const int kidx = j * TBN2 + half * 32 + lane;
          bool ok = kidx < p.T;
          if (ok && vg) ok = vg[kidx] != 0;
          const uint32_t word = __ballot_sync(0xffffffffu, ok);
          if (lane == j) step_masks = word;
        }
      }
      tmem_st_32x32b_x16(t_lane + COL_Q + half * 16, qr);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(q_full);
    }
    const uint8_t* valid_g = p.key_valid ? p.key_valid + (int64_t)b * p.T : nullptr;
    uint32_t* vwords = reinterpret_cast<uint32_t*>(valid_smem);  // [2 parities][2 words]: validity bit per key
    const bool live = q0 + quad * 32 + 31 >= q_tile * TBM;  // see the first kernel
    auto publish_valid = [&](int j) {  // threads r < 64 of half 0: validity bit of key j*64 + r
      if (half == 0 && quad < 2) {
        const int kidx = j * TBN2 + r;
        bool ok = kidx < p.T;
        if (ok && valid_g) ok = valid_g[kidx] != 0;
        const uint32_t word = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) vwords[(j & 1) * 2 + quad] = word;
      }
    };
    auto pair_sync = [&]() {  // warps quad and quad + 4; literal ids so that ptxas reserves 5 barriers, not all 16
      if (quad == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
      else if (quad == 1) asm volatile("bar.sync 2, 64;" ::: "memory");
      else if (quad == 2) asm volatile("bar.sync 3, 64;" ::: "memory");
      else asm volatile("bar.sync 4, 64;" ::: "memory");
    };
    auto own_mask = [&](int j) -> uint32_t {  // PAIR: validity bits of this warp's own 32 keys of step j
      const int kidx = j * TBN2 + half * 32 + lane;
      bool ok = kidx < p.T;
      if (ok && valid_g) ok = valid_g[kidx] != 0;
      return __ballot_sync(0xffffffffu, ok);
    };
    if (!live) {
      // (PAIR: both warps of a pair are dead together and nobody else waits for them at a bar.sync; the mbarrier
      // waits alone keep their arrivals one per phase)
      for (int j = 0; j < n_kv; ++j) {
        if (!PAIR) publish_valid(j);
        mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
        if (!PAIR) {
          asm volatile("bar.sync 1, 256;" ::: "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        if (j > 0) mbar_wait(o_full, (j - 1) & 1);
        mbar_arrive(p_ready0 + 8 * (j & 1));
      }
      mbar_wait(o_full, (n_kv - 1) & 1);
      if (!PAIR) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    } else {
      // the running output row as 16 fp32 PAIRS: rescale + accumulate is one FFMA2 per two columns
      uint64_t o2[OH / 2];
#pragma unroll
      for (int i = 0; i < OH / 2; ++i) o2[i] = 0ull;
      float m_run = -INFINITY, l_run = 0.f, corr_prev = 1.f;
      auto fold_o = [&](int jj, float corr) {  // o = o * corr + P.V tile of step jj (this thread's 32 columns)
        mbar_wait(o_full, jj & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + COL_O + half * OH, v);
        tc_wait_ld();
        const uint64_t c2 = f2_bcast(corr);
#pragma unroll
        for (int i = 0; i < OH / 2; ++i)
          o2[i] = f2_fma(o2[i], c2, f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])));
      };
      // new running maximum and the factor that brings the row sum (and O) from the old offset to the new one
      auto raise_max = [&](float mx, float& m_new, float& corr) {
        const float m_cand = fmaxf(m_run, mx);
        if (OACC) {
          const bool first = (m_run == -INFINITY);
          const float d = first ? 0.f : (m_cand - m_run) * p.scale_log2;
          const bool grow = (m_cand != -INFINITY) && (first || d > 8.f);
          m_new = grow ? m_cand : m_run;
          corr = (grow && !first) ? exp2f(-d) : 1.f;
        } else {
          m_new = m_cand;
          corr = (m_new == -INFINITY) ? 1.f : exp2f((m_run - m_new) * p.scale_log2);
        }
      };
      // what has to happen to O between the P store of step j and the release of P(j) to the MMA thread
      auto settle_o = [&](int j, float corr) {
        if (OACC) {
          // Every thread passes EVERY phase of o_full in order, needed or not: a parity wait can only tell the current
          // phase from the one before it, and a thread that skipped phases would, at the end, take "P.V(n-2) still
          // running" for "P.V(n-1) done" (seen on B200 as run-to-run differences in rows whose last step is a fast one).
          // P.V(j-1) has had a whole step to finish, so this wait is almost always a single successful poll.
          if (j > 0) mbar_wait(o_full, (j - 1) & 1);
          if (j > 0 && __any_sync(0xffffffffu, corr != 1.f)) {
            tc_fence_after();
            uint32_t ov[32];
            tmem_ld_32x32b_x32(t_lane + COL_O + half * OH, ov);
            tc_wait_ld();
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              lo[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
              hi[i] = __float_as_uint(__uint_as_float(ov[16 + i]) * corr);
            }
            tmem_st_32x32b_x16(t_lane + COL_O + half * OH, lo);
            tmem_st_32x32b_x16(t_lane + COL_O + half * OH + 16, hi);
          }
        } else {
          // the P.V tile of the previous step (complete for about a step by now) must be folded in before P.V(j),
          // which overwrites it, can be released
          if (j > 0) fold_o(j - 1, corr_prev);
          corr_prev = corr;
        }
      };
      for (int j = 0; j < n_kv; ++j) {
        uint32_t mw;
        if (PAIR) mw = (n_kv <= 32) ? __shfl_sync(0xffffffffu, step_masks, j) : own_mask(j);
        else publish_valid(j);
        mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        if (PAIR && mw == 0u) {
          // None of this warp's 32 keys is valid (the upper half of the last step when T % 64 <= 32 — 512 residues +
          // BOS + EOS leave 2 keys for step 9 — or a block of padding): no scores to read, no exponentials; only the
          // row maximum of the partner warp (for the rescale of this thread's O columns) and the appointments.
          const int xo = (j & 1) * 2 * TBM;
          xchg[xo + half * TBM + r] = __float2bfloat16_ru(-INFINITY);
          pair_sync();
          const float mx = __bfloat162float(xchg[xo + (half ^ 1) * TBM + r]);
          float m_new, corr;
          raise_max(mx, m_new, corr);
          if (j * TBN2 + half * 32 < p.T) {  // P.V(j) reads this warp's P columns (K = keys of the step up to T): zeros
            uint32_t zeros[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) zeros[i] = 0u;
            tmem_st_32x32b_x16(t_lane + (uint32_t)((j & 1) * TBN2 + half * 16), zeros);
          }
          l_run *= corr;
          m_run = m_new;
          settle_o(j, corr);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(p_ready0 + 8 * (j & 1));
          continue;
        }
        uint32_t v[32];  // this thread's 32 scores of the step stay in registers for both passes
        tmem_ld_32x32b_x32(t_lane + (uint32_t)((j & 1) * TBN2 + half * 32), v);
        tc_wait_ld();
        if (!PAIR) {
          asm volatile("bar.sync 1, 256;" ::: "memory");  // validity words visible; every thread holds its scores
          mw = vwords[(j & 1) * 2 + half];
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (mw == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            mx4[i & 3] = fmaxf(mx4[i & 3], ((mw >> i) & 1u) ? __uint_as_float(v[i]) : -INFINITY);
        }
        const __nv_bfloat16 mx_own_b = __float2bfloat16_ru(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])));
        // (PAIR: slots double-buffered by step parity — the partner may read this step's value after this thread has
        // gone on to the next step; the barrier also tells each thread that its partner holds its scores, whose
        // TMEM columns the two P stores below overwrite)
        const int xo = PAIR ? (j & 1) * 2 * TBM : 0;
        xchg[xo + half * TBM + r] = mx_own_b;
        if (PAIR) pair_sync();
        else asm volatile("bar.sync 1, 256;" ::: "memory");
        const float mx = fmaxf(__bfloat162float(mx_own_b), __bfloat162float(xchg[xo + (half ^ 1) * TBM + r]));
        float m_new, corr;
        raise_max(mx, m_new, corr);
        const float moff = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
        // p = exp2(s*scale - moff) -> bf16 pairs -> TMEM columns [16*half, 16*half + 16) of this step's S buffer
        // (all 256 threads passed the first bar.sync with their scores in registers, so the buffer is free to reuse)
        float ls4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t packed[16];
        if (PAIR && mw == 0xffffffffu) {  // every key of this warp's 32 is valid: no select per score
          // exponent and row sum on pairs of scores (FFMA2 / FADD2: half the FMA-pipe instructions)
          const uint64_t sc2 = f2_bcast(p.scale_log2), nm2 = f2_bcast(-moff);
          uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0, e1, p0, p1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), e0, e1);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e0));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e1));
            ls2[(i >> 1) & 3] = f2_add(ls2[(i >> 1) & 3], f2_pack(p0, p1));
            packed[i >> 1] = pack_bf16x2(p0, p1);
          }
          float a0, a1;
          f2_unpack(f2_add(f2_add(ls2[0], ls2[1]), f2_add(ls2[2], ls2[3])), a0, a1);
          ls4[0] = a0 + a1;
        } else
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0, p1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v[i]), p.scale_log2, -moff)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -moff)));
          if (mw != 0xffffffffu) {
            p0 = ((mw >> i) & 1u) ? p0 : 0.f;
            p1 = ((mw >> (i + 1)) & 1u) ? p1 : 0.f;
          }
          ls4[(i >> 1) & 3] += p0 + p1;
          if (ALU_PACK)
            packed[i >> 1] = __byte_perm(__float_as_uint(p0) + 0x8000u, __float_as_uint(p1) + 0x8000u, 0x7632);
          else
            packed[i >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x16(t_lane + (uint32_t)((j & 1) * TBN2 + half * 16), packed);
        l_run = l_run * corr + ((ls4[0] + ls4[1]) + (ls4[2] + ls4[3]));
        m_run = m_new;
        settle_o(j, corr);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(p_ready0 + 8 * (j & 1));
      }
      float o[OH];
      if (OACC) {
        mbar_wait(o_full, (n_kv - 1) & 1);
        tc_fence_after();
        uint32_t ov[32];
        tmem_ld_32x32b_x32(t_lane + COL_O + half * OH, ov);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < OH; ++i) o[i] = __uint_as_float(ov[i]);
      } else {
        fold_o(n_kv - 1, corr_prev);
#pragma unroll
        for (int i = 0; i < OH / 2; ++i) f2_unpack(o2[i], o[2 * i], o[2 * i + 1]);
      }
      // ---- finalize: row sum = both halves ----
      if (PAIR) pair_sync();
      else asm volatile("bar.sync 1, 256;" ::: "memory");
      xchg_f[half * TBM + r] = l_run;
      if (PAIR) pair_sync();
      else asm volatile("bar.sync 1, 256;" ::: "memory");
      const float l_tot = l_run + xchg_f[(half ^ 1) * TBM + r];
      const int qrow = q0 + r;
      if (qrow < p.T && qrow >= q_tile * TBM) {  // rows below q_tile*TBM belong to the previous tile
        const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
        bf16* op = p.o + (int64_t)(row_base + qrow) * p.o_rs + h * THD + half * OH;
#pragma unroll
        for (int c = 0; c < OH; c += 8) {
          uint4 u;
          u.x = pack_bf16x2(o[c] * inv, o[c + 1] * inv);
          u.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
          u.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv);
          u.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
          *reinterpret_cast<uint4*>(op + c) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SM_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS_ATT);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 6: ONE thread per query row.
//
// In kernels 2-5 a row's 64 scores of a step are split over two threads in two warps, which then have to agree on the
// row maximum (shared-memory exchange + a named barrier in every step), and eight softmax warps per CTA each pay the
// fixed cost of a step (waits, fences, mask word, rescale factor, loop) for 32 scores per thread.  ncu of kernel 4 / 5
// (profiles/r02_ncu_esm_attention_pair.txt): 40 % of the issue slots used, no pipe saturated, the softmax threads'
// serial chain per step is what the CTA waits for.  Here four softmax warps own one TMEM lane quadrant each and a thread
// does its row's whole step: 64 scores in registers (two tcgen05.ld), the maximum without any exchange (and exact in
// fp32: no bf16 rounding of an exchanged value), 64 exponentials, one P store of 32 packed columns, and — as in kernel
// 5 — O accumulated in TMEM and rescaled only when the maximum grows by more than 2^8.  No thread ever touches another
// thread's TMEM lane, so the only synchronisation left is with the MMA thread (s_full / p_ready / o_full).  Five warps per
// CTA leave 204 registers per thread at two CTAs per SM.  The MMA / TMA thread is kernel 5's.
// ---------------------------------------------------------------------------------------------------------------
constexpr int RW_WARPS = 4;
constexpr int RW_THREADS = (RW_WARPS + 1) * 32;

__global__ void __launch_bounds__(RW_THREADS, 2)
esm_attention_row_kernel(const __grid_constant__ CUtensorMap tmap, const TcAttnParams p, const bf16* __restrict__ qkv) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t sK = base;                             // KV2_STAGES stages
  const uint32_t sV = sK + KV2_STAGES * KV2_BYTES;      // KV2_STAGES stages
  const uint32_t bars = sV + KV2_STAGES * KV2_BYTES;
  const uint32_t q_full = bars, kv_full0 = bars + 8, kv_empty0 = kv_full0 + 8 * KV2_STAGES,
                 s_full0 = kv_empty0 + 8 * KV2_STAGES, p_ready0 = s_full0 + 16, o_full = p_ready0 + 16,
                 tmem_slot = o_full + 8;
  constexpr uint32_t COL_O = 2 * TBN2, COL_Q = 2 * TBN2 + THD;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = min(q_tile * TBM, p.T - TBM);  // shifted last tile, see the first kernel
  const int row_base = b * p.T;
  const int n_kv = (p.T + TBN2 - 1) / TBN2;

  // first global loads (query row, validity words) before the CTA's setup, see esm_attention_ts_kernel
  uint32_t qa[16], qb[16];
  uint32_t masks_lo = 0, masks_hi = 0;  // lane j: validity words of keys [64 j, 64 j + 32) and [64 j + 32, 64 j + 64)
  if (warp < RW_WARPS) {
    const uint4* qp = reinterpret_cast<const uint4*>(qkv + (int64_t)(row_base + q0 + warp * 32 + lane) * (3 * p.d) + h * THD);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = __ldg(qp + i), w = __ldg(qp + 4 + i);
      qa[4 * i] = u.x; qa[4 * i + 1] = u.y; qa[4 * i + 2] = u.z; qa[4 * i + 3] = u.w;
      qb[4 * i] = w.x; qb[4 * i + 1] = w.y; qb[4 * i + 2] = w.z; qb[4 * i + 3] = w.w;
    }
    if (p.valid_words != nullptr && lane < n_kv) {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(p.valid_words + (int64_t)b * (2 * n_kv)) + lane);
      masks_lo = w.x;
      masks_hi = w.y;
    }
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, RW_WARPS * 32);
    for (int s = 0; s < KV2_STAGES; ++s) {
      mbar_init(kv_full0 + 8 * s, 1);
      mbar_init(kv_empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full0 + 8 * s, 1);
      mbar_init(p_ready0 + 8 * s, RW_WARPS * 32);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == RW_WARPS) {
    tmem_alloc(tmem_slot, TMEM_COLS_ATT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == RW_WARPS) {
    if (lane == 0) {
      // ---------------- TMA producer + MMA issuer (one thread), as in kernel 5 ----------------
      auto load_kv = [&](int t) {
        const int st = t & (KV2_STAGES - 1);
        mbar_arrive_expect_tx(kv_full0 + 8 * st, 2 * KV2_BYTES);
        tma_load_2d(sK + st * KV2_BYTES, &tmap, kv_full0 + 8 * st, p.d + h * THD, row_base + t * TBN2);
        tma_load_2d(sV + st * KV2_BYTES, &tmap, kv_full0 + 8 * st, 2 * p.d + h * THD, row_base + t * TBN2);
      };
      auto issue_s = [&](int t) {  // S(t) = Q K(t)^T into S buffer t & 1; A = Q in TMEM
        const int st = t & (KV2_STAGES - 1);
        mbar_wait(kv_full0 + 8 * st, (t / KV2_STAGES) & 1);
        tc_fence_after();
        const int keys = min(TBN2, p.T - t * TBN2);
        const uint32_t idesc_s = make_idesc_bf16(TBM, (keys + 15) & ~15);
        const uint64_t kd = make_desc_kmajor_sw128(sK + st * KV2_BYTES);
#pragma unroll
        for (int k = 0; k < THD / 16; ++k)
          tc_mma_bf16_ts(tmem_base + (uint32_t)((t & 1) * TBN2), tmem_base + COL_Q + 8 * k, kd + 2 * k, idesc_s,
                         k > 0 ? 1u : 0u);
        tc_commit(s_full0 + 8 * (t & 1));
      };
      for (int t = 0; t < min(KV2_AHEAD, n_kv); ++t) load_kv(t);
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      const uint32_t idesc_o = make_idesc_bf16(TBM, THD, 0, 1);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int t = j + KV2_AHEAD;  // refill the stage tile j-1 used, once P.V(j-1) has retired (see KV2_AHEAD)
        if (t < n_kv) {
          if (t >= KV2_STAGES) mbar_wait(kv_empty0 + 8 * (t & (KV2_STAGES - 1)), ((t / KV2_STAGES) - 1) & 1);
          load_kv(t);
        }
        // O += P(j) V(j) : M=128, N=64, K = keys of this step rounded up to 16; A = P in TMEM, B = V (MN-major)
        mbar_wait(p_ready0 + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        const int st = j & (KV2_STAGES - 1);
        const int n_mma = (min(TBN2, p.T - j * TBN2) + 15) & ~15;
        for (int k = 0; k < n_mma / 16; ++k) {
          const uint64_t vd = make_desc_mnmajor_sw128(sV + st * KV2_BYTES + k * 2048, 1024);
          tc_mma_bf16_ts(tmem_base + COL_O, tmem_base + (uint32_t)((j & 1) * TBN2 + 8 * k), vd, idesc_o,
                         (k > 0 || j > 0) ? 1u : 0u);
        }
        tc_commit(kv_empty0 + 8 * st);
        tc_commit(o_full);
      }
    }
  } else {
    // ---------------- softmax: one thread per query row, all 64 keys of a step ----------------
    const int r = warp * 32 + lane;  // query row within the tile = TMEM lane (warp w may touch lanes 32 w .. 32 w + 31)
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    {
      // the query row (64 bf16 = 32 packed columns, loaded before the setup) -> TMEM: the A operand of every S = Q K^T;
      // without packed validity words: the bits of every step's keys from the bytes (n_kv <= 32 is checked on the host)
      const uint8_t* vg = p.key_valid ? p.key_valid + (int64_t)b * p.T : nullptr;
      if (p.valid_words == nullptr)
#pragma unroll 2
      for (int j = 0; j < n_kv; ++j) {
        const int k0 = j * TBN2 + lane, k1 = k0 + 32;
        bool ok0 = k0 < p.T, ok1 = k1 < p.T;
        if (ok0 && vg) ok0 = vg[k0] != 0;
        if (ok1 && vg) ok1 = vg[k1] != 0;
        const uint32_t w0 = __ballot_sync(0xffffffffu, ok0), w1 = __ballot_sync(0xffffffffu, ok1);
        if (lane == j) { masks_lo = w0; masks_hi = w1; }
      }
      tmem_st_32x32b_x16(t_lane + COL_Q, qa);
      tmem_st_32x32b_x16(t_lane + COL_Q + 16, qb);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(q_full);
    }
    const bool live = q0 + warp * 32 + 31 >= q_tile * TBM;  // see the first kernel
    if (!live) {
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
        if (j > 0) mbar_wait(o_full, (j - 1) & 1);
        mbar_arrive(p_ready0 + 8 * (j & 1));
      }
      mbar_wait(o_full, (n_kv - 1) & 1);
    } else {
      float m_run = -INFINITY, l_run = 0.f;
      // row maximum of 32 scores (invalid keys and the stale columns past the step's keys count as -inf)
      auto max32 = [&](const uint32_t (&v)[32], uint32_t mw) -> float {
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (mw == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            mx4[i & 3] = fmaxf(mx4[i & 3], ((mw >> i) & 1u) ? __uint_as_float(v[i]) : -INFINITY);
        }
        return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      };
      // p = exp2(s * scale - moff) of 32 scores -> 16 packed bf16 pairs; returns their sum
      auto exp32 = [&](const uint32_t (&v)[32], uint32_t mw, float moff, uint32_t (&packed)[16]) -> float {
        const uint64_t sc2 = f2_bcast(p.scale_log2), nm2 = f2_bcast(-moff);
        uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
        if (mw == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0, e1, p0, p1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), e0, e1);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e0));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e1));
            ls2[(i >> 1) & 3] = f2_add(ls2[(i >> 1) & 3], f2_pack(p0, p1));
            packed[i >> 1] = pack_bf16x2(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0, e1, p0, p1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), e0, e1);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e0));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e1));
            p0 = ((mw >> i) & 1u) ? p0 : 0.f;  // select, not multiply: a stale column may hold anything
            p1 = ((mw >> (i + 1)) & 1u) ? p1 : 0.f;
            ls2[(i >> 1) & 3] = f2_add(ls2[(i >> 1) & 3], f2_pack(p0, p1));
            packed[i >> 1] = pack_bf16x2(p0, p1);
          }
        }
        float a0, a1;
        f2_unpack(f2_add(f2_add(ls2[0], ls2[1]), f2_add(ls2[2], ls2[3])), a0, a1);
        return a0 + a1;
      };
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t mlo = __shfl_sync(0xffffffffu, masks_lo, j), mhi = __shfl_sync(0xffffffffu, masks_hi, j);
        const bool hi_read = j * TBN2 + 32 < p.T;  // P.V(j) reads the P columns of keys 32..63 (K = keys up to T)
        mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        const uint32_t s_col = t_lane + (uint32_t)((j & 1) * TBN2);
        uint32_t va[32], vb[32];
        if (mlo != 0u) tmem_ld_32x32b_x32(s_col, va);
        if (mhi != 0u) tmem_ld_32x32b_x32(s_col + 32, vb);
        tc_wait_ld();
        float mx = -INFINITY;
        if (mlo != 0u) mx = max32(va, mlo);
        if (mhi != 0u) mx = fmaxf(mx, max32(vb, mhi));
        // running maximum: raised only by more than 2^8 (see kernel 5); corr brings l and O to the new offset
        const float m_cand = fmaxf(m_run, mx);
        const bool first = (m_run == -INFINITY);
        const float dlt = first ? 0.f : (m_cand - m_run) * p.scale_log2;
        const bool grow = (m_cand != -INFINITY) && (first || dlt > 8.f);
        const float m_new = grow ? m_cand : m_run;
        const float corr = (grow && !first) ? exp2f(-dlt) : 1.f;
        const float moff = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
        // P(j) over this row's own scores: columns [0, 16) = keys 0..31, [16, 32) = keys 32..63 of S buffer j & 1
        float lsum = 0.f;
        {
          uint32_t packed[16];
          if (mlo != 0u) {
            lsum = exp32(va, mlo, moff, packed);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[i] = 0u;
          }
          tmem_st_32x32b_x16(s_col, packed);
        }
        if (hi_read) {
          uint32_t packed[16];
          if (mhi != 0u) {
            lsum += exp32(vb, mhi, moff, packed);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[i] = 0u;
          }
          tmem_st_32x32b_x16(s_col + 16, packed);
        }
        l_run = l_run * corr + lsum;
        m_run = m_new;
        // every phase of o_full is passed in order (see kernel 5); O is rescaled in TMEM only when a maximum was raised
        if (j > 0) mbar_wait(o_full, (j - 1) & 1);
        if (j > 0 && __any_sync(0xffffffffu, corr != 1.f)) {
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < THD; c += 32) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(t_lane + COL_O + c, ov);
            tc_wait_ld();
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              lo[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
              hi[i] = __float_as_uint(__uint_as_float(ov[16 + i]) * corr);
            }
            tmem_st_32x32b_x16(t_lane + COL_O + c, lo);
            tmem_st_32x32b_x16(t_lane + COL_O + c + 16, hi);
          }
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(p_ready0 + 8 * (j & 1));
      }
      // ---- finalize: O / row sum ----
      mbar_wait(o_full, (n_kv - 1) & 1);
      tc_fence_after();
      const int qrow = q0 + r;
      const bool write = qrow < p.T && qrow >= q_tile * TBM;  // rows below q_tile*TBM belong to the previous tile
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* op = p.o + (int64_t)(row_base + qrow) * p.o_rs + h * THD;
#pragma unroll
      for (int c = 0; c < THD; c += 32) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(t_lane + COL_O + c, ov);
        tc_wait_ld();
        if (write) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
            *reinterpret_cast<uint4*>(op + c + i) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RW_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS_ATT);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 7: kernel 6 as a PERSISTENT CTA.
//
// A (protein, head, query tile) item is nine key steps: with one CTA per item, launching the CTA, initialising its
// barriers, allocating TMEM, the first global round trip for Q and draining the pipeline at the end are a large part
// of its life (ncu of kernel 6: 11 % of the warp samples in the prologue, 6 % at the final barrier).  Here 2 x #SM CTAs
// walk the items round-robin.  All mbarriers keep running phase counters (global step g, global K/V tile count), TMEM
// is allocated once, the MMA thread runs ahead into the next item (its K/V tiles stream in while the softmax threads
// still finish the current one), and the softmax threads request the next item's query row and validity words before
// they read out the current item's O.  One more barrier, o_free: the first P.V of an item overwrites O (use_acc = 0), so
// it waits until every softmax thread has read the previous item's result.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RW_THREADS, 2)
esm_attention_row_persistent_kernel(const __grid_constant__ CUtensorMap tmap, const TcAttnParams p,
                                    const bf16* __restrict__ qkv, int n_items) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t sK = base;                             // KV2_STAGES stages
  const uint32_t sV = sK + KV2_STAGES * KV2_BYTES;      // KV2_STAGES stages
  const uint32_t bars = sV + KV2_STAGES * KV2_BYTES;
  const uint32_t q_full = bars, kv_full0 = bars + 8, kv_empty0 = kv_full0 + 8 * KV2_STAGES,
                 s_full0 = kv_empty0 + 8 * KV2_STAGES, p_ready0 = s_full0 + 16, o_full = p_ready0 + 16,
                 o_free = o_full + 8, tmem_slot = o_free + 8;
  constexpr uint32_t COL_O = 2 * TBN2, COL_Q = 2 * TBN2 + THD;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.T + TBN2 - 1) / TBN2;
  auto decode = [&](int it, int& q_tile, int& h, int& b) {  // query tile fastest: neighbours share K / V in L2
    q_tile = it % p.n_q_tiles;
    const int t = it / p.n_q_tiles;
    h = t % p.H;
    b = t / p.H;
  };
  // this thread's query row (64 bf16) and the validity words of the item's sequence (lane j: step j)
  uint32_t qa[16], qb[16];
  uint32_t masks_lo = 0, masks_hi = 0;
  auto request_item = [&](int it) {
    int q_tile, h, b;
    decode(it, q_tile, h, b);
    const int q0 = min(q_tile * TBM, p.T - TBM);
    const uint4* qp =
        reinterpret_cast<const uint4*>(qkv + (int64_t)(b * p.T + q0 + warp * 32 + lane) * (3 * p.d) + h * THD);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = __ldg(qp + i), w = __ldg(qp + 4 + i);
      qa[4 * i] = u.x; qa[4 * i + 1] = u.y; qa[4 * i + 2] = u.z; qa[4 * i + 3] = u.w;
      qb[4 * i] = w.x; qb[4 * i + 1] = w.y; qb[4 * i + 2] = w.z; qb[4 * i + 3] = w.w;
    }
    masks_lo = masks_hi = 0;
    if (lane < n_kv) {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(p.valid_words + (int64_t)b * (2 * n_kv)) + lane);
      masks_lo = w.x;
      masks_hi = w.y;
    }
  };
  if (warp < RW_WARPS && (int)blockIdx.x < n_items) request_item(blockIdx.x);  // before the setup, see kernel 6

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, RW_WARPS * 32);
    for (int s = 0; s < KV2_STAGES; ++s) {
      mbar_init(kv_full0 + 8 * s, 1);
      mbar_init(kv_empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full0 + 8 * s, 1);
      mbar_init(p_ready0 + 8 * s, RW_WARPS * 32);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, RW_WARPS * 32);
    fence_barrier_init();
  }
  if (warp == RW_WARPS) {
    tmem_alloc(tmem_slot, TMEM_COLS_ATT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == RW_WARPS) {
    if (lane == 0) {
      // ---------------- TMA producer + MMA issuer (one thread) ----------------
      uint32_t g0 = 0;  // global step of the item's step 0 (S / P buffer and barrier parities run across items)
      uint32_t kl = 0;  // K / V tiles requested so far (stage = count & 3)
      uint32_t ks = 0;  // K / V tiles consumed by an S = Q K^T so far
      int n = 0;        // items done by this CTA
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
        int q_tile, h, b;
        decode(it, q_tile, h, b);
        const int row_base = b * p.T;
        auto load_kv = [&](int t) {  // tile t of this item
          const uint32_t st = kl & (KV2_STAGES - 1);
          if (kl >= KV2_STAGES) mbar_wait(kv_empty0 + 8 * st, ((kl / KV2_STAGES) - 1) & 1);  // its last user's P.V retired
          mbar_arrive_expect_tx(kv_full0 + 8 * st, 2 * KV2_BYTES);
          tma_load_2d(sK + st * KV2_BYTES, &tmap, kv_full0 + 8 * st, p.d + h * THD, row_base + t * TBN2);
          tma_load_2d(sV + st * KV2_BYTES, &tmap, kv_full0 + 8 * st, 2 * p.d + h * THD, row_base + t * TBN2);
          ++kl;
        };
        const uint32_t k0 = ks;  // global count of this item's tile 0
        auto issue_s = [&](int t) {  // S(t) = Q K(t)^T into S buffer (g0 + t) & 1; A = Q in TMEM
          const uint32_t st = ks & (KV2_STAGES - 1);
          mbar_wait(kv_full0 + 8 * st, (ks / KV2_STAGES) & 1);
          tc_fence_after();
          const int keys = min(TBN2, p.T - t * TBN2);
          const uint32_t idesc_s = make_idesc_bf16(TBM, (keys + 15) & ~15);
          const uint64_t kd = make_desc_kmajor_sw128(sK + st * KV2_BYTES);
          const uint32_t gs = g0 + (uint32_t)t;
#pragma unroll
          for (int k = 0; k < THD / 16; ++k)
            tc_mma_bf16_ts(tmem_base + (gs & 1u) * TBN2, tmem_base + COL_Q + 8 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          tc_commit(s_full0 + 8 * (gs & 1u));
          ++ks;
        };
        for (int t = 0; t < min(KV2_AHEAD, n_kv); ++t) load_kv(t);
        mbar_wait(q_full, n & 1);
        tc_fence_after();
        issue_s(0);
        const uint32_t idesc_o = make_idesc_bf16(TBM, THD, 0, 1);
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) issue_s(j + 1);
          const int t = j + KV2_AHEAD;
          if (t < n_kv) load_kv(t);
          const uint32_t gj = g0 + (uint32_t)j;
          mbar_wait(p_ready0 + 8 * (gj & 1u), (gj >> 1) & 1u);
          if (j == 0 && n > 0) mbar_wait(o_free, (n - 1) & 1);  // the previous item's O has been read out
          tc_fence_after();
          const uint32_t st = (k0 + (uint32_t)j) & (KV2_STAGES - 1);
          const int n_mma = (min(TBN2, p.T - j * TBN2) + 15) & ~15;
          for (int k = 0; k < n_mma / 16; ++k) {
            const uint64_t vd = make_desc_mnmajor_sw128(sV + st * KV2_BYTES + k * 2048, 1024);
            tc_mma_bf16_ts(tmem_base + COL_O, tmem_base + (gj & 1u) * TBN2 + 8 * k, vd, idesc_o,
                           (k > 0 || j > 0) ? 1u : 0u);
          }
          tc_commit(kv_empty0 + 8 * st);
          tc_commit(o_full);
        }
        g0 += (uint32_t)n_kv;
      }
    }
  } else {
    // ---------------- softmax: one thread per query row, all 64 keys of a step ----------------
    const int r = warp * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    auto max32 = [&](const uint32_t (&v)[32], uint32_t mw) -> float {
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (mw == 0xffffffffu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          mx4[i & 3] = fmaxf(mx4[i & 3], ((mw >> i) & 1u) ? __uint_as_float(v[i]) : -INFINITY);
      }
      return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    };
    auto exp32 = [&](const uint32_t (&v)[32], uint32_t mw, float moff, uint32_t (&packed)[16]) -> float {
      const uint64_t sc2 = f2_bcast(p.scale_log2), nm2 = f2_bcast(-moff);
      uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
      if (mw == 0xffffffffu) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0, e1, p0, p1;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), e0, e1);
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e0));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e1));
          ls2[(i >> 1) & 3] = f2_add(ls2[(i >> 1) & 3], f2_pack(p0, p1));
          packed[i >> 1] = pack_bf16x2(p0, p1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0, e1, p0, p1;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), e0, e1);
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e0));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e1));
          p0 = ((mw >> i) & 1u) ? p0 : 0.f;
          p1 = ((mw >> (i + 1)) & 1u) ? p1 : 0.f;
          ls2[(i >> 1) & 3] = f2_add(ls2[(i >> 1) & 3], f2_pack(p0, p1));
          packed[i >> 1] = pack_bf16x2(p0, p1);
        }
      }
      float a0, a1;
      f2_unpack(f2_add(f2_add(ls2[0], ls2[1]), f2_add(ls2[2], ls2[3])), a0, a1);
      return a0 + a1;
    };
    uint32_t g0 = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      int q_tile, h, b;
      decode(it, q_tile, h, b);
      const int q0 = min(q_tile * TBM, p.T - TBM);
      const int row_base = b * p.T;
      const uint32_t mlo_all = masks_lo, mhi_all = masks_hi;  // (the registers are reloaded for the next item below)
      // the query row -> TMEM (every S MMA of the previous item has completed: this thread passed its last s_full)
      tmem_st_32x32b_x16(t_lane + COL_Q, qa);
      tmem_st_32x32b_x16(t_lane + COL_Q + 16, qb);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(q_full);
      const int it_next = it + (int)gridDim.x;
      const uint32_t g_last = g0 + (uint32_t)n_kv - 1u;
      const bool live = q0 + warp * 32 + 31 >= q_tile * TBM;
      if (!live) {
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t gj = g0 + (uint32_t)j;
          mbar_wait(s_full0 + 8 * (gj & 1u), (gj >> 1) & 1u);
          if (j > 0) mbar_wait(o_full, (gj - 1u) & 1u);
          mbar_arrive(p_ready0 + 8 * (gj & 1u));
        }
        if (it_next < n_items) request_item(it_next);
        mbar_wait(o_full, g_last & 1u);
        mbar_arrive(o_free);
      } else {
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t gj = g0 + (uint32_t)j;
          const uint32_t mlo = __shfl_sync(0xffffffffu, mlo_all, j), mhi = __shfl_sync(0xffffffffu, mhi_all, j);
          const bool hi_read = j * TBN2 + 32 < p.T;
          mbar_wait(s_full0 + 8 * (gj & 1u), (gj >> 1) & 1u);
          tc_fence_after();
          const uint32_t s_col = t_lane + (gj & 1u) * TBN2;
          uint32_t va[32], vb[32];
          if (mlo != 0u) tmem_ld_32x32b_x32(s_col, va);
          if (mhi != 0u) tmem_ld_32x32b_x32(s_col + 32, vb);
          tc_wait_ld();
          float mx = -INFINITY;
          if (mlo != 0u) mx = max32(va, mlo);
          if (mhi != 0u) mx = fmaxf(mx, max32(vb, mhi));
          const float m_cand = fmaxf(m_run, mx);
          const bool first = (m_run == -INFINITY);
          const float dlt = first ? 0.f : (m_cand - m_run) * p.scale_log2;
          const bool grow = (m_cand != -INFINITY) && (first || dlt > 8.f);
          const float m_new = grow ? m_cand : m_run;
          const float corr = (grow && !first) ? exp2f(-dlt) : 1.f;
          const float moff = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
          float lsum = 0.f;
          {
            uint32_t packed[16];
            if (mlo != 0u) {
              lsum = exp32(va, mlo, moff, packed);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) packed[i] = 0u;
            }
            tmem_st_32x32b_x16(s_col, packed);
          }
          if (hi_read) {
            uint32_t packed[16];
            if (mhi != 0u) {
              lsum += exp32(vb, mhi, moff, packed);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) packed[i] = 0u;
            }
            tmem_st_32x32b_x16(s_col + 16, packed);
          }
          l_run = l_run * corr + lsum;
          m_run = m_new;
          if (j > 0) mbar_wait(o_full, (gj - 1u) & 1u);  // every phase, in order (see kernel 5)
          if (j > 0 && __any_sync(0xffffffffu, corr != 1.f)) {
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < THD; c += 32) {
              uint32_t ov[32];
              tmem_ld_32x32b_x32(t_lane + COL_O + c, ov);
              tc_wait_ld();
              uint32_t lo[16], hi[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                lo[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
                hi[i] = __float_as_uint(__uint_as_float(ov[16 + i]) * corr);
              }
              tmem_st_32x32b_x16(t_lane + COL_O + c, lo);
              tmem_st_32x32b_x16(t_lane + COL_O + c + 16, hi);
            }
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(p_ready0 + 8 * (gj & 1u));
        }
        // the next item's query row and validity words are on their way while this item's result is read out
        if (it_next < n_items) request_item(it_next);
        mbar_wait(o_full, g_last & 1u);
        tc_fence_after();
        const int qrow = q0 + r;
        const bool write = qrow < p.T && qrow >= q_tile * TBM;  // rows below q_tile*TBM belong to the previous tile
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        bf16* op = p.o + (int64_t)(row_base + qrow) * p.o_rs + h * THD;
#pragma unroll
        for (int c = 0; c < THD; c += 32) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(t_lane + COL_O + c, ov);
          tc_wait_ld();
          if (c + 32 == THD) {  // all of O is in registers: the next item's first P.V may overwrite it
            tc_fence_before();
            mbar_arrive(o_free);
          }
          if (write) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
              *reinterpret_cast<uint4*>(op + c + i) = u;
            }
          }
        }
      }
      g0 += (uint32_t)n_kv;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RW_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS_ATT);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_qkv_map(const bf16* qkv, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
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
  cuuint32_t box[2] = {(cuuint32_t)THD, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(qkv), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PCY_ERR_CUDA, "cuTensorMapEncodeTiled(qkv) failed (%d)", (int)r);
  return 0;
}

}  // namespace

// validity bytes [B][T] -> bits: one warp per 32-key word, 2 * ceil(T / 64) words per sequence (an even count, so that the
// two words of a 64-key step are one aligned 8-byte load)
__global__ void pack_key_valid_kernel(const uint8_t* __restrict__ valid, uint32_t* __restrict__ words, int T, int nw,
                                      int total) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= total) return;
  const int b = w / nw, k = (w - b * nw) * 32 + lane;
  const bool ok = k < T && valid[(int64_t)b * T + k] != 0;
  const uint32_t word = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) words[w] = word;
}

// qkv bf16 [B*T, 3d] (q pre-scaled, RoPE applied), out bf16 [B*T, d]; covers all T query rows of every sequence when
// T >= 128 and head_dim == 64 (*rows_done = T), otherwise does nothing (*rows_done = 0).
bool esm_attention_tc_ropes_q(int T, int n_heads, int d) {
  return g_esm_attention_kernel >= 2 && g_esm_attention_kernel <= 5 && g_esm_attention_q_rope && T >= TBM &&
         d / n_heads == THD;
}

// q_rope: RoPE table for Q when the caller left Q un-rotated (only if esm_attention_tc_ropes_q() said so), else null
int esm_pack_key_valid(const uint8_t* key_valid, uint32_t* words, int B, int T, cudaStream_t stream) {
  PCY_REQUIRE(key_valid && words, "esm_pack_key_valid: null argument");
  const int nw = esm_key_valid_words(T), total = B * nw;
  if (total == 0) return 0;
  pack_key_valid_kernel<<<ceil_div(total, 8), 256, 0, stream>>>(key_valid, words, T, nw, total);
  PCY_LAUNCH_CHECK();
  return 0;
}

int esm_attention_tc(const bf16* qkv, const uint8_t* key_valid, const uint32_t* key_valid_words, bf16* out, int B, int T,
                     int n_heads, int d, float scale, const float* q_rope, int* rows_done, cudaStream_t stream) {
  *rows_done = 0;
  if (T < TBM || d / n_heads != THD) return 0;
  int n_q_tiles = (T + TBM - 1) / TBM;
  int covered = T;
  if (T % TBM != 0 && T % TBM <= g_esm_attention_tail_rows) {  // the caller runs the few rows left (rows_done < T)
    n_q_tiles = T / TBM;
    covered = n_q_tiles * TBM;
  }
  static SmemOptIn opt;
  if (opt.need(1)) {
    PCY_CUDA(cudaFuncSetAttribute(esm_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    PCY_CUDA(cudaFuncSetAttribute(esm_attention_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM));
    PCY_CUDA(cudaFuncSetAttribute(esm_attention_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
    PCY_CUDA(cudaFuncSetAttribute(esm_attention_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
    PCY_CUDA(cudaFuncSetAttribute(esm_attention_ts_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TS_SMEM));
    PCY_CUDA(cudaFuncSetAttribute(esm_attention_ts_kernel<false, true, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
    PCY_CUDA(cudaFuncSetAttribute(esm_attention_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
    PCY_CUDA(cudaFuncSetAttribute(esm_attention_row_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TS_SMEM));
  }
  int kern = g_esm_attention_kernel;
  if (kern >= 6 && (T + TBN2 - 1) / TBN2 > 32) kern = 5;  // kernels 6 / 7 keep a step's validity words in one lane each
  if (kern == 7 && (key_valid == nullptr || key_valid_words == nullptr)) kern = 6;  // the persistent kernel reads words
  const bool steps64 = kern != 0;
  CUtensorMap tmap;
  PCY_TRY(make_qkv_map(qkv, (int64_t)B * T, 3 * d, 3 * d, steps64 ? TBN2 : TBN, &tmap));
  TcAttnParams p;
  p.o = out; p.o_rs = d; p.key_valid = key_valid; p.valid_words = key_valid ? key_valid_words : nullptr;
  p.B = B; p.H = n_heads; p.T = T; p.d = d; p.n_q_tiles = n_q_tiles;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.q_rope = q_rope;
  PCY_REQUIRE(q_rope == nullptr || (kern >= 2 && kern <= 5),
              "esm_attention_tc: only the TMEM-operand kernels 2-5 rotate Q themselves");
  dim3 grid(n_q_tiles, n_heads, B);
  if (kern == 7) {
    const int n_items = n_q_tiles * n_heads * B;
    esm_attention_row_persistent_kernel<<<std::min(n_items, 2 * num_sms()), RW_THREADS, TS_SMEM, stream>>>(tmap, p, qkv,
                                                                                                          n_items);
  } else if (kern == 6) esm_attention_row_kernel<<<grid, RW_THREADS, TS_SMEM, stream>>>(tmap, p, qkv);
  else if (kern == 5) esm_attention_ts_kernel<false, true, true><<<grid, TC_THREADS, TS_SMEM, stream>>>(tmap, p, qkv);
  else if (kern == 4) esm_attention_ts_kernel<false, true><<<grid, TC_THREADS, TS_SMEM, stream>>>(tmap, p, qkv);
  else if (kern == 3) esm_attention_ts_kernel<true><<<grid, TC_THREADS, TS_SMEM, stream>>>(tmap, p, qkv);
  else if (kern == 2) esm_attention_ts_kernel<false><<<grid, TC_THREADS, TS_SMEM, stream>>>(tmap, p, qkv);
  else if (kern == 1) esm_attention_tc64_kernel<<<grid, TC_THREADS, TC2_SMEM, stream>>>(tmap, p);
  else esm_attention_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(tmap, p);
  PCY_LAUNCH_CHECK();
  *rows_done = covered;
  return 0;
}

}  // namespace pcy
