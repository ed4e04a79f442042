// Weight-streaming skinny GEMM (M <= 16 activation rows): the Llama decode step and the projector MLPs at
// tiny batch.  HBM-bound: every weight row is read exactly once with 16-byte streaming loads, fp32
// accumulation, warp-shuffle reduction.  Optional fused RMSNorm of the activation rows (HF LlamaRMSNorm
// semantics: fp32 variance, normalise, round to bf16, multiply by bf16 weight) and fused SwiGLU epilogue on
// the packed gate/up weight layout (see ops.h).
//
// Replaces cuBLAS GEMV calls under HF LlamaDecoderLayer at decode time (procyon/model/model_unified.py:769)
// and create_mlp at M = #proteins (procyon/model/model_unified.py:402-405).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ops.h"

namespace pcy {

bool g_pdl = true;
bool g_skinny_mma = true;  // pcy_set_skinny_mma(0): scalar-FMA kernel for every M <= 16 (A/B measurements, tests)
int skinny_mma_min_rows() {  // rows from which the tensor-core kernel is used (tuning knob)
  static const int v = [] { const char* e = getenv("PCY_SKINNY_MMA_MIN_M"); return e ? atoi(e) : 3; }();
  return v;
}

namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;

struct SkinnyParams {
  const bf16* A;
  int64_t lda;
  const bf16* W;
  int64_t ldw;
  void* C;
  int64_t ldc;
  const float* bias;
  const bf16* residual;
  int64_t ldr;
  const bf16* rms_weight;
  float rms_eps;
  int M, N, K;
  int c_fp32;
  int act;
  float scale;
  int scale_ncols;
  int num_units;
};

__device__ __forceinline__ float dot8(const uint4& a, const uint4& w, float s) {
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  const float2 w0 = unpack_bf16x2(w.x), w1 = unpack_bf16x2(w.y), w2 = unpack_bf16x2(w.z), w3 = unpack_bf16x2(w.w);
  s = fmaf(a0.x, w0.x, s); s = fmaf(a0.y, w0.y, s);
  s = fmaf(a1.x, w1.x, s); s = fmaf(a1.y, w1.y, s);
  s = fmaf(a2.x, w2.x, s); s = fmaf(a2.y, w2.y, s);
  s = fmaf(a3.x, w3.x, s); s = fmaf(a3.y, w3.y, s);
  return s;
}

// MT   = activation rows held in registers
// RPW  = weight rows per warp "unit"; KU = k-steps unrolled (RPW*KU independent 16-byte loads in flight per lane)
// ASMEM= activations staged in shared memory (required for the fused RMSNorm), else read through L1
template <int MT, int RPW, int KU, bool ASMEM>
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(const SkinnyParams p) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  bf16* sA = reinterpret_cast<bf16*>(sk_smem);  // [MT][K]
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int K = p.K;

  if (ASMEM) {
    for (int m = warp; m < MT; m += SK_WARPS) {
      bf16* dst = sA + (int64_t)m * K;
      if (m < p.M) {
        const bf16* src = p.A + (int64_t)m * p.lda;
        if (p.rms_weight != nullptr) {
          float ss = 0.f;
          for (int k = lane * 8; k < K; k += 256) {
            const uint4 u = *reinterpret_cast<const uint4*>(src + k);
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
                         d = unpack_bf16x2(u.w);
            ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
          }
          ss = warp_sum(ss);
          const float rstd = rsqrtf(ss / (float)K + p.rms_eps);
          for (int k = lane * 8; k < K; k += 256) {
            const uint4 u = *reinterpret_cast<const uint4*>(src + k);
            const uint4 w = *reinterpret_cast<const uint4*>(p.rms_weight + k);
            const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            uint32_t oo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 x = unpack_bf16x2(uu[i]);
              const float2 g = unpack_bf16x2(ww[i]);
              // HF LlamaRMSNorm: weight * (x_fp32 * rstd).to(bf16)
              oo[i] = pack_bf16x2(g.x * bf16_round(x.x * rstd), g.y * bf16_round(x.y * rstd));
            }
            *reinterpret_cast<uint4*>(dst + k) = make_uint4(oo[0], oo[1], oo[2], oo[3]);
          }
        } else {
          for (int k = lane * 8; k < K; k += 256)
            *reinterpret_cast<uint4*>(dst + k) = *reinterpret_cast<const uint4*>(src + k);
        }
      } else {
        for (int k = lane * 8; k < K; k += 256) *reinterpret_cast<uint4*>(dst + k) = make_uint4(0, 0, 0, 0);
      }
    }
    __syncthreads();
  }

  const bool swiglu = (p.act == ACT_SWIGLU);
  constexpr int H = RPW / 2;

  for (int unit = blockIdx.x * SK_WARPS + warp; unit < p.num_units; unit += gridDim.x * SK_WARPS) {
    // weight rows of this unit
    int rows[RPW];
    if (!swiglu) {
#pragma unroll
      for (int r = 0; r < RPW; ++r) rows[r] = min(unit * RPW + r, p.N - 1);
    } else {
      // packed groups of 32 rows: 16 gate then 16 up; a unit takes H gate rows and the matching H up rows
      const int per_group = 16 / (H > 0 ? H : 1);
      const int g = unit / per_group, sub = unit % per_group;
#pragma unroll
      for (int r = 0; r < RPW; ++r)
        rows[r] = g * 32 + (r < H ? 0 : 16) + sub * H + (r < H ? r : r - H);
    }
    float acc[RPW][MT];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[r][m] = 0.f;

    for (int k0 = lane * 8; k0 < K; k0 += 256 * KU) {
      uint4 w[KU][RPW];
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        const int k = k0 + u * 256;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          if (k < K) w[u][r] = ldg_nc_v4(p.W + (int64_t)rows[r] * p.ldw + k);
          else w[u][r] = make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        const int k = k0 + u * 256;
        if (k < K) {
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            uint4 a;
            if (ASMEM) a = *reinterpret_cast<const uint4*>(sA + (int64_t)m * K + k);
            else a = (m < p.M) ? __ldg(reinterpret_cast<const uint4*>(p.A + (int64_t)m * p.lda + k))
                               : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int r = 0; r < RPW; ++r) acc[r][m] = dot8(a, w[u][r], acc[r][m]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[r][m] = warp_sum(acc[r][m]);

    if (!swiglu) {
      for (int idx = lane; idx < RPW * MT; idx += 32) {
        const int r = idx / MT, m = idx % MT;
        const int n = unit * RPW + r;
        if (n >= p.N || m >= p.M) continue;
        float v = 0.f;
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
          for (int mm = 0; mm < MT; ++mm)
            if (rr == r && mm == m) v = acc[rr][mm];
        if (p.bias) v += p.bias[n];
        if (n < p.scale_ncols) v *= p.scale;
        if (p.act == ACT_GELU) v = gelu_erf(v);
        if (p.residual) v += __bfloat162float(p.residual[(int64_t)m * p.ldr + n]);
        if (p.c_fp32) reinterpret_cast<float*>(p.C)[(int64_t)m * p.ldc + n] = v;
        else reinterpret_cast<bf16*>(p.C)[(int64_t)m * p.ldc + n] = __float2bfloat16_rn(v);
      }
    } else if (H > 0) {
      const int per_group = 16 / (H > 0 ? H : 1);
      const int g = unit / per_group, sub = unit % per_group;
      for (int idx = lane; idx < H * MT; idx += 32) {
        const int r = idx / MT, m = idx % MT;
        if (m >= p.M) continue;
        float gv = 0.f, uv = 0.f;
#pragma unroll
        for (int rr = 0; rr < H; ++rr)
#pragma unroll
          for (int mm = 0; mm < MT; ++mm)
            if (rr == r && mm == m) { gv = acc[rr][mm]; uv = acc[rr + H][mm]; }
        if (p.bias) { gv += p.bias[rows[0] + r]; uv += p.bias[rows[0] + 16 + r]; }
        float v = silu(gv) * uv;
        const int ocol = g * 16 + sub * H + r;
        if (p.residual) v += __bfloat162float(p.residual[(int64_t)m * p.ldr + ocol]);
        reinterpret_cast<bf16*>(p.C)[(int64_t)m * p.ldc + ocol] = __float2bfloat16_rn(v);
      }
    }
  }
}

template <int MT, int RPW, int KU, bool ASMEM>
int launch_skinny(SkinnyParams p, cudaStream_t stream) {
  const size_t smem = ASMEM ? (size_t)MT * p.K * sizeof(bf16) : 0;
  static SmemOptIn opt;
  if (smem > 48 * 1024 && opt.need(smem))
    PCY_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<MT, RPW, KU, ASMEM>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.num_units = (p.act == ACT_SWIGLU) ? p.N / RPW : ceil_div(p.N, RPW);
  int grid = ceil_div(p.num_units, SK_WARPS);
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  gemm_skinny_kernel<MT, RPW, KU, ASMEM><<<grid, SK_THREADS, smem, stream>>>(p);
  PCY_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// 3..16 activation rows (beam search: the reference's evaluation default is beam_size = 10): the scalar kernel above
// needs ~2 FMA-pipe instructions per weight element and row and is ALU-bound there (18.8 ms per Llama-3-8B decode step
// at 10 rows against 3.2 ms at 1 row).  Here the legacy tensor cores do the math: a CTA of 4 warps owns 16 weight rows
// (32 for SwiGLU: the 16 gate rows and their 16 up rows), warp w streams k-blocks w, w+4, ... of 64 elements through
// its own cp.async ring (weights 16 x 128 B and the matching slice of the <= 16 activation rows, XOR-swizzled),
// ldmatrix + mma.sync.m16n8k16 with the weights as A and the activations as B, then the four partial accumulators
// meet in shared memory for the fused epilogue.  Every weight byte is still read exactly once.
constexpr int TK = 64;          // k elements per stage
constexpr int T_WARPS = 4;
constexpr int T_THREADS = T_WARPS * 32;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool SWIGLU, int STG>
__global__ void __launch_bounds__(T_THREADS)
gemm_skinny_mma_kernel(const SkinnyParams p) {
  constexpr int WT = SWIGLU ? 2 : 1;                 // 16-row weight tiles per CTA tile
  constexpr int STAGE_BYTES = (WT + 1) * 16 * 128;   // weights + activations
  extern __shared__ __align__(128) uint8_t tk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = smem_u32(tk_smem) + warp * (STG * STAGE_BYTES);
  float* red = reinterpret_cast<float*>(tk_smem + T_WARPS * STG * STAGE_BYTES);  // [T_WARPS][32][8 WT]
  const int n_tiles = (p.N + 16 * WT - 1) / (16 * WT);
  const int k_blocks = p.K / TK;
  const int my_blocks = (k_blocks - warp + T_WARPS - 1) / T_WARPS;  // k-blocks warp, warp + 4, ...

  pdl_launch_dependents();  // the next op of the chain may start ITS weight prefetch under this kernel's tail
  bool dep_pending = true;  // the activations (and C / residual) belong to the preceding kernel until pdl_wait()
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * 16 * WT;
    auto issue_w = [&](int i) {  // weights of k-block warp + 4 i of this tile -> ring slot i % STG
      const int k0 = (warp + i * T_WARPS) * TK;
      const uint32_t st = ring + (i % STG) * STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < 4 * WT; ++j) {  // 16 WT rows x 8 chunks of 16 B
        const int c = lane + 32 * j, r = c >> 3, ch = c & 7;
        const int row = min(row0 + r, p.N - 1);  // (rows past N are computed and dropped)
        cp_async16(st + r * 128 + ((ch ^ (r & 7)) << 4), p.W + (int64_t)row * p.ldw + k0 + ch * 8, true);
      }
    };
    auto issue_a = [&](int i) {  // the matching slice of the <= 16 activation rows (zero past M)
      const int k0 = (warp + i * T_WARPS) * TK;
      const uint32_t st = ring + (i % STG) * STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = lane + 32 * j, r = c >> 3, ch = c & 7;
        const bool ok = r < p.M;
        cp_async16(st + WT * 2048 + r * 128 + ((ch ^ (r & 7)) << 4), p.A + (int64_t)(ok ? r : 0) * p.lda + k0 + ch * 8,
                   ok);
      }
    };
    auto issue = [&](int i) { issue_w(i); issue_a(i); };
    float acc[WT][2][4];
#pragma unroll
    for (int t = 0; t < WT; ++t)
#pragma unroll
      for (int n = 0; n < 2; ++n) acc[t][n][0] = acc[t][n][1] = acc[t][n][2] = acc[t][n][3] = 0.f;
    if (dep_pending) {
      // first tile: the weights of the whole prologue go out before the dependency wait (group 0 then holds every
      // prologue weight stage plus activation stage 0: a group is complete only when all of them have landed)
#pragma unroll
      for (int i = 0; i < STG - 1; ++i)
        if (i < my_blocks) issue_w(i);
      pdl_wait();
      dep_pending = false;
#pragma unroll
      for (int i = 0; i < STG - 1; ++i) {
        if (i < my_blocks) issue_a(i);
        cp_async_commit();
      }
    } else {
#pragma unroll
      for (int i = 0; i < STG - 1; ++i) {
        if (i < my_blocks) issue(i);
        cp_async_commit();
      }
    }
    for (int i = 0; i < my_blocks; ++i) {
      cp_async_wait<STG - 2>();
      __syncwarp();
      if (i + STG - 1 < my_blocks) issue(i + STG - 1);  // slot (i - 1) % STG: read by every lane before the sync above
      cp_async_commit();
      const uint32_t st = ring + (i % STG) * STAGE_BYTES;
#pragma unroll
      for (int ks = 0; ks < TK / 16; ++ks) {
        uint32_t bf[4];
        {  // activations as B: matrices (m 0-7, k lo) (m 0-7, k hi) (m 8-15, k lo) (m 8-15, k hi)
          const int r = (lane & 7) + (lane >> 4) * 8, ch = ks * 2 + ((lane >> 3) & 1);
          ldsm_x4(st + WT * 2048 + r * 128 + ((ch ^ (r & 7)) << 4), bf);
        }
#pragma unroll
        for (int t = 0; t < WT; ++t) {
          uint32_t af[4];  // weights as A: (rows 0-7, k lo) (rows 8-15, k lo) (rows 0-7, k hi) (rows 8-15, k hi)
          const int r = (lane & 7) + ((lane >> 3) & 1) * 8, ch = ks * 2 + (lane >> 4);
          ldsm_x4(st + t * 2048 + r * 128 + ((ch ^ (r & 7)) << 4), af);
          mma16816(acc[t][0], af, bf[0], bf[1]);
          mma16816(acc[t][1], af, bf[2], bf[3]);
        }
      }
    }
    cp_async_wait<0>();
    // ---- the four warps' partial sums meet in shared memory ----
#pragma unroll
    for (int t = 0; t < WT; ++t)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[(warp * 32 + lane) * (8 * WT) + t * 8 + n * 4 + j] = acc[t][n][j];
    __syncthreads();
    // fragment entry (lane', e): row = lane'/4 + 8 (e>>1 & 1), m = 2 (lane'%4) + (e & 1) + 8 (e >> 2)
    for (int idx = threadIdx.x; idx < 32 * 8; idx += T_THREADS) {
      const int ln = idx >> 3, e = idx & 7;
      const int r = (ln >> 2) + ((e >> 1) & 1) * 8, m = (ln & 3) * 2 + (e & 1) + (e >> 2) * 8;
      float v[WT];
#pragma unroll
      for (int t = 0; t < WT; ++t) {
        v[t] = 0.f;
#pragma unroll
        for (int w = 0; w < T_WARPS; ++w) v[t] += red[(w * 32 + ln) * (8 * WT) + t * 8 + e];
      }
      if (m >= p.M) continue;
      if (!SWIGLU) {
        const int n = row0 + r;
        if (n >= p.N) continue;
        float x = v[0];
        if (p.bias) x += p.bias[n];
        if (n < p.scale_ncols) x *= p.scale;
        if (p.act == ACT_GELU) x = gelu_erf(x);
        if (p.residual) x += __bfloat162float(p.residual[(int64_t)m * p.ldr + n]);
        if (p.c_fp32) reinterpret_cast<float*>(p.C)[(int64_t)m * p.ldc + n] = x;
        else reinterpret_cast<bf16*>(p.C)[(int64_t)m * p.ldc + n] = __float2bfloat16_rn(x);
      } else {
        float g = v[0], u = v[WT - 1];
        if (p.bias) { g += p.bias[row0 + r]; u += p.bias[row0 + 16 + r]; }
        float x = silu(g) * u;
        const int ocol = tile * 16 + r;
        if (p.residual) x += __bfloat162float(p.residual[(int64_t)m * p.ldr + ocol]);
        reinterpret_cast<bf16*>(p.C)[(int64_t)m * p.ldc + ocol] = __float2bfloat16_rn(x);
      }
    }
    __syncthreads();  // red is reused by the next tile
  }
  if (dep_pending) pdl_wait();  // (a CTA without a tile: keep the chain's ordering anyway)
}

template <bool SWIGLU, int STG>
int launch_skinny_mma(const SkinnyParams& p, cudaStream_t stream) {
  constexpr int WT = SWIGLU ? 2 : 1;
  constexpr size_t smem = (size_t)T_WARPS * STG * (WT + 1) * 2048 + (size_t)T_WARPS * 32 * 8 * WT * sizeof(float);
  static SmemOptIn opt;
  if (opt.need(smem))
    PCY_CUDA(cudaFuncSetAttribute(gemm_skinny_mma_kernel<SWIGLU, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  const int n_tiles = ceil_div(p.N, 16 * WT);
  const int grid = std::min(n_tiles, num_sms() * 8);
  PCY_CUDA(launch_pdl(gemm_skinny_mma_kernel<SWIGLU, STG>, dim3(grid), dim3(T_THREADS), smem, stream, p));
  PCY_LAUNCH_CHECK();
  return 0;
}

template <int MT>
int dispatch_mt(const SkinnyParams& p, bool asmem, cudaStream_t stream) {
  // enough units to give every SM several warps; wide rows-per-warp only when N is large
  const bool wide = (p.N >= 8192) && MT <= 4;
  if (asmem) {
    if (wide) return launch_skinny<MT, 4, 2, true>(p, stream);
    return launch_skinny<MT, 2, 4, true>(p, stream);
  }
  if (wide) return launch_skinny<MT, 4, 2, false>(p, stream);
  return launch_skinny<MT, 2, 4, false>(p, stream);
}

}  // namespace

int gemm_bf16_skinny(const GemmArgs& a, const bf16* rms_weight, float rms_eps, cudaStream_t stream) {
  PCY_REQUIRE(a.M >= 1 && a.M <= 16, "skinny gemm: M=%d out of range [1,16]", a.M);
  PCY_REQUIRE(a.K % 8 == 0 && a.lda % 8 == 0 && a.ldw % 8 == 0, "skinny gemm: K/lda/ldw must be multiples of 8");
  PCY_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0,
              "skinny gemm: A and W must be 16-byte aligned");
  if (a.act == ACT_SWIGLU)
    PCY_REQUIRE(a.N % 32 == 0 && !a.c_fp32, "skinny gemm: SwiGLU needs N %% 32 == 0 and bf16 output");
  SkinnyParams p;
  p.A = a.A; p.lda = a.lda; p.W = a.W; p.ldw = a.ldw; p.C = a.C; p.ldc = a.ldc; p.bias = a.bias;
  p.residual = a.residual; p.ldr = a.ldr; p.rms_weight = rms_weight; p.rms_eps = rms_eps;
  p.M = a.M; p.N = a.N; p.K = a.K; p.c_fp32 = a.c_fp32; p.act = a.act; p.scale = a.scale;
  p.scale_ncols = a.scale_ncols; p.num_units = 0;
  // 3..16 rows without a fused norm: tensor-core kernel (needs whole 64-element k-blocks)
  if (a.M >= skinny_mma_min_rows() && rms_weight == nullptr && a.K % TK == 0 && g_skinny_mma) {
    if (a.act == ACT_SWIGLU) return launch_skinny_mma<true, 4>(p, stream);
    return launch_skinny_mma<false, 6>(p, stream);
  }
  const int mt = a.M <= 1 ? 1 : a.M <= 2 ? 2 : a.M <= 4 ? 4 : a.M <= 8 ? 8 : 16;
  const bool fits = (size_t)mt * a.K * 2 <= 160 * 1024;
  if (rms_weight != nullptr)
    PCY_REQUIRE(fits, "skinny gemm: fused RMSNorm needs M*K*2 <= 160 KB (M=%d K=%d)", a.M, a.K);
  const bool asmem = fits;
  switch (mt) {
    case 1: return dispatch_mt<1>(p, asmem, stream);
    case 2: return dispatch_mt<2>(p, asmem, stream);
    case 4: return dispatch_mt<4>(p, asmem, stream);
    case 8: return dispatch_mt<8>(p, asmem, stream);
    default: return dispatch_mt<16>(p, asmem, stream);
  }
}

int gemm_bf16(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 16) return gemm_bf16_skinny(a, nullptr, 0.f, stream);
  return gemm_bf16_tc(a, stream);
}

}  // namespace pcy
