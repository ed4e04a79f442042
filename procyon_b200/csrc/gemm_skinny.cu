// Weight-streaming skinny GEMM (M <= 16 activation rows): the Llama decode step and the projector MLPs at
// tiny batch.  HBM-bound: every weight row is read exactly once with 16-byte streaming loads, fp32
// accumulation, warp-shuffle reduction.  Optional fused RMSNorm of the activation rows (HF LlamaRMSNorm
// semantics: fp32 variance, normalise, round to bf16, multiply by bf16 weight) and fused SwiGLU epilogue on
// the packed gate/up weight layout (see ops.h).
//
// Replaces cuBLAS GEMV calls under HF LlamaDecoderLayer at decode time (procyon/model/model_unified.py:769)
// and create_mlp at M = #proteins (procyon/model/model_unified.py:402-405).
#include "common.cuh"
#include "ops.h"

namespace pcy {

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
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    PCY_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<MT, RPW, KU, ASMEM>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  p.num_units = (p.act == ACT_SWIGLU) ? p.N / RPW : ceil_div(p.N, RPW);
  int grid = ceil_div(p.num_units, SK_WARPS);
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  gemm_skinny_kernel<MT, RPW, KU, ASMEM><<<grid, SK_THREADS, smem, stream>>>(p);
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
