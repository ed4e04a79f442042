// Internal C++ API of the kernels (one function per op). The C ABI in capi.cu and the model
// drivers (esm_model.cu, llama_model.cu) are thin layers over these.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcy {

typedef __nv_bfloat16 bf16;

enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_SWIGLU = 2 };

// C[M,N] = epi(A[M,K] @ W[N,K]^T).  A, W bf16 row-major (K contiguous).
//   v = acc + bias[n];  if (n < scale_ncols) v *= scale;  v = act(v);  v += residual[m,n];  store (bf16 | fp32)
// ACT_SWIGLU: W rows are packed as 16 gate rows then 16 up rows per 32-row group; output has N/2 columns
// out[m, g*16 + j] = silu(v[m, g*32 + j]) * v[m, g*32 + 16 + j]; bias/scale apply before, residual after.
struct GemmArgs {
  const bf16* A = nullptr;
  int64_t lda = 0;
  const bf16* W = nullptr;
  int64_t ldw = 0;
  void* C = nullptr;
  int64_t ldc = 0;
  int M = 0, N = 0, K = 0;
  int c_fp32 = 0;
  const float* bias = nullptr;
  const bf16* residual = nullptr;
  int64_t ldr = 0;
  int act = ACT_NONE;
  float scale = 1.0f;
  int scale_ncols = 0;
};

// tcgen05 + TMA GEMM (any M; pads with TMA zero fill). Requires K % 8 == 0 and 16-byte aligned rows.
int gemm_bf16_tc(const GemmArgs& a, cudaStream_t stream);
// Weight-streaming skinny GEMM for M <= 16 rows (decode, projector at tiny batch). Optionally applies
// RMSNorm (fp32 statistics) to the A rows first: A' = A * rsqrt(mean(A^2)+eps) * rms_weight.
int gemm_bf16_skinny(const GemmArgs& a, const bf16* rms_weight, float rms_eps, cudaStream_t stream);
// Dispatch: skinny for M <= 16, tensor-core otherwise.
int gemm_bf16(const GemmArgs& a, cudaStream_t stream);

// ---- normalisation ------------------------------------------------------------------------------
int layernorm_bf16(const bf16* x, const bf16* gamma, const bf16* beta, bf16* y, int64_t rows, int d, float eps,
                   cudaStream_t stream);
int rmsnorm_bf16(const bf16* x, const bf16* weight, bf16* y, int64_t rows, int d, float eps, cudaStream_t stream);

// ---- ESM2 pieces -------------------------------------------------------------------------------
// x[b,t,:] = E[tok] * (1-0.12)/(1-mask_ratio_b); mask tokens -> 0; pad rows -> 0  (fair-esm ESM2.forward)
int esm_embed(const int32_t* tokens, const bf16* table, bf16* x, int B, int T, int d, int pad_idx, int mask_idx,
              int token_dropout, cudaStream_t stream);
// rotate-half RoPE on q and k inside a fused [rows, 3*d] qkv buffer; position = row % T.
int rope_qk_inplace(bf16* qkv, int64_t rows, int T, int n_heads, int head_dim, int64_t ld, const float* cos_sin,
                    cudaStream_t stream);
// fills cos_sin[T][head_dim/2][2] fp32
int rope_table(float* cos_sin, int T, int head_dim, float theta, int pos0, cudaStream_t stream);
// bidirectional attention with key-padding mask. qkv [B*T, 3d] (q pre-scaled), out [B*T, d].
int esm_attention(const bf16* qkv, const int32_t* tokens, bf16* out, int B, int T, int n_heads, int head_dim,
                  int pad_idx, cudaStream_t stream);
// ProteinPooler: segmented mean/max over non-pad rows grouped by batch_keys.
int pool_segments(const bf16* x, const int32_t* tokens, const int32_t* batch_keys, void* out, int out_fp32,
                  int n_rows, int T, int d, int n_out, int pad_idx, int mode /*0 mean,1 max*/, int correction,
                  cudaStream_t stream);

// ---- Llama pieces -------------------------------------------------------------------------------
int llama_embed_splice(const int32_t* ids, const bf16* table, const bf16* soft_tokens, const int32_t* soft_index,
                       bf16* x, int64_t n_tok, int d, cudaStream_t stream);

}  // namespace pcy
