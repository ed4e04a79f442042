// extern "C" surface of libprocyon_b200.so (declared in include/procyon_b200.h).
#include "../../include/procyon_b200.h"

#include "common.cuh"
#include "ops.h"

namespace pcy {
const char* last_error();
long long launch_count();
void reset_launch_count();

__global__ void pack_gate_up_kernel(const uint4* __restrict__ gate, const uint4* __restrict__ up,
                                    uint4* __restrict__ packed, int F, int K8) {
  // one thread per 16-byte chunk of the packed matrix
  const int64_t total = (int64_t)2 * F * K8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / K8;
    const int c = (int)(i % K8);
    const int g = (int)(row / 32), r = (int)(row % 32);
    const int src_row = g * 16 + (r & 15);
    packed[i] = (r < 16 ? gate : up)[(int64_t)src_row * K8 + c];
  }
}
}  // namespace pcy

using namespace pcy;

extern "C" {

const char* pcy_last_error(void) { return pcy::last_error(); }
int pcy_version(void) { return 100; }
long long pcy_launch_count(void) { return pcy::launch_count(); }
void pcy_reset_launch_count(void) { pcy::reset_launch_count(); }

static GemmArgs make_args(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                          int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                          int scale_ncols, int c_fp32) {
  GemmArgs a;
  a.A = (const bf16*)A; a.lda = lda; a.W = (const bf16*)W; a.ldw = ldw; a.C = C; a.ldc = ldc;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.residual = (const bf16*)residual; a.ldr = ldr;
  a.act = act; a.scale = scale; a.scale_ncols = scale_ncols; a.c_fp32 = c_fp32;
  return a;
}

int pcy_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                    int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                    int scale_ncols, int c_fp32, void* stream) {
  return gemm_bf16(make_args(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, act, scale, scale_ncols, c_fp32),
                   (cudaStream_t)stream);
}

int pcy_linear_bf16_ex(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                       int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                       int scale_ncols, int c_fp32, int force_tc, const void* rms_weight, float rms_eps,
                       void* stream) {
  GemmArgs a = make_args(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, act, scale, scale_ncols, c_fp32);
  if (force_tc) {
    if (rms_weight) return set_error(PCY_ERR_UNSUPPORTED, "fused RMSNorm is only available on the skinny path");
    return gemm_bf16_tc(a, (cudaStream_t)stream);
  }
  return gemm_bf16_skinny(a, (const bf16*)rms_weight, rms_eps, (cudaStream_t)stream);
}

int pcy_pack_gate_up(const void* gate, const void* up, void* packed, int F, int K, void* stream) {
  PCY_REQUIRE(F % 16 == 0 && K % 8 == 0, "pack_gate_up: F %% 16 and K %% 8 must be 0 (F=%d K=%d)", F, K);
  pack_gate_up_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)stream>>>((const uint4*)gate, (const uint4*)up,
                                                                      (uint4*)packed, F, K / 8);
  PCY_LAUNCH_CHECK();
  return 0;
}

int pcy_layernorm_bf16(const void* x, const void* gamma, const void* beta, void* y, int64_t rows, int d, float eps,
                       void* stream) {
  return layernorm_bf16((const bf16*)x, (const bf16*)gamma, (const bf16*)beta, (bf16*)y, rows, d, eps,
                        (cudaStream_t)stream);
}
int pcy_rmsnorm_bf16(const void* x, const void* weight, void* y, int64_t rows, int d, float eps, void* stream) {
  return rmsnorm_bf16((const bf16*)x, (const bf16*)weight, (bf16*)y, rows, d, eps, (cudaStream_t)stream);
}
int pcy_rope_inplace(void* x, int64_t rows, int T, int n_heads, int head_dim, int64_t ld, int col0,
                     const float* cos_sin, int pos0, void* stream) {
  return rope_inplace((bf16*)x, rows, T, n_heads, head_dim, ld, col0, cos_sin, nullptr, pos0, (cudaStream_t)stream);
}
int pcy_attention_bf16(const void* q, const void* k, const void* v, void* o, int64_t q_bs, int64_t q_rs, int q_hs,
                       int64_t k_bs, int64_t k_rs, int k_hs, int64_t v_bs, int64_t v_rs, int v_hs, int64_t o_bs,
                       int64_t o_rs, int o_hs, int B, int H, int KVH, int Tq, int Tk, int head_dim,
                       const uint8_t* key_valid, int64_t key_valid_bs, float scale, int causal, void* stream) {
  AttnArgs a;
  a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.o = (bf16*)o;
  a.q_bs = q_bs; a.q_rs = q_rs; a.q_hs = q_hs; a.k_bs = k_bs; a.k_rs = k_rs; a.k_hs = k_hs;
  a.v_bs = v_bs; a.v_rs = v_rs; a.v_hs = v_hs; a.o_bs = o_bs; a.o_rs = o_rs; a.o_hs = o_hs;
  a.B = B; a.H = H; a.KVH = KVH; a.Tq = Tq; a.Tk = Tk; a.head_dim = head_dim;
  a.key_valid = key_valid; a.key_valid_bs = key_valid_bs; a.scale = scale; a.causal = causal;
  return flash_attention(a, (cudaStream_t)stream);
}

int pcy_embed_splice(const int32_t* ids, const void* table, const void* soft_tokens, const int32_t* soft_index,
                     void* out, int64_t n_tok, int d, void* stream) {
  return llama_embed_splice(ids, (const bf16*)table, (const bf16*)soft_tokens, soft_index, (bf16*)out, n_tok, d,
                            (cudaStream_t)stream);
}

}  // extern "C"
