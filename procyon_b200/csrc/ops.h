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
  // tensor-core path only: fused rotate-half RoPE on columns [0, rope_ncols) (heads of rope_hd = 64 | 128 columns),
  // position = row % rope_T, table fp32 [P][rope_hd/2][2]; applied after bias/scale
  const float* rope = nullptr;
  int rope_hd = 0, rope_T = 0, rope_ncols = 0;
};

// tcgen05 + TMA GEMM (any M; pads with TMA zero fill). Requires K % 8 == 0 and 16-byte aligned rows.
int gemm_bf16_tc(const GemmArgs& a, cudaStream_t stream);
// Weight-streaming skinny GEMM for M <= 16 rows (decode, projector at tiny batch). Optionally applies
// RMSNorm (fp32 statistics) to the A rows first: A' = A * rsqrt(mean(A^2)+eps) * rms_weight.
int gemm_bf16_skinny(const GemmArgs& a, const bf16* rms_weight, float rms_eps, cudaStream_t stream);
// Dispatch: skinny for M <= 16, tensor-core otherwise.
int gemm_bf16(const GemmArgs& a, cudaStream_t stream);
int skinny_mma_min_rows();

// ---- normalisation ------------------------------------------------------------------------------
int layernorm_bf16(const bf16* x, const bf16* gamma, const bf16* beta, bf16* y, int64_t rows, int d, float eps,
                   cudaStream_t stream);
int rmsnorm_bf16(const bf16* x, const bf16* weight, bf16* y, int64_t rows, int d, float eps, cudaStream_t stream);

// ---- embeddings / rotary / pooling -------------------------------------------------------------
// x[b,t,:] = E[tok] * (1-0.12)/(1-mask_ratio_b); mask tokens -> 0; pad rows -> 0  (fair-esm ESM2.forward)
int esm_embed(const int32_t* tokens, const bf16* table, bf16* x, int B, int T, int d, int pad_idx, int mask_idx,
              int token_dropout, cudaStream_t stream);
int llama_embed_splice(const int32_t* ids, const bf16* table, const bf16* soft_tokens, const int32_t* soft_index,
                       bf16* x, int64_t n_tok, int d, cudaStream_t stream);
// rotate-half RoPE, in place, on n_heads heads starting at column col0 of each row; position =
// (pos_ptr ? *pos_ptr : pos0) + row % T; cos_sin fp32 [P][head_dim/2][2].
int rope_inplace(bf16* x, int64_t rows, int T, int n_heads, int head_dim, int64_t ld, int col0,
                 const float* cos_sin, const int32_t* pos_ptr, int pos0, cudaStream_t stream);
int rope_table(float* cos_sin, int P, int head_dim, float theta, cudaStream_t stream);
// ProteinPooler: segmented mean/max over non-pad rows. seg_ptr [n_out+1], seg_rows = chunk rows per output.
int pool_segments(const bf16* x, const int32_t* tokens, const int32_t* seg_ptr, const int32_t* seg_rows, void* out,
                  int out_fp32, int T, int d, int n_out, int pad_idx, int mode /*0 mean,1 max*/, int correction,
                  cudaStream_t stream);
// key_valid[i] = tokens[i] != pad_idx
int make_key_valid(const int32_t* tokens, uint8_t* valid, int64_t n, int pad_idx, cudaStream_t stream);

// ---- attention ------------------------------------------------------------------------------------
struct AttnArgs {
  const bf16* q = nullptr;
  const bf16* k = nullptr;
  const bf16* v = nullptr;
  bf16* o = nullptr;
  // element strides: batch, row (token), head
  int64_t q_bs = 0, q_rs = 0, k_bs = 0, k_rs = 0, v_bs = 0, v_rs = 0, o_bs = 0, o_rs = 0;
  int q_hs = 0, k_hs = 0, v_hs = 0, o_hs = 0;
  int B = 0, H = 0, KVH = 0, Tq = 0, Tk = 0, head_dim = 0;
  const uint8_t* key_valid = nullptr;  // [B, Tk] 1 = attend; may be null
  int64_t key_valid_bs = 0;
  float scale = 1.0f;  // multiplies q.k before softmax
  int causal = 0;      // key j visible to query i iff j <= i + (Tk - Tq)
};
int flash_attention(const AttnArgs& a, cudaStream_t stream);
// tcgen05 bidirectional attention for head_dim 64 on a fused [B*T, 3d] qkv buffer; covers the full 128-row query
// tiles of every sequence (rows_done = floor(T/128)*128, 0 if the shape is unsupported)
bool esm_attention_tc_ropes_q(int T, int n_heads, int d);
// key_valid_words: the same validity as bits, 2 * ceil(T / 64) words per sequence (esm_pack_key_valid), or null
int esm_attention_tc(const bf16* qkv, const uint8_t* key_valid, const uint32_t* key_valid_words, bf16* out, int B, int T,
                     int n_heads, int d, float scale, const float* q_rope, int* rows_done, cudaStream_t stream);
inline int esm_key_valid_words(int T) { return 2 * ((T + 63) / 64); }
int esm_pack_key_valid(const uint8_t* key_valid, uint32_t* words, int B, int T, cudaStream_t stream);

// ---- decode-step attention (RoPE + KV append + split-KV attention + combine, one launch) ------------
struct DecodeAttnArgs {
  const bf16* qkv = nullptr;  // [rows, (H + 2 KVH) * head_dim], un-roped output of the fused QKV projection
  int64_t qkv_ld = 0;
  const float* cos_sin = nullptr;
  const bf16* k_prompt = nullptr;  // [n_inputs][S][KVH*head_dim]
  const bf16* v_prompt = nullptr;
  bf16* k_gen = nullptr;  // [rows][max_gen][KVH*head_dim]
  bf16* v_gen = nullptr;
  const int32_t* slots = nullptr;         // [rows][max_gen]
  const uint8_t* prompt_valid = nullptr;  // [n_inputs][S] or null
  const int32_t* state = nullptr;         // device: state[0] = t
  float* partials = nullptr;
  int32_t* tickets = nullptr;  // [rows*KVH], zero-initialised once
  bf16* out = nullptr;         // [rows, H*head_dim]
  int rows = 0, beams = 1, H = 0, KVH = 0, head_dim = 0, S = 0, max_gen = 0;
};
int decode_attention(const DecodeAttnArgs& a, cudaStream_t stream);
int decode_attention_splits(int S, int max_gen);
int64_t decode_attention_partial_floats(int rows, int H, int KVH, int S, int max_gen);

int llama_attention_tc(const bf16* qkv, int64_t qkv_ld, const uint8_t* key_valid, bf16* out, int64_t o_rs, int B, int S,
                       int H, int KVH, int head_dim, float scale, int* done, cudaStream_t stream);

// ---- decode-step token selection ---------------------------------------------------------------------
struct DecodeSelectArgs {
  const float* logits = nullptr;  // [rows][vocab]
  float* logits_hist = nullptr;   // [max_gen][rows][vocab] or null
  int32_t* tokens = nullptr;      // [rows][max_gen]
  int32_t* slots = nullptr;       // [rows][max_gen]
  float* logprobs = nullptr;      // [rows]
  int32_t* state = nullptr;       // [8]: t, -, finished, finish_step, scratch
  float* workspace = nullptr;     // topk_workspace_floats(rows)
  int n_inputs = 0, beams = 1, group = 1, max_gen = 0, vocab = 0, eos_id = -1;
  float diversity_penalty = 0.f;
  int greedy = 0;
  int stop_on_all_eos = 0;
  int32_t* group_state = nullptr;  // int32 [4] shared by a lock-step group of sessions (see sampling.cu) or null
  int group_last = 1;
};
int decode_select(const DecodeSelectArgs& a, cudaStream_t stream);
int topk_workspace_floats(int rows);

// ---- persistent decode-step kernel (rows <= 4) --------------------------------------------------------------
struct LlamaLayerPtrs {
  const bf16 *ln1, *ln2, *wqkv, *wo, *wgu, *wdown;
};
}  // namespace pcy
#include "../../include/procyon_b200.h"
namespace pcy {
void decode_megakernel_set_timing(unsigned long long* dev_buf);
void decode_megakernel_set_self_refill(int enabled);
int decode_megakernel_set_shares(const float* shares, int n);
bool decode_megakernel_supported(const pcy_llama_config& c, int rows);
int64_t decode_megakernel_scratch_bytes(const pcy_llama_config& c, int rows, int S, int max_gen);
int decode_megakernel(const pcy_llama_config& c, const LlamaLayerPtrs* layers_dev, const bf16* embed,
                      const bf16* lm_head, const bf16* norm, const float* rope, const pcy_decode_buffers* b,
                      void* scratch, cudaStream_t stream);

// ---- persistent decode-step kernel for 3..16 rows (beam search), decode_rows_megakernel.cu ------------------------
void decode_rows_megakernel_set_timing(unsigned long long* dev_buf);
int decode_rows_build_maps(const pcy_llama_config& c, const LlamaLayerPtrs* layers_host, const bf16* lm_head,
                           void** maps_dev);
bool decode_rows_megakernel_supported(const pcy_llama_config& c, int rows, int beams);
int64_t decode_rows_megakernel_scratch_bytes(const pcy_llama_config& c, int rows, int beams, int S, int max_gen);
int decode_rows_megakernel(const pcy_llama_config& c, const LlamaLayerPtrs* layers_dev, const void* maps_dev,
                           const bf16* embed, const bf16* norm, const float* rope, const pcy_decode_buffers* b,
                           void* scratch, cudaStream_t stream);

// ---- losses / scoring -----------------------------------------------------------------------------------
int cross_entropy_rows(const float* logits, const int32_t* labels, int rows, int V, int64_t ld, float* acc,
                       cudaStream_t stream);
int normalize_rows(const float* x, float* out, int rows, int d, cudaStream_t stream);
int infonce_loss(const float* zs, const float* zt, const float* all_s, const float* all_t, const uint8_t* mask,
                 float* sims_scratch, float* loss, int b, int G, int d, int rank_off, float temperature,
                 cudaStream_t stream);
int retrieval_scores_topk(const float* Q, const void* D, int db_bf16, float* scores, int nq, int N, int d, int64_t lds,
                          int k, int index_base, float* top_val, int32_t* top_idx, int* ticket, cudaStream_t stream);
int cosine_scores(const float* Q, const void* D, int db_bf16, float* out, int nq, int N, int d, int64_t ldo,
                  cudaStream_t stream);

}  // namespace pcy
