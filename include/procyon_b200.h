/* procyon_b200 — C ABI of the B200 (sm_100a) hot path of ProCyon.
 *
 * The reference (mims-harvard/ProCyon) is pure Python: there is no FFI to mirror, so every entry point below
 * names the Python call site(s) of the reference it replaces (paths relative to the reference repo root).
 * Conventions: plain pointers + sizes, all tensors are DEVICE pointers unless stated, row-major, bf16 =
 * uint16 storage; every function takes the CUDA stream (as void*) it must enqueue on and returns 0 on
 * success or a negative pcy status (PCY_ERR_*), with a message available from pcy_last_error().
 * Nothing here allocates per call except the model handles (weights) — workspaces are caller-owned.
 */
#ifndef PROCYON_B200_H_
#define PROCYON_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCY_OK 0
#define PCY_ERR_INVALID_ARG (-1)
#define PCY_ERR_UNSUPPORTED (-2)
#define PCY_ERR_CUDA (-3)
#define PCY_ERR_WORKSPACE (-4)

/* activation codes of pcy_linear_bf16 */
#define PCY_ACT_NONE 0
#define PCY_ACT_GELU 1   /* exact erf GELU (torch.nn.GELU(), fair-esm gelu) */
#define PCY_ACT_SWIGLU 2 /* packed gate/up rows, see pcy_pack_gate_up */

/* ---- library state ------------------------------------------------------------------------------ */
const char* pcy_last_error(void);
int pcy_version(void);
/* number of CUDA kernels this library has launched since the last reset (bench.py: gpu_launches) */
long long pcy_launch_count(void);
void pcy_reset_launch_count(void);

/* ---- dense layers --------------------------------------------------------------------------------
 * C[M,N] = epi(A[M,K] @ W[N,K]^T): torch.nn.Linear as used by fair-esm ESM2 (procyon/model/esm.py:536),
 * HF LlamaDecoderLayer (procyon/model/pmc_llama.py:571) and create_mlp (procyon/model/model_utils.py:13-41).
 * epi: v = acc + bias[n]; if (n < scale_ncols) v *= scale; v = act(v); v += residual[m,n]; store.
 * bias is fp32 [N] or NULL; residual bf16 [M,ldr] or NULL; C is bf16 (c_fp32 = 0) or fp32.
 * M <= 16 takes the weight-streaming path, otherwise tcgen05 tensor cores. */
int pcy_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                    int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                    int scale_ncols, int c_fp32, void* stream);
/* same, but forces the tensor-core (force_tc = 1) or the weight-streaming (force_tc = 0) kernel */
int pcy_linear_bf16_ex(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                       int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                       int scale_ncols, int c_fp32, int force_tc, const void* rms_weight, float rms_eps,
                       void* stream);
/* interleave gate_proj / up_proj rows ([F,K] each) into the packed [2F,K] layout PCY_ACT_SWIGLU expects:
 * rows [32g, 32g+16) = gate rows [16g, 16g+16), rows [32g+16, 32g+32) = up rows [16g, 16g+16). F % 16 == 0. */
int pcy_pack_gate_up(const void* gate, const void* up, void* packed, int F, int K, void* stream);

/* ---- row kernels (exposed for parity tests and for host code that composes ops) ---------------------------
 * LayerNorm = torch.nn.LayerNorm (fair-esm ESM1bLayerNorm); RMSNorm = HF LlamaRMSNorm. bf16 in/out. */
int pcy_layernorm_bf16(const void* x, const void* gamma, const void* beta, void* y, int64_t rows, int d, float eps,
                       void* stream);
int pcy_rmsnorm_bf16(const void* x, const void* weight, void* y, int64_t rows, int d, float eps, void* stream);
/* rotate-half RoPE in place on n_heads heads from column col0; position = pos0 + row % T;
 * cos_sin fp32 [n_pos][head_dim/2][2] (fair-esm RotaryEmbedding / HF apply_rotary_pos_emb). */
int pcy_rope_inplace(void* x, int64_t rows, int T, int n_heads, int head_dim, int64_t ld, int col0,
                     const float* cos_sin, int pos0, void* stream);
/* softmax(scale * q k^T + mask) v without materialising scores. Strides in elements (batch, row, head).
 * key_valid: uint8 [B,Tk] (1 = attend) or NULL; causal: key j visible to query i iff j <= i + Tk - Tq.
 * fair-esm MultiheadAttention (procyon/model/esm.py:536), HF LlamaAttention (procyon/model/pmc_llama.py:221-247). */
int pcy_attention_bf16(const void* q, const void* k, const void* v, void* o, int64_t q_bs, int64_t q_rs, int q_hs,
                       int64_t k_bs, int64_t k_rs, int k_hs, int64_t v_bs, int64_t v_rs, int v_hs, int64_t o_bs,
                       int64_t o_rs, int o_hs, int B, int H, int KVH, int Tq, int Tk, int head_dim,
                       const uint8_t* key_valid, int64_t key_valid_bs, float scale, int causal, void* stream);

/* ---- ESM2 encoder -------------------------------------------------------------------------------------------
 * Replaces fair-esm `ESM2.forward(tokens, repr_layers=[L])` as called by ESM_PLM.forward
 * (procyon/model/esm.py:504-541): returns representations[L] (after emb_layer_norm_after), bf16 [B*T, d]. */
typedef struct {
  int n_layers, d_model, n_heads, ffn_dim, vocab;
  int pad_idx, mask_idx, token_dropout;
  float ln_eps;
} pcy_esm_config;

enum {
  PCY_ESM_EMBED = 0, /* bf16 [vocab,d]        embed_tokens.weight */
  PCY_ESM_LNF_G = 1, /* bf16 [d]              emb_layer_norm_after.weight */
  PCY_ESM_LNF_B = 2,
  PCY_ESM_LN1_G = 3, /* bf16 [d]              layers.N.self_attn_layer_norm */
  PCY_ESM_LN1_B = 4,
  PCY_ESM_WQKV = 5,  /* bf16 [3d,d]           cat(q_proj, k_proj, v_proj).weight */
  PCY_ESM_BQKV = 6,  /* fp32 [3d] */
  PCY_ESM_WO = 7,    /* bf16 [d,d]            out_proj */
  PCY_ESM_BO = 8,    /* fp32 [d] */
  PCY_ESM_LN2_G = 9, /* bf16 [d]              final_layer_norm */
  PCY_ESM_LN2_B = 10,
  PCY_ESM_W1 = 11,   /* bf16 [ffn,d]          fc1 */
  PCY_ESM_B1 = 12,   /* fp32 [ffn] */
  PCY_ESM_W2 = 13,   /* bf16 [d,ffn]          fc2 */
  PCY_ESM_B2 = 14    /* fp32 [d] */
};

/* head_dim-64 encoders run attention on tcgen05 by default; 0 selects the mma.sync kernel for every row (tests) */
int pcy_set_esm_tc_attention(int enabled);
/* tcgen05 ESM attention kernel: 0 = 128-key steps, MMAs and softmax of a CTA taking turns; 1 = 64-key steps with the
   S, P and P.V tiles double-buffered so that the MMAs of step j+1 run under the softmax of step j; 2 = the
   same schedule with both A operands in TMEM (Q written once per CTA, P written over the scores it came from); 3 = as
   2 with the bf16 pairs of P built on the ALU pipe (round half up) instead of the XU pipe's F2FP; 4 = as 2 with one
   64-thread named barrier per step between the two warps of a row pair instead of two CTA-wide bar.sync; 5 = as 4 with
   the output tile accumulated in TMEM across the steps (use_acc) and rescaled there only when a row's running maximum
   grows by more than 2^8, instead of a TMEM read + register fold of the P.V tile in every step (default); 6 = ONE thread
   per query row (four softmax warps, 64 scores per thread and step, no row-maximum exchange, no barrier between
   softmax warps; otherwise as 5): 21 % fewer instructions than 5 at the same speed on B200; 7 = kernel 6 as persistent
   CTAs (2 x #SM CTAs walk the (protein, head, query tile) items; barriers keep running phases, TMEM is allocated once,
   the next item's Q and K / V are requested under the current item's last steps): same speed again */
int pcy_set_esm_attention_kernel(int kernel);
/* rows: when the sequence length leaves at most `rows` query rows beyond the last full 128-row tile (512 residues +
   BOS + EOS = 4 tiles + 2 rows), the mma.sync kernel takes those rows instead of one more tcgen05 CTA per (protein,
   head) that streams all keys for them; 0: every row on the tcgen05 kernel */
int pcy_set_esm_attention_tail_rows(int rows);
/* 1: attention kernels 2 / 3 / 4 apply RoPE to Q while they move it into TMEM and the RoPE pass only rotates K;
   0 (default; measured faster overall on B200): the RoPE pass rotates Q and K in place before the attention kernel */
int pcy_set_esm_attention_q_rope(int enabled);
/* 1: apply RoPE in the QKV GEMM epilogue (head_dim 64/128, tensor-core path); 0 (default): separate vectorised pass */
int pcy_set_fused_rope(int enabled);
/* 1: tcgen05 GEMMs with >= 2 row-blocks run as 2-CTA clusters sharing the weight tile by TMA multicast;
   0 (default): independent CTAs. Both paths are bit-identical (tests); measured equal speed on B200 */
int pcy_set_gemm_cluster(int enabled);
/* tcgen05 GEMMs as 2-CTA clusters issuing ONE tcgen05.mma.cta_group::2 of M = 256 per k-step (each CTA stages its 128
   rows of A and half of the W tile; the leader CTA issues for both SMs). mode 0: never; 1 (default): for problems of
   at least three waves of tiles, two or more row-blocks and N > 128; 2: whenever there are two row-blocks and N > 128
   (tests). Same chain of k-steps per output element as the one-CTA kernel: results are bit-identical (tests) */
int pcy_set_gemm_pair_mma(int mode);
/* tile width of the one-CTA tcgen05 GEMM: 0 (default) = heuristic over 128x256 / 128x192 / 128x128 by wave efficiency,
   128 / 192 / 256 = that width whenever legal (tests, A/B) */
int pcy_set_gemm_tile(int width);
/* 1 (default): linears with 5..16 activation rows and no fused norm stream the weights through mma.sync (tensor
   cores); 0: the scalar-FMA weight-streaming kernel for every M <= 16 */
int pcy_set_skinny_mma(int enabled);
/* 1 (default): the kernels of the one-launch-per-op decode step (RMSNorm, weight-streaming GEMV, decode attention,
   token embedding) are launched with the programmatic-dependent-launch attribute: each may become resident while its
   predecessor drains and prefetches its first weight stages before `griddepcontrol.wait` (see csrc/common.cuh).
   Results are identical to 0 (plain stream edges); tests compare both. */
int pcy_set_pdl(int enabled);
/* Profiling aid: pcy_esm_profile(1) makes every pcy_esm_encode record CUDA events between its kernels and sync at
   the end; pcy_esm_profile_read returns the milliseconds accumulated since, per kernel class
   [embed, layernorm, qkv, rope, attention, out_proj, fc1, fc2] (n >= 8). pcy_esm_profile(0) turns it off. */
int pcy_esm_profile(int enabled);
int pcy_esm_profile_read(double* ms, int n);
int pcy_esm_create(const pcy_esm_config* cfg, void** handle);
int pcy_esm_destroy(void* handle);
/* src may be a host or a device pointer; the library keeps its own packed copy */
int pcy_esm_load_tensor(void* handle, int kind, int layer, const void* src, int64_t nbytes);
/* cos/sin table fp32 [n_pos][head_dim/2][2] (host or device pointer); n_pos >= the longest T encoded */
int pcy_esm_set_rope_table(void* handle, const float* cos_sin, int n_pos);
int64_t pcy_esm_workspace_bytes(void* handle, int B, int T);
/* tokens int32 [B,T] (device) -> out_states bf16 [B*T, d] (device) */
int pcy_esm_encode(void* handle, const int32_t* tokens, int B, int T, void* out_states, void* workspace,
                   int64_t workspace_bytes, void* stream);
/* ProteinPooler.forward (procyon/model/esm.py:154-217): out[o] = mean|max over the non-pad token rows of the chunk
 * rows seg_rows[seg_ptr[o] : seg_ptr[o+1]] (concatenated in that order). mode 0 = mean (nanmean), 1 = max;
 * correction = protein_pooling_correction_option (drop first and last non-pad row, mean only).
 * states bf16 [n_rows*T, d]; tokens int32 [n_rows, T]; out bf16 or fp32 [n_out, d]. */
int pcy_pool_segments(const void* states, const int32_t* tokens, const int32_t* seg_ptr, const int32_t* seg_rows,
                      void* out, int out_fp32, int T, int d, int n_out, int pad_idx, int mode, int correction,
                      void* stream);

/* ---- Llama decoder ------------------------------------------------------------------------------------------
 * Replaces LlamaPostTokenization.forward -> HF LlamaForCausalLM.forward (procyon/model/pmc_llama.py:546-596),
 * the prefill / decode calls of UnifiedProCyon._generate_beam_search and _generate_sampling
 * (procyon/model/model_unified.py:762-769, :885-887) and their token selection (:782-833, :891-911). */
typedef struct {
  int n_layers, d_model, n_heads, n_kv_heads, head_dim, ffn_dim, vocab;
  float rms_eps;
} pcy_llama_config;

enum {
  PCY_LLAMA_EMBED = 0,   /* bf16 [V,d]            model.embed_tokens.weight */
  PCY_LLAMA_LM_HEAD = 1, /* bf16 [V,d]            lm_head.weight */
  PCY_LLAMA_NORM = 2,    /* bf16 [d]              model.norm.weight */
  PCY_LLAMA_LN1 = 3,     /* bf16 [d]              layers.N.input_layernorm.weight */
  PCY_LLAMA_LN2 = 4,     /* bf16 [d]              layers.N.post_attention_layernorm.weight */
  PCY_LLAMA_WQKV = 5,    /* bf16 [(H+2KVH)hd,d]   cat(q_proj, k_proj, v_proj).weight */
  PCY_LLAMA_WO = 6,      /* bf16 [d,H*hd]         o_proj.weight */
  PCY_LLAMA_WGATEUP = 7, /* bf16 [2F,d]           pcy_pack_gate_up(gate_proj.weight, up_proj.weight) */
  PCY_LLAMA_WDOWN = 8    /* bf16 [d,F]            down_proj.weight */
};

/* 1 (default): prefill attention of prompts with >= 128 positions and head_dim 128 runs on tcgen05 (causal, GQA,
   left-pad key mask; csrc/attention_tc_causal.cu); 0: the mma.sync kernel for every prefill (tests, A/B) */
int pcy_set_llama_tc_attention(int enabled);
int pcy_llama_create(const pcy_llama_config* cfg, void** handle);
int pcy_llama_destroy(void* handle);
int pcy_llama_load_tensor(void* handle, int kind, int layer, const void* src, int64_t nbytes);
/* cos/sin fp32 [n_pos][head_dim/2][2] (host or device); n_pos >= prompt length + generated length */
int pcy_llama_set_rope_table(void* handle, const float* cos_sin, int n_pos);
int64_t pcy_llama_prefill_workspace_bytes(void* handle, int B, int S);
/* Full-sequence forward over B sequences of S positions (left- or right-padded; key_valid uint8 [B,S], 1 = attend,
 * NULL = no padding). input_embeds bf16 [B*S,d]. Optional outputs: kv_prompt bf16 [L][2][B][S][KVH*hd] (the KV
 * cache, one copy per input), hidden_out bf16 [B*S,d] (= hidden_states[-1], after the final RMSNorm),
 * sel_logits fp32 [n_sel,V] = LM-head logits of the flat rows sel_rows[0..n_sel). */
int pcy_llama_prefill(void* handle, const void* input_embeds, const uint8_t* key_valid, int B, int S, void* kv_prompt,
                      void* hidden_out, const int32_t* sel_rows, int n_sel, float* sel_logits, void* workspace,
                      int64_t workspace_bytes, void* stream);
/* same, plus: acc_out fp32 [n_acc, d] = sum over HF's L+1 `hidden_states` (embeddings, every layer's output, the last
 * after the final norm) of the token rows acc_rows[n_acc] — `ret_token_access='all'`
 * (procyon/model/model_unified.py:560-563); needs hidden_out */
int pcy_llama_prefill_ex(void* handle, const void* input_embeds, const uint8_t* key_valid, int B, int S, void* kv_prompt,
                         void* hidden_out, const int32_t* sel_rows, int n_sel, float* sel_logits,
                         const int32_t* acc_rows, int n_acc, float* acc_out, void* workspace, int64_t workspace_bytes,
                         void* stream);

/* Device-resident state of one generate() call. rows = n_inputs * beams (<= 16). All pointers are device memory
 * owned by the caller. state[0] = t (tokens generated so far), state[2] = 1 once every beam of every input holds an
 * EOS (beam mode with stop_on_all_eos), state[3] = the step at which that happened. */
typedef struct {
  int n_inputs, beams, S, max_gen;
  const void* kv_prompt;       /* bf16 [L][2][n_inputs][S][KVH*hd], from pcy_llama_prefill */
  const uint8_t* prompt_valid; /* uint8 [n_inputs][S] or NULL */
  void* kv_gen;                /* bf16 [L][2][rows][max_gen][KVH*hd] */
  int32_t* tokens;             /* [rows][max_gen] generated token ids */
  int32_t* slots;              /* [rows][max_gen] physical KV row of every generated position (beam ancestry) */
  float* logprobs;             /* [rows] running log-probabilities */
  float* logits_cur;           /* fp32 [rows][V] logits of the current step */
  float* logits_hist;          /* fp32 [max_gen][rows][V] per-step logits (physical rows) or NULL */
  int32_t* state;              /* int32 [8] */
  void* workspace;
  int64_t workspace_bytes;
} pcy_decode_buffers;

#define PCY_SELECT_GREEDY 0
#define PCY_SELECT_BEAM 1

/* Decode steps with up to `max_rows` rows (inputs x beams) run as the greedy persistent kernel (one weight row per ring
   slot); default 1 (measured fastest only there), supported up to 4; 0 selects the one-launch-per-op path for EVERY row
   count (it also switches the persistent beam kernel below off); -1 sends every row count, 1 included, to the beam
   kernel (experiment: 3.50 vs 3.07 ms per token at Llama-3-8B size). */
int pcy_set_decode_megakernel(int max_rows);
/* Decode steps with more rows than that (beam search: rows = inputs x beam_size, the reference's evaluation default is
   beam_size 10, procyon/evaluate/framework/procyon.py:71-76) also run as ONE persistent kernel - weight tiles of 16
   rows x 256 k streamed by TMA into a shared-memory ring, mma.sync against up to 16 staged activation rows, stream-K
   work split, attention over the union of the beams' keys (csrc/decode_rows_megakernel.cu).  1 (default) enables it,
   0 selects the one-launch-per-op path (A/B measurements, tests).  Shapes it does not support (head_dim != 128,
   H != 4 KVH, d_model or ffn_dim not a multiple of 256) always take the per-op path. */
int pcy_set_decode_rows_megakernel(int enabled);
/* 1: the greedy persistent kernel runs without its producer warp (12 warps, 168 registers, every warp refills the ring
   slots it owns); bit-identical results, measured equally fast on B200 (3.05 ms per token either way), default 0 */
int pcy_set_decode_self_refill(int enabled);
/* profiling aid for that kernel: device uint64 buffer (zeroed, >= 64 + 40 * n_layers words) that receives %globaltimer
   at every phase boundary of CTA 0 (NULL disables) */
int pcy_set_decode_rows_timing_buffer(void* dev_u64);
/* profiling aid: device uint64 buffer of >= 4096 words (zeroed) that receives %globaltimer at every phase boundary of
 * the persistent decode kernel: 24 stamps per layer + 4, written by CTA 0 (NULL disables).  If word 4095 holds the tag
 * 0x534B4557 the buffer must have 4096 + (4 L + 1) * n_sms * 2 words, and every CTA also records (time, SM id) when it
 * finishes streaming each weight phase (skew analysis, scripts/profile_decode_skew.py). */
int pcy_set_decode_timing_buffer(void* dev_u64);
/* Relative weight-streaming rates of the SMs (host array, one entry per SM id, SM ids must be 0..n-1), measured by the
   caller with the timing buffer's per-CTA stamps: the persistent decode kernel then cuts every weight matrix into
   slices proportional to them (clamped to 0.85..1.15 of the equal slice) instead of equal ones.  Results do not
   change: a row's dot product is the same wherever it is computed.  NULL / 0 restores equal slices. */
int pcy_set_decode_sm_shares(const float* shares, int n);
int64_t pcy_llama_decode_workspace_bytes(void* handle, int rows, int S, int max_gen);
/* clears state/tokens/slots/log-probs/workspace; copies prefill_logits fp32 [n_inputs,V] to every beam row */
int pcy_decode_reset(void* handle, const pcy_decode_buffers* b, const float* prefill_logits, void* stream);
/* one model step on token t-1 of every row (KV-cache append + attention through `slots`) -> logits_cur */
int pcy_llama_decode_forward(void* handle, const pcy_decode_buffers* b, void* stream);
/* token selection from logits_cur: greedy arg-max, or diverse beam search (Hamming penalty across groups) with
 * beam reorder of tokens / log-probs / KV ancestry; appends logits to logits_hist; t += 1 */
int pcy_decode_select(void* handle, const pcy_decode_buffers* b, int mode, int group_size, float diversity_penalty,
                      int eos_id, int stop_on_all_eos, void* stream);

/* Same for a batch split over several decode sessions (<= 16 beam rows each) that step in lock-step on one stream:
 * the reference stops the WHOLE batch at the first step where every beam of every input holds an EOS
 * (procyon/model/model_unified.py:833). group_state: device int32 [4] shared by the sessions, zeroed by the caller
 * except [3] = total number of inputs; [1] = finished, [2] = the step at which that happened (they replace state[2] /
 * state[3]); group_last = 1 for the session that runs last within a step. */
int pcy_decode_select_group(void* handle, const pcy_decode_buffers* b, int mode, int group_size,
                            float diversity_penalty, int eos_id, int stop_on_all_eos, int32_t* group_state,
                            int group_last, void* stream);

/* ---- losses and retrieval scoring -------------------------------------------------------------------------
 * Row-wise cross-entropy of fp32 logits [rows, V] (row stride ld) against int32 labels: acc[0] += sum_i
 * (logsumexp(x_i) - x_i[label_i]), acc[1] += rows. HF LlamaForCausalLM loss (procyon/model/pmc_llama.py:576) is
 * acc[0]/acc[1] over the shifted, non-ignored positions. */
int pcy_cross_entropy_rows(const float* logits, const int32_t* labels, int rows, int V, int64_t ld, float* acc,
                           void* stream);
/* out[q][n] = cos(queries[q], db[n]) with F.normalize semantics (eps 1e-12): get_proteins_from_embedding
 * (procyon/data/inference_utils.py:955-961), ProcyonRetrievalEval cosine sims
 * (procyon/evaluate/framework/procyon.py:400-406). queries fp32 [nq,d]; db fp32 or bf16 [n_db,d]. */
int pcy_cosine_scores(const float* queries, const void* db, int db_is_bf16, float* out, int n_queries, int n_db,
                      int d, int64_t ld_out, void* stream);

/* The same scores plus the ranking get_proteins_from_embedding takes from them (procyon/data/inference_utils.py:964-
 * 970: full argsort, then the first top_k), in ONE launch: the CTA that finishes last selects the k <= 32 best rows of
 * every query from the scores (still in L2).  scores fp32 [n_queries, ld_scores] is written as by pcy_cosine_scores;
 * top_val fp32 / top_idx int32 [n_queries, k] = the k largest scores in descending order (ties: smaller row first)
 * and their row numbers + index_base (the first global row of a database shard, so that per-shard results of a
 * row-sharded database merge by value); missing entries (n_db < k) are (-inf, -1).  workspace: device memory of
 * pcy_retrieval_workspace_bytes() bytes whose first word is zero on entry (the kernel leaves it zero; the rest holds
 * the per-CTA candidate lists).  k = 0 skips the ranking (workspace may be NULL). */
int pcy_retrieval_scores_topk(const float* queries, const void* db, int db_is_bf16, float* scores, int n_queries,
                              int n_db, int d, int64_t ld_scores, int k, int index_base, float* top_val,
                              int32_t* top_idx, int32_t* workspace, void* stream);
int64_t pcy_retrieval_workspace_bytes(void);

/* Merge of per-shard top-k candidates of a row-sharded database (the all-gather of every rank's top_val / top_idx):
 * cand_val fp32 / cand_idx int32 [n_queries, m] (idx < 0 = padding) -> the k <= 32 best per query, same order. */
int pcy_topk_merge(const float* cand_val, const int32_t* cand_idx, int n_queries, int m, int k, float* out_val,
                   int32_t* out_idx, void* stream);

/* F.normalize(x, dim=-1) on fp32 rows */
int pcy_normalize_rows(const float* x, float* out, int rows, int d, void* stream);
/* InfoNCEInBatch.forward (procyon/model/contrastive.py:120-204) on L2-normalised fp32 embeddings:
 * sim_st = zs @ all_t^T / tau, sim_ts = zt @ all_s^T / tau, logits multiplied by mask[rank_off+i][j] (uint8 [G,G] or
 * NULL), targets rank_off + i, loss = (CE_st + CE_ts) / 2 -> loss[0]. sims_scratch: fp32 [2*b*G]. Without gathering,
 * pass all_s = zs, all_t = zt, G = b, rank_off = 0. */
int pcy_infonce_loss(const float* zs, const float* zt, const float* all_s, const float* all_t, const uint8_t* mask,
                     float* sims_scratch, float* loss, int b, int G, int d, int rank_off, float temperature,
                     void* stream);

/* Token embedding + soft-token splice: UnifiedProCyon._prepare_input_embeddings
 * (procyon/model/model_unified.py:1147-1167). out[i] = soft_tokens[soft_index[i]] if soft_index[i] >= 0 else
 * table[ids[i]]; ids/soft_index int32 [n_tok] (soft_index may be NULL), table/soft/out bf16 rows of width d. */
int pcy_embed_splice(const int32_t* ids, const void* table, const void* soft_tokens, const int32_t* soft_index,
                     void* out, int64_t n_tok, int d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROCYON_B200_H_ */
