// Llama decoder driver: packed weights + prefill and KV-cache decode step as fixed kernel sequences.
//
// prefill  : embeds -> L x [RMSNorm -> fused QKV GEMM (tcgen05) -> RoPE -> causal flash attention (left-pad
//            mask, GQA) -> o_proj GEMM + residual -> RMSNorm -> gate/up GEMM with SwiGLU epilogue -> down GEMM +
//            residual] -> final RMSNorm -> (last-token | selected-row) LM head.  K/V rows are copied once into the
//            prompt cache, stored per INPUT (beams share it; the reference repeats the prompt beam_size times,
//            procyon/model/model_unified.py:751-752).
// decode   : 5 launches per layer, all weight-streaming (HBM-bound) kernels with the RMSNorms fused into the
//            GEMV prologues; no host sync, step index read from device memory so the sequence can be replayed
//            as a CUDA graph.
// Semantics: HF transformers 4.31 LlamaForCausalLM as wrapped by LlamaPostTokenization.forward
// (procyon/model/pmc_llama.py:546-596, :287-406): positions arange(past, past+S), fp32 softmax, RMSNorm in fp32.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace pcy {

struct LlamaLayer {
  bf16 *ln1, *ln2, *wqkv, *wo, *wgu, *wdown;
};

struct LlamaModel {
  pcy_llama_config cfg;
  bf16 *embed = nullptr, *lm_head = nullptr, *norm = nullptr;
  std::vector<LlamaLayer> layers;
  std::vector<void*> slabs;
  float* rope = nullptr;
  int rope_pos = 0;
  LlamaLayerPtrs* layers_dev = nullptr;
  void* rows_maps = nullptr;  // device array of tensor maps for the 3..16-row persistent decode kernel (or null)
  int qkv_dim() const { return (cfg.n_heads + 2 * cfg.n_kv_heads) * cfg.head_dim; }
  int kv_dim() const { return cfg.n_kv_heads * cfg.head_dim; }
};

namespace {

template <typename T>
T* carve(uint8_t*& p, int64_t n) {
  T* r = reinterpret_cast<T*>(p);
  p += round_up(n * (int64_t)sizeof(T), 256);
  return r;
}

int alloc_bf16(LlamaModel* m, bf16** dst, int64_t n) {
  void* p = nullptr;
  PCY_CUDA(cudaMalloc(&p, round_up(n * 2, 256)));
  m->slabs.push_back(p);
  *dst = reinterpret_cast<bf16*>(p);
  return 0;
}

// x[row] = table[tokens[row][t-1]]
__global__ void embed_last_token_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ state,
                                        const bf16* __restrict__ table, bf16* __restrict__ x, int max_gen, int d) {
  const int row = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();  // tokens / state: written by the selection kernels of the previous step
  const int t = state[0];
  const int tok = tokens[(int64_t)row * max_gen + (t - 1)];
  const uint4* src = reinterpret_cast<const uint4*>(table + (int64_t)tok * d);
  uint4* dst = reinterpret_cast<uint4*>(x + (int64_t)row * d);
  for (int i = threadIdx.x; i < d / 8; i += blockDim.x) dst[i] = src[i];
}

// out[i] = in[rows[i]]  (16-byte vectors)
__global__ void gather_rows_kernel(const bf16* __restrict__ in, const int32_t* __restrict__ rows,
                                   bf16* __restrict__ out, int d) {
  const int i = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(in + (int64_t)rows[i] * d);
  uint4* dst = reinterpret_cast<uint4*>(out + (int64_t)i * d);
  for (int k = threadIdx.x; k < d / 8; k += blockDim.x) dst[k] = src[k];
}

__global__ void broadcast_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int beams, int V) {
  // out[(i*beams + b)][:] = in[i][:]
  const int i = blockIdx.y;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < V; k += gridDim.x * blockDim.x) {
    const float v = in[(int64_t)i * V + k];
    for (int b = 0; b < beams; ++b) out[((int64_t)i * beams + b) * V + k] = v;
  }
}

}  // namespace
}  // namespace pcy

using namespace pcy;

// rows (inputs x beams) up to which a decode step runs as the GREEDY persistent kernel (decode_megakernel.cu: one weight
// row per ring slot, dot products on the MMA diagonals).  It supports 4, but measured on B200
// (scripts/bench_decode_rows.py) it wins only at 1 row: 3.07 ms, against 4.58 / 6.1 / 6.5 ms at 2 / 3 / 4 rows, where the
// tile-streaming kernel of decode_rows_megakernel.cu takes 3.51 / 3.64 / 3.8 ms.
static int g_megakernel_max_rows = 1;
// 3..16 rows (beam search): the tile-streaming persistent kernel of decode_rows_megakernel.cu (1), or the
// one-launch-per-op path (0: A/B measurements, tests)
static int g_rows_megakernel = 1;
static bool use_rows_megakernel(const LlamaModel* m, int rows, int beams) {
  static const bool force_multi = getenv("PCY_DECODE_MULTIKERNEL") != nullptr;
  // (pcy_set_decode_megakernel(0) selects the per-op path for EVERY row count, as documented)
  return !force_multi && g_rows_megakernel != 0 && g_megakernel_max_rows != 0 && rows > g_megakernel_max_rows &&
         m->rows_maps != nullptr && decode_rows_megakernel_supported(m->cfg, rows, beams);
}

extern "C" {

int pcy_set_decode_timing_buffer(void* dev_u64) {
  decode_megakernel_set_timing(reinterpret_cast<unsigned long long*>(dev_u64));
  return 0;
}

int pcy_set_decode_sm_shares(const float* shares, int n) { return decode_megakernel_set_shares(shares, n); }

int pcy_set_decode_self_refill(int enabled) {
  decode_megakernel_set_self_refill(enabled);
  return 0;
}

int pcy_set_decode_rows_megakernel(int enabled) {
  g_rows_megakernel = enabled != 0;
  return 0;
}

int pcy_set_decode_rows_timing_buffer(void* dev_u64) {
  decode_rows_megakernel_set_timing(reinterpret_cast<unsigned long long*>(dev_u64));
  return 0;
}

int pcy_set_decode_megakernel(int max_rows) {
  // (-1: experiment - the tile-streaming kernel for every row count, 1 included)
  g_megakernel_max_rows = max_rows < 0 ? -1 : (max_rows > 4 ? 4 : max_rows);
  return 0;
}

int pcy_llama_create(const pcy_llama_config* cfg, void** handle) {
  PCY_REQUIRE(cfg && handle, "llama_create: null argument");
  PCY_REQUIRE(cfg->head_dim == 128, "llama_create: head_dim must be 128 (got %d)", cfg->head_dim);
  PCY_REQUIRE(cfg->n_heads % cfg->n_kv_heads == 0, "llama_create: n_heads %% n_kv_heads != 0");
  PCY_REQUIRE(cfg->d_model % 8 == 0 && cfg->ffn_dim % 16 == 0, "llama_create: d_model %% 8 / ffn_dim %% 16 != 0");
  LlamaModel* m = new LlamaModel();
  m->cfg = *cfg;
  const int64_t d = cfg->d_model, f = cfg->ffn_dim, V = cfg->vocab;
  int rc = 0;
  rc |= alloc_bf16(m, &m->embed, V * d);
  rc |= alloc_bf16(m, &m->lm_head, V * d);
  rc |= alloc_bf16(m, &m->norm, d);
  m->layers.resize(cfg->n_layers);
  for (int l = 0; l < cfg->n_layers && rc == 0; ++l) {
    LlamaLayer& y = m->layers[l];
    rc |= alloc_bf16(m, &y.ln1, d);
    rc |= alloc_bf16(m, &y.ln2, d);
    rc |= alloc_bf16(m, &y.wqkv, (int64_t)m->qkv_dim() * d);
    rc |= alloc_bf16(m, &y.wo, d * cfg->n_heads * cfg->head_dim);
    rc |= alloc_bf16(m, &y.wgu, 2 * f * d);
    rc |= alloc_bf16(m, &y.wdown, d * f);
  }
  if (rc != 0) {
    for (void* p : m->slabs) cudaFree(p);
    delete m;
    return PCY_ERR_CUDA;
  }
  {
    std::vector<LlamaLayerPtrs> host(cfg->n_layers);
    for (int l = 0; l < cfg->n_layers; ++l) {
      const LlamaLayer& y = m->layers[l];
      host[l] = LlamaLayerPtrs{y.ln1, y.ln2, y.wqkv, y.wo, y.wgu, y.wdown};
    }
    void* p = nullptr;
    PCY_CUDA(cudaMalloc(&p, sizeof(LlamaLayerPtrs) * cfg->n_layers));
    m->slabs.push_back(p);
    PCY_CUDA(cudaMemcpy(p, host.data(), sizeof(LlamaLayerPtrs) * cfg->n_layers, cudaMemcpyHostToDevice));
    m->layers_dev = reinterpret_cast<LlamaLayerPtrs*>(p);
    // tensor maps of the weights for the 3..16-row persistent decode kernel (they hold addresses and shapes only);
    // a failure just leaves that kernel unavailable
    if (decode_rows_build_maps(*cfg, host.data(), m->lm_head, &m->rows_maps) == 0) m->slabs.push_back(m->rows_maps);
    else m->rows_maps = nullptr;
  }
  *handle = m;
  return 0;
}

int pcy_llama_destroy(void* handle) {
  if (!handle) return 0;
  LlamaModel* m = reinterpret_cast<LlamaModel*>(handle);
  for (void* p : m->slabs) cudaFree(p);
  if (m->rope) cudaFree(m->rope);
  delete m;
  return 0;
}

int pcy_llama_load_tensor(void* handle, int kind, int layer, const void* src, int64_t nbytes) {
  PCY_REQUIRE(handle && src, "llama_load_tensor: null argument");
  LlamaModel* m = reinterpret_cast<LlamaModel*>(handle);
  const int64_t d = m->cfg.d_model, f = m->cfg.ffn_dim, V = m->cfg.vocab;
  void* dst = nullptr;
  int64_t want = 0;
  if (kind == PCY_LLAMA_EMBED) { dst = m->embed; want = V * d * 2; }
  else if (kind == PCY_LLAMA_LM_HEAD) { dst = m->lm_head; want = V * d * 2; }
  else if (kind == PCY_LLAMA_NORM) { dst = m->norm; want = d * 2; }
  else {
    PCY_REQUIRE(layer >= 0 && layer < m->cfg.n_layers, "llama_load_tensor: layer %d out of range", layer);
    LlamaLayer& y = m->layers[layer];
    switch (kind) {
      case PCY_LLAMA_LN1: dst = y.ln1; want = d * 2; break;
      case PCY_LLAMA_LN2: dst = y.ln2; want = d * 2; break;
      case PCY_LLAMA_WQKV: dst = y.wqkv; want = (int64_t)m->qkv_dim() * d * 2; break;
      case PCY_LLAMA_WO: dst = y.wo; want = d * m->cfg.n_heads * m->cfg.head_dim * 2; break;
      case PCY_LLAMA_WGATEUP: dst = y.wgu; want = 2 * f * d * 2; break;
      case PCY_LLAMA_WDOWN: dst = y.wdown; want = d * f * 2; break;
      default: return set_error(PCY_ERR_INVALID_ARG, "llama_load_tensor: unknown kind %d", kind);
    }
  }
  PCY_REQUIRE(nbytes == want, "llama_load_tensor: kind %d expects %lld bytes, got %lld", kind, (long long)want,
              (long long)nbytes);
  PCY_CUDA(cudaMemcpy(dst, src, nbytes, cudaMemcpyDefault));
  return 0;
}

int pcy_llama_set_rope_table(void* handle, const float* cos_sin, int n_pos) {
  PCY_REQUIRE(handle && cos_sin && n_pos > 0, "llama_set_rope_table: bad argument");
  LlamaModel* m = reinterpret_cast<LlamaModel*>(handle);
  if (m->rope) cudaFree(m->rope);
  m->rope = nullptr;
  const int64_t bytes = (int64_t)n_pos * (m->cfg.head_dim / 2) * 2 * sizeof(float);
  PCY_CUDA(cudaMalloc(&m->rope, bytes));
  PCY_CUDA(cudaMemcpy(m->rope, cos_sin, bytes, cudaMemcpyDefault));
  m->rope_pos = n_pos;
  return 0;
}

int64_t pcy_llama_prefill_workspace_bytes(void* handle, int B, int S) {
  if (!handle) return -1;
  LlamaModel* m = reinterpret_cast<LlamaModel*>(handle);
  const int64_t n = (int64_t)B * S, d = m->cfg.d_model;
  const int64_t wide = std::max<int64_t>(m->qkv_dim(), m->cfg.ffn_dim);
  return 2 * round_up(n * d * 2, 256) + round_up(n * wide * 2, 256) + round_up(n * (int64_t)m->qkv_dim() * 2, 256) +
         4096;
}

// acc[i][:] (+)= x[rows[i]][:] in fp32: the sum over all L+1 hidden states of selected token rows
// (ret_token_access = 'all', procyon/model/model_unified.py:560-563)
__global__ void accumulate_rows_kernel(const bf16* __restrict__ x, const int32_t* __restrict__ rows, float* acc, int d,
                                       int first) {
  const int i = blockIdx.x;
  const bf16* src = x + (int64_t)rows[i] * d;
  float* dst = acc + (int64_t)i * d;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const float v = __bfloat162float(src[k]);
    dst[k] = first ? v : dst[k] + v;
  }
}

int pcy_llama_prefill_ex(void* handle, const void* input_embeds, const uint8_t* key_valid, int B, int S, void* kv_prompt,
                         void* hidden_out, const int32_t* sel_rows, int n_sel, float* sel_logits,
                         const int32_t* acc_rows, int n_acc, float* acc_out, void* workspace, int64_t workspace_bytes,
                         void* stream_);

int pcy_llama_prefill(void* handle, const void* input_embeds, const uint8_t* key_valid, int B, int S, void* kv_prompt,
                      void* hidden_out, const int32_t* sel_rows, int n_sel, float* sel_logits, void* workspace,
                      int64_t workspace_bytes, void* stream_) {
  return pcy_llama_prefill_ex(handle, input_embeds, key_valid, B, S, kv_prompt, hidden_out, sel_rows, n_sel, sel_logits,
                              nullptr, 0, nullptr, workspace, workspace_bytes, stream_);
}

// input_embeds bf16 [B*S, d]; key_valid uint8 [B,S] or NULL; kv_prompt bf16 [L][2][B][S][kv_dim] or NULL;
// hidden_out bf16 [B*S, d] (post final norm) or NULL; sel_rows int32 [n_sel] flat token rows whose logits are
// wanted -> sel_logits fp32 [n_sel, V] (NULL/0 to skip).  acc_rows int32 [n_acc] flat token rows whose L+1 hidden
// states (embeddings, every layer's output, the last one after the final norm - HF's `hidden_states` tuple) are summed
// in fp32 into acc_out [n_acc, d] (NULL/0 to skip).
int pcy_llama_prefill_ex(void* handle, const void* input_embeds, const uint8_t* key_valid, int B, int S, void* kv_prompt,
                         void* hidden_out, const int32_t* sel_rows, int n_sel, float* sel_logits,
                         const int32_t* acc_rows, int n_acc, float* acc_out, void* workspace, int64_t workspace_bytes,
                         void* stream_) {
  if (B == 0 || S == 0) return 0;
  PCY_REQUIRE(handle && input_embeds && workspace, "llama_prefill: null argument");
  LlamaModel* m = reinterpret_cast<LlamaModel*>(handle);
  cudaStream_t stream = (cudaStream_t)stream_;
  const pcy_llama_config& c = m->cfg;
  const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, KVH = c.n_kv_heads, hd = c.head_dim;
  const int qkv_dim = m->qkv_dim(), kvd = m->kv_dim();
  PCY_REQUIRE(m->rope && S <= m->rope_pos, "llama_prefill: rope table has %d positions, need %d", m->rope_pos, S);
  if (workspace_bytes < pcy_llama_prefill_workspace_bytes(handle, B, S))
    return set_error(PCY_ERR_WORKSPACE, "llama_prefill: workspace too small");
  const int64_t n = (int64_t)B * S;
  PCY_REQUIRE(n < (1ll << 31), "llama_prefill: B*S too large");
  PCY_REQUIRE(n_sel == 0 || (sel_rows && sel_logits), "llama_prefill: sel_rows/sel_logits missing");
  uint8_t* p = reinterpret_cast<uint8_t*>(round_up(reinterpret_cast<int64_t>(workspace), 256));
  bf16* x = carve<bf16>(p, n * d);
  bf16* h = carve<bf16>(p, n * d);
  const int64_t wide = std::max<int64_t>(qkv_dim, f);
  bf16* act = carve<bf16>(p, n * wide);
  bf16* qkv = carve<bf16>(p, n * qkv_dim);

  PCY_REQUIRE(n_acc == 0 || (acc_rows && acc_out), "llama_prefill: acc_rows/acc_out missing");
  PCY_CUDA(cudaMemcpyAsync(x, input_embeds, n * d * 2, cudaMemcpyDeviceToDevice, stream));
  for (int l = 0; l < c.n_layers; ++l) {
    const LlamaLayer& y = m->layers[l];
    if (n_acc > 0) {  // hidden_states[l]: the embeddings (l = 0) or the output of layer l - 1
      accumulate_rows_kernel<<<n_acc, 256, 0, stream>>>(x, acc_rows, acc_out, d, l == 0);
      PCY_LAUNCH_CHECK();
    }
    PCY_TRY(rmsnorm_bf16(x, y.ln1, h, n, d, c.rms_eps, stream));
    GemmArgs g;
    g.A = h; g.lda = d; g.W = y.wqkv; g.ldw = d; g.C = qkv; g.ldc = qkv_dim; g.M = (int)n; g.N = qkv_dim; g.K = d;
    const bool fuse_rope = g_fused_rope && n > 16;  // optional: RoPE of the q and k heads in the GEMM epilogue
    if (fuse_rope) { g.rope = m->rope; g.rope_hd = hd; g.rope_T = S; g.rope_ncols = (H + KVH) * hd; }
    PCY_TRY(gemm_bf16(g, stream));
    if (!fuse_rope) PCY_TRY(rope_inplace(qkv, n, S, H + KVH, hd, qkv_dim, 0, m->rope, nullptr, 0, stream));
    if (kv_prompt) {
      bf16* kdst = reinterpret_cast<bf16*>(kv_prompt) + ((int64_t)l * 2 + 0) * n * kvd;
      bf16* vdst = reinterpret_cast<bf16*>(kv_prompt) + ((int64_t)l * 2 + 1) * n * kvd;
      PCY_CUDA(cudaMemcpy2DAsync(kdst, (size_t)kvd * 2, qkv + H * hd, (size_t)qkv_dim * 2, (size_t)kvd * 2, n,
                                 cudaMemcpyDeviceToDevice, stream));
      PCY_CUDA(cudaMemcpy2DAsync(vdst, (size_t)kvd * 2, qkv + H * hd + kvd, (size_t)qkv_dim * 2, (size_t)kvd * 2, n,
                                 cudaMemcpyDeviceToDevice, stream));
    }
    AttnArgs a;
    a.q = qkv; a.k = qkv + H * hd; a.v = qkv + H * hd + kvd; a.o = h;
    a.q_bs = a.k_bs = a.v_bs = (int64_t)S * qkv_dim; a.q_rs = a.k_rs = a.v_rs = qkv_dim;
    a.q_hs = a.k_hs = a.v_hs = hd;
    a.o_bs = (int64_t)S * d; a.o_rs = d; a.o_hs = hd;
    a.B = B; a.H = H; a.KVH = KVH; a.Tq = S; a.Tk = S; a.head_dim = hd;
    a.key_valid = key_valid; a.key_valid_bs = S; a.scale = 1.0f / sqrtf((float)hd); a.causal = 1;
    int tc_done = 0;  // prompts of >= 128 positions, head_dim 128: tcgen05 kernel (attention_tc_causal.cu)
    PCY_TRY(llama_attention_tc(qkv, qkv_dim, key_valid, h, d, B, S, H, KVH, hd, a.scale, &tc_done, stream));
    if (!tc_done) PCY_TRY(flash_attention(a, stream));
    GemmArgs o;
    o.A = h; o.lda = d; o.W = y.wo; o.ldw = H * hd; o.C = x; o.ldc = d; o.M = (int)n; o.N = d; o.K = H * hd;
    o.residual = x; o.ldr = d;
    PCY_TRY(gemm_bf16(o, stream));
    PCY_TRY(rmsnorm_bf16(x, y.ln2, h, n, d, c.rms_eps, stream));
    GemmArgs gu;
    gu.A = h; gu.lda = d; gu.W = y.wgu; gu.ldw = d; gu.C = act; gu.ldc = f; gu.M = (int)n; gu.N = 2 * f; gu.K = d;
    gu.act = ACT_SWIGLU;
    PCY_TRY(gemm_bf16(gu, stream));
    GemmArgs dn;
    dn.A = act; dn.lda = f; dn.W = y.wdown; dn.ldw = f; dn.C = x; dn.ldc = d; dn.M = (int)n; dn.N = d; dn.K = f;
    dn.residual = x; dn.ldr = d;
    PCY_TRY(gemm_bf16(dn, stream));
  }
  if (hidden_out) PCY_TRY(rmsnorm_bf16(x, m->norm, reinterpret_cast<bf16*>(hidden_out), n, d, c.rms_eps, stream));
  if (n_acc > 0) {  // hidden_states[L]: the last layer's output AFTER the final norm
    PCY_REQUIRE(hidden_out != nullptr, "llama_prefill: acc_rows needs hidden_out");
    accumulate_rows_kernel<<<n_acc, 256, 0, stream>>>(reinterpret_cast<const bf16*>(hidden_out), acc_rows, acc_out, d,
                                                      c.n_layers == 0);
    PCY_LAUNCH_CHECK();
  }
  if (n_sel > 0) {
    // gather the selected rows (pre-norm), then LM head with the final RMSNorm fused (or separate for many rows)
    bf16* sel = h;  // h is free here
    gather_rows_kernel<<<n_sel, 128, 0, stream>>>(x, sel_rows, sel, d);
    PCY_LAUNCH_CHECK();
    for (int r0 = 0; r0 < n_sel; r0 += 16) {
      const int mrows = std::min(16, n_sel - r0);
      GemmArgs lm;
      lm.A = sel + (int64_t)r0 * d; lm.lda = d; lm.W = m->lm_head; lm.ldw = d;
      lm.C = sel_logits + (int64_t)r0 * c.vocab; lm.ldc = c.vocab; lm.M = mrows; lm.N = c.vocab; lm.K = d;
      lm.c_fp32 = 1;
      PCY_TRY(gemm_bf16_skinny(lm, m->norm, c.rms_eps, stream));
    }
  }
  return 0;
}

int64_t pcy_llama_decode_workspace_bytes(void* handle, int rows, int S, int max_gen) {
  if (!handle) return -1;
  LlamaModel* m = reinterpret_cast<LlamaModel*>(handle);
  const int64_t d = m->cfg.d_model;
  int64_t b = 3 * round_up(rows * d * 2, 256) + round_up((int64_t)rows * m->qkv_dim() * 2, 256) +
              round_up((int64_t)rows * m->cfg.ffn_dim * 2, 256);
  b += round_up(decode_attention_partial_floats(rows, m->cfg.n_heads, m->cfg.n_kv_heads, S, max_gen) * 4, 256);
  b += round_up((int64_t)rows * m->cfg.n_kv_heads * 4, 256);
  b += round_up((int64_t)topk_workspace_floats(rows) * 4, 256);
  if (decode_megakernel_supported(m->cfg, rows)) b += decode_megakernel_scratch_bytes(m->cfg, rows, S, max_gen) + 256;
  // (the beam count is not known here: size for the worst case over the divisors of rows)
  if (m->rows_maps != nullptr) {
    int64_t extra = 0;
    for (int beams = 1; beams <= rows; ++beams)
      if (rows % beams == 0 && decode_rows_megakernel_supported(m->cfg, rows, beams))
        extra = std::max(extra, decode_rows_megakernel_scratch_bytes(m->cfg, rows, beams, S, max_gen));
    b += extra + 256;
  }
  return b + 4096;
}

// One decode forward: consumes tokens[row][t-1] (t = state[0] >= 1) and writes logits_cur [rows, V].
int pcy_llama_decode_forward(void* handle, const pcy_decode_buffers* b, void* stream_) {
  PCY_REQUIRE(handle && b, "llama_decode_forward: null argument");
  LlamaModel* m = reinterpret_cast<LlamaModel*>(handle);
  cudaStream_t stream = (cudaStream_t)stream_;
  const pcy_llama_config& c = m->cfg;
  const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, KVH = c.n_kv_heads, hd = c.head_dim;
  const int qkv_dim = m->qkv_dim(), kvd = m->kv_dim();
  const int rows = b->n_inputs * b->beams;
  PCY_REQUIRE(rows >= 1 && rows <= 16, "llama_decode_forward: rows=%d must be in [1,16]", rows);
  PCY_REQUIRE(m->rope && b->S + b->max_gen <= m->rope_pos, "llama_decode_forward: rope table too short");
  if (b->workspace_bytes < pcy_llama_decode_workspace_bytes(handle, rows, b->S, b->max_gen))
    return set_error(PCY_ERR_WORKSPACE, "llama_decode_forward: workspace too small");
  uint8_t* p = reinterpret_cast<uint8_t*>(round_up(reinterpret_cast<int64_t>(b->workspace), 256));
  carve<float>(p, topk_workspace_floats(rows));  // selection scratch comes first (see pcy_decode_select)
  bf16* x = carve<bf16>(p, (int64_t)rows * d);
  bf16* attn = carve<bf16>(p, (int64_t)rows * d);
  bf16* qkv = carve<bf16>(p, (int64_t)rows * qkv_dim);
  bf16* act = carve<bf16>(p, (int64_t)rows * f);
  float* partials = carve<float>(p, decode_attention_partial_floats(rows, H, KVH, b->S, b->max_gen));
  int32_t* tickets = carve<int32_t>(p, (int64_t)rows * KVH);  // zeroed by pcy_decode_reset
  bf16* xn = carve<bf16>(p, (int64_t)rows * d);                // normalised rows (3+ rows only)
  static const bool force_multi = getenv("PCY_DECODE_MULTIKERNEL") != nullptr;
  if (!force_multi && rows <= g_megakernel_max_rows && decode_megakernel_supported(c, rows)) {
    // persistent single-launch step: see decode_megakernel.cu
    return decode_megakernel(c, m->layers_dev, m->embed, m->lm_head, m->norm, m->rope, b, p, stream);
  }
  if (use_rows_megakernel(m, rows, b->beams)) {
    if (decode_megakernel_supported(c, rows)) p += round_up(decode_megakernel_scratch_bytes(c, rows, b->S, b->max_gen) + 256, 256);
    return decode_rows_megakernel(c, m->layers_dev, m->rows_maps, m->embed, m->norm, m->rope, b, p, stream);
  }

  PCY_CUDA(launch_pdl(embed_last_token_kernel, dim3(rows), dim3(128), 0, stream, (const int32_t*)b->tokens,
                      (const int32_t*)b->state, (const bf16*)m->embed, x, b->max_gen, d));
  PCY_LAUNCH_CHECK();
  const int64_t n_prompt = (int64_t)b->n_inputs * b->S;
  const int64_t n_gen = (int64_t)rows * b->max_gen;
  // 3..16 rows (beam search): the tensor-core weight-streaming kernel has no fused norm, so RMSNorm runs as its own
  // (tiny) kernel; 1-2 rows keep the fused prologue of the scalar kernel
  const bool split_norm = rows >= skinny_mma_min_rows() && g_skinny_mma;
  auto normed_linear = [&](GemmArgs& g, const bf16* ln) -> int {
    if (!split_norm) return gemm_bf16_skinny(g, ln, c.rms_eps, stream);
    PCY_TRY(rmsnorm_bf16(x, ln, xn, rows, d, c.rms_eps, stream));
    g.A = xn;
    return gemm_bf16_skinny(g, nullptr, 0.f, stream);
  };
  for (int l = 0; l < c.n_layers; ++l) {
    const LlamaLayer& y = m->layers[l];
    GemmArgs g;
    g.A = x; g.lda = d; g.W = y.wqkv; g.ldw = d; g.C = qkv; g.ldc = qkv_dim; g.M = rows; g.N = qkv_dim; g.K = d;
    PCY_TRY(normed_linear(g, y.ln1));
    DecodeAttnArgs a;
    a.qkv = qkv; a.qkv_ld = qkv_dim; a.cos_sin = m->rope;
    a.k_prompt = reinterpret_cast<const bf16*>(b->kv_prompt) + ((int64_t)l * 2 + 0) * n_prompt * kvd;
    a.v_prompt = reinterpret_cast<const bf16*>(b->kv_prompt) + ((int64_t)l * 2 + 1) * n_prompt * kvd;
    a.k_gen = reinterpret_cast<bf16*>(b->kv_gen) + ((int64_t)l * 2 + 0) * n_gen * kvd;
    a.v_gen = reinterpret_cast<bf16*>(b->kv_gen) + ((int64_t)l * 2 + 1) * n_gen * kvd;
    a.slots = b->slots; a.prompt_valid = b->prompt_valid; a.state = b->state; a.partials = partials;
    a.tickets = tickets; a.out = attn; a.rows = rows; a.beams = b->beams; a.H = H; a.KVH = KVH; a.head_dim = hd;
    a.S = b->S; a.max_gen = b->max_gen;
    PCY_TRY(decode_attention(a, stream));
    GemmArgs o;
    o.A = attn; o.lda = d; o.W = y.wo; o.ldw = H * hd; o.C = x; o.ldc = d; o.M = rows; o.N = d; o.K = H * hd;
    o.residual = x; o.ldr = d;
    PCY_TRY(gemm_bf16_skinny(o, nullptr, 0.f, stream));
    GemmArgs gu;
    gu.A = x; gu.lda = d; gu.W = y.wgu; gu.ldw = d; gu.C = act; gu.ldc = f; gu.M = rows; gu.N = 2 * f; gu.K = d;
    gu.act = ACT_SWIGLU;
    PCY_TRY(normed_linear(gu, y.ln2));
    GemmArgs dn;
    dn.A = act; dn.lda = f; dn.W = y.wdown; dn.ldw = f; dn.C = x; dn.ldc = d; dn.M = rows; dn.N = d; dn.K = f;
    dn.residual = x; dn.ldr = d;
    PCY_TRY(gemm_bf16_skinny(dn, nullptr, 0.f, stream));
  }
  GemmArgs lm;
  lm.A = x; lm.lda = d; lm.W = m->lm_head; lm.ldw = d; lm.C = b->logits_cur; lm.ldc = c.vocab; lm.M = rows;
  lm.N = c.vocab; lm.K = d; lm.c_fp32 = 1;
  PCY_TRY(normed_linear(lm, m->norm));
  return 0;
}

// zero the step state / tickets / log-probs and seed logits_cur of every beam row with its input's prefill logits
int pcy_decode_reset(void* handle, const pcy_decode_buffers* b, const float* prefill_logits, void* stream_) {
  PCY_REQUIRE(handle && b, "decode_reset: null argument");
  LlamaModel* m = reinterpret_cast<LlamaModel*>(handle);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int rows = b->n_inputs * b->beams;
  PCY_CUDA(cudaMemsetAsync(b->state, 0, 8 * sizeof(int32_t), stream));
  PCY_CUDA(cudaMemsetAsync(b->logprobs, 0, rows * sizeof(float), stream));
  PCY_CUDA(cudaMemsetAsync(b->tokens, 0, (size_t)rows * b->max_gen * sizeof(int32_t), stream));
  PCY_CUDA(cudaMemsetAsync(b->slots, 0, (size_t)rows * b->max_gen * sizeof(int32_t), stream));
  PCY_CUDA(cudaMemsetAsync(b->workspace, 0, (size_t)b->workspace_bytes, stream));
  if (prefill_logits) {
    dim3 grid(ceil_div(m->cfg.vocab, 256 * 4), b->n_inputs);
    broadcast_rows_kernel<<<grid, 256, 0, stream>>>(prefill_logits, b->logits_cur, b->beams, m->cfg.vocab);
    PCY_LAUNCH_CHECK();
  }
  return 0;
}

// token selection for the step whose logits are in logits_cur; advances state[0]
int pcy_decode_select_group(void* handle, const pcy_decode_buffers* b, int mode, int group_size, float diversity_penalty,
                            int eos_id, int stop_on_all_eos, int32_t* group_state, int group_last, void* stream_) {
  PCY_REQUIRE(handle && b, "decode_select: null argument");
  LlamaModel* m = reinterpret_cast<LlamaModel*>(handle);
  const int rows = b->n_inputs * b->beams;
  // the top-k scratch sits after the forward-pass buffers in the shared workspace
  uint8_t* p = reinterpret_cast<uint8_t*>(round_up(reinterpret_cast<int64_t>(b->workspace), 256));
  float* tk = carve<float>(p, topk_workspace_floats(rows));
  DecodeSelectArgs a;
  a.logits = b->logits_cur; a.logits_hist = b->logits_hist; a.tokens = b->tokens; a.slots = b->slots;
  a.logprobs = b->logprobs; a.state = b->state; a.workspace = tk; a.n_inputs = b->n_inputs; a.beams = b->beams;
  a.group = group_size; a.max_gen = b->max_gen; a.vocab = m->cfg.vocab; a.eos_id = eos_id;
  a.diversity_penalty = diversity_penalty; a.greedy = (mode == PCY_SELECT_GREEDY); a.stop_on_all_eos = stop_on_all_eos;
  a.group_state = group_state; a.group_last = group_last;
  return decode_select(a, (cudaStream_t)stream_);
}

int pcy_decode_select(void* handle, const pcy_decode_buffers* b, int mode, int group_size, float diversity_penalty,
                      int eos_id, int stop_on_all_eos, void* stream_) {
  return pcy_decode_select_group(handle, b, mode, group_size, diversity_penalty, eos_id, stop_on_all_eos, nullptr, 1,
                                 stream_);
}

}  // extern "C"
