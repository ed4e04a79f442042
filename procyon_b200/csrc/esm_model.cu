// ESM2 encoder driver: owns an immutable packed copy of the weights and runs the layer stack
//   embed -> L x [ x += out_proj(MHA(rope(qkv(LN(x))))) ; x += fc2(gelu(fc1(LN(x)))) ] -> final LN
// as a fixed sequence of kernels on one stream (graph-capturable: no host sync, no allocation).
// Semantics follow fair-esm 2.0.0 `ESM2.forward` as called from procyon/model/esm.py:526,536
// (representations[repr_layer] = output of emb_layer_norm_after); the LM head is skipped because the pooled
// path discards logits (procyon/model/esm.py:547, model_unified.py:391).
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace pcy {

struct EsmLayer {
  bf16 *ln1_g, *ln1_b, *wqkv, *wo, *ln2_g, *ln2_b, *w1, *w2;
  float *bqkv, *bo, *b1, *b2;
};

struct EsmModel {
  pcy_esm_config cfg;
  bf16* embed = nullptr;
  bf16 *lnf_g = nullptr, *lnf_b = nullptr;
  std::vector<EsmLayer> layers;
  float* rope = nullptr;  // [rope_pos][head_dim/2][2]
  int rope_pos = 0;
  void* slab = nullptr;
};

namespace {
// In-situ time per kernel class of the encoder (pcy_esm_profile): CUDA events between the ops of a real encode, so
// the shares are taken at the clocks the step actually runs at (ncu serialises launches and measures them cold).
enum { PC_EMBED = 0, PC_LN, PC_QKV, PC_ROPE, PC_ATTN, PC_OUT, PC_FC1, PC_FC2, PC_N };
struct EsmProfile {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  std::vector<int> cls;  // class of the interval that ends at event i + 1
  size_t used = 0;
  double ms[PC_N] = {};
  void mark(int c, cudaStream_t st) {
    if (!on) return;
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    cudaEventRecord(pool[used++], st);
    if (c >= 0) cls.push_back(c);
  }
  void collect() {
    if (!on || used < 2) return;
    cudaEventSynchronize(pool[used - 1]);
    for (size_t i = 0; i + 1 < used; ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, pool[i], pool[i + 1]);
      ms[cls[i]] += t;
    }
    used = 0;
    cls.clear();
  }
};
EsmProfile g_prof;

template <typename T>
T* carve(uint8_t*& p, int64_t n) {
  T* r = reinterpret_cast<T*>(p);
  p += round_up(n * (int64_t)sizeof(T), 256);
  return r;
}
}  // namespace

}  // namespace pcy

using namespace pcy;

static bool g_esm_tc_attention = true;

extern "C" {

int pcy_set_esm_tc_attention(int enabled) {
  g_esm_tc_attention = enabled != 0;
  return 0;
}

int pcy_set_esm_attention_kernel(int kernel) {
  PCY_REQUIRE(kernel >= 0 && kernel <= 7, "set_esm_attention_kernel: %d not in {0, ..., 7}", kernel);
  pcy::g_esm_attention_kernel = kernel;
  return 0;
}

int pcy_set_esm_attention_tail_rows(int rows) {
  PCY_REQUIRE(rows >= 0 && rows < 128, "set_esm_attention_tail_rows: %d not in [0, 128)", rows);
  pcy::g_esm_attention_tail_rows = rows;
  return 0;
}

int pcy_set_esm_attention_q_rope(int enabled) {
  pcy::g_esm_attention_q_rope = enabled != 0;
  return 0;
}

int pcy_set_fused_rope(int enabled) {
  g_fused_rope = enabled != 0;
  return 0;
}

int pcy_set_gemm_tile(int width) {
  pcy::g_gemm_force_tile = (width == 128 || width == 192 || width == 256) ? width : 0;
  return 0;
}
int pcy_set_pdl(int enabled) {
  pcy::g_pdl = enabled != 0;
  return 0;
}
int pcy_set_skinny_mma(int enabled) {
  pcy::g_skinny_mma = enabled != 0;
  return 0;
}

int pcy_set_gemm_cluster(int enabled) {
  pcy::g_gemm_cluster = enabled != 0;
  return 0;
}

int pcy_set_gemm_pair_mma(int mode) {
  PCY_REQUIRE(mode >= 0 && mode <= 2, "set_gemm_pair_mma: mode %d not in {0, 1, 2}", mode);
  pcy::g_gemm_pair_mma = mode;
  return 0;
}

int pcy_esm_profile(int enabled) {
  g_prof.collect();
  g_prof.on = enabled != 0;
  for (double& v : g_prof.ms) v = 0.0;
  return 0;
}

// ms accumulated since pcy_esm_profile(1): [embed, layernorm, qkv, rope, attention, out_proj, fc1, fc2]
int pcy_esm_profile_read(double* ms, int n) {
  PCY_REQUIRE(ms && n >= PC_N, "esm_profile_read: need room for %d values", (int)PC_N);
  for (int i = 0; i < PC_N; ++i) ms[i] = g_prof.ms[i];
  return 0;
}

int pcy_esm_create(const pcy_esm_config* cfg, void** handle) {
  PCY_REQUIRE(cfg && handle, "esm_create: null argument");
  PCY_REQUIRE(cfg->d_model % cfg->n_heads == 0, "esm_create: d_model %% n_heads != 0");
  const int d = cfg->d_model, f = cfg->ffn_dim, hd = d / cfg->n_heads;
  PCY_REQUIRE(d % 8 == 0 && f % 8 == 0 && hd % 8 == 0 && hd <= 128,
              "esm_create: unsupported shape d=%d ffn=%d head_dim=%d", d, f, hd);
  EsmModel* m = new EsmModel();
  m->cfg = *cfg;
  const int L = cfg->n_layers;
  auto pad = [](int64_t n, int64_t sz) { return round_up(n * sz, 256); };
  int64_t bytes = pad((int64_t)cfg->vocab * d, 2) + 2 * pad(d, 2);
  const int64_t per_layer = 4 * pad(d, 2) + pad((int64_t)3 * d * d, 2) + pad((int64_t)d * d, 2) +
                            2 * pad((int64_t)f * d, 2) + pad(3 * d, 4) + 2 * pad(d, 4) + pad(f, 4);
  bytes += per_layer * L;
  cudaError_t e = cudaMalloc(&m->slab, bytes);
  if (e != cudaSuccess) {
    delete m;
    return cuda_error(e, "cudaMalloc(esm weights)", __FILE__, __LINE__);
  }
  cudaMemset(m->slab, 0, bytes);
  uint8_t* p = reinterpret_cast<uint8_t*>(m->slab);
  m->embed = carve<bf16>(p, (int64_t)cfg->vocab * d);
  m->lnf_g = carve<bf16>(p, d);
  m->lnf_b = carve<bf16>(p, d);
  m->layers.resize(L);
  for (int l = 0; l < L; ++l) {
    EsmLayer& y = m->layers[l];
    y.ln1_g = carve<bf16>(p, d); y.ln1_b = carve<bf16>(p, d);
    y.ln2_g = carve<bf16>(p, d); y.ln2_b = carve<bf16>(p, d);
    y.wqkv = carve<bf16>(p, (int64_t)3 * d * d);
    y.wo = carve<bf16>(p, (int64_t)d * d);
    y.w1 = carve<bf16>(p, (int64_t)f * d);
    y.w2 = carve<bf16>(p, (int64_t)f * d);
    y.bqkv = carve<float>(p, 3 * d);
    y.bo = carve<float>(p, d);
    y.b2 = carve<float>(p, d);
    y.b1 = carve<float>(p, f);
  }
  *handle = m;
  return 0;
}

int pcy_esm_destroy(void* handle) {
  if (!handle) return 0;
  EsmModel* m = reinterpret_cast<EsmModel*>(handle);
  if (m->slab) cudaFree(m->slab);
  if (m->rope) cudaFree(m->rope);
  delete m;
  return 0;
}

int pcy_esm_load_tensor(void* handle, int kind, int layer, const void* src, int64_t nbytes) {
  PCY_REQUIRE(handle && src, "esm_load_tensor: null argument");
  EsmModel* m = reinterpret_cast<EsmModel*>(handle);
  const int d = m->cfg.d_model, f = m->cfg.ffn_dim;
  void* dst = nullptr;
  int64_t want = 0;
  if (kind == PCY_ESM_EMBED) { dst = m->embed; want = (int64_t)m->cfg.vocab * d * 2; }
  else if (kind == PCY_ESM_LNF_G) { dst = m->lnf_g; want = d * 2; }
  else if (kind == PCY_ESM_LNF_B) { dst = m->lnf_b; want = d * 2; }
  else {
    PCY_REQUIRE(layer >= 0 && layer < m->cfg.n_layers, "esm_load_tensor: layer %d out of range", layer);
    EsmLayer& y = m->layers[layer];
    switch (kind) {
      case PCY_ESM_LN1_G: dst = y.ln1_g; want = d * 2; break;
      case PCY_ESM_LN1_B: dst = y.ln1_b; want = d * 2; break;
      case PCY_ESM_LN2_G: dst = y.ln2_g; want = d * 2; break;
      case PCY_ESM_LN2_B: dst = y.ln2_b; want = d * 2; break;
      case PCY_ESM_WQKV: dst = y.wqkv; want = (int64_t)3 * d * d * 2; break;
      case PCY_ESM_BQKV: dst = y.bqkv; want = (int64_t)3 * d * 4; break;
      case PCY_ESM_WO: dst = y.wo; want = (int64_t)d * d * 2; break;
      case PCY_ESM_BO: dst = y.bo; want = (int64_t)d * 4; break;
      case PCY_ESM_W1: dst = y.w1; want = (int64_t)f * d * 2; break;
      case PCY_ESM_B1: dst = y.b1; want = (int64_t)f * 4; break;
      case PCY_ESM_W2: dst = y.w2; want = (int64_t)f * d * 2; break;
      case PCY_ESM_B2: dst = y.b2; want = (int64_t)d * 4; break;
      default: return set_error(PCY_ERR_INVALID_ARG, "esm_load_tensor: unknown kind %d", kind);
    }
  }
  PCY_REQUIRE(nbytes == want, "esm_load_tensor: kind %d expects %lld bytes, got %lld", kind, (long long)want,
              (long long)nbytes);
  PCY_CUDA(cudaMemcpy(dst, src, nbytes, cudaMemcpyDefault));
  return 0;
}

int pcy_esm_set_rope_table(void* handle, const float* cos_sin, int n_pos) {
  PCY_REQUIRE(handle && cos_sin && n_pos > 0, "esm_set_rope_table: bad argument");
  EsmModel* m = reinterpret_cast<EsmModel*>(handle);
  const int hd = m->cfg.d_model / m->cfg.n_heads;
  if (m->rope) cudaFree(m->rope);
  m->rope = nullptr;
  const int64_t bytes = (int64_t)n_pos * (hd / 2) * 2 * sizeof(float);
  PCY_CUDA(cudaMalloc(&m->rope, bytes));
  PCY_CUDA(cudaMemcpy(m->rope, cos_sin, bytes, cudaMemcpyDefault));
  m->rope_pos = n_pos;
  return 0;
}

int64_t pcy_esm_workspace_bytes(void* handle, int B, int T) {
  if (!handle) return -1;
  EsmModel* m = reinterpret_cast<EsmModel*>(handle);
  const int64_t n = (int64_t)B * T, d = m->cfg.d_model;
  const int64_t wide = std::max<int64_t>(3 * d, m->cfg.ffn_dim);
  return round_up(n * d * 2, 256) * 2 + round_up(n * wide * 2, 256) + round_up(n, 256) +
         round_up((int64_t)B * esm_key_valid_words(T) * 4, 256) + 1024;
}

int pcy_esm_encode(void* handle, const int32_t* tokens, int B, int T, void* out_states, void* workspace,
                   int64_t workspace_bytes, void* stream_) {
  if (B == 0 || T == 0) return 0;
  PCY_REQUIRE(handle && tokens && out_states && workspace, "esm_encode: null argument");
  EsmModel* m = reinterpret_cast<EsmModel*>(handle);
  cudaStream_t stream = (cudaStream_t)stream_;
  const pcy_esm_config& c = m->cfg;
  const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, hd = d / H;
  if (B == 0 || T == 0) return 0;
  PCY_REQUIRE(m->rope && T <= m->rope_pos, "esm_encode: rope table has %d positions, need %d", m->rope_pos, T);
  if (workspace_bytes < pcy_esm_workspace_bytes(handle, B, T))
    return set_error(PCY_ERR_WORKSPACE, "esm_encode: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                     (long long)pcy_esm_workspace_bytes(handle, B, T));
  const int64_t n = (int64_t)B * T;
  PCY_REQUIRE(n < (1ll << 31), "esm_encode: B*T too large");
  uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
  p = reinterpret_cast<uint8_t*>(round_up(reinterpret_cast<int64_t>(p), 256));
  bf16* x = carve<bf16>(p, n * d);
  bf16* h = carve<bf16>(p, n * d);
  const int64_t wide = std::max<int64_t>(3 * d, f);
  bf16* big = carve<bf16>(p, n * wide);
  uint8_t* valid = carve<uint8_t>(p, n);
  uint32_t* valid_words = carve<uint32_t>(p, (int64_t)B * esm_key_valid_words(T));

  g_prof.mark(-1, stream);
  PCY_TRY(esm_embed(tokens, m->embed, x, B, T, d, c.pad_idx, c.mask_idx, c.token_dropout, stream));
  PCY_TRY(make_key_valid(tokens, valid, n, c.pad_idx, stream));
  PCY_TRY(esm_pack_key_valid(valid, valid_words, B, T, stream));  // the same as bits, for the attention prologue
  g_prof.mark(PC_EMBED, stream);
  const float q_scale = 1.0f / sqrtf((float)hd);

  for (int l = 0; l < c.n_layers; ++l) {
    const EsmLayer& y = m->layers[l];
    // --- self attention block ---
    PCY_TRY(layernorm_bf16(x, y.ln1_g, y.ln1_b, h, n, d, c.ln_eps, stream));
    g_prof.mark(PC_LN, stream);
    GemmArgs g;
    g.A = h; g.lda = d; g.W = y.wqkv; g.ldw = d; g.C = big; g.ldc = 3 * d; g.M = (int)n; g.N = 3 * d; g.K = d;
    g.bias = y.bqkv; g.scale = q_scale; g.scale_ncols = d;  // q = (x Wq + bq) * head_dim^-0.5
    const bool fuse_rope = g_fused_rope && (hd == 64 || hd == 128) && n > 16;
    if (fuse_rope) { g.rope = m->rope; g.rope_hd = hd; g.rope_T = T; g.rope_ncols = 2 * d; }  // q and k heads
    PCY_TRY(gemm_bf16(g, stream));
    g_prof.mark(PC_QKV, stream);
    // the TMEM-operand attention kernels rotate Q while they move it into TMEM: the RoPE pass then only covers K
    const bool q_in_attn = !fuse_rope && g_esm_tc_attention && esm_attention_tc_ropes_q(T, H, d);
    if (q_in_attn) PCY_TRY(rope_inplace(big, n, T, H, hd, 3 * d, d, m->rope, nullptr, 0, stream));
    else if (!fuse_rope) PCY_TRY(rope_inplace(big, n, T, 2 * H, hd, 3 * d, 0, m->rope, nullptr, 0, stream));
    g_prof.mark(PC_ROPE, stream);
    int rows_done = 0;
    if (g_esm_tc_attention)
      PCY_TRY(esm_attention_tc(big, valid, valid_words, h, B, T, H, d, 1.0f, q_in_attn ? m->rope : nullptr, &rows_done, stream));
    AttnArgs a;
    a.q = big + (int64_t)rows_done * 3 * d; a.k = big + d; a.v = big + 2 * d; a.o = h + (int64_t)rows_done * d;
    a.q_bs = a.k_bs = a.v_bs = (int64_t)T * 3 * d; a.q_rs = a.k_rs = a.v_rs = 3 * d;
    a.q_hs = a.k_hs = a.v_hs = hd;
    a.o_bs = (int64_t)T * d; a.o_rs = d; a.o_hs = hd;
    a.B = B; a.H = H; a.KVH = H; a.Tq = T - rows_done; a.Tk = T; a.head_dim = hd;
    a.key_valid = valid; a.key_valid_bs = T; a.scale = 1.0f; a.causal = 0;
    if (rows_done < T) PCY_TRY(flash_attention(a, stream));  // ragged tail rows (or everything when hd != 64)
    g_prof.mark(PC_ATTN, stream);
    GemmArgs o;
    o.A = h; o.lda = d; o.W = y.wo; o.ldw = d; o.C = x; o.ldc = d; o.M = (int)n; o.N = d; o.K = d;
    o.bias = y.bo; o.residual = x; o.ldr = d;
    PCY_TRY(gemm_bf16(o, stream));
    g_prof.mark(PC_OUT, stream);
    // --- feed forward block ---
    PCY_TRY(layernorm_bf16(x, y.ln2_g, y.ln2_b, h, n, d, c.ln_eps, stream));
    g_prof.mark(PC_LN, stream);
    GemmArgs f1;
    f1.A = h; f1.lda = d; f1.W = y.w1; f1.ldw = d; f1.C = big; f1.ldc = f; f1.M = (int)n; f1.N = f; f1.K = d;
    f1.bias = y.b1; f1.act = ACT_GELU;
    PCY_TRY(gemm_bf16(f1, stream));
    g_prof.mark(PC_FC1, stream);
    GemmArgs f2;
    f2.A = big; f2.lda = f; f2.W = y.w2; f2.ldw = f; f2.C = x; f2.ldc = d; f2.M = (int)n; f2.N = d; f2.K = f;
    f2.bias = y.b2; f2.residual = x; f2.ldr = d;
    PCY_TRY(gemm_bf16(f2, stream));
    g_prof.mark(PC_FC2, stream);
  }
  PCY_TRY(layernorm_bf16(x, m->lnf_g, m->lnf_b, reinterpret_cast<bf16*>(out_states), n, d, c.ln_eps, stream));
  g_prof.mark(PC_LN, stream);
  g_prof.collect();
  return 0;
}

int pcy_pool_segments(const void* states, const int32_t* tokens, const int32_t* seg_ptr, const int32_t* seg_rows,
                      void* out, int out_fp32, int T, int d, int n_out, int pad_idx, int mode, int correction,
                      void* stream) {
  return pool_segments(reinterpret_cast<const bf16*>(states), tokens, seg_ptr, seg_rows, out, out_fp32, T, d, n_out,
                       pad_idx, mode, correction, (cudaStream_t)stream);
}

}  // extern "C"
