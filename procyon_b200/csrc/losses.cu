// Loss kernels: row-wise cross-entropy over fp32 logits (LM loss), F.normalize, InfoNCE (retrieval scoring:
// retrieval.cu).  HBM-bound streaming reductions: one pass over the logits / database rows with 16-byte loads.
//
// Replaces: HF LlamaForCausalLM's CrossEntropyLoss over (B*S, V) logits (procyon/model/pmc_llama.py:576),
// get_proteins_from_embedding's normalize + matmul (procyon/data/inference_utils.py:955-961),
// InfoNCEInBatch.forward (procyon/model/contrastive.py:120-204).
#include "common.cuh"
#include "ops.h"

namespace pcy {

namespace {

// one CTA per row: loss = logsumexp(x) - x[label]; acc[0] += loss, acc[1] += 1
__global__ void __launch_bounds__(512)
ce_rows_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, int V, int64_t ld,
               float* __restrict__ acc) {
  __shared__ float s_f[16];
  const int row = blockIdx.x;
  const float* x = logits + (int64_t)row * ld;
  const int lab = labels[row];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) m = fmaxf(m, x[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_f[threadIdx.x >> 5] = m;
  __syncthreads();
  m = s_f[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_f[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(x[i] - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_f[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_f[w];
    const float loss = (m + logf(t)) - x[lab];
    atomicAdd(&acc[0], loss);
    atomicAdd(&acc[1], 1.0f);
  }
}

// out[r] = x[r] / max(|x[r]|, 1e-12)  (F.normalize), one warp per row
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) { const float v = x[(int64_t)r * d + k]; ss += v * v; }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int k = lane; k < d; k += 32) out[(int64_t)r * d + k] = x[(int64_t)r * d + k] * inv;
}

// sims[0][i][j] = <zs[i], all_t[j]> / tau * mask ; sims[1][i][j] = <zt[i], all_s[j]> / tau * mask
// mask row = mask[(rank_off + i) * G + j] (multiplicative 0/1, as the reference: masked logits become 0)
__global__ void __launch_bounds__(256)
infonce_sims_kernel(const float* __restrict__ zs, const float* __restrict__ zt, const float* __restrict__ all_s,
                    const float* __restrict__ all_t, const uint8_t* __restrict__ mask, float* __restrict__ sims,
                    int b, int G, int d, int rank_off, float inv_tau) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = 2 * b * G;
  for (int idx = blockIdx.x * 8 + warp; idx < total; idx += gridDim.x * 8) {
    const int which = idx / (b * G), i = (idx / G) % b, j = idx % G;
    const float* a = (which == 0 ? zs : zt) + (int64_t)i * d;
    const float* c = (which == 0 ? all_t : all_s) + (int64_t)j * d;
    float acc = 0.f;
    for (int k = lane; k < d; k += 32) acc = fmaf(a[k], c[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc * inv_tau;
      if (mask) v *= (float)mask[(int64_t)(rank_off + i) * G + j];
      sims[idx] = v;
    }
  }
}

// loss = 0.5 * (mean_i CE(sims[0][i], rank_off + i) + mean_i CE(sims[1][i], rank_off + i)); single CTA
__global__ void __launch_bounds__(256)
infonce_ce_kernel(const float* __restrict__ sims, float* __restrict__ loss, int b, int G, int rank_off) {
  __shared__ float s_acc[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float part = 0.f;
  for (int r = warp; r < 2 * b; r += 8) {
    const float* x = sims + (int64_t)r * G;
    const int i = r % b;
    float m = -INFINITY;
    for (int j = lane; j < G; j += 32) m = fmaxf(m, x[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < G; j += 32) s += expf(x[j] - m);
    s = warp_sum(s);
    part += (m + logf(s)) - x[rank_off + i];
  }
  if (lane == 0) s_acc[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_acc[w];
    loss[0] = 0.5f * t / (float)b;
  }
}

}  // namespace

int cross_entropy_rows(const float* logits, const int32_t* labels, int rows, int V, int64_t ld, float* acc,
                       cudaStream_t stream) {
  if (rows == 0) return 0;
  ce_rows_kernel<<<rows, 512, 0, stream>>>(logits, labels, V, ld, acc);
  PCY_LAUNCH_CHECK();
  return 0;
}

int normalize_rows(const float* x, float* out, int rows, int d, cudaStream_t stream) {
  if (rows == 0) return 0;
  normalize_rows_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(x, out, rows, d);
  PCY_LAUNCH_CHECK();
  return 0;
}

int infonce_loss(const float* zs, const float* zt, const float* all_s, const float* all_t, const uint8_t* mask,
                 float* sims_scratch, float* loss, int b, int G, int d, int rank_off, float temperature,
                 cudaStream_t stream) {
  PCY_REQUIRE(b > 0 && G >= b && rank_off >= 0 && rank_off + b <= G, "infonce: bad batch geometry b=%d G=%d off=%d", b,
              G, rank_off);
  int grid = ceil_div(2 * b * G, 8);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  infonce_sims_kernel<<<grid, 256, 0, stream>>>(zs, zt, all_s, all_t, mask, sims_scratch, b, G, d, rank_off,
                                               1.0f / temperature);
  PCY_LAUNCH_CHECK();
  infonce_ce_kernel<<<1, 256, 0, stream>>>(sims_scratch, loss, b, G, rank_off);
  PCY_LAUNCH_CHECK();
  return 0;
}

}  // namespace pcy

using namespace pcy;

extern "C" {

int pcy_cross_entropy_rows(const float* logits, const int32_t* labels, int rows, int V, int64_t ld, float* acc,
                           void* stream) {
  return cross_entropy_rows(logits, labels, rows, V, ld, acc, (cudaStream_t)stream);
}

int pcy_normalize_rows(const float* x, float* out, int rows, int d, void* stream) {
  return normalize_rows(x, out, rows, d, (cudaStream_t)stream);
}

int pcy_infonce_loss(const float* zs, const float* zt, const float* all_s, const float* all_t, const uint8_t* mask,
                     float* sims_scratch, float* loss, int b, int G, int d, int rank_off, float temperature,
                     void* stream) {
  return infonce_loss(zs, zt, all_s, all_t, mask, sims_scratch, loss, b, G, d, rank_off, temperature,
                      (cudaStream_t)stream);
}

}  // extern "C"
