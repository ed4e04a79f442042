// Loss / scoring kernels: row-wise cross-entropy over fp32 logits (LM loss), cosine retrieval scores,
// InfoNCE.  HBM-bound streaming reductions: one pass over the logits / database rows with 16-byte loads.
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

// Cosine scores: out[q][n] = <Q[q], D[n]> / (max(|Q[q]|, eps) * max(|D[n]|, eps)), eps = 1e-12 (F.normalize).
// One warp per database row, QT queries at a time held in shared memory; fp32 database rows are read once.
template <int QT, typename DT>
__global__ void __launch_bounds__(256)
cosine_scores_kernel(const float* __restrict__ Q, const DT* __restrict__ D, float* __restrict__ out, int nq, int N,
                     int d, int64_t ldo) {
  extern __shared__ float s_q[];  // [QT][d] then [QT] inverse norms
  float* s_inv = s_q + (size_t)QT * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < QT * d; i += blockDim.x) {
    const int q = i / d;
    s_q[i] = (q < nq) ? Q[(int64_t)q * d + (i % d)] : 0.f;
  }
  __syncthreads();
  for (int q = warp; q < QT; q += 8) {
    float ss = 0.f;
    for (int k = lane; k < d; k += 32) ss += s_q[q * d + k] * s_q[q * d + k];
    ss = warp_sum(ss);
    if (lane == 0) s_inv[q] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  }
  __syncthreads();
  for (int n = blockIdx.x * 8 + warp; n < N; n += gridDim.x * 8) {
    const DT* row = D + (int64_t)n * d;
    float dot[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) dot[q] = 0.f;
    float nn = 0.f;
    for (int k = lane * 4; k < d; k += 128) {
      float r[4];
      if (sizeof(DT) == 4) {
        const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + k);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(row) + k);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
      }
      nn += r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        const float4 qq = *reinterpret_cast<const float4*>(s_q + q * d + k);
        dot[q] += r[0] * qq.x + r[1] * qq.y + r[2] * qq.z + r[3] * qq.w;
      }
    }
    nn = warp_sum(nn);
    const float inv_n = 1.0f / fmaxf(sqrtf(nn), 1e-12f);
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      const float v = warp_sum(dot[q]);
      if (lane == 0 && q < nq) out[(int64_t)q * ldo + n] = v * s_inv[q] * inv_n;
    }
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

int cosine_scores(const float* Q, const void* D, int db_bf16, float* out, int nq, int N, int d, int64_t ldo,
                  cudaStream_t stream) {
  PCY_REQUIRE(d % 4 == 0, "cosine_scores: d %% 4 != 0");
  if (nq == 0 || N == 0) return 0;
  int grid = ceil_div(N, 8);
  const int max_grid = num_sms() * 8;
  if (grid > max_grid) grid = max_grid;
  for (int q0 = 0; q0 < nq; q0 += 4) {
    const int cnt = std::min(4, nq - q0);
    const size_t smem = ((size_t)4 * d + 4) * sizeof(float);
    PCY_REQUIRE(smem <= 48 * 1024, "cosine_scores: d=%d too large", d);
    if (db_bf16)
      cosine_scores_kernel<4, bf16><<<grid, 256, smem, stream>>>(Q + (int64_t)q0 * d, (const bf16*)D,
                                                                out + (int64_t)q0 * ldo, cnt, N, d, ldo);
    else
      cosine_scores_kernel<4, float><<<grid, 256, smem, stream>>>(Q + (int64_t)q0 * d, (const float*)D,
                                                                 out + (int64_t)q0 * ldo, cnt, N, d, ldo);
    PCY_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace pcy

using namespace pcy;

extern "C" {

int pcy_cross_entropy_rows(const float* logits, const int32_t* labels, int rows, int V, int64_t ld, float* acc,
                           void* stream) {
  return cross_entropy_rows(logits, labels, rows, V, ld, acc, (cudaStream_t)stream);
}

int pcy_cosine_scores(const float* queries, const void* db, int db_is_bf16, float* out, int n_queries, int n_db,
                      int d, int64_t ld_out, void* stream) {
  return cosine_scores(queries, db, db_is_bf16, out, n_queries, n_db, d, ld_out, (cudaStream_t)stream);
}

}  // extern "C"
