// Retrieval scoring over a protein-embedding database: cosine similarity of a few queries against N database rows and
// the top-k hits, in ONE launch.
//
// Replaces get_proteins_from_embedding (procyon/data/inference_utils.py:955-976: F.normalize of the whole database on
// the host, upload, matmul, full argsort, slice) and ProcyonRetrievalEval's cosine scores
// (procyon/evaluate/framework/procyon.py:400-406).  HBM-bound: the database (N*d*4 bytes, 102 MB at N = 20 000,
// d = 1280) is read exactly once, row norms and dot products come out of the same pass; the query lives in registers.
// The top-k costs no second pass: every warp keeps a sorted list of its 32 best (score, row) pairs spread over its
// lanes (one insertion per row that beats the list's last entry); lists meet by bitonic merges in a tree: the 8 warps
// of a CTA, then the 16 CTAs of a group, then the 19 groups - at each level the CTA that arrives last (ticket) merges.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace pcy {

namespace {

constexpr int RT_THREADS = 256;  // 8 warps, two CTAs per SM; every warp owns ROWS database rows per step
constexpr int RT_WARPS = RT_THREADS / 32;
constexpr int RT_MAXK = 32;
constexpr int RT_GROUP = 16;  // CTAs per group of the ranking tree

// order of the ranking: larger score first; equal scores: smaller row index first
__device__ __forceinline__ bool ranks_before(float v, int i, float w, int j) { return v > w || (v == w && i < j); }

// the warp's running top-32, lane l holding the l-th best: insert (v, idx) (same values in all lanes)
__device__ __forceinline__ void warp_topk_insert(float& tv, int& ti, float v, int idx, int lane) {
  const unsigned ahead = __ballot_sync(0xffffffffu, ranks_before(tv, ti, v, idx));
  const int pos = __popc(ahead);  // the list is sorted: exactly the lanes [0, pos) rank before the newcomer
  const float uv = __shfl_up_sync(0xffffffffu, tv, 1);
  const int ui = __shfl_up_sync(0xffffffffu, ti, 1);
  if (lane == pos) { tv = v; ti = idx; }
  else if (lane > pos) { tv = uv; ti = ui; }
}

// one candidate per lane (s, idx): merge those that beat the current 32nd into the list
__device__ __forceinline__ void warp_topk_offer(float& tv, int& ti, float s, int idx, int lane) {
  const float kv = __shfl_sync(0xffffffffu, tv, 31);
  const int ki = __shfl_sync(0xffffffffu, ti, 31);
  unsigned cand = __ballot_sync(0xffffffffu, ranks_before(s, idx, kv, ki));
  while (cand) {
    const int src = __ffs(cand) - 1;
    cand &= cand - 1;
    warp_topk_insert(tv, ti, __shfl_sync(0xffffffffu, s, src), __shfl_sync(0xffffffffu, idx, src), lane);
  }
}

// top-32 of the union of two sorted lists (each spread over the lanes, best first): reverse one, take the better of
// each pair -> a bitonic sequence holding exactly the 32 best -> 5 compare-exchange stages sort it.  ~40 instructions.
__device__ __forceinline__ void warp_topk_merge(float& tv, int& ti, float ov, int oi, int lane) {
  const float rv = __shfl_sync(0xffffffffu, ov, 31 - lane);
  const int ri = __shfl_sync(0xffffffffu, oi, 31 - lane);
  if (ranks_before(rv, ri, tv, ti)) { tv = rv; ti = ri; }
#pragma unroll
  for (int stride = 16; stride > 0; stride >>= 1) {
    const float pv = __shfl_xor_sync(0xffffffffu, tv, stride);
    const int pi = __shfl_xor_sync(0xffffffffu, ti, stride);
    const bool keep_better = (lane & stride) == 0;
    const bool partner_better = ranks_before(pv, pi, tv, ti);
    if (partner_better == keep_better) { tv = pv; ti = pi; }
  }
}

template <typename DT>
__device__ __forceinline__ void load4(const DT* row, int k, float (&r)[4]) {
  if constexpr (sizeof(DT) == 4) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + k));  // streamed once
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  } else {
    const uint2 u = __ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(row) + k));
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
  }
}

// scores[q][n] = <Q[q], D[n]> / (max(|Q[q]|, eps) * max(|D[n]|, eps)), eps = 1e-12 (F.normalize), q < QT queries.
// NV > 0: d == NV * 128 and a warp's ROWS rows of a step sit in registers as ROWS * NV independent 16-byte loads per
// lane, requested one step ahead (software pipeline); NV == 0: any d % 4 == 0, plain loop.  The queries are read from
// shared memory (one 16-byte LDS per 16-byte global load: 1/5 of the shared-memory bandwidth at the HBM rate).
// k > 0: the last CTA to finish ranks all N scores of every query and writes the top k (value, row + index_base).
template <int QT, int NV, int ROWS, typename DT>
__global__ void __launch_bounds__(RT_THREADS, 2)
retrieval_scores_topk_kernel(const float* __restrict__ Q, const DT* __restrict__ D, float* __restrict__ scores, int nq,
                             int N, int d, int64_t lds, int k, int index_base, float* __restrict__ top_val,
                             int32_t* __restrict__ top_idx, int* __restrict__ ws) {
  extern __shared__ float s_q[];  // [QT][d] queries, then [QT] inverse norms
  float* s_inv = s_q + (size_t)QT * d;
  __shared__ float s_tv[RT_WARPS][RT_MAXK];
  __shared__ int s_ti[RT_WARPS][RT_MAXK];
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_step = gridDim.x * RT_WARPS * ROWS;
  int n0 = (blockIdx.x * RT_WARPS + warp) * ROWS;

  // the first rows are requested before the query is staged: their latency covers the prologue
  constexpr int NVR = NV > 0 ? NV : 1;
  float pre[ROWS][NVR][4];
  if constexpr (NV > 0) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (n0 + r < N) load4<DT>(D + (int64_t)(n0 + r) * d, j * 128 + lane * 4, pre[r][j]);
        else pre[r][j][0] = pre[r][j][1] = pre[r][j][2] = pre[r][j][3] = 0.f;
      }
  }
  for (int i = threadIdx.x; i < QT * d; i += RT_THREADS) {
    const int q = i / d;
    s_q[i] = (q < nq) ? Q[(int64_t)q * d + (i % d)] : 0.f;
  }
  __syncthreads();
  for (int q = warp; q < QT; q += RT_WARPS) {
    float ss = 0.f;
    for (int j = lane; j < d; j += 32) ss += s_q[q * d + j] * s_q[q * d + j];
    ss = warp_sum(ss);
    if (lane == 0) s_inv[q] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  }
  __syncthreads();
  // every warp ranks the rows it scores as it goes: QT sorted lists of 32 (score, row), lane l holding the l-th best
  float tv[QT];
  int ti[QT];
#pragma unroll
  for (int q = 0; q < QT; ++q) { tv[q] = -INFINITY; ti[q] = 0x7fffffff; }
  for (; n0 < N; n0 += rows_per_step) {
    float dot[ROWS][QT], nn[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) nn[r] = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int q = 0; q < QT; ++q) dot[r][q] = 0.f;
    if constexpr (NV > 0) {
      // consume the rows already in registers, then request the next pair
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const float* x = pre[r][j];
          nn[r] += x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
#pragma unroll
          for (int q = 0; q < QT; ++q) {
            const float4 qq = *reinterpret_cast<const float4*>(s_q + q * d + j * 128 + lane * 4);
            dot[r][q] += x[0] * qq.x + x[1] * qq.y + x[2] * qq.z + x[3] * qq.w;
          }
        }
      const int nx = n0 + rows_per_step;
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (nx + r < N) load4<DT>(D + (int64_t)(nx + r) * d, j * 128 + lane * 4, pre[r][j]);
    } else {
      for (int kk = lane * 4; kk < d; kk += 128) {
        float x[ROWS][4];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          if (n0 + r < N) load4<DT>(D + (int64_t)(n0 + r) * d, kk, x[r]);
          else x[r][0] = x[r][1] = x[r][2] = x[r][3] = 0.f;
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          nn[r] += x[r][0] * x[r][0] + x[r][1] * x[r][1] + x[r][2] * x[r][2] + x[r][3] * x[r][3];
#pragma unroll
          for (int q = 0; q < QT; ++q) {
            const float4 qq = *reinterpret_cast<const float4*>(s_q + q * d + kk);
            dot[r][q] += x[r][0] * qq.x + x[r][1] * qq.y + x[r][2] * qq.z + x[r][3] * qq.w;
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const float inv_n = 1.0f / fmaxf(sqrtf(warp_sum(nn[r])), 1e-12f);
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        float v = warp_sum(dot[r][q]) * s_inv[q] * inv_n;
        if (lane == 0 && q < nq && n0 + r < N) scores[(int64_t)q * lds + n0 + r] = v;
        if (k > 0 && n0 + r < N) {  // (warp-uniform) one insertion, only if the row beats the list's last entry
          if (!(v == v)) v = -INFINITY;  // a NaN in the inputs ranks last
          const float kv = __shfl_sync(0xffffffffu, tv[q], 31);
          const int ki = __shfl_sync(0xffffffffu, ti[q], 31);
          if (ranks_before(v, n0 + r, kv, ki)) warp_topk_insert(tv[q], ti[q], v, n0 + r, lane);
        }
      }
    }
  }
  if (k <= 0) return;

  // ---- ranking: warps -> CTA -> group of 16 CTAs -> all groups; whoever arrives last at a level merges it ----------
  // ws: [0] global ticket, [1 + g] ticket of group g, then (from word 64) the CTA lists [gridDim.x][QT] and the group
  // lists [n_groups][QT], each a sorted list of 32 scores followed by their 32 rows
  const int n_groups = ((int)gridDim.x + RT_GROUP - 1) / RT_GROUP;
  float* cta_lists = reinterpret_cast<float*>(ws + 64);
  float* grp_lists = cta_lists + (size_t)gridDim.x * QT * 64;
  // the 8 warp lists of this CTA -> warp 0, as a tree of bitonic merges (3 levels)
  auto cta_merge = [&](float& mv, int& mi) {
#pragma unroll
    for (int half = RT_WARPS / 2; half >= 1; half >>= 1) {
      s_tv[warp][lane] = mv;
      s_ti[warp][lane] = mi;
      __syncthreads();
      if (warp < half) warp_topk_merge(mv, mi, s_tv[warp + half][lane], s_ti[warp + half][lane], lane);
      __syncthreads();
    }
  };
  auto store_list = [&](float* dst, float mv, int mi) {
    dst[lane] = mv;
    reinterpret_cast<int*>(dst)[32 + lane] = mi;
  };
  auto load_merge = [&](const float* src, bool in, float& mv, int& mi) {
    const float lv = in ? __ldcg(src + lane) : -INFINITY;  // L2: written by other SMs
    const int li = in ? __ldcg(reinterpret_cast<const int*>(src) + 32 + lane) : 0x7fffffff;
    warp_topk_merge(mv, mi, lv, li, lane);
  };
#pragma unroll
  for (int q = 0; q < QT; ++q) {
    cta_merge(tv[q], ti[q]);
    if (warp == 0) store_list(cta_lists + ((size_t)blockIdx.x * QT + q) * 64, tv[q], ti[q]);
  }
  const int grp = blockIdx.x / RT_GROUP;
  const int grp_size = min(RT_GROUP, (int)gridDim.x - grp * RT_GROUP);
  __threadfence();  // this CTA's lists are visible device-wide before its ticket is
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ws + 1 + grp, 1) == grp_size - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int q = 0; q < nq; ++q) {  // the group's 16 CTA lists: two per warp
    float mv = -INFINITY;
    int mi = 0x7fffffff;
#pragma unroll
    for (int u = 0; u < RT_GROUP / RT_WARPS; ++u) {
      const int c = warp + u * RT_WARPS;
      load_merge(cta_lists + ((size_t)(grp * RT_GROUP + c) * QT + q) * 64, c < grp_size, mv, mi);
    }
    cta_merge(mv, mi);
    if (warp == 0) store_list(grp_lists + ((size_t)grp * QT + q) * 64, mv, mi);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ws, 1) == n_groups - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int q = 0; q < nq; ++q) {  // the <= 19 group lists
    float mv = -INFINITY;
    int mi = 0x7fffffff;
    for (int g0 = warp; g0 < n_groups; g0 += RT_WARPS) load_merge(grp_lists + ((size_t)g0 * QT + q) * 64, true, mv, mi);
    cta_merge(mv, mi);
    if (warp == 0 && lane < k) {
      const bool real = mi != 0x7fffffff;  // fewer than k rows in the database: pad with (-inf, -1)
      top_val[(int64_t)q * k + lane] = real ? mv : -INFINITY;
      top_idx[(int64_t)q * k + lane] = real ? mi + index_base : -1;
    }
  }
  if (threadIdx.x <= n_groups) ws[threadIdx.x] = 0;  // every ticket back to zero: ready for the next launch
}

// Final merge of per-shard candidates (row-sharded database, one shard per rank): out = the k best of m (value,
// index) pairs per query, same order as the ranking above.  One warp per query.
__global__ void __launch_bounds__(32)
topk_merge_kernel(const float* __restrict__ cand_val, const int32_t* __restrict__ cand_idx, int m, int k,
                  float* __restrict__ out_val, int32_t* __restrict__ out_idx) {
  const int q = blockIdx.x, lane = threadIdx.x;
  float tv = -INFINITY;
  int ti = 0x7fffffff;
  for (int base = 0; base < m; base += 32) {
    const int j = base + lane;
    float s = -INFINITY;
    int idx = 0x7fffffff;
    if (j < m) {
      s = cand_val[(int64_t)q * m + j];
      idx = cand_idx[(int64_t)q * m + j];
      if (idx < 0) { s = -INFINITY; idx = 0x7fffffff; }  // padding of a short shard
      if (!(s == s)) s = -INFINITY;
    }
    warp_topk_offer(tv, ti, s, idx, lane);
  }
  if (lane < k) {
    const bool real = ti != 0x7fffffff;
    out_val[(int64_t)q * k + lane] = real ? tv : -INFINITY;
    out_idx[(int64_t)q * k + lane] = real ? ti : -1;
  }
}

template <int QT, typename DT>
int launch_retrieval(const float* Q, const DT* D, float* scores, int nq, int N, int d, int64_t lds, int k,
                     int index_base, float* top_val, int32_t* top_idx, int* ws, cudaStream_t stream) {
  const size_t smem = ((size_t)QT * d + QT) * sizeof(float);
  PCY_REQUIRE(smem <= 200 * 1024, "retrieval_scores: d=%d too large", d);
  // enough CTAs for every SM to keep >= 64 KB of row data in flight, never more rows than the database has
  int grid = std::min(2 * num_sms(), std::max(1, ceil_div(N, RT_WARPS * 2)));
#define PCY_RT_LAUNCH(NV_, ROWS_)                                                                                        \
  do {                                                                                                            \
    auto kern = retrieval_scores_topk_kernel<QT, NV_, ROWS_, DT>;                                                        \
    static SmemOptIn opt;                                                                                         \
    if (smem > 48 * 1024 && opt.need(smem))                                                                       \
      PCY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    kern<<<grid, RT_THREADS, smem, stream>>>(Q, D, scores, nq, N, d, lds, k, index_base, top_val, top_idx, ws);         \
  } while (0)
  if (d == 1280) PCY_RT_LAUNCH(10, 2);
  else if (d == 2560) PCY_RT_LAUNCH(20, 1);
  else if (d == 640) PCY_RT_LAUNCH(5, 2);
  else PCY_RT_LAUNCH(0, 2);
#undef PCY_RT_LAUNCH
  PCY_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int retrieval_scores_topk(const float* Q, const void* D, int db_bf16, float* scores, int nq, int N, int d, int64_t lds,
                          int k, int index_base, float* top_val, int32_t* top_idx, int* ws, cudaStream_t stream) {
  PCY_REQUIRE(d % 4 == 0, "retrieval_scores: d %% 4 != 0");
  PCY_REQUIRE(k >= 0 && k <= RT_MAXK, "retrieval_scores: top-k of %d > %d is not fused (rank the scores instead)", k,
              RT_MAXK);
  PCY_REQUIRE(k == 0 || (top_val && top_idx && ws), "retrieval_scores: top-k outputs / workspace missing");
  if (nq == 0) return 0;
  if (N == 0) {
    if (k > 0) {  // empty shard: nothing to rank
      std::vector<float> v((size_t)nq * k, -INFINITY);
      std::vector<int32_t> idx((size_t)nq * k, -1);
      PCY_CUDA(cudaMemcpyAsync(top_val, v.data(), v.size() * 4, cudaMemcpyHostToDevice, stream));
      PCY_CUDA(cudaMemcpyAsync(top_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, stream));
      PCY_CUDA(cudaStreamSynchronize(stream));
    }
    return 0;
  }
  // up to 4 queries share one pass over the database; more queries take further passes (the ranking of a pass only
  // covers its own queries, so every pass is complete in itself)
  for (int q0 = 0; q0 < nq; q0 += 4) {
    const int cnt = std::min(4, nq - q0);
    const float* q = Q + (int64_t)q0 * d;
    float* sc = scores + (int64_t)q0 * lds;
    float* tv = top_val ? top_val + (int64_t)q0 * k : nullptr;
    int32_t* ti = top_idx ? top_idx + (int64_t)q0 * k : nullptr;
#define PCY_RT_QT(QT_)                                                                                              \
  do {                                                                                                              \
    if (db_bf16)                                                                                                    \
      PCY_TRY((launch_retrieval<QT_, bf16>(q, (const bf16*)D, sc, cnt, N, d, lds, k, index_base, tv, ti, ws,       \
                                           stream)));                                                              \
    else                                                                                                            \
      PCY_TRY((launch_retrieval<QT_, float>(q, (const float*)D, sc, cnt, N, d, lds, k, index_base, tv, ti, ws,     \
                                            stream)));                                                             \
  } while (0)
    if (cnt == 1) PCY_RT_QT(1);
    else if (cnt == 2) PCY_RT_QT(2);
    else PCY_RT_QT(4);
#undef PCY_RT_QT
  }
  return 0;
}

int topk_merge(const float* cand_val, const int32_t* cand_idx, int nq, int m, int k, float* out_val, int32_t* out_idx,
               cudaStream_t stream) {
  PCY_REQUIRE(k >= 1 && k <= RT_MAXK && m >= 0, "topk_merge: k=%d m=%d out of range", k, m);
  if (nq == 0) return 0;
  topk_merge_kernel<<<nq, 32, 0, stream>>>(cand_val, cand_idx, m, k, out_val, out_idx);
  PCY_LAUNCH_CHECK();
  return 0;
}

int cosine_scores(const float* Q, const void* D, int db_bf16, float* out, int nq, int N, int d, int64_t ldo,
                  cudaStream_t stream) {
  return retrieval_scores_topk(Q, D, db_bf16, out, nq, N, d, ldo, 0, 0, nullptr, nullptr, nullptr, stream);
}

}  // namespace pcy

using namespace pcy;

extern "C" {

int pcy_cosine_scores(const float* queries, const void* db, int db_is_bf16, float* out, int n_queries, int n_db,
                      int d, int64_t ld_out, void* stream) {
  return cosine_scores(queries, db, db_is_bf16, out, n_queries, n_db, d, ld_out, (cudaStream_t)stream);
}

int pcy_retrieval_scores_topk(const float* queries, const void* db, int db_is_bf16, float* scores, int n_queries,
                              int n_db, int d, int64_t ld_scores, int k, int index_base, float* top_val,
                              int32_t* top_idx, int32_t* workspace, void* stream) {
  return retrieval_scores_topk(queries, db, db_is_bf16, scores, n_queries, n_db, d, ld_scores, k, index_base, top_val,
                               top_idx, workspace, (cudaStream_t)stream);
}

int64_t pcy_retrieval_workspace_bytes(void) {
  const int64_t grid = 2 * num_sms();
  return 256 + (grid + (grid + 15) / 16) * 4 * 64 * 4;
}

int pcy_topk_merge(const float* cand_val, const int32_t* cand_idx, int n_queries, int m, int k, float* out_val,
                   int32_t* out_idx, void* stream) {
  return topk_merge(cand_val, cand_idx, n_queries, m, k, out_val, out_idx, (cudaStream_t)stream);
}

}  // extern "C"
