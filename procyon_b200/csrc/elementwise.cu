// HBM-bound row kernels: LayerNorm / RMSNorm, ESM2 token embedding, Llama embedding + soft-token splice,
// rotary embedding, ProteinPooler.  All use 16-byte vector loads, one warp per row (rows are 0.6-10 KB),
// fp32 statistics and a single rounding to bf16 at the store.
#include "common.cuh"
#include "ops.h"

namespace pcy {

namespace {

constexpr int ROW_THREADS = 256;
constexpr int ROW_WARPS = ROW_THREADS / 32;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (torch.nn.LayerNorm / fair-esm ESM1bLayerNorm, eps inside sqrt, biased variance)
// NV = 16-byte vectors per lane (d <= NV*256)
// ---------------------------------------------------------------------------------------------
template <int NV, bool RMS>
__global__ void __launch_bounds__(ROW_THREADS)
norm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
            bf16* __restrict__ y, int64_t rows, int d, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROW_WARPS + warp;
  pdl_launch_dependents();
  pdl_wait();  // x is the preceding kernel's output
  if (row >= rows) return;
  const bf16* xr = x + row * d;
  uint4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int k = (i * 32 + lane) * 8;
    if (k < d) {
      v[i] = *reinterpret_cast<const uint4*>(xr + k);
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += RMS ? f[j] * f[j] : f[j];
    } else {
      v[i] = make_uint4(0, 0, 0, 0);
    }
  }
  sum = warp_sum(sum);
  float mean = 0.f, rstd;
  if (RMS) {
    rstd = rsqrtf(sum / (float)d + eps);
  } else {
    mean = sum / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int k = (i * 32 + lane) * 8;
      if (k < d) {
        float f[8];
        unpack8(v[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float c = f[j] - mean; sq += c * c; }
      }
    }
    sq = warp_sum(sq);
    rstd = rsqrtf(sq / (float)d + eps);
  }
  bf16* yr = y + row * d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int k = (i * 32 + lane) * 8;
    if (k < d) {
      float f[8], g[8], o[8];
      unpack8(v[i], f);
      unpack8(*reinterpret_cast<const uint4*>(gamma + k), g);
      if (RMS) {
        // HF LlamaRMSNorm: weight * (x_fp32 * rstd).to(input_dtype)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = g[j] * bf16_round(f[j] * rstd);
      } else {
        float b[8];
        unpack8(*reinterpret_cast<const uint4*>(beta + k), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (f[j] - mean) * rstd * g[j] + b[j];
      }
      *reinterpret_cast<uint4*>(yr + k) = pack8(o);
    }
  }
}

template <bool RMS>
int launch_norm(const bf16* x, const bf16* gamma, const bf16* beta, bf16* y, int64_t rows, int d, float eps,
                cudaStream_t stream) {
  PCY_REQUIRE(d % 8 == 0 && d <= 8192, "norm: d=%d must be a multiple of 8 and <= 8192", d);
  if (rows == 0) return 0;
  const int grid = ceil_div(rows, ROW_WARPS);
  const int nv = ceil_div(d, 256);
#define PCY_NORM_CASE(NV)                                                                          \
  launch_pdl(norm_kernel<NV, RMS>, dim3(grid), dim3(ROW_THREADS), 0, stream, x, gamma, beta, y, rows, d, eps)
  if (nv <= 2) PCY_NORM_CASE(2);
  else if (nv <= 5) PCY_NORM_CASE(5);
  else if (nv <= 10) PCY_NORM_CASE(10);
  else if (nv <= 16) PCY_NORM_CASE(16);
  else if (nv <= 20) PCY_NORM_CASE(20);
  else PCY_NORM_CASE(32);
#undef PCY_NORM_CASE
  PCY_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// ESM2 embedding (fair-esm ESM2.forward, reached from procyon/model/esm.py:536):
//   x = E[tok]; x[tok == mask] = 0; x = x * (1 - 0.12) / (1 - n_mask/src_len); x[pad] = 0
// The reference runs this in the module dtype, so each step rounds to bf16; reproduced here.
// grid = (ceil(T / ROW_WARPS), B)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
esm_embed_kernel(const int32_t* __restrict__ tokens, const bf16* __restrict__ table, bf16* __restrict__ x, int T,
                 int d, int pad_idx, int mask_idx, int token_dropout) {
  __shared__ int s_cnt[2];
  const int b = blockIdx.y;
  const int32_t* tok = tokens + (int64_t)b * T;
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  float inv_keep = 1.f;
  if (token_dropout) {
    int n_mask = 0, n_nonpad = 0;
    for (int t = threadIdx.x; t < T; t += ROW_THREADS) {
      const int v = tok[t];
      n_mask += (v == mask_idx);
      n_nonpad += (v != pad_idx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_mask += __shfl_xor_sync(0xffffffffu, n_mask, o);
      n_nonpad += __shfl_xor_sync(0xffffffffu, n_nonpad, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&s_cnt[0], n_mask);
      atomicAdd(&s_cnt[1], n_nonpad);
    }
    __syncthreads();
    // mask_ratio_observed = n_mask.to(bf16) / src_lengths  (bf16 result); denominator (1 - ratio) in bf16
    const float ratio = bf16_round(bf16_round((float)s_cnt[0]) / (float)s_cnt[1]);
    inv_keep = bf16_round(1.0f - ratio);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * ROW_WARPS + warp;
  if (t >= T) return;
  const int v = tok[t];
  bf16* xr = x + ((int64_t)b * T + t) * d;
  const bool zero = (v == pad_idx) || (token_dropout && v == mask_idx);
  const bf16* er = table + (int64_t)v * d;
  for (int k = lane * 8; k < d; k += 256) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (!zero) {
      u = *reinterpret_cast<const uint4*>(er + k);
      if (token_dropout) {
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j] * 0.88f) / inv_keep;
        u = pack8(f);
      }
    }
    *reinterpret_cast<uint4*>(xr + k) = u;
  }
}

// ---------------------------------------------------------------------------------------------
// Llama embedding + soft-token splice (procyon/model/model_unified.py:1147-1167):
//   z = embed_tokens(ids); z[ids == placeholder] = soft tokens in row-major order.
// soft_index[i] = row of `soft` to write at flat position i, or -1 to take the embedding row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
llama_embed_splice_kernel(const int32_t* __restrict__ ids, const bf16* __restrict__ table,
                          const bf16* __restrict__ soft, const int32_t* __restrict__ soft_index,
                          bf16* __restrict__ x, int64_t n_tok, int d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * ROW_WARPS + warp;
  if (i >= n_tok) return;
  const int si = soft_index ? soft_index[i] : -1;
  const bf16* src = (si >= 0) ? soft + (int64_t)si * d : table + (int64_t)ids[i] * d;
  bf16* dst = x + i * d;
  for (int k = lane * 8; k < d; k += 256)
    *reinterpret_cast<uint4*>(dst + k) = *reinterpret_cast<const uint4*>(src + k);
}

// ---------------------------------------------------------------------------------------------
// Rotary embedding, rotate-half convention (fair-esm RotaryEmbedding; HF apply_rotary_pos_emb):
//   out[i] = x[i]*cos[i] - x[i+h]*sin[i];  out[i+h] = x[i+h]*cos[i] + x[i]*sin[i],  h = head_dim/2
// Applied in place to `n_heads` heads starting at column col0 of every row; position = pos0 + (row % T).
// cos_sin: fp32 [P][h][2].  One warp per (row, head).
// ---------------------------------------------------------------------------------------------
// IdxT: uint32_t whenever rows * threads-per-row fits (the two divisions per unit are 64-bit otherwise: ~200
// instructions in front of 32 bytes of traffic)
template <typename IdxT>
__global__ void __launch_bounds__(ROW_THREADS)
rope_kernel(bf16* __restrict__ x, int64_t rows, int T, int n_heads, int head_dim, int64_t ld, int col0,
            const float* __restrict__ cos_sin, const int32_t* __restrict__ pos_ptr, int pos0) {
  // one thread = 8 consecutive columns of the low half of a head and their partners in the high half
  const int half = head_dim >> 1;
  const int per_head = half >> 3;  // threads per head (head_dim % 16 == 0 on this path)
  const IdxT per_row = (IdxT)n_heads * (IdxT)per_head;
  const IdxT total = (IdxT)rows * per_row;
  for (IdxT idx = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (IdxT)gridDim.x * blockDim.x) {
    const IdxT row = idx / per_row;
    const int rem = (int)(idx - row * per_row);
    const int h = rem / per_head, i = (rem - h * per_head) * 8;
    const int pos = (pos_ptr ? pos_ptr[0] : pos0) + (int)(row % (IdxT)T);
    bf16* p = x + (int64_t)row * ld + col0 + h * head_dim + i;
    const float4* cs = reinterpret_cast<const float4*>(cos_sin + ((int64_t)pos * half + i) * 2);
    float lo[8], hi[8];
    unpack8(*reinterpret_cast<const uint4*>(p), lo);
    unpack8(*reinterpret_cast<const uint4*>(p + half), hi);
    float ol[8], oh[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float4 t = __ldg(cs + (j >> 1));  // (cos_j, sin_j, cos_j+1, sin_j+1)
      ol[j] = lo[j] * t.x - hi[j] * t.y;
      oh[j] = hi[j] * t.x + lo[j] * t.y;
      ol[j + 1] = lo[j + 1] * t.z - hi[j + 1] * t.w;
      oh[j + 1] = hi[j + 1] * t.z + lo[j + 1] * t.w;
    }
    *reinterpret_cast<uint4*>(p) = pack8(ol);
    *reinterpret_cast<uint4*>(p + half) = pack8(oh);
  }
}

// scalar variant for head dims that are not multiples of 16 (ESM2-35M: 24): one warp per (row, head)
__global__ void __launch_bounds__(ROW_THREADS)
rope_kernel_generic(bf16* __restrict__ x, int64_t rows, int T, int n_heads, int head_dim, int64_t ld, int col0,
                    const float* __restrict__ cos_sin, const int32_t* __restrict__ pos_ptr, int pos0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t idx = (int64_t)blockIdx.x * ROW_WARPS + warp;
  if (idx >= rows * n_heads) return;
  const int64_t row = idx / n_heads;
  const int h = (int)(idx % n_heads);
  const int half = head_dim >> 1;
  const int pos = (pos_ptr ? pos_ptr[0] : pos0) + (int)(row % T);
  bf16* p = x + row * ld + col0 + h * head_dim;
  const float2* cs = reinterpret_cast<const float2*>(cos_sin) + (int64_t)pos * half;
  for (int i = lane * 2; i < half; i += 64) {
    const float2 lo = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p + i));
    const float2 hi = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p + i + half));
    const float2 c0 = cs[i], c1 = cs[i + 1];
    const float o0 = lo.x * c0.x - hi.x * c0.y, o1 = lo.y * c1.x - hi.y * c1.y;
    const float q0 = hi.x * c0.x + lo.x * c0.y, q1 = hi.y * c1.x + lo.y * c1.y;
    *reinterpret_cast<uint32_t*>(p + i) = pack_bf16x2(o0, o1);
    *reinterpret_cast<uint32_t*>(p + i + half) = pack_bf16x2(q0, q1);
  }
}

// ---------------------------------------------------------------------------------------------
// ProteinPooler (procyon/model/esm.py:131-217): for each output protein o, reduce over all rows of all
// chunks whose batch key == o (seg_rows lists them), skipping pad tokens.
//   mean: nanmean over non-pad rows (CLS/EOS included; `correction` drops the first and last non-pad row of
//         the concatenated sequence, esm.py:144-145)
//   max : max over non-pad rows
// seg_ptr [n_out+1] offsets into seg_rows (chunk-row indices, in concatenation order).
// grid = (n_out, ceil(d / 256)); block 256 threads = 8 warps striding over tokens, 32 lanes x 8 columns.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
pool_kernel(const bf16* __restrict__ x, const int32_t* __restrict__ tokens, const int32_t* __restrict__ seg_ptr,
            const int32_t* __restrict__ seg_rows, void* __restrict__ out, int out_fp32, int T, int d, int pad_idx,
            int mode, int correction) {
  __shared__ float s_part[ROW_WARPS][256];
  __shared__ int s_first, s_last, s_count;
  const int o = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * 256 + lane * 8;
  const int beg = seg_ptr[o], end = seg_ptr[o + 1];
  const int n_pos = (end - beg) * T;  // concatenated positions
  // first / last non-pad position and count (needed for the correction option and the mean divisor)
  if (threadIdx.x == 0) { s_first = n_pos; s_last = -1; s_count = 0; }
  __syncthreads();
  {
    int first = n_pos, last = -1, cnt = 0;
    for (int p = threadIdx.x; p < n_pos; p += ROW_THREADS) {
      const int r = seg_rows[beg + p / T];
      if (tokens[(int64_t)r * T + (p % T)] != pad_idx) { first = min(first, p); last = max(last, p); ++cnt; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      first = min(first, __shfl_xor_sync(0xffffffffu, first, s));
      last = max(last, __shfl_xor_sync(0xffffffffu, last, s));
      cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    }
    if (lane == 0) { atomicMin(&s_first, first); atomicMax(&s_last, last); atomicAdd(&s_count, cnt); }
  }
  __syncthreads();
  const int first = s_first, last = s_last;
  int count = s_count;
  const bool drop_ends = (mode == 0 && correction);
  if (drop_ends) count = max(count - 2, 0);

  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (mode == 1) ? -INFINITY : 0.f;
  if (c0 < d) {
    for (int p = warp; p < n_pos; p += ROW_WARPS) {
      const int r = seg_rows[beg + p / T];
      const int t = p % T;
      if (tokens[(int64_t)r * T + t] == pad_idx) continue;
      if (drop_ends && (p == first || p == last)) continue;
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + ((int64_t)r * T + t) * d + c0), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = (mode == 1) ? fmaxf(acc[j], f[j]) : acc[j] + f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_part[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  if (warp == 0 && c0 < d) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = s_part[0][lane * 8 + j];
      for (int w = 1; w < ROW_WARPS; ++w) {
        const float u = s_part[w][lane * 8 + j];
        v = (mode == 1) ? fmaxf(v, u) : v + u;
      }
      if (mode == 0) v = (count > 0) ? v / (float)count : __int_as_float(0x7fc00000);  // nanmean of nothing = nan
      acc[j] = v;
    }
    if (out_fp32) {
      float* op = reinterpret_cast<float*>(out) + (int64_t)o * d + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) op[j] = acc[j];
    } else {
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out) + (int64_t)o * d + c0) = pack8(acc);
    }
  }
}

__global__ void key_valid_kernel(const int32_t* __restrict__ tokens, uint8_t* __restrict__ valid, int64_t n,
                                 int pad_idx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    valid[i] = tokens[i] != pad_idx;
}

__global__ void rope_table_kernel(float* __restrict__ cos_sin, int P, int half, float theta, int head_dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * half) return;
  const int pos = i / half, j = i % half;
  // inv_freq = theta^(-2j/dim) in fp32, angle = pos * inv_freq in fp32 (torch semantics)
  const float inv_freq = 1.0f / powf(theta, (float)(2 * j) / (float)head_dim);
  const float ang = (float)pos * inv_freq;
  cos_sin[2 * i] = cosf(ang);
  cos_sin[2 * i + 1] = sinf(ang);
}

}  // namespace

int layernorm_bf16(const bf16* x, const bf16* gamma, const bf16* beta, bf16* y, int64_t rows, int d, float eps,
                   cudaStream_t stream) {
  return launch_norm<false>(x, gamma, beta, y, rows, d, eps, stream);
}
int rmsnorm_bf16(const bf16* x, const bf16* weight, bf16* y, int64_t rows, int d, float eps, cudaStream_t stream) {
  return launch_norm<true>(x, weight, nullptr, y, rows, d, eps, stream);
}

int esm_embed(const int32_t* tokens, const bf16* table, bf16* x, int B, int T, int d, int pad_idx, int mask_idx,
              int token_dropout, cudaStream_t stream) {
  PCY_REQUIRE(d % 8 == 0, "esm_embed: d %% 8 != 0");
  if (B == 0 || T == 0) return 0;
  dim3 grid(ceil_div(T, ROW_WARPS), B);
  esm_embed_kernel<<<grid, ROW_THREADS, 0, stream>>>(tokens, table, x, T, d, pad_idx, mask_idx, token_dropout);
  PCY_LAUNCH_CHECK();
  return 0;
}

int llama_embed_splice(const int32_t* ids, const bf16* table, const bf16* soft_tokens, const int32_t* soft_index,
                       bf16* x, int64_t n_tok, int d, cudaStream_t stream) {
  PCY_REQUIRE(d % 8 == 0, "embed_splice: d %% 8 != 0");
  if (n_tok == 0) return 0;
  llama_embed_splice_kernel<<<ceil_div(n_tok, ROW_WARPS), ROW_THREADS, 0, stream>>>(ids, table, soft_tokens,
                                                                                   soft_index, x, n_tok, d);
  PCY_LAUNCH_CHECK();
  return 0;
}

int rope_inplace(bf16* x, int64_t rows, int T, int n_heads, int head_dim, int64_t ld, int col0,
                 const float* cos_sin, const int32_t* pos_ptr, int pos0, cudaStream_t stream) {
  PCY_REQUIRE(head_dim % 4 == 0 && col0 % 2 == 0 && ld % 2 == 0, "rope: head_dim %% 4, col0 %% 2, ld %% 2 must be 0");
  if (rows == 0) return 0;
  const int64_t units = rows * n_heads;
  if (head_dim % 16 == 0 && ld % 8 == 0 && col0 % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int64_t threads = units * (head_dim / 16);
    int64_t grid = (threads + ROW_THREADS - 1) / ROW_THREADS;
    if (grid > (int64_t)num_sms() * 32) grid = (int64_t)num_sms() * 32;
    if (threads + grid * ROW_THREADS < (int64_t)1 << 31)  // (the grid-stride increment must not wrap either)
      rope_kernel<uint32_t><<<(int)grid, ROW_THREADS, 0, stream>>>(x, rows, T, n_heads, head_dim, ld, col0, cos_sin,
                                                                  pos_ptr, pos0);
    else
      rope_kernel<int64_t><<<(int)grid, ROW_THREADS, 0, stream>>>(x, rows, T, n_heads, head_dim, ld, col0, cos_sin,
                                                                 pos_ptr, pos0);
  } else {
    rope_kernel_generic<<<ceil_div(units, ROW_WARPS), ROW_THREADS, 0, stream>>>(x, rows, T, n_heads, head_dim, ld,
                                                                                col0, cos_sin, pos_ptr, pos0);
  }
  PCY_LAUNCH_CHECK();
  return 0;
}

int rope_table(float* cos_sin, int P, int head_dim, float theta, cudaStream_t stream) {
  const int half = head_dim / 2;
  rope_table_kernel<<<ceil_div((int64_t)P * half, 256), 256, 0, stream>>>(cos_sin, P, half, theta, head_dim);
  PCY_LAUNCH_CHECK();
  return 0;
}

int make_key_valid(const int32_t* tokens, uint8_t* valid, int64_t n, int pad_idx, cudaStream_t stream) {
  if (n == 0) return 0;
  int grid = ceil_div(n, 256);
  if (grid > 4096) grid = 4096;
  key_valid_kernel<<<grid, 256, 0, stream>>>(tokens, valid, n, pad_idx);
  PCY_LAUNCH_CHECK();
  return 0;
}

int pool_segments(const bf16* x, const int32_t* tokens, const int32_t* seg_ptr, const int32_t* seg_rows, void* out,
                  int out_fp32, int T, int d, int n_out, int pad_idx, int mode, int correction,
                  cudaStream_t stream) {
  PCY_REQUIRE(d % 8 == 0, "pool: d %% 8 != 0");
  PCY_REQUIRE(mode == 0 || mode == 1, "pool: mode must be 0 (mean) or 1 (max)");
  if (n_out == 0) return 0;
  dim3 grid(n_out, ceil_div(d, 256));
  pool_kernel<<<grid, ROW_THREADS, 0, stream>>>(x, tokens, seg_ptr, seg_rows, out, out_fp32, T, d, pad_idx, mode,
                                                correction);
  PCY_LAUNCH_CHECK();
  return 0;
}

}  // namespace pcy
