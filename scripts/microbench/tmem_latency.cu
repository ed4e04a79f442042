// Microbenchmark: round-trip latency of the TMEM accesses a softmax thread of the attention kernels makes in every key
// step, on an otherwise idle B200 SM: tcgen05.ld 32x32b.x32 + wait::ld, tcgen05.st 32x32b.x16 + wait::st, and the
// x32 load while another warp keeps the tensor pipe busy with N = 64 MMAs (the situation inside the kernel).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I procyon_b200/csrc -o scripts/microbench/tmem_latency scripts/microbench/tmem_latency.cu
#include <cstdio>

#include "common.cuh"

using namespace pcy;

__global__ void __launch_bounds__(128, 1) tmem_latency_kernel(int busy, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB = base;  // 64 rows x 64 bf16
  const uint32_t bar = base + 8192, slot = bar + 8;
  for (uint32_t i = threadIdx.x; i < 8192 / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 4 * i), "r"(0u) : "memory");
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0 && busy) {
    // keep the tensor pipe busy: N = 64 MMAs with a TMEM A operand into columns [256, 320)
    const uint32_t idesc = make_idesc_bf16(128, 64);
    const uint64_t bd = make_desc_kmajor_sw128(sB);
    for (int i = 0; i < 4000; ++i)
      tc_mma_bf16_ts(tmem_base + 256, tmem_base + 384 + 8 * (i & 3), bd + 2 * (i & 3), idesc, i > 0 ? 1u : 0u);
    tc_commit(bar);
    mbar_wait(bar, 0);
  }
  if (warp == 0) {
    const uint32_t t_lane = tmem_base;  // lanes 0..31
    long long ld_best = 1ll << 60, st_best = 1ll << 60, ld8 = 1ll << 60, pair_best = 1ll << 60;
    uint32_t v[32], w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = i;
    uint32_t sink = 0;
    for (int r = 0; r < 200; ++r) {
      long long t0 = clock64();
      tmem_ld_32x32b_x32(t_lane, v);
      tc_wait_ld();
      long long t1 = clock64();
      sink += v[r & 31];
      if (t1 - t0 < ld_best) ld_best = t1 - t0;
      t0 = clock64();
      tmem_st_32x32b_x16(t_lane + 64, w);
      tc_wait_st();
      t1 = clock64();
      if (t1 - t0 < st_best) st_best = t1 - t0;
      t0 = clock64();
#pragma unroll 1
      for (int k = 0; k < 8; ++k) {  // eight dependent load round trips
        tmem_ld_32x32b_x32(t_lane + (sink & 1), v);
        tc_wait_ld();
        sink += v[k];
      }
      t1 = clock64();
      if (t1 - t0 < ld8) ld8 = t1 - t0;
      t0 = clock64();  // what a step does: read 64 columns, write 32, both waited for
      tmem_ld_32x32b_x32(t_lane, v);
      tc_wait_ld();
      sink += v[3];
      tmem_st_32x32b_x16(t_lane, w);
      tmem_st_32x32b_x16(t_lane + 16, w);
      tc_wait_st();
      t1 = clock64();
      if (t1 - t0 < pair_best) pair_best = t1 - t0;
    }
    if (lane == 0 && blockIdx.x == 0) {
      out[0] = ld_best; out[1] = st_best; out[2] = ld8; out[3] = pair_best; out[4] = sink;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  const int smem = 8192 + 64 + 1024;
  for (int busy = 0; busy < 2; ++busy) {
    tmem_latency_kernel<<<1, 128, smem>>>(busy, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[8];
    cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
    printf("{\"tensor_pipe_busy\": %s, \"ld_x32_wait_cycles\": %lld, \"st_x16_wait_cycles\": %lld, "
           "\"ld_x32_dependent_chain_cycles_each\": %.1f, \"ld_x32_then_2_st_x16_cycles\": %lld}\n",
           busy ? "true" : "false", h[0], h[1], h[2] / 8.0, h[3]);
  }
  return 0;
}
