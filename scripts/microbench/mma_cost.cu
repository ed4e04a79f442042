// Microbenchmark: what does ONE tcgen05.mma (kind::f16, bf16 operands, M = 128, K = 16) cost on the tensor pipe of a
// B200 SM as a function of the tile width N and of where the A operand lives (shared-memory descriptor vs TMEM)?
// The ESM2 attention kernels (head_dim 64, 64-key steps) issue sixteen N = 64 instructions per 128 keys; their
// diagnosis needs this number.  One CTA per SM, one thread issues CHAIN instructions back to back
// into the same accumulator (they serialise on the tensor pipe), commits, waits; cycles = clock64 around the chain.
// Operands are zero-filled shared memory (values do not matter for timing).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I procyon_b200/csrc -o gpurun_out/mma_cost scripts/microbench/mma_cost.cu
#include <cstdio>

#include "common.cuh"

using namespace pcy;

template <bool A_TMEM>
__global__ void __launch_bounds__(128, 1) mma_cost_kernel(int N, int chain, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;               // 128 rows x 64 bf16 (one 128-byte swizzle atom wide): 16 KB
  const uint32_t sB = base + 16384;       // 256 rows x 64 bf16: 32 KB
  const uint32_t bar = base + 16384 + 32768;
  const uint32_t slot = bar + 8;
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
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
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t ad = make_desc_kmajor_sw128(sA), bd = make_desc_kmajor_sw128(sB);
    long long best = 1ll << 60;
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int i = 0; i < chain; ++i) {
        // accumulator in columns [0, N); the TMEM A operand (K = 16 -> 8 columns) sits at column 256 + 8 (i & 3)
        if (A_TMEM) tc_mma_bf16_ts(tmem_base, tmem_base + 256 + 8 * (i & 3), bd + 2 * (i & 3), idesc, i > 0 ? 1u : 0u);
        else tc_mma_bf16(tmem_base, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i > 0 ? 1u : 0u);
      }
      tc_commit(bar);
      mbar_wait(bar, phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (blockIdx.x == 0) out[0] = best;
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
  cudaMalloc(&d_out, 8);
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(mma_cost_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_cost_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int grid : {1, sms})
      for (int N : Ns) {
        long long c[2];
        const int chains[2] = {64, 320};
        for (int k = 0; k < 2; ++k) {
          if (a_tmem) mma_cost_kernel<true><<<grid, 128, smem>>>(N, chains[k], 20, d_out);
          else mma_cost_kernel<false><<<grid, 128, smem>>>(N, chains[k], 20, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&c[k], d_out, 8, cudaMemcpyDeviceToHost);
        }
        // slope between the two chain lengths = cycles per instruction without the fixed issue / commit / wait cost
        printf("{\"a_operand\": \"%s\", \"ctas\": %d, \"M\": 128, \"N\": %d, \"K\": 16, \"cycles_per_mma\": %.1f, "
               "\"ideal_cycles\": %.1f, \"fixed_cycles\": %.0f}\n",
               a_tmem ? "tmem" : "smem", grid, N, (double)(c[1] - c[0]) / (chains[1] - chains[0]), N / 2.0,
               c[0] - (double)(c[1] - c[0]) / (chains[1] - chains[0]) * chains[0]);
      }
  return 0;
}
