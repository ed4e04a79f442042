// Microbenchmark: what does ONE grid-wide barrier of the persistent decode kernel cost on this GPU when nothing else
// is going on (no skew, no memory traffic)?  148 CTAs x 384 threads, cooperative launch, N barriers back to back.
//   v0: the kernel's barrier (red.release.gpu + relaxed poll by thread 0, CTA barrier before and after)
//   v1: arrivals spread over 8 counters (64-byte apart), thread 0..7 poll one each
//   v2: v0 with the poll by a whole warp on the same word (more requests in flight)
//   v3: v0 without the first CTA barrier's participation of all warps (named barrier of 64 threads only) - lower bound
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/grid_barrier scripts/microbench/grid_barrier.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int V>
__global__ void __launch_bounds__(384, 1) barrier_kernel(unsigned* counters, int n_iter, float* sink) {
  unsigned target = 0;
  const unsigned nb = gridDim.x;
  float acc = 0.f;
  for (int it = 0; it < n_iter; ++it) {
    acc += __sinf((float)it + acc);  // a little dependent work so that the loop is not collapsed
    if (V == 3) {
      if (threadIdx.x < 64) asm volatile("bar.sync 1, 64;" ::: "memory");
    } else {
      __syncthreads();
    }
    if (V == 0 || V == 3) {
      if (threadIdx.x == 0) {
        target += nb;
        red_release(counters, 1u);
        while (ld_relaxed(counters) < target) {}
      }
    } else if (V == 1) {
      target += nb / 8 + ((blockIdx.x & 7) < (nb & 7) ? 0 : 0);
      if (threadIdx.x == 0) red_release(counters + 16 * (blockIdx.x & 7), 1u);
      if (threadIdx.x < 8) {
        // CTAs are dealt round-robin onto 8 counters: counter c sees ceil((nb - c) / 8) arrivals per barrier
        const unsigned per = (nb - threadIdx.x + 7) / 8;
        const unsigned tgt = per * (unsigned)(it + 1);
        while (ld_relaxed(counters + 16 * threadIdx.x) < tgt) {}
      }
    } else if (V == 2) {
      if (threadIdx.x < 32) {
        target += nb;
        if (threadIdx.x == 0) red_release(counters, 1u);
        while (ld_relaxed(counters) < target) {}
      }
    }
    if (V == 3) {
      if (threadIdx.x < 64) asm volatile("bar.sync 1, 64;" ::: "memory");
    } else {
      __syncthreads();
    }
  }
  if (acc == 12345.f) sink[0] = acc;
}

template <int V>
void run(const char* name, int n_sms) {
  unsigned* c;
  float* sink;
  cudaMalloc(&c, 4096);
  cudaMalloc(&sink, 4);
  int n_iter = 2000;
  void* args[] = {&c, &n_iter, &sink};
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(c, 0, 4096);
    cudaEventRecord(a);
    cudaLaunchCooperativeKernel((void*)barrier_kernel<V>, dim3(n_sms), dim3(384), args, 0, 0);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (rep == 2) printf("%-58s %.3f us per barrier (%s)\n", name, ms * 1e3f / n_iter, cudaGetErrorString(e));
  }
  cudaFree(c);
  cudaFree(sink);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>("v0 red.release + relaxed poll by one thread", p.multiProcessorCount);
  run<1>("v1 arrivals over 8 counters, 8 polling threads", p.multiProcessorCount);
  run<2>("v2 poll by a whole warp", p.multiProcessorCount);
  run<3>("v3 as v0, only 64 threads join the CTA barriers", p.multiProcessorCount);
  return 0;
}
