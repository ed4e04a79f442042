// extern "C" surface of libprocyon_b200.so (declared in include/procyon_b200.h).
#include "../../include/procyon_b200.h"

#include "common.cuh"
#include "ops.h"

namespace pcy {
const char* last_error();
long long launch_count();
void reset_launch_count();

__global__ void pack_gate_up_kernel(const uint4* __restrict__ gate, const uint4* __restrict__ up,
                                    uint4* __restrict__ packed, int F, int K8) {
  // one thread per 16-byte chunk of the packed matrix
  const int64_t total = (int64_t)2 * F * K8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / K8;
    const int c = (int)(i % K8);
    const int g = (int)(row / 32), r = (int)(row % 32);
    const int src_row = g * 16 + (r & 15);
    packed[i] = (r < 16 ? gate : up)[(int64_t)src_row * K8 + c];
  }
}
}  // namespace pcy

using namespace pcy;

extern "C" {

const char* pcy_last_error(void) { return pcy::last_error(); }
int pcy_version(void) { return 100; }
long long pcy_launch_count(void) { return pcy::launch_count(); }
void pcy_reset_launch_count(void) { pcy::reset_launch_count(); }

static GemmArgs make_args(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                          int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                          int scale_ncols, int c_fp32) {
  GemmArgs a;
  a.A = (const bf16*)A; a.lda = lda; a.W = (const bf16*)W; a.ldw = ldw; a.C = C; a.ldc = ldc;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.residual = (const bf16*)residual; a.ldr = ldr;
  a.act = act; a.scale = scale; a.scale_ncols = scale_ncols; a.c_fp32 = c_fp32;
  return a;
}

int pcy_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                    int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                    int scale_ncols, int c_fp32, void* stream) {
  return gemm_bf16(make_args(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, act, scale, scale_ncols, c_fp32),
                   (cudaStream_t)stream);
}

int pcy_linear_bf16_ex(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                       int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                       int scale_ncols, int c_fp32, int force_tc, const void* rms_weight, float rms_eps,
                       void* stream) {
  GemmArgs a = make_args(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, act, scale, scale_ncols, c_fp32);
  if (force_tc) {
    if (rms_weight) return set_error(PCY_ERR_UNSUPPORTED, "fused RMSNorm is only available on the skinny path");
    return gemm_bf16_tc(a, (cudaStream_t)stream);
  }
  return gemm_bf16_skinny(a, (const bf16*)rms_weight, rms_eps, (cudaStream_t)stream);
}

int pcy_pack_gate_up(const void* gate, const void* up, void* packed, int F, int K, void* stream) {
  PCY_REQUIRE(F % 16 == 0 && K % 8 == 0, "pack_gate_up: F %% 16 and K %% 8 must be 0 (F=%d K=%d)", F, K);
  pack_gate_up_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)stream>>>((const uint4*)gate, (const uint4*)up,
                                                                      (uint4*)packed, F, K / 8);
  PCY_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
