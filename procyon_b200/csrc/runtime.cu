// Error state, device queries and launch accounting shared by all kernels.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace pcy {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

const char* last_error() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_error(cudaError_t e, const char* what, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return PCY_ERR_CUDA;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

// RoPE in the QKV GEMM epilogue (1) or as a separate vectorised pass (0, default: measured faster on B200 because
// the table look-ups lengthen the epilogue past the MMA time of the tile)
bool g_fused_rope = false;

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
void reset_launch_count() { g_launches.store(0, std::memory_order_relaxed); }

}  // namespace pcy
