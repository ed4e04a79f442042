/* procyon_b200 — C ABI of the B200 (sm_100a) hot path of ProCyon.
 *
 * The reference (mims-harvard/ProCyon) is pure Python: there is no FFI to mirror, so every entry point below
 * names the Python call site(s) of the reference it replaces (paths relative to the reference repo root).
 * Conventions: plain pointers + sizes, all tensors are DEVICE pointers unless stated, row-major, bf16 =
 * uint16 storage; every function takes the CUDA stream (as void*) it must enqueue on and returns 0 on
 * success or a negative pcy status (PCY_ERR_*), with a message available from pcy_last_error().
 * Nothing here allocates per call except the model handles (weights) — workspaces are caller-owned.
 */
#ifndef PROCYON_B200_H_
#define PROCYON_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCY_OK 0
#define PCY_ERR_INVALID_ARG (-1)
#define PCY_ERR_UNSUPPORTED (-2)
#define PCY_ERR_CUDA (-3)
#define PCY_ERR_WORKSPACE (-4)

/* activation codes of pcy_linear_bf16 */
#define PCY_ACT_NONE 0
#define PCY_ACT_GELU 1   /* exact erf GELU (torch.nn.GELU(), fair-esm gelu) */
#define PCY_ACT_SWIGLU 2 /* packed gate/up rows, see pcy_pack_gate_up */

/* ---- library state ------------------------------------------------------------------------------ */
const char* pcy_last_error(void);
int pcy_version(void);
/* number of CUDA kernels this library has launched since the last reset (bench.py: gpu_launches) */
long long pcy_launch_count(void);
void pcy_reset_launch_count(void);

/* ---- dense layers --------------------------------------------------------------------------------
 * C[M,N] = epi(A[M,K] @ W[N,K]^T): torch.nn.Linear as used by fair-esm ESM2 (procyon/model/esm.py:536),
 * HF LlamaDecoderLayer (procyon/model/pmc_llama.py:571) and create_mlp (procyon/model/model_utils.py:13-41).
 * epi: v = acc + bias[n]; if (n < scale_ncols) v *= scale; v = act(v); v += residual[m,n]; store.
 * bias is fp32 [N] or NULL; residual bf16 [M,ldr] or NULL; C is bf16 (c_fp32 = 0) or fp32.
 * M <= 16 takes the weight-streaming path, otherwise tcgen05 tensor cores. */
int pcy_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                    int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                    int scale_ncols, int c_fp32, void* stream);
/* same, but forces the tensor-core (force_tc = 1) or the weight-streaming (force_tc = 0) kernel */
int pcy_linear_bf16_ex(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                       int K, const float* bias, const void* residual, int64_t ldr, int act, float scale,
                       int scale_ncols, int c_fp32, int force_tc, const void* rms_weight, float rms_eps,
                       void* stream);
/* interleave gate_proj / up_proj rows ([F,K] each) into the packed [2F,K] layout PCY_ACT_SWIGLU expects:
 * rows [32g, 32g+16) = gate rows [16g, 16g+16), rows [32g+16, 32g+32) = up rows [16g, 16g+16). F % 16 == 0. */
int pcy_pack_gate_up(const void* gate, const void* up, void* packed, int F, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROCYON_B200_H_ */
