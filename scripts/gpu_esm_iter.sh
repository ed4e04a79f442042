#!/bin/bash
# ESM iteration: parity tests of the encoder / GEMM / fused model paths (own processes + timeouts: a hang must not take
# the rest of the call with it), then the in-situ breakdown of the batch encode with both attention kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_esm.py -m gpu -q -x > gpurun_out/pytest_esm.log 2>&1; rc=$?; echo "pytest esm rc=$rc"
tail -25 gpurun_out/pytest_esm.log
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pair_mma.py tests/test_gpu_unified.py -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; echo "pytest kernels/unified rc=$?"
tail -8 gpurun_out/pytest_k.log
if [ $rc -eq 0 ]; then
  timeout 300 python scripts/profile_esm_breakdown.py > gpurun_out/esm_breakdown.log 2>&1; echo "breakdown rc=$?"
  tail -8 gpurun_out/esm_breakdown.log
fi
