#!/bin/bash
# ESM iteration: parity tests of the encoder / fused model paths, in-situ breakdown of the batch encode, then the
# cta_group::2 GEMM (own process + timeout: a hang must not take the rest of the call with it)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_esm.py tests/test_gpu_unified.py -m gpu -q -x > gpurun_out/pytest_esm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_esm.log
tail -15 gpurun_out/pytest_esm.log
timeout 300 python scripts/profile_esm_breakdown.py > gpurun_out/esm_breakdown.log 2>&1; echo "breakdown rc=$?"
cat gpurun_out/esm_breakdown.log | tail -12
timeout 240 python -m pytest tests/test_gpu_pair_mma.py -m gpu -q -x > gpurun_out/pytest_pair.log 2>&1; rc=$?; echo "pair rc=$rc"
tail -25 gpurun_out/pytest_pair.log
if [ $rc -eq 0 ]; then
  timeout 300 python scripts/bench_gemm_shapes.py > gpurun_out/gemm_shapes_pair.log 2>&1; echo "shapes rc=$?"
  cat gpurun_out/gemm_shapes_pair.log
  PCY_PAIR_MMA=1 timeout 300 python scripts/profile_esm_breakdown.py > gpurun_out/esm_breakdown_pair.log 2>&1; echo "breakdown pair rc=$?"
  tail -8 gpurun_out/esm_breakdown_pair.log
fi
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
