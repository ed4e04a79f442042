#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_esm.py -m gpu -q -x > gpurun_out/pytest_kernels.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_kernels.log
tail -4 gpurun_out/pytest_kernels.log
timeout 600 python scripts/bench_gemm_shapes.py > gpurun_out/gemm_shapes.log 2>&1
cat gpurun_out/gemm_shapes.log | tail -12
