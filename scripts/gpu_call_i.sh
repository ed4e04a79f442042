#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llama.py tests/test_gpu_unified.py -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k.log
tail -12 gpurun_out/pytest_k.log
timeout 600 python scripts/bench_decode_rows.py 6,10,16 2>&1 | tail -4
