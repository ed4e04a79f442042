#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q 2>&1 | tail -8
timeout 300 python scripts/bench_retrieval.py 2>&1 | tee gpurun_out/retrieval.log | cut -c1-260
