#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_llama.py -m gpu -q -k "programmatic or beam" 2>&1 | tail -5
./scripts/microbench/grid_barrier 2>&1 | tail -6 | tee gpurun_out/grid_barrier.log
timeout 300 python scripts/bench_retrieval.py 2>&1 | tee gpurun_out/retrieval.log | cut -c1-260
timeout 300 python scripts/bench_decode_rows.py 2>&1 | tee gpurun_out/decode_rows.log | tail -12
