#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pair_mma.py -m gpu -q 2>&1 | tail -8
for t in 0 192 256 128; do echo "== tile $t"; PCY_GEMM_FORCE_TILE=$t timeout 300 python scripts/bench_gemm_shapes.py 2>&1 | grep -E "llama (qkv|o)" | cut -c1-130; done
timeout 600 python bench.py --quick --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print(d['value'],d['phases'])"; tail -3 gpurun_out/bench_quick.err
