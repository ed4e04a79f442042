#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pair_mma.py tests/test_gpu_llama.py tests/test_gpu_fullshape.py -m gpu -q -k "gemm or linear or pair or prefill or llama8b_width_prefill" 2>&1 | tail -4
timeout 300 python scripts/bench_gemm_shapes.py 2>&1 | grep llama | cut -c1-200 | tee gpurun_out/gemm_shapes.log
timeout 600 python bench.py --quick --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print(d['value'],d['phases'],d['roofline']['frac'],d['decode_beam10'])"
