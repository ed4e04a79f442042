#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llama.py -m gpu -q -x -k "tcgen05_causal" 2>&1 | tail -15
timeout 600 python bench.py --quick --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print(d['value'],d['phases'])"; tail -3 gpurun_out/bench_quick.err
