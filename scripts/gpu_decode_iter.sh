#!/bin/bash
# decode megakernel iteration: parity tests for the Llama paths, phase stamps, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_llama.py tests/test_gpu_unified.py -m gpu -q -x > gpurun_out/pytest_llama.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_llama.log
tail -3 gpurun_out/pytest_llama.log
timeout 600 python scripts/profile_decode_phases.py > gpurun_out/decode_phases.log 2>&1
tail -30 gpurun_out/decode_phases.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['ms_per_launch'],d['phases'])"; tail -5 gpurun_out/bench.err
