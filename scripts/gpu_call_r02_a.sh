#!/bin/bash
# round 2, call A: the north-star denominator (HF eager on the GPU) and the config-4 timing, both written in round 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python scripts/bench_hf_gpu_baseline.py --steps 2 --warmup 1 > gpurun_out/hf_gpu_baseline.json 2> gpurun_out/hf_gpu_baseline.err; echo "hf gpu baseline rc=$?"; cat gpurun_out/hf_gpu_baseline.json; tail -3 gpurun_out/hf_gpu_baseline.err
timeout 400 python scripts/bench_retrieval.py > gpurun_out/retrieval.log 2>&1; echo "retrieval rc=$?"; tail -5 gpurun_out/retrieval.log
