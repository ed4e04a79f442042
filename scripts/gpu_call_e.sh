#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_llama.py tests/test_gpu_fullshape.py -m gpu -q -x -k "dynamic or persistent or llama8b or greedy" 2>&1 | tail -15
for dyn in 1 0; do
  PCY_DYN=$dyn timeout 300 python scripts/profile_decode_phases.py > gpurun_out/decode_phases_dyn$dyn.log 2>&1
  echo "== dynamic=$dyn"; tail -32 gpurun_out/decode_phases_dyn$dyn.log
done
