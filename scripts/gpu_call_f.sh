#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --launch-timeout 120 python -m pytest tests/test_gpu_llama.py -m gpu -q -x -k "persistent_and_per_op and gq4wide-2" -p no:cacheprovider > gpurun_out/san_dyn.log 2>&1; grep -E "Invalid|at .*decode_megakernel|by thread|Address|ERROR SUMMARY|passed|failed|timeout|trap" gpurun_out/san_dyn.log | head -30
for cfg in "12 32768" "14 32768" "15 32768" "14 65536" "15 65536" "16 32768"; do
  set -- $cfg
  echo "== static16=$1 block=$2"
  PCY_DYN=1 PCY_DYN_STATIC16=$1 PCY_DYN_BLOCK=$2 timeout 200 python scripts/profile_decode_phases.py 2>&1 | grep -E "stream|barrier|total" | grep -v "^P2 "
done
