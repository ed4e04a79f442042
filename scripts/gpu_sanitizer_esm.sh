#!/bin/bash
# compute-sanitizer memcheck + racecheck over the kernels the last session of round 2 changed: the tcgen05 GEMM epilogue
# (packed fp32 pairs) and the ESM2 attention kernels (kernel 5: O accumulated in TMEM, prologue masks, light path)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 \
    python -m pytest tests/test_gpu_esm.py tests/test_gpu_kernels.py tests/test_gpu_pair_mma.py -m gpu -q -x \
    -k "esm or linear_parity or tile_widths or pair_mma" -p no:cacheprovider > gpurun_out/sanitizer_esm_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitizer_esm_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_esm_$tool.log | tail -4
done
