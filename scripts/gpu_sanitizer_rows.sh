#!/bin/bash
# compute-sanitizer memcheck + racecheck over the persistent beam-search decode kernel (decode_rows_megakernel.cu):
# self-refilled TMA ring, shared-memory pool, ticket merges, attention tiles.  Logs -> gpurun_out/sanitizer_rows_*.log
mkdir -p gpurun_out
SEL='teacher_forced and (gq4-3 or gq4-9 or gq4wide-10) or identical_state and (1-10-300 or 3-4-100)'
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 \
    python -m pytest tests/test_gpu_decode_rows.py -m gpu -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_rows_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitizer_rows_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_rows_$tool.log | tail -4
done
grep -E "Race reported|Invalid|and (Read|Write) access" gpurun_out/sanitizer_rows_racecheck.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -30
