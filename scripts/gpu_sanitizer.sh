#!/bin/bash
# compute-sanitizer memcheck + racecheck over the kernels with hand-rolled synchronisation: the persistent decode
# kernel (cooperative grid barrier, single-consumer mbarrier ring), the tcgen05 GEMMs (pair MMA), the tcgen05 ESM
# attention kernels, the fused retrieval ranking (ticket), the beam selection.  Logs -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
SEL='persistent or pair_mma or tcgen05 or fused_topk or selection_is_bit_exact or lockstep'
FILES="tests/test_gpu_llama.py tests/test_gpu_pair_mma.py tests/test_gpu_esm.py tests/test_gpu_retrieval.py tests/test_gpu_beam_strict.py tests/test_gpu_kernels.py"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 \
    python -m pytest $FILES -m gpu -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -4
done
