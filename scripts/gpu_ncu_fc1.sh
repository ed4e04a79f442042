#!/bin/bash
# ncu --set full (source-level stall sampling) of the fc1 GEMM (bias + GELU epilogue) of layer 1 of an ESM2-650M
# encode, with LayerNorm folded into it and without
mkdir -p gpurun_out
PCY_LN_FOLD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 6 -c 1 \
  -o gpurun_out/prof_fc1_fold -f python scripts/profile_esm_kernels.py > gpurun_out/ncu_fc1_fold.log 2>&1; echo "fold rc=$?"
PCY_LN_FOLD=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 6 -c 1 \
  -o gpurun_out/prof_fc1_nofold -f python scripts/profile_esm_kernels.py > gpurun_out/ncu_fc1_nofold.log 2>&1; echo "nofold rc=$?"
tail -3 gpurun_out/ncu_fc1_fold.log gpurun_out/ncu_fc1_nofold.log
