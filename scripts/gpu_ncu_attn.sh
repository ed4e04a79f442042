#!/bin/bash
# ncu --set full (source-level stall sampling) of an ESM2 attention kernel inside an ESM2-650M encode
# (PCY_ESM_ATTN selects the kernel: 5 = two threads per row, 6 = one thread per row)
mkdir -p gpurun_out
K=${PCY_ESM_ATTN:-6}
PCY_ESM_ATTN=$K timeout 600 ncu --set full --clock-control none --import-source on -k regex:esm_attention -s 2 -c 1 \
  -o gpurun_out/prof_esm_attn_k$K -f python scripts/profile_esm_kernels.py > gpurun_out/ncu_attn.log 2>&1; echo "attn rc=$?"
tail -n 3 gpurun_out/ncu_attn.log
