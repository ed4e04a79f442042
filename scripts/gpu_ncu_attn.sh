#!/bin/bash
# ncu --set full (source-level stall sampling) of the default ESM2 attention kernel inside an ESM2-650M encode
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:esm_attention_ts -s 2 -c 1 \
  -o gpurun_out/prof_esm_attn_r02b -f python scripts/profile_esm_kernels.py > gpurun_out/ncu_attn.log 2>&1; echo "attn rc=$?"
tail -n 3 gpurun_out/ncu_attn.log
