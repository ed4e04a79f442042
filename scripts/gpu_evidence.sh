#!/bin/bash
# ncu --set full captures of the dominant kernels of both decode paths (one launch each)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -c 1 -o gpurun_out/prof_decode_final -f python scripts/profile_paths.py --what decode --decode-steps 2 > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny_mma|shared_prompt" -s 6 -c 6 -o gpurun_out/prof_beam_final -f python scripts/profile_paths.py --what decode --decode-steps 1 --beams 10 > gpurun_out/ncu_b.log 2>&1
tail -n 2 gpurun_out/ncu_a.log; tail -n 2 gpurun_out/ncu_b.log
