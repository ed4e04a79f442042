#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -c 1 -o gpurun_out/prof_decode2 -f python scripts/profile_paths.py --what decode --decode-steps 2 > gpurun_out/ncu_decode2.log 2>&1
tail -3 gpurun_out/ncu_decode2.log
