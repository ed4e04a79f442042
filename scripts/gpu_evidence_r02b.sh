#!/bin/bash
# round-2 evidence for the persistent beam-search decode kernel: ncu launch list of 10-beam decode steps and one
# ncu --set full capture of the kernel (summaries are copied to profiles/ afterwards)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_beam10.csv \
  python scripts/profile_paths.py --what decode --decode-steps 3 --beams 10 > gpurun_out/ncu_beam.log 2>&1; echo "beam launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows_megakernel -s 1 -c 1 -o gpurun_out/prof_decode_rows_r02 -f \
  python scripts/profile_paths.py --what decode --decode-steps 3 --beams 10 > gpurun_out/ncu_rows.log 2>&1; echo "ncu rows kernel rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_beam10.csv
