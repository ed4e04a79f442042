#!/bin/bash
# round-2 evidence: ncu launch lists (bench command; prefill + decode paths), ncu --set full captures of the dominant
# kernels (one launch each), all under gpurun_out/ (summaries are copied to profiles/ afterwards)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_paths.csv \
  python scripts/profile_paths.py --what prefill,decode --decode-steps 2 > gpurun_out/ncu_paths.log 2>&1; echo "paths rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_beam10.csv \
  python scripts/profile_paths.py --what decode --decode-steps 2 --beams 10 > gpurun_out/ncu_beam.log 2>&1; echo "beam rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -c 1 -o gpurun_out/prof_decode_r02 -f \
  python scripts/profile_paths.py --what decode --decode-steps 2 > gpurun_out/ncu_a.log 2>&1; echo "ncu decode rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:llama_attention_tc -s 4 -c 1 -o gpurun_out/prof_prefill_attn_r02 -f \
  python scripts/profile_paths.py --what prefill > gpurun_out/ncu_b.log 2>&1; echo "ncu prefill attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:retrieval_scores -s 8 -c 2 -o gpurun_out/prof_retrieval_r02 -f \
  python scripts/bench_retrieval.py > gpurun_out/ncu_c.log 2>&1; echo "ncu retrieval rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
