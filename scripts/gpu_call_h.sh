#!/bin/bash
# full validation + evidence for the round: parity tests, bench (with CPU baseline), reference arm, launch list, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
timeout 600 python scripts/profile_decode_phases.py > gpurun_out/decode_phases.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -c 1 -o gpurun_out/prof_decode3 -f python scripts/profile_paths.py --what decode --decode-steps 2 > gpurun_out/ncu_decode3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:esm_attention -s 4 -c 1 -o gpurun_out/prof_esm_attn -f python scripts/profile_paths.py --what esm --proteins 64 > gpurun_out/ncu_esm_attn.log 2>&1
ls -la gpurun_out | tail -20
