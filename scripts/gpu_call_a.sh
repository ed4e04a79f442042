#!/bin/bash
# One GPU call: parity tests, bench line, ncu launch list, ncu --set full of the dominant kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
timeout 600 python scripts/profile_decode_phases.py > gpurun_out/decode_phases.log 2>&1
tail -30 gpurun_out/decode_phases.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_paths.csv python scripts/profile_paths.py --proteins 64 --decode-steps 2 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 40 -c 3 -o gpurun_out/prof_gemm -f python scripts/profile_paths.py --what esm --proteins 64 > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn -s 10 -c 2 -o gpurun_out/prof_attn -f python scripts/profile_paths.py --what esm --proteins 64 > gpurun_out/ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -c 2 -o gpurun_out/prof_decode -f python scripts/profile_paths.py --what decode --decode-steps 2 > gpurun_out/ncu_decode.log 2>&1
ls -la gpurun_out
