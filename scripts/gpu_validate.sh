#!/bin/bash
# full validation + evidence for the round: parity tests, bench (with CPU baseline), reference arm, smoke, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python scripts/profile_decode_phases.py > gpurun_out/decode_phases.log 2>&1
timeout 600 python scripts/bench_decode_rows.py 1,2,3,4,5,8,10,16 > gpurun_out/decode_rows.log 2>&1
cat gpurun_out/decode_rows.log | tail -8
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_beam10.csv python scripts/profile_paths.py --what decode --decode-steps 2 --beams 10 > /dev/null 2>&1
