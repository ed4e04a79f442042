#!/bin/bash
# full validation + evidence for the round: parity tests, smoke, bench (with CPU baseline), reference arm, ESM
# breakdown, decode phase stamps, ncu captures of the new ESM kernels, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python scripts/bench_hf_gpu_baseline.py --steps 2 --warmup 1 > gpurun_out/hf_gpu_baseline.json 2> gpurun_out/hf_gpu_baseline.err; echo "hf gpu baseline rc=$?"; cat gpurun_out/hf_gpu_baseline.json
timeout 300 python scripts/bench_retrieval.py > gpurun_out/retrieval.log 2>&1; echo "retrieval rc=$?"; tail -3 gpurun_out/retrieval.log
timeout 300 python scripts/profile_esm_breakdown.py > gpurun_out/esm_breakdown.log 2>&1; echo "breakdown rc=$?"
timeout 300 python scripts/profile_decode_phases.py > gpurun_out/decode_phases.log 2>&1
timeout 300 python scripts/bench_gemm_shapes.py > gpurun_out/gemm_shapes_pair.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:esm_attention_ts -s 3 -c 1 -o gpurun_out/prof_attn_ts -f python scripts/profile_esm_kernels.py > gpurun_out/ncu_attn_ts.log 2>&1; echo "ncu attn rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 8 -c 4 -o gpurun_out/prof_gemm_pair -f python scripts/profile_esm_kernels.py > gpurun_out/ncu_gemm_pair.log 2>&1; echo "ncu gemm rc=$?"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['phases'],d['esm2_encode'],d['decode_beam10'])"
