#!/bin/bash
# full GPU validation on 2 GPUs: parity tests, 1-GPU bench with CPU baseline, reference arm, 2-GPU bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_2gpu.json') if l.startswith('{')][-1]);print(d['value'],d['n_gpus'],d['esm2_encode'])"; tail -3 gpurun_out/bench_2gpu.err
