#!/bin/bash
# 2 GPUs: NCCL test of the gathered InfoNCE forward, then the bench line at N = 2 (sharded paths under NCCL)
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multirank.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_multirank_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
tail -4 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_2gpu.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"])
for k in ("esm2_encode", "esm2_encode_8192", "retrieval", "it_forward_loss"):
    print(k, json.dumps(d.get(k))[:700])
PY
