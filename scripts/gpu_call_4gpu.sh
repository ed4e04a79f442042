mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench4 rc=$?"
tail -3 gpurun_out/bench_4gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_4gpu.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"])
for k in ("esm2_encode", "esm2_encode_8192", "it_forward_loss"):
    print(k, json.dumps(d.get(k))[:400])
print("query", json.dumps(d["retrieval"]["query"])[:500])
PY
