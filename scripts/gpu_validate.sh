#!/bin/bash
# full validation of the round: parity tests, smoke, bench (with CPU baseline + HF-eager GPU reference), reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "launches", d["gpu_launches"])
for k in ("roofline", "phases", "decode_beam10", "e2e_beam10", "esm2_encode", "esm2_encode_8192", "retrieval",
          "it_forward_loss", "gpu_reference", "vs_gpu_reference", "cpu_baseline", "clocks"):
    print(k, json.dumps(d.get(k)))
print("reference arm:", open('gpurun_out/bench_ref.json').read()[:400])
PY
