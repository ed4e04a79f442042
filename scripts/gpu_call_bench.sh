#!/bin/bash
# full bench line (N = 1) + reference arm; stdout/stderr kept under gpurun_out/
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench.json'))
except Exception as e:
    print("no bench line:", e); raise SystemExit
pr = lambda k: print(k, json.dumps(d.get(k)))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
for k in ("roofline", "phases", "decode_beam10", "e2e_beam10", "esm2_encode", "esm2_encode_8192", "retrieval",
          "it_forward_loss", "gpu_reference", "vs_gpu_reference", "cpu_baseline", "clocks"):
    pr(k)
PY
