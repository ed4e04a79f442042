#!/bin/bash
# ESM iteration: parity tests of the encoder (own process + timeout: a hang must not take the rest of the call with
# it), then the in-situ breakdown of the batch encode with the three attention kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_esm.py -m gpu -q -x > gpurun_out/pytest_esm.log 2>&1; rc=$?; echo "pytest esm rc=$rc"
tail -25 gpurun_out/pytest_esm.log
timeout 300 python scripts/profile_esm_breakdown.py > gpurun_out/esm_breakdown.log 2>&1; echo "breakdown rc=$?"
tail -8 gpurun_out/esm_breakdown.log
