#!/bin/bash
# ESM iteration: parity tests of the encoder and of the GEMM epilogues (own process + timeout: a hang must not take the
# rest of the call with it), the full-size ESM2-650M parity tests, then the in-situ breakdown of the batch encode
# (LayerNorm folded / passes)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_esm.py -m gpu -q > gpurun_out/pytest_esm.log 2>&1; rc=$?; echo "pytest esm rc=$rc"
tail -15 gpurun_out/pytest_esm.log
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pair_mma.py -m gpu -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"
tail -8 gpurun_out/pytest_kernels.log
timeout 400 python -m pytest tests/test_gpu_fullshape.py -m gpu -q -x -k esm2 > gpurun_out/pytest_esm_full.log 2>&1; echo "pytest esm full rc=$?"
tail -5 gpurun_out/pytest_esm_full.log
timeout 300 python scripts/profile_esm_breakdown.py > gpurun_out/esm_breakdown.log 2>&1; echo "breakdown rc=$?"
tail -8 gpurun_out/esm_breakdown.log
