#!/bin/bash
# run a pytest selection on the GPU box: bash scripts/gpu_call_tests.sh <log name> <pytest args...>
mkdir -p gpurun_out
name=$1; shift
timeout 1500 python -m pytest "$@" -m gpu -q --durations=8 > gpurun_out/$name.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/$name.log
tail -40 gpurun_out/$name.log
