#!/bin/bash
mkdir -p gpurun_out
echo "== uncalibrated"; PCY_DECODE_CALIBRATE=0 timeout 600 python scripts/profile_decode_skew.py 2>&1 | grep -E "phase|spread|Error" | head -16
PCY_DECODE_CALIBRATE=0 timeout 600 python scripts/bench_decode_rows.py 1 2>&1 | tail -1
echo "== calibrated"; PCY_DECODE_CALIBRATE=1 timeout 600 python scripts/profile_decode_skew.py 2>&1 | grep -E "phase|spread|Error|error" | head -16
PCY_DECODE_CALIBRATE=1 timeout 600 python scripts/bench_decode_rows.py 1 2>&1 | tail -1
