#!/bin/bash
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --print-limit 6 python scripts/debug_mega.py 1 mega > gpurun_out/dbg_san.log 2>&1
grep -E "=========" gpurun_out/dbg_san.log | head -60
