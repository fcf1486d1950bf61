#!/bin/bash
# Round 2, session 13 (1 GPU): 64 x 16 warp tiles for register-scaled weights (default on full general tiles; bit 2048
# switches them off) -- A/B and the -m gpu suite (parity of all modes, bitwise against gram_dmma_kernel).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/bench_kernels.py --shapes 10000000x1000,2000000x500 --no-lincomb --opts 16,2064 > $O/r02_s13_kernels_wide.jsonl 2>&1
grep -E '^\{' $O/r02_s13_kernels_wide.jsonl | cut -c1-420
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method thread -p no:cacheprovider > $O/r02_s13_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r02_s13_pytest_gpu.log
tail -n 4 $O/r02_s13_pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > $O/r02_s13_smoke.log 2>&1; echo "smoke exit $?" >> $O/r02_s13_smoke.log; tail -n 2 $O/r02_s13_smoke.log | cut -c1-200
