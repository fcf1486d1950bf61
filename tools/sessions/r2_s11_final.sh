#!/bin/bash
# Round 2, last session (1 GPU): smoke() and the -m gpu suite on the final tree.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python __graft_entry__.py smoke > $O/r02_s11_smoke.log 2>&1; echo "smoke exit $?" >> $O/r02_s11_smoke.log; tail -n 2 $O/r02_s11_smoke.log | cut -c1-200
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method thread -p no:cacheprovider > $O/r02_s11_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r02_s11_pytest_gpu.log
tail -n 4 $O/r02_s11_pytest_gpu.log | cut -c1-300
