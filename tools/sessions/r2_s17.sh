#!/bin/bash
# Round 2, session 17 (1 GPU): triangular plan on the diagonal tiles of lower-only general products (symmetric CSR
# two-forms): parity tests, config 5 per-GPU shard before/after (bit 4096 switches it off).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 --timeout-method thread -p no:cacheprovider -k "gram2 or csr or golden" > $O/r02_s17_pytest.log 2>&1; echo "pytest exit $?" >> $O/r02_s17_pytest.log
tail -n 3 $O/r02_s17_pytest.log | cut -c1-200
timeout 300 python tools/bench_project.py --nx 2500 --ny 2500 --k 2000 --reps 2 > $O/r02_s17_project_tri.log 2>&1; grep -hE '^\{' $O/r02_s17_project_tri.log | cut -c1-700
PMC_GRAM_OPTS=4112 timeout 300 python tools/bench_project.py --nx 2500 --ny 2500 --k 2000 --reps 2 > $O/r02_s17_project_notri.log 2>&1; grep -hE '^\{' $O/r02_s17_project_notri.log | cut -c1-400
