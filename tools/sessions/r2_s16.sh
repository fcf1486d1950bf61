#!/bin/bash
# Round 2, session 16 (1 GPU): ncu launch list of the bench command (10M x 1000), after the same command ran plain.
set -u
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-gs --no-csr-pod"
timeout 300 $CMD > $O/r02_s16_plain.log 2>&1; echo "plain exit $?" >> $O/r02_s16_plain.log; tail -n 1 $O/r02_s16_plain.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_pod_10Mx1000.csv $CMD > $O/r02_s16_ncu.log 2>&1
echo "ncu exit $?"; grep -c '^"[0-9]' $O/r02_launches_pod_10Mx1000.csv
