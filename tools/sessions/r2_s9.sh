#!/bin/bash
# Round 2, session 9 (1 GPU): why are the shared-memory weight modes of gram2 slow?  ncu --set full with source
# counters of the warp-specialised mode (bit 1024) next to the register mode, at 2M x 500 (short kernels).
set -u
mkdir -p gpurun_out
O=gpurun_out
for OPT in 1040 16; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram2_dmma --launch-skip 2 -c 1 \
      -o $O/r02_prof_gram2_opt$OPT -f python tools/bench_kernels.py --shapes 4000000x1000 --no-lincomb --opts $OPT --reps 1 > $O/r02_s9_ncu_$OPT.log 2>&1
  ncu -i $O/r02_prof_gram2_opt$OPT.ncu-rep --page details > $O/r02_s9_details_$OPT.txt 2>&1
  ncu -i $O/r02_prof_gram2_opt$OPT.ncu-rep --page source --csv > $O/r02_s9_source_$OPT.csv 2>&1
  grep -E "Duration|No Eligible|Issue Slots Busy|Warp Cycles Per Issued|Stall" $O/r02_s9_details_$OPT.txt | cut -c1-150
done
