#!/usr/bin/env bash
# ncu --set full of the non-solver kernels at 512^3 (kernel-name filters: run_stage.py's torch set-up kernels are skipped)
set -u
O=gpurun_out
for pair in "warp:warp_derivatives" "resample:resample" "blur:conv_axis" "median:median5"; do
  st=${pair%%:*}; k=${pair##*:}
  python scripts/run_stage.py $st --reps 5 > $O/r2s_plain_$st.txt 2>&1; cat $O/r2s_plain_$st.txt
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 3 -f -o $O/r2s_prof_$st \
     python scripts/run_stage.py $st --reps 2 > $O/r2s_ncu_$st.log 2>&1
  ls -la $O/r2s_prof_$st.ncu-rep
done
