#!/usr/bin/env bash
# round-2 GPU session B: ncu --set full of the TMA sweep (512^3), and of warp+derivatives / resample / blur
set -u
O=gpurun_out
python scripts/run_stage.py sweep --variant 1 --nchunks 8 --reps 10 > $O/r2b_tma_plain.txt 2>&1
python scripts/run_stage.py sweep --variant 0 --nchunks 10 --reps 10 >> $O/r2b_tma_plain.txt 2>&1
cat $O/r2b_tma_plain.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_tma -s 1 -c 1 -f -o $O/r2b_prof_sweep_tma \
   python scripts/run_stage.py sweep --variant 1 --nchunks 8 --reps 2 > $O/r2b_ncu_sweep_tma.log 2>&1
for st in warp resample blur; do
  python scripts/run_stage.py $st --reps 5 >> $O/r2b_tma_plain.txt 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -s 1 -c 4 -f -o $O/r2b_prof_$st \
     python scripts/run_stage.py $st --reps 1 > $O/r2b_ncu_$st.log 2>&1
done
tail -4 $O/r2b_tma_plain.txt
ls -la $O/r2b_prof_*.ncu-rep
