#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the two finest levels of the 512^3 solve (every kernel of a
# level appears; the full 40-level list takes >25 min under ncu), (2) --set full captures of the solver kernels
# inside that same run, (3) the median and resample kernels from the stage runner.
O=gpurun_out; T=${1:-r02}
CMD="python bench.py --steps 1 --warmup 1 --levels 2 --no-e2e --no-cpu-baseline"
$CMD > $O/${T}_plain_levels2.json 2> $O/${T}_plain_levels2.err || exit 1
cut -c1-200 $O/${T}_plain_levels2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_levels2.csv $CMD > $O/${T}_ncu_launch.log 2>&1
wc -l $O/${T}_launches_levels2.csv
for k in "sweep_kernel" "phi_ksi_kernel" "median" "resample_axis_vec4" "warp_derivatives"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 60 -c 2 -f -o $O/${T}_prof_$k \
     env FLOW3D_AUTOTUNE=0 $CMD > $O/${T}_ncu_$k.log 2>&1
  ls -la $O/${T}_prof_$k.ncu-rep
done
