#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, the reference arm, the launch list and the
# per-kernel ncu captures.  Outputs land in gpurun_out/ with the given tag.
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1
tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json | cut -c1-600
if [ "$2" = "full" ]; then
  # NOTE: the full 40-level launch list takes >25 min under ncu; scripts/gpu_profiles.sh profiles the two finest levels instead
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/${TAG}_plain_launch.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 10400 -c 10500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1
  for st in sweep phi_ksi median; do
    python scripts/run_stage.py $st --size 512 --reps 3 > $O/${TAG}_plain_$st.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel|phi_ksi_kernel|median" -c 1 -f \
      -o $O/${TAG}_prof_$st python scripts/run_stage.py $st --size 512 --reps 1 > $O/${TAG}_ncu_$st.log 2>&1
    cat $O/${TAG}_plain_$st.log
  done
fi
