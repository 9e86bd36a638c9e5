#!/bin/bash
# Profiling recipe (B200_PROFILING.md), run on the GPU box through gpurun.  Outputs land in gpurun_out/<tag>/.
#   profiles/capture.sh <tag> <mesh> <particles>
TAG=${1:-r1}; MESH=${2:-256}; NP=${3:-1e9}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --mesh $MESH --particles $NP --steps 2 --warmup 2 --skip_cpu_baseline --init_max_it 60 --subcycled_steps 0 --fp32_steps 0 --poisson_full_max_it 0"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches.csv $CMD > $OUT/launches.log 2>&1
python profiles/launch_list.py $OUT/launches.csv "$CMD" > $OUT/launches.md
# (2) full-set capture of the hot kernels, a few launches each, taken from the timed region (skip set-up + warm-up launches)
for K in k_run:16:6 k_cell_deposit:6:3 k_sor_row:300:2 'k_mcc<':2:1 k_sort_permute:14:2; do
  IFS=: read NAME SKIP COUNT <<< "$K"
  ncu --set full --clock-control none --import-source on -k regex:"^$NAME" -s $SKIP -c $COUNT -o $OUT/$NAME -f $CMD > $OUT/$NAME.log 2>&1
  python profiles/ncu_summary.py $OUT/$NAME.ncu-rep 0 > $OUT/$NAME.summary.txt 2>&1
done
python profiles/ncu_traffic.py $TAG $OUT/k_run.ncu-rep $OUT/k_cell_deposit.ncu-rep $OUT/k_sor_row.ncu-rep $OUT/k_mcc.ncu-rep $OUT/k_sort_permute.ncu-rep > $OUT/traffic.log 2>&1
cp profiles/ncu_traffic_$TAG.json $OUT/
ls -la $OUT
