#!/bin/bash
# Profiling recipe (B200_PROFILING.md), run on the GPU box through gpurun.  Outputs land in gpurun_out/.
#   profiles/capture.sh <tag> <mesh> <particles>
TAG=${1:-r1}; MESH=${2:-128}; NP=${3:-1.2e8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --mesh $MESH --particles $NP --steps 2 --warmup 1 --skip_cpu_baseline --init_max_it 60"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv $CMD > $OUT/launches.log 2>&1
# (2) full-set capture of the hot kernels, one launch each, taken from the timed region (skip set-up + warm-up launches)
for K in k_run k_sor_row k_find_movers k_mcc; do
  ncu --set full --clock-control none --import-source on -k regex:"^$K" -s 12 -c 6 -o $OUT/$K -f $CMD > $OUT/$K.log 2>&1
done
ls -la $OUT
