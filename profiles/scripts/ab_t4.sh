mkdir -p gpurun_out/t4
run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --skip_cpu_baseline --subcycled_steps 0 --fp32_steps 0 --poisson_full_max_it 0 > gpurun_out/t4/$name.json 2> gpurun_out/t4/$name.err; echo "$name rc=$?"; }
run A PICG_MERGE_FRACTION=0.05 PICG_MOVER_FRACTION=0.10
run B PICG_MERGE_FRACTION=0.08 PICG_MOVER_FRACTION=0.10
run C PICG_MERGE_FRACTION=0.12 PICG_MOVER_FRACTION=0.15
run D PICG_MERGE_FRACTION=0.05 PICG_MOVER_FRACTION=0.20
run E PICG_MERGE_FRACTION=0.03 PICG_MOVER_FRACTION=0.10
