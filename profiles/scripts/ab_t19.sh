mkdir -p gpurun_out/t19
run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --skip_cpu_baseline --subcycled_steps 0 --fp32_steps 0 --poisson_full_max_it 0 > gpurun_out/t19/$name.json 2> gpurun_out/t19/$name.err; echo "$name rc=$?"; }
run lg3 PICG_CELL_LG=3
run lg2 PICG_CELL_LG=2
run base X=1
