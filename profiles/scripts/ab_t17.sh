mkdir -p gpurun_out/t17
run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --skip_cpu_baseline --subcycled_steps 0 --fp32_steps 0 --poisson_full_max_it 0 > gpurun_out/t17/$name.json 2> gpurun_out/t17/$name.err; echo "$name rc=$?"; }
for v in c208 c176; do run $v PICGPU_SO=engineering-degree-in-plasma-simulations_b200/libpicgpu_$v.so; done
run lg4 PICG_CELL_LG=4
