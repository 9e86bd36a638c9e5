mkdir -p gpurun_out/t6
run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --skip_cpu_baseline --subcycled_steps 0 --fp32_steps 0 --poisson_full_max_it 0 > gpurun_out/t6/$name.json 2> gpurun_out/t6/$name.err; echo "$name rc=$?"; }
run base X=1
run w128 PICGPU_SO=engineering-degree-in-plasma-simulations_b200/libpicgpu_w128.so
run m5 PICGPU_SO=engineering-degree-in-plasma-simulations_b200/libpicgpu_m5.so
