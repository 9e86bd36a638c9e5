mkdir -p gpurun_out/t16
run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --skip_cpu_baseline --subcycled_steps 0 --fp32_steps 0 --poisson_full_max_it 0 > gpurun_out/t16/$name.json 2> gpurun_out/t16/$name.err; echo "$name rc=$?"; }
run base X=1
for v in c128b3 c96b4 c128b2; do run $v PICGPU_SO=engineering-degree-in-plasma-simulations_b200/libpicgpu_$v.so; done
