mkdir -p gpurun_out/t15
run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --skip_cpu_baseline --subcycled_steps 0 --fp32_steps 0 --poisson_full_max_it 0 > gpurun_out/t15/$name.json 2> gpurun_out/t15/$name.err; echo "$name rc=$?"; }
run base X=1
for v in e25 e26 h24 h25e26; do run $v PICGPU_SO=engineering-degree-in-plasma-simulations_b200/libpicgpu_$v.so; done
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
