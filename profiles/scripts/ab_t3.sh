mkdir -p gpurun_out/t3
run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 5 --skip_cpu_baseline --subcycled_steps 0 --fp32_steps 0 --poisson_full_max_it 0 > gpurun_out/t3/$name.json 2> gpurun_out/t3/$name.err; echo "$name rc=$?"; }
run A PICG_L2_SCOPE=0 PICG_MCC_BALANCE=0
run B PICG_L2_SCOPE=1 PICG_MCC_BALANCE=0
run C PICG_L2_SCOPE=1 PICG_MCC_BALANCE=1
run D PICG_L2_SCOPE=0 PICG_MCC_BALANCE=1 PICG_L2_FETCH=32
run E PICG_L2_SCOPE=0 PICG_MCC_BALANCE=0 PICG_L2_FETCH=64
python -m pytest tests/test_facade.py -m gpu -x -q 2>&1 | tail -3
