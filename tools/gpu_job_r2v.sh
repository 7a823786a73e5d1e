set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2v_pytest.log 2>&1; tail -3 gpurun_out/r2v_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ll_angle|order_kernel|rect_nfa' -c 30 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2v_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ll_angle|order_kernel|rect_nfa' -c 12 --csv --log-file gpurun_out/r2v_launches_4096.csv python bench.py --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2v_ncu2.log 2>&1
