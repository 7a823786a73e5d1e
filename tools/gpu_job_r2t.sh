set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edlines.py -x -q > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'blur5_sobel|scale08|ll_angle|order_kernel' -c 40 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2t_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ed_grad' -c 10 --csv --log-file gpurun_out/r2t_launches_E1.csv python bench.py --workload E1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2t_ncu2.log 2>&1
