set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "primitives or three_octaves or bench_path" > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
for T in 0 1; do
if [ $T = 1 ]; then export VPL_NO_TMA=1; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'blur5_sobel|scale08|ll_angle|order_kernel' -c 40 --csv --log-file gpurun_out/r2r_launches_notma$T.csv python bench.py --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2r_ncu$T.log 2>&1
done
