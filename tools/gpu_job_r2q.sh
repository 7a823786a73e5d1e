set -x
mkdir -p gpurun_out
timeout 600 python bench.py --batch 4096 --steps 2 --warmup 2 --no-cpu-baseline --no-parity > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:region_engine -s 2 -c 1 -o gpurun_out/r2q_engine_b4096 -f python bench.py --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2q_ncu.log 2>&1; tail -2 gpurun_out/r2q_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2q_launches_b4096.csv python bench.py --batch 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2q_ncu2.log 2>&1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2q_pytest_gpu.log
