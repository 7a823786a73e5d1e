set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2aj_pytest.log 2>&1; tail -3 gpurun_out/r2aj_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2aj_bench.log 2>&1; tail -1 gpurun_out/r2aj_bench.log | cut -c1-900
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 6 --e2e-only --e2e-trace > gpurun_out/r2aj_trace.log 2>&1; tail -1 gpurun_out/r2aj_trace.log
timeout 300 python bench.py --gpus 1 --steps 4 --warmup 3 --timeline > gpurun_out/r2aj_tl.log 2>&1; tail -1 gpurun_out/r2aj_tl.log
