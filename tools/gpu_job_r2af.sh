set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2af_pytest.log 2>&1; tail -3 gpurun_out/r2af_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2af_bench.log 2>&1; tail -1 gpurun_out/r2af_bench.log | cut -c1-1200
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-only --e2e-staggered > gpurun_out/r2af_stag.log 2>&1; tail -1 gpurun_out/r2af_stag.log | cut -c1-300
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-only --batch 4096 > gpurun_out/r2af_b4096.log 2>&1; tail -1 gpurun_out/r2af_b4096.log | cut -c1-300
