set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "bench_path or upload_ahead or dense_collect" > gpurun_out/r2aa_pytest.log 2>&1; tail -3 gpurun_out/r2aa_pytest.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-only > gpurun_out/r2aa_e2e.log 2>&1; tail -1 gpurun_out/r2aa_e2e.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-only --stress-upload-passes 2 > gpurun_out/r2aa_e2e_p2.log 2>&1; tail -1 gpurun_out/r2aa_e2e_p2.log
