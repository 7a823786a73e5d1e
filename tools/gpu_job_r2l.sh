set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_vp.py -x -q > gpurun_out/r2l_pytest.log 2>&1; tail -15 gpurun_out/r2l_pytest.log
