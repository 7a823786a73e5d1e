# GPU job: vanishing-point parity (incl. every hypothesis' score) + V1 bench
set -x
timeout 300 python -m pytest tests/test_gpu_vp.py -x -q -m gpu > gpurun_out/pytest_vp7.log 2>&1; tail -5 gpurun_out/pytest_vp7.log
timeout 300 python bench.py --workload V1 --no-cpu-baseline > gpurun_out/bench_V1_v7.json 2> gpurun_out/bench_V1_v7.err; python tools/bench_summary.py gpurun_out/bench_V1_v7.json 2>/dev/null | head -3; tail -3 gpurun_out/bench_V1_v7.err
