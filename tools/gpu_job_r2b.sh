# round 2, job B: the speculative region engine -- parity first, then a short bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2b_pytest_parity.log 2>&1; tail -15 gpurun_out/r2b_pytest_parity.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; python tools/bench_summary.py gpurun_out/r2b_bench.json | head -8; tail -3 gpurun_out/r2b_bench.err
