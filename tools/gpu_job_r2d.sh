# round 2, job D: compact layout + both region engines -- parity, then benches at several batch sizes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2d_pytest_parity.log 2>&1; tail -15 gpurun_out/r2d_pytest_parity.log
for B in 64 512 4096; do
timeout 300 python bench.py --batch $B --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2d_bench_b$B.json 2> gpurun_out/r2d_bench_b$B.err; python tools/bench_summary.py gpurun_out/r2d_bench_b$B.json | head -3; tail -2 gpurun_out/r2d_bench_b$B.err
done
