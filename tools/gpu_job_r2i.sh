set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2i_pytest_parity.log 2>&1; tail -5 gpurun_out/r2i_pytest_parity.log
for B in 64 4096; do
timeout 400 python bench.py --batch $B --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_b$B.json 2> gpurun_out/r2i_bench_b$B.err; python tools/bench_summary.py gpurun_out/r2i_bench_b$B.json 2>/dev/null | head -3; tail -2 gpurun_out/r2i_bench_b$B.err
done
