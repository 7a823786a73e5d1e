set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "engines or lsd_segments or bench_path" > gpurun_out/r2o_pytest.log 2>&1; tail -3 gpurun_out/r2o_pytest.log
VPL_ENGINE_VARIANT=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "engines or lsd_segments or bench_path" > gpurun_out/r2o_pytest_v1.log 2>&1; tail -3 gpurun_out/r2o_pytest_v1.log
for V in 0 1; do for B in 4096 6144; do
VPL_ENGINE_VARIANT=$V timeout 400 python bench.py --batch $B --steps 4 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2o_bench_v${V}_b$B.json 2> gpurun_out/r2o_bench_v${V}_b$B.err; python tools/bench_summary.py gpurun_out/r2o_bench_v${V}_b$B.json 2>/dev/null | head -2; tail -2 gpurun_out/r2o_bench_v${V}_b$B.err
done; done
