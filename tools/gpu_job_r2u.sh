set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "engines or lsd_segments or bench_path or c4" > gpurun_out/r2u_pytest.log 2>&1; tail -3 gpurun_out/r2u_pytest.log
for V in 1 0 4; do
VPL_ENGINE_VARIANT=$V timeout 400 python bench.py --batch 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2u_bench_v${V}.json 2> gpurun_out/r2u_bench_v${V}.err; python tools/bench_summary.py gpurun_out/r2u_bench_v${V}.json 2>/dev/null | head -2; tail -2 gpurun_out/r2u_bench_v${V}.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'region_engine' -c 6 --csv --log-file gpurun_out/r2u_launches_engine.csv python bench.py --batch 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2u_ncu.log 2>&1
