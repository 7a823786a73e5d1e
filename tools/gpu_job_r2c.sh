# round 2, job C: ncu of the speculative region engine (after the plain run exits 0)
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2c_bench_b512.json 2> gpurun_out/r2c_bench_b512.err && python tools/bench_summary.py gpurun_out/r2c_bench_b512.json | head -4
timeout 300 python bench.py --batch 64 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2c_bench_b64.json 2> gpurun_out/r2c_bench_b64.err && python tools/bench_summary.py gpurun_out/r2c_bench_b64.json | head -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:region_engine -s 1 -c 1 -o gpurun_out/r2c_engine_b512 -f python bench.py --batch 512 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2c_ncu.log 2>&1; tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/*.ncu-rep
