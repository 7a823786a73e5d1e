set -x
mkdir -p gpurun_out
timeout 300 python bench.py --batch 512 --steps 4 --warmup 2 --no-cpu-baseline --no-parity > gpurun_out/r2m_bench_b512.json 2>gpurun_out/r2m_bench_b512.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches_b512.csv python bench.py --batch 512 --steps 4 --warmup 2 --no-cpu-baseline --no-parity > gpurun_out/r2m_ncu.log 2>&1
timeout 300 python bench.py --workload E2 --batch 512 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2m_bench_E2_b512.json 2>gpurun_out/r2m_bench_E2_b512.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches_E2_b512.csv python bench.py --workload E2 --batch 512 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2m_ncu2.log 2>&1
ls -la gpurun_out/r2m*
