set -x
mkdir -p gpurun_out
for cfg in "4096 3" "2048 3" "2048 4"; do set -- $cfg
timeout 400 python bench.py --batch $1 --slots $2 --steps 6 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2k_bench_b$1_s$2.json 2> gpurun_out/r2k_bench_b$1_s$2.err; python tools/bench_summary.py gpurun_out/r2k_bench_b$1_s$2.json 2>/dev/null | head -1; tail -2 gpurun_out/r2k_bench_b$1_s$2.err
done
