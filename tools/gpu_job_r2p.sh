set -x
mkdir -p gpurun_out
for V in 0 1 2 3 4; do for B in 4096 8192; do
VPL_ENGINE_VARIANT=$V timeout 400 python bench.py --batch $B --steps 4 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2p_bench_v${V}_b$B.json 2> gpurun_out/r2p_bench_v${V}_b$B.err; python tools/bench_summary.py gpurun_out/r2p_bench_v${V}_b$B.json 2>/dev/null | head -2; tail -2 gpurun_out/r2p_bench_v${V}_b$B.err
done; done
