set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:region_engine -s 2 -c 1 -o gpurun_out/r2g_engine_b4096 -f python bench.py --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2g_ncu.log 2>&1; tail -2 gpurun_out/r2g_ncu.log
