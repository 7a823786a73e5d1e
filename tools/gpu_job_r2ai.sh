set -x
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 6 --e2e-only --e2e-trace > gpurun_out/r2ai_trace.log 2>&1; tail -1 gpurun_out/r2ai_trace.log
