set -x
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python tools/tma_probe.py > gpurun_out/r2s_sanitizer.log 2>&1; grep -v "^=========     at\|^=========         Host" gpurun_out/r2s_sanitizer.log | head -40
