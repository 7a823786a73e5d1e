# GPU job: the fused readImage pipeline: parity, then the R1 / R1s benches
set -x
timeout 300 python -m pytest tests/test_gpu_vp.py -x -q -m gpu -k "readimage" > gpurun_out/pytest_r1.log 2>&1; tail -12 gpurun_out/pytest_r1.log
timeout 400 python bench.py --workload R1 > gpurun_out/bench_R1.json 2> gpurun_out/bench_R1.err; python tools/bench_summary.py gpurun_out/bench_R1.json 2>/dev/null | head -3; tail -4 gpurun_out/bench_R1.err
timeout 400 python bench.py --workload R1s > gpurun_out/bench_R1s.json 2> gpurun_out/bench_R1s.err; python tools/bench_summary.py gpurun_out/bench_R1s.json 2>/dev/null | head -3; tail -4 gpurun_out/bench_R1s.err
