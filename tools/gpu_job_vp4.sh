# GPU job: parity (whole suite) + V1 / C2 benches after a change of the shared double-double functions
set -x
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all4.log 2>&1; tail -3 gpurun_out/pytest_gpu_all4.log
timeout 300 python bench.py --workload V1 --no-cpu-baseline > gpurun_out/bench_V1_v4.json 2> gpurun_out/bench_V1_v4.err; python tools/bench_summary.py gpurun_out/bench_V1_v4.json; tail -3 gpurun_out/bench_V1_v4.err
timeout 200 python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_V1_b512_v4.json 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_V1_b512_v4.csv python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1.log 2>&1
python tools/launch_summary.py gpurun_out/launches_V1_b512_v4.csv
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_C2_v4.json 2> gpurun_out/bench_C2_v4.err; python tools/bench_summary.py gpurun_out/bench_C2_v4.json
