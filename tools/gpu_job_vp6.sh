# GPU job: whole GPU suite (shared sincos changed again) + V1 / C2 benches + V1 launch list
set -x
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all6.log 2>&1; tail -3 gpurun_out/pytest_gpu_all6.log
timeout 300 python bench.py --workload V1 --no-cpu-baseline > gpurun_out/bench_V1_v6.json 2> gpurun_out/bench_V1_v6.err; python tools/bench_summary.py gpurun_out/bench_V1_v6.json | head -3; tail -3 gpurun_out/bench_V1_v6.err
timeout 200 python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_V1_b512_v6.json 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_V1_b512_v6.csv python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1.log 2>&1
python tools/launch_summary.py gpurun_out/launches_V1_b512_v6.csv
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_C2_v6.json 2> gpurun_out/bench_C2_v6.err; python tools/bench_summary.py gpurun_out/bench_C2_v6.json | head -3
