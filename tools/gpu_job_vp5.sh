# GPU job: vanishing-point parity + bench + launch list after a kernel change
set -x
timeout 300 python -m pytest tests/test_gpu_vp.py tests/test_cpp_facade.py -x -q -m gpu > gpurun_out/pytest_vp5.log 2>&1; tail -3 gpurun_out/pytest_vp5.log
timeout 300 python bench.py --workload V1 --no-cpu-baseline > gpurun_out/bench_V1_v5.json 2> gpurun_out/bench_V1_v5.err; python tools/bench_summary.py gpurun_out/bench_V1_v5.json; tail -3 gpurun_out/bench_V1_v5.err
timeout 200 python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_V1_b512_v5.json 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_V1_b512_v5.csv python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1.log 2>&1
python tools/launch_summary.py gpurun_out/launches_V1_b512_v5.csv
