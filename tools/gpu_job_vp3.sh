# GPU job: whole GPU suite (the shared sincos changed), V1 bench, launch list, full capture of the vp kernels
set -x
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all3.log 2>&1; tail -3 gpurun_out/pytest_gpu_all3.log
timeout 300 python bench.py --workload V1 --no-cpu-baseline > gpurun_out/bench_V1_v3.json 2> gpurun_out/bench_V1_v3.err; python tools/bench_summary.py gpurun_out/bench_V1_v3.json; tail -3 gpurun_out/bench_V1_v3.err
timeout 200 python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_V1_b512_v3.json 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_V1_b512_v3.csv python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1.log 2>&1
python tools/launch_summary.py gpurun_out/launches_V1_b512_v3.csv
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -c 5 -o gpurun_out/prof_vp_b512_v3 python bench.py --workload V1 --batch 512 --steps 1 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1_full.log 2>&1
ls -la gpurun_out/prof_vp_b512_v3.ncu-rep
