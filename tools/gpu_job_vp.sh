# GPU job of the vanishing-point stage: benches, launch list, one full ncu capture (each command under its own timeout)
set -x
timeout 200 python -m pytest tests/test_cpp_facade.py -m gpu -x -q > gpurun_out/pytest_cpp_vp.log 2>&1; tail -3 gpurun_out/pytest_cpp_vp.log
timeout 300 python bench.py --workload V1 > gpurun_out/bench_V1.json 2> gpurun_out/bench_V1.err; tail -c 2500 gpurun_out/bench_V1.json; tail -5 gpurun_out/bench_V1.err
timeout 300 python bench.py --workload V1r > gpurun_out/bench_V1r.json 2> gpurun_out/bench_V1r.err; tail -c 900 gpurun_out/bench_V1r.json
timeout 200 python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_V1_b512.json 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_V1_b512.csv python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -c 5 -o gpurun_out/prof_vp_b512 python bench.py --workload V1 --batch 512 --steps 1 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1_full.log 2>&1
ls -la gpurun_out/prof_vp_b512.ncu-rep
