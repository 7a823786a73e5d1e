# GPU job: parity of the LSD path + C2 / C4 benches after a change in lbd_kernel's launch
set -x
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity_lbd.log 2>&1; tail -3 gpurun_out/pytest_parity_lbd.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_C2_lbd3.json 2> gpurun_out/bench_C2_lbd3.err; python tools/bench_summary.py gpurun_out/bench_C2_lbd3.json | head -4
timeout 200 python bench.py --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_C2_b512_lbd.json 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C2_b512_lbd.csv python bench.py --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_C2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_C2_b512_lbd.csv | grep -E "lbd|angle|kernel "
timeout 300 python bench.py --workload C4 --batch 640 --no-cpu-baseline > gpurun_out/bench_C4_lbd.json 2> gpurun_out/bench_C4_lbd.err; python tools/bench_summary.py gpurun_out/bench_C4_lbd.json | head -4
