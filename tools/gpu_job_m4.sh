# GPU job: the matcher alone (M4) with the measured POPC peak, + the match parity tests
set -x
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "match or hamming or knn" > gpurun_out/pytest_match.log 2>&1; tail -3 gpurun_out/pytest_match.log
timeout 300 python bench.py --workload M4 > gpurun_out/bench_M4.json 2> gpurun_out/bench_M4.err; tail -c 2600 gpurun_out/bench_M4.json; tail -5 gpurun_out/bench_M4.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hamming_knn -c 2 -o gpurun_out/prof_hamming_m4 python bench.py --workload M4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_M4.log 2>&1
ls -la gpurun_out/prof_hamming_m4.ncu-rep
