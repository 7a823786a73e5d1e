set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2ap_pytest.log 2>&1; tail -2 gpurun_out/r2ap_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ap_bench.log 2>&1; tail -1 gpurun_out/r2ap_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['parity_checked'], {k:v for k,v in d['roofline']['stage_ms_per_step'].items() if v>0})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rect_nfa|lbd_kernel' -c 8 --csv --log-file gpurun_out/r2ap_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2ap_ncu.log 2>&1
