set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2as_pytest.log 2>&1; tail -2 gpurun_out/r2as_pytest.log
for V in 1 0 1; do
VPL_GROUP_SERIAL_BACK=$V timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2as_bench_$V.log 2>&1; tail -1 gpurun_out/r2as_bench_$V.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($V, d['value'], d['e2e']['value'], d['parity_checked'], {k:v for k,v in d['roofline']['stage_ms_per_step'].items() if v>0})"
done
