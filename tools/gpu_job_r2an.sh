# engine FIFO window in shared memory: 256 (default) / 128 / 64 entries -> more of the SM's 256 KB left as L1
set -x
mkdir -p gpurun_out
for R in 128 64; do
  VPL_EXTRA_NVCC="-DVPL_ENGINE_RING=$R" python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2an_build_$R.log 2>&1
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "bench_path or both_region" > gpurun_out/r2an_pytest_$R.log 2>&1; tail -1 gpurun_out/r2an_pytest_$R.log
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2an_bench_$R.log 2>&1; tail -1 gpurun_out/r2an_bench_$R.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($R, d['value'], d['e2e']['value'], d['parity_checked'], d['roofline']['stage_ms_per_step']['region'])"
done
