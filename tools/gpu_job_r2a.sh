# round 2, job A: GPU suite + both bench arms at N=1 after the C-ABI / bench changes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; tail -c 600 gpurun_out/r2a_bench_ref.json
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
