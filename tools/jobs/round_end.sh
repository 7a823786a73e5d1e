# GPU job: what the driver runs at round end (smoke, GPU suite, both bench arms at N=1)
set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_default_final.json 2> gpurun_out/bench_ref_default_final.err; tail -c 400 gpurun_out/bench_ref_default_final.json
timeout 500 python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; python tools/bench_summary.py gpurun_out/bench_default_final.json 2>/dev/null | head -6; tail -3 gpurun_out/bench_default_final.err
