set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2ao_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2ao_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ao_smoke.log 2>&1; tail -1 gpurun_out/r2ao_smoke.log
