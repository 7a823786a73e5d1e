# upload-ahead: parity, then the e2e figure on N GPUs with and without it
set -x
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "bench_path or upload_ahead or dense_collect" > gpurun_out/r2x_pytest.log 2>&1; tail -3 gpurun_out/r2x_pytest.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x_bench_1.log 2>&1; tail -1 gpurun_out/r2x_bench_1.log | cut -c1-900
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-only --no-upload-ahead > gpurun_out/r2x_e2e_1_noahead.log 2>&1; tail -1 gpurun_out/r2x_e2e_1_noahead.log
timeout 400 $TR 29522 bench.py --gpus $N --steps 20 --warmup 5 --e2e-only > gpurun_out/r2x_e2e_${N}_ahead.log 2>&1; tail -1 gpurun_out/r2x_e2e_${N}_ahead.log
timeout 400 $TR 29523 bench.py --gpus $N --steps 20 --warmup 5 --e2e-only --no-upload-ahead > gpurun_out/r2x_e2e_${N}_noahead.log 2>&1; tail -1 gpurun_out/r2x_e2e_${N}_noahead.log
timeout 400 $TR 29524 bench.py --gpus $N --steps 20 --warmup 5 --e2e-only --slots 3 > gpurun_out/r2x_e2e_${N}_ahead_s3.log 2>&1; tail -1 gpurun_out/r2x_e2e_${N}_ahead_s3.log
