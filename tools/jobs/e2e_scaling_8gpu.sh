# 8-GPU check: host upload rates with every GPU copying, the full bench line, and e2e without upload-ahead
set -x
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 300 $TR 29511 tools/h2d_probe.py > gpurun_out/r2y_h2d_$N.log 2>&1; tail -1 gpurun_out/r2y_h2d_$N.log
timeout 500 $TR 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2y_bench_$N.log 2>&1; tail -1 gpurun_out/r2y_bench_$N.log | cut -c1-1000
timeout 400 $TR 29523 bench.py --gpus $N --steps 20 --warmup 5 --e2e-only --no-upload-ahead > gpurun_out/r2y_e2e_${N}_noahead.log 2>&1; tail -1 gpurun_out/r2y_e2e_${N}_noahead.log
timeout 400 $TR 29524 bench.py --gpus $N --steps 20 --warmup 5 --e2e-only > gpurun_out/r2y_e2e_${N}_ahead.log 2>&1; tail -1 gpurun_out/r2y_e2e_${N}_ahead.log
