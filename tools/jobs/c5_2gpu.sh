# C5: one 1280x720 sequence (2 octaves, k = 2) sharded over the ranks, results compared with the 1-GPU run
set -x
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 900 $TR 29542 bench.py --gpus $N --workload C5 --total-frames 9472 --steps 2 --warmup 1 > gpurun_out/c5_${N}gpu.log 2>&1; tail -1 gpurun_out/c5_${N}gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d.get('identical_to_1gpu'), d.get('parity_checked'), d['config']['frames_per_batch'])" || tail -5 gpurun_out/c5_${N}gpu.log
timeout 900 python bench.py --gpus 1 --workload C5 --total-frames 9472 --steps 2 --warmup 1 > gpurun_out/c5_1gpu.log 2>&1; tail -1 gpurun_out/c5_1gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d.get('identical_to_1gpu'), d.get('parity_checked'))" || tail -5 gpurun_out/c5_1gpu.log
