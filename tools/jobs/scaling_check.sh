# N-GPU weak-scaling check of the default bench (N as argument), one process per GPU
set -x
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 600 $TR 29551 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scaling_${N}gpu.log 2>&1; tail -1 gpurun_out/scaling_${N}gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['parity_checked'])"

