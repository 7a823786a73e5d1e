set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 600 $TR 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2aq_bench_2gpu.log 2>&1; tail -1 gpurun_out/r2aq_bench_2gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['parity_checked'], d['n_gpus'], d['config']['parallelism'])"
timeout 900 $TR 29542 bench.py --gpus 2 --workload C5 --total-frames 2048 --steps 1 --warmup 1 > gpurun_out/r2aq_C5_2gpu.log 2>&1; tail -1 gpurun_out/r2aq_C5_2gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d.get('identical_to_1gpu'), d.get('parity_checked'))"
