# round 2, job J (2 GPUs): C5 sharded sequence with the 1-GPU identity check, C2 weak scaling with distinct frames
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload C5 --total-frames 2048 --steps 2 --warmup 1 > gpurun_out/r2j_bench_C5_1gpu.json 2> gpurun_out/r2j_bench_C5_1gpu.err; tail -c 900 gpurun_out/r2j_bench_C5_1gpu.json; tail -3 gpurun_out/r2j_bench_C5_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload C5 --total-frames 2048 --steps 2 --warmup 1 > gpurun_out/r2j_bench_C5_2gpu.json 2> gpurun_out/r2j_bench_C5_2gpu.err; tail -c 900 gpurun_out/r2j_bench_C5_2gpu.json; tail -3 gpurun_out/r2j_bench_C5_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2j_bench_C2_2gpu.json 2> gpurun_out/r2j_bench_C2_2gpu.err; python tools/bench_summary.py gpurun_out/r2j_bench_C2_2gpu.json | head -2; tail -3 gpurun_out/r2j_bench_C2_2gpu.err
