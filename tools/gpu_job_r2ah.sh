set -x
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 4 --warmup 3 --timeline > gpurun_out/r2ah_tl.log 2>&1; tail -1 gpurun_out/r2ah_tl.log
timeout 300 python bench.py --gpus 1 --steps 2 --warmup 2 --timeline --batch 9472 --slots 1 > gpurun_out/r2ah_tl_s1.log 2>&1; tail -1 gpurun_out/r2ah_tl_s1.log
