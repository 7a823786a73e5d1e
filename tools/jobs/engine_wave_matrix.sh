set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" > gpurun_out/r2ae_$name.log 2>&1; echo "$name $(tail -1 gpurun_out/r2ae_$name.log | cut -c1-260)"; }
P="timeout 300 python bench.py --gpus 1 --e2e-only"
export VPL_ENGINE_VARIANT=2
run b8192_s1 X=0 $P --steps 10 --warmup 3 --batch 8192 --slots 1
run b9472_s1 X=0 $P --steps 10 --warmup 3 --batch 9472 --slots 1
run b6144_s2 X=0 $P --steps 14 --warmup 4 --batch 6144 --slots 2
run b6144_s2_tog X=0 $P --steps 14 --warmup 4 --batch 6144 --slots 2 --e2e-together
run b4736_s2_tog X=0 $P --steps 20 --warmup 6 --batch 4736 --slots 2 --e2e-together
run b4736_s2 X=0 $P --steps 20 --warmup 6 --batch 4736 --slots 2
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
