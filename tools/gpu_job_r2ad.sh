set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" > gpurun_out/r2ad_$name.log 2>&1; echo "$name $(tail -1 gpurun_out/r2ad_$name.log | cut -c1-330)"; }
B="timeout 300 python bench.py --gpus 1 --steps 20 --warmup 6 --e2e-only"
run eng2_together VPL_ENGINE_VARIANT=2 $B --e2e-together
run base_together X=0 $B --e2e-together
run eng2 VPL_ENGINE_VARIANT=2 $B
run eng2_together_b8192_s1 VPL_ENGINE_VARIANT=2 timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --e2e-only --batch 8192 --slots 2 --e2e-together
run eng2_b8192 VPL_ENGINE_VARIANT=2 timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --e2e-only --batch 8192 --slots 2
