set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" > gpurun_out/r2ac_$name.log 2>&1; echo "$name $(tail -1 gpurun_out/r2ac_$name.log | cut -c1-330)"; }
B="timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-only"
run eng2 VPL_ENGINE_VARIANT=2 $B
run eng2_nochain VPL_ENGINE_VARIANT=2 $B --e2e-no-chain
run eng2_noprofile VPL_ENGINE_VARIANT=2 $B --e2e-no-profile
run eng2_s3 VPL_ENGINE_VARIANT=2 $B --slots 3
run eng2_noahead VPL_ENGINE_VARIANT=2 $B --no-upload-ahead
run base_noprofile X=0 $B --e2e-no-profile
run base X=0 $B
