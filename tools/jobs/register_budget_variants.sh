# register-budget variants measured where they matter: in the two-slot pipeline (bench value / e2e), same box, interleaved
set -x
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2ab_$name.log 2>&1
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2ab_$name.log").read().strip().splitlines()[-1])
print("$name", round(d["value"]), round(d["e2e"]["value"]), {k:round(v,1) for k,v in d["roofline"]["stage_ms_per_step"].items() if v>0})
PY
}
run base1 X=0
run eng2 VPL_ENGINE_VARIANT=2
run nfa1 VPL_NFA_VARIANT=1
run nfa2 VPL_NFA_VARIANT=2
run lla1 VPL_LLA_VARIANT=1
run base2 X=0
run eng2_nfa1 VPL_ENGINE_VARIANT=2 VPL_NFA_VARIANT=1
run eng2_nfa2_lla1 VPL_ENGINE_VARIANT=2 VPL_NFA_VARIANT=2 VPL_LLA_VARIANT=1
run eng3 VPL_ENGINE_VARIANT=3
run eng2b VPL_ENGINE_VARIANT=2
