# -DVPL_NFA_CHECK build: the float32 early-exit test of nfa() against the double sequence, on every decision of a bench run
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2ak_check_C2.log 2>&1; tail -1 gpurun_out/r2ak_check_C2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('nfa_check'), d['parity_checked'])"
timeout 600 python bench.py --gpus 1 --workload C1 --batch 1024 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2ak_check_C1.log 2>&1; tail -1 gpurun_out/r2ak_check_C1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('nfa_check'), d['parity_checked'])"
timeout 600 python bench.py --gpus 1 --workload C3 --batch 512 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2ak_check_C3.log 2>&1; tail -1 gpurun_out/r2ak_check_C3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('nfa_check'), d['parity_checked'])"
