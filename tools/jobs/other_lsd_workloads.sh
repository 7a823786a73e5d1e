# the other LSD workloads of BASELINE.json through the default (grouped) pipeline, default batch sizes
set -x
mkdir -p gpurun_out
for WL in C1 C3 C4; do
timeout 600 python bench.py --gpus 1 --workload $WL --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/wl_$WL.log 2>&1; tail -1 gpurun_out/wl_$WL.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$WL', d['config']['frames_per_step'], round(d['value']), round(d['e2e']['value']), d['parity_checked'], d['lines_per_frame'])" || tail -3 gpurun_out/wl_$WL.log
done
