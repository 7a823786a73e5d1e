# GPU job: the CTA-per-frame vote variant (VPL_VP_VOTE=1) against the warp-per-frame one: parity, then timing
set -x
VPL_VP_VOTE=1 timeout 150 python -m pytest tests/test_gpu_vp.py -x -q -m gpu > gpurun_out/pytest_vp_vote1.log 2>&1; tail -3 gpurun_out/pytest_vp_vote1.log
for v in 0 1; do for b in 64 512 4096; do
VPL_VP_VOTE=$v timeout 120 python bench.py --workload V1 --batch $b --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_V1_vote${v}_b${b}.json 2> gpurun_out/bench_V1_vote.err; python tools/bench_summary.py gpurun_out/bench_V1_vote${v}_b${b}.json | head -3
done; done
