# GPU job: end-of-round bench lines and captures of the vanishing-point stage (1 GPU)
set -x
timeout 400 python bench.py --workload V1 > gpurun_out/bench_V1_final.json 2> gpurun_out/bench_V1_final.err; python tools/bench_summary.py gpurun_out/bench_V1_final.json 2>/dev/null | head -3; tail -3 gpurun_out/bench_V1_final.err
timeout 400 python bench.py --workload V1r > gpurun_out/bench_V1r_final.json 2> gpurun_out/bench_V1r_final.err; python tools/bench_summary.py gpurun_out/bench_V1r_final.json 2>/dev/null | head -3
timeout 200 python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_V1_b512_final.json 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_V1_b512_final.csv python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1.log 2>&1
python tools/launch_summary.py gpurun_out/launches_V1_b512_final.csv
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -c 5 -o gpurun_out/prof_vp_b512_final python bench.py --workload V1 --batch 512 --steps 1 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1_full.log 2>&1
VPL_VP_VOTE=0 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vp_vote -c 1 -o gpurun_out/prof_vp_vote_warp_b4096 python bench.py --workload V1 --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1_vote.log 2>&1
ls -la gpurun_out/prof_vp_b512_final.ncu-rep gpurun_out/prof_vp_vote_warp_b4096.ncu-rep
