# GPU job: the bench lines and captures that go into profiles/ (1 GPU)
set -x
timeout 400 python bench.py --workload V1 > gpurun_out/bench_V1_final.json 2> gpurun_out/bench_V1_final.err; python tools/bench_summary.py gpurun_out/bench_V1_final.json; tail -3 gpurun_out/bench_V1_final.err
timeout 400 python bench.py --workload V1r > gpurun_out/bench_V1r_final.json 2> gpurun_out/bench_V1r_final.err; python tools/bench_summary.py gpurun_out/bench_V1r_final.json
timeout 300 python bench.py --workload V1 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_V1.json 2> gpurun_out/bench_ref_V1.err; tail -c 700 gpurun_out/bench_ref_V1.json
timeout 200 python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_V1_b512_final.json 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_V1_b512_final.csv python bench.py --workload V1 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1.log 2>&1
python tools/launch_summary.py gpurun_out/launches_V1_b512_final.csv
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -c 5 -o gpurun_out/prof_vp_b512_final python bench.py --workload V1 --batch 512 --steps 1 --warmup 1 --no-cpu-baseline --profile-region > gpurun_out/ncu_V1_full.log 2>&1
ls -la gpurun_out/prof_vp_b512_final.ncu-rep
timeout 500 python bench.py > gpurun_out/bench_C2_final3.json 2> gpurun_out/bench_C2_final3.err; python tools/bench_summary.py gpurun_out/bench_C2_final3.json; tail -3 gpurun_out/bench_C2_final3.err
