set -x
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2ar_ref.log 2>&1; tail -1 gpurun_out/r2ar_ref.log | cut -c1-600
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ar_bench.log 2>&1; tail -1 gpurun_out/r2ar_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['parity_checked'], d['roofline']['frac'], d['roofline']['frac_all_concurrent_launches'], d['roofline']['issue_rate'], d['cpu_baseline']['value'], d['latency'], d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2ar_launches_b4736.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2ar_ncu2.log 2>&1
