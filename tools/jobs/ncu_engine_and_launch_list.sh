set -x
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ag_bench.log 2>&1; tail -1 gpurun_out/r2ag_bench.log | cut -c1-700
timeout 900 ncu --set full --clock-control none --import-source on -k regex:region_engine -s 2 -c 1 -o gpurun_out/r2ag_engine_b9472 -f python bench.py --batch 9472 --slots 1 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2ag_ncu.log 2>&1; tail -2 gpurun_out/r2ag_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2ag_launches_b4736.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2ag_ncu2.log 2>&1
