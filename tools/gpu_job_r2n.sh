set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ll_angle|blur5_sobel|scale08|order_kernel|rect_nfa|lbd_kernel' -s 12 -c 6 -o gpurun_out/r2n_small_b512 -f python bench.py --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2n_ncu.log 2>&1; tail -2 gpurun_out/r2n_ncu.log
