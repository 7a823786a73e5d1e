# does the length of the upload (copy-engine time per step) change the e2e step time on one GPU?
set -x
mkdir -p gpurun_out
for P in 0 1 2; do
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-only --stress-upload-passes $P > gpurun_out/r2z_e2e_p$P.log 2>&1; tail -1 gpurun_out/r2z_e2e_p$P.log
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-only --no-upload-ahead > gpurun_out/r2z_e2e_noahead.log 2>&1; tail -1 gpurun_out/r2z_e2e_noahead.log
