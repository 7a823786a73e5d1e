# e2e scaling experiments on N GPUs (N = $1): upload rates, then the e2e figure with 2, 3 and 4 slots
set -x
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
nproc; free -g | head -2
timeout 300 $TR 29511 tools/h2d_probe.py > gpurun_out/r2w_h2d_$N.log 2>&1; tail -1 gpurun_out/r2w_h2d_$N.log
for S in 2 3; do
  timeout 400 $TR 2952$S bench.py --gpus $N --steps 20 --warmup 5 --slots $S --e2e-only > gpurun_out/r2w_e2e_${N}_s$S.log 2>&1
  tail -1 gpurun_out/r2w_e2e_${N}_s$S.log
done
timeout 400 $TR 29531 bench.py --gpus $N --steps 40 --warmup 8 --slots 4 --batch 2048 --e2e-only > gpurun_out/r2w_e2e_${N}_s4_b2048.log 2>&1
tail -1 gpurun_out/r2w_e2e_${N}_s4_b2048.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --slots 2 --e2e-only > gpurun_out/r2w_e2e_1_s2.log 2>&1
tail -1 gpurun_out/r2w_e2e_1_s2.log
