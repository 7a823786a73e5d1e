"""Host->device copy rate per GPU, alone and with every rank copying at once (torchrun, one rank per GPU).

Explains the e2e scaling of bench.py: a step uploads 1.48 GB of frames per GPU, so the pipeline hides the upload only
while every GPU still gets more than (bytes per step / step time) from the host.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py
"""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 1478492160  # one C2 step: 4096 frames of 752x480
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(7)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    small_h = torch.empty(100 << 20, dtype=torch.uint8).pin_memory()
    small_d = torch.empty(100 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rate(reps=6, d2h=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
            if d2h:
                small_h.copy_(small_d, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    out = {"world": world}
    # every rank at once
    barrier()
    r = rate()
    barrier()
    t = torch.tensor([r], device="cuda", dtype=torch.float64)
    allr = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, t)
    else:
        allr = [t]
    out["concurrent_gbs"] = [round(float(x[0]), 2) for x in allr]
    # one rank at a time
    alone = []
    for k in range(world):
        barrier()
        v = rate() if rank == k else 0.0
        barrier()
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        alone.append(round(float(t[0]), 2))
    out["alone_gbs"] = alone
    if rank == 0:
        out["concurrent_sum_gbs"] = round(sum(out["concurrent_gbs"]), 1)
        out["need_per_gpu_gbs_at_80ms"] = round(nbytes / 0.080 / 1e9, 2)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
