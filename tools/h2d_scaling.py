"""Host->device copy bandwidth per GPU when N ranks copy at once (the end-to-end arm's ceiling, SURVEY 8e / bench `e2e`).

Run under torchrun with N ranks; every rank copies a pinned 1 GiB buffer to its GPU 20 times between two barriers and
reports GB/s from CUDA events (rank 0 prints min / mean / max over ranks). BIND=0 skips the NUMA-local CPU binding.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_scaling.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import bind_to_gpu_cpus


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    bind = os.environ.get("BIND", "1") != "0"
    binding = bind_to_gpu_cpus(local) if bind else None
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 30
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)                                    # first touch on the bound cores
    devb = torch.empty(n, dtype=torch.uint8, device="cuda")
    back = torch.empty(n, dtype=torch.uint8).pin_memory()
    out = {}
    for name, fn in (("h2d", lambda: devb.copy_(host, non_blocking=True)), ("d2h", lambda: back.copy_(devb, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        gbs = 20 * n / (a.elapsed_time(b) * 1e-3) / 1e9
        t = torch.tensor([gbs], device="cuda")
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            vals = [float(v.item()) for v in allv]
        else:
            vals = [gbs]
        out[name] = dict(per_gpu_gbs=[round(v, 2) for v in vals], mean=round(sum(vals) / len(vals), 2), aggregate=round(sum(vals), 1))
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, bound_to_gpu_cpus=bind, binding_rank0=binding, **out)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
