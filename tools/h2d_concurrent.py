#!/usr/bin/env python
"""Concurrent pinned host -> device copy bandwidth, one process per GPU ("measure first").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/h2d_concurrent.py [--mb 128] [--iters 20] [--bind-numa]

Every rank copies `mb` MB (the per-rank image bytes of one survey step at N = 8 are 128 MB) from page-locked host memory
to its GPU `iters` times; all ranks start together.  Prints one JSON line on rank 0: per-rank GB/s alone (ranks take
turns) and together, the aggregate, and where the pinned buffers live.  --bind-numa binds each rank to the CPU affinity that
`nvidia-smi topo -m` reports for its GPU BEFORE allocating (first touch decides the NUMA node of the pinned pages).
"""
import argparse
import json
import os
import re
import subprocess

import torch
import torch.distributed as dist


def gpu_cpu_affinity(index):
    try:
        txt = subprocess.check_output(["nvidia-smi", "topo", "-m"], text=True)
    except Exception:      # noqa: BLE001
        return None, None
    for ln in txt.splitlines():
        if ln.startswith("GPU%d" % index + "\t") or ln.startswith("GPU%d " % index):
            cols = [c for c in re.split(r"\t+", ln.strip()) if c]
            rng = [c for c in cols if re.fullmatch(r"[0-9,\-]+", c) and ("-" in c or "," in c)]
            numa = cols[-2] if len(cols) >= 2 else None
            if rng:
                cpus = set()
                for part in rng[0].split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                return sorted(cpus), numa
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=128)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--bind-numa", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus, numa = gpu_cpu_affinity(local)
    bound = False
    if a.bind_numa and cpus:
        try:
            os.sched_setaffinity(0, cpus)
            bound = True
        except Exception:      # noqa: BLE001
            pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(rank + 1)                                   # first touch under the (optional) binding
    d = torch.empty(n, dtype=torch.uint8, device=dev)

    def run():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return n * a.iters / (e0.elapsed_time(e1) * 1e-3) / 1e9

    together = run()
    alone = 0.0
    for r in range(world):                              # ranks take turns
        if world > 1:
            dist.barrier()
        if r == rank:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                d.copy_(h, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            alone = n * a.iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([together, alone], device=dev, dtype=torch.float64)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
    else:
        allt = [t]
    info = [None] * world
    mine = dict(rank=rank, gpu=local, cpu_affinity_of_gpu="%s" % (("%d-%d" % (cpus[0], cpus[-1])) if cpus else None), numa=numa, bound=bound)
    if world > 1:
        dist.all_gather_object(info, mine)
    else:
        info = [mine]
    if rank == 0:
        tg = [float(x[0]) for x in allt]
        al = [float(x[1]) for x in allt]
        print(json.dumps(dict(tool="h2d_concurrent", n_gpus=world, mb=a.mb, iters=a.iters, numa_binding=a.bind_numa,
                              gbs_together_per_rank=[round(v, 2) for v in tg], gbs_together_aggregate=round(sum(tg), 2),
                              gbs_alone_per_rank=[round(v, 2) for v in al], host_cpus=os.cpu_count(), ranks=info)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
