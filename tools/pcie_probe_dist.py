"""All ranks copy device->host (and host->device) at the same time: what the host side of an N-GPU box sustains, with
and without binding every rank to the CPUs / memory next to its GPU (NVML cpu affinity).
usage: torchrun --nproc-per-node N tools/pcie_probe_dist.py [GB per rank] [affinity 0|1]"""
import os
import sys
import time

import torch
import torch.distributed as dist


def bind_to_gpu_cpus(index):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    return sorted(os.sched_getaffinity(0))


def main():
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    aff = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    cpus = sorted(os.sched_getaffinity(0))
    note = ""
    if aff:
        try:
            cpus = bind_to_gpu_cpus(local)
        except Exception as e:      # noqa: BLE001
            note = " (affinity failed: %s)" % e
    dist.init_process_group("nccl")
    n = int(gb * 1e9)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    out = []
    for direction in ("d2h", "h2d"):
        best = 0.0
        for _ in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if direction == "d2h":
                host.copy_(dev, non_blocking=True)
            else:
                dev.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, n / (time.perf_counter() - t0) / 1e9)
        out.append(best)
    t = torch.tensor(out, dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    print("rank %d cpus %s..%s (%d)%s" % (rank, cpus[0], cpus[-1], len(cpus), note), flush=True)
    dist.barrier()
    if rank == 0:
        d2h = [float(v[0]) for v in allv]
        h2d = [float(v[1]) for v in allv]
        print("affinity=%d  d2h per rank %s  sum %.1f GB/s | h2d per rank %s  sum %.1f GB/s"
              % (aff, ["%.1f" % v for v in d2h], sum(d2h), ["%.1f" % v for v in h2d], sum(h2d)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
