#!/usr/bin/env python
"""Aggregate host<->device bandwidth of the box with all N GPUs copying at once (torchrun, one process per GPU):
H2D alone, D2H alone, both directions together, from page-locked memory; plus the same with every rank bound to the CPUs
NVML reports as local to its GPU.  The end-to-end packed-batch numbers at N GPUs are bound by these figures.  Measurement
tool, not product."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")
    nbytes = 1 << 29

    def bind_local():
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [i for i in range(os.cpu_count()) if (mask[i // 64] >> (i % 64)) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
            return cpus
        except Exception as e:  # noqa: BLE001
            return str(e)

    def run(tag):
        h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h_in.fill_(1)  # first touch on the current CPU set
        h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h_out.fill_(2)
        d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        out = {}
        for name in ("h2d", "d2h", "both"):
            for rep in range(3):
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier(group=cpu_group)
                t0 = time.perf_counter()
                for _ in range(4):
                    if name in ("h2d", "both"):
                        with torch.cuda.stream(s1):
                            d_a.copy_(h_in, non_blocking=True)
                    if name in ("d2h", "both"):
                        with torch.cuda.stream(s2):
                            h_out.copy_(d_b, non_blocking=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                moved = 4 * nbytes * world * (2 if name == "both" else 1)
                out[name] = max(out.get(name, 0.0), moved / float(t.item()) / 1e9)
        if rank == 0:
            print(json.dumps({"tag": tag, "n_gpus": world, "aggregate_GBps": out}), flush=True)

    run("default placement")
    cpus = bind_local()
    if rank == 0:
        print(json.dumps({"rank0_local_cpus": cpus if isinstance(cpus, str) else f"{len(cpus)} cpus: {cpus[:4]}..."}), flush=True)
    run("process bound to the GPU's local CPUs (NVML affinity), buffers first-touched there")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
