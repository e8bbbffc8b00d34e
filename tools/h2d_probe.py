#!/usr/bin/env python
"""Host<->device copy bandwidth with all ranks of a box copying at once (developer tool): what bounds the e2e
figure at N = 8. Run under torchrun. Reports per-rank and aggregate GB/s for H2D alone, D2H alone and both,
without and with the NUMA binding of parallel.bind_to_gpu_numa, and where the pinned pages landed."""
import os
import re
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def node_of(t):
    """NUMA node histogram of the pages backing pinned tensor t (from /proc/self/numa_maps)."""
    addr = t.data_ptr()
    best = None
    try:
        for line in open("/proc/self/numa_maps"):
            a = int(line.split()[0], 16)
            if a <= addr and (best is None or a > best[0]):
                best = (a, line)
        return " ".join(re.findall(r"N\d+=\d+", best[1])) if best else "?"
    except Exception as e:  # noqa: BLE001
        return "err %s" % e


def bw(fn, nbytes, iters=10):
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from prometheus_b200 import parallel

    if rank == 0:
        os.system("nvidia-smi topo -m 2>&1 | head -24; lscpu | grep -i 'numa\\|socket\\|model name\\|^CPU(s)'; "
                  "cat /sys/devices/system/node/node*/meminfo 2>/dev/null | grep MemTotal")
    n = 128 << 20
    s_out = torch.cuda.Stream()
    for mode in ("nobind", "bind"):
        if mode == "bind":
            ok = parallel.bind_to_gpu_numa(lr)
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_in.fill_(1)
        h_out = torch.empty(n // 3, dtype=torch.uint8).pin_memory()
        h_out.fill_(1)
        d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n // 3, dtype=torch.uint8, device="cuda")

        def h2d():
            d_in.copy_(h_in, non_blocking=True)

        def d2h():
            h_out.copy_(d_out, non_blocking=True)

        def both():
            d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_out)

        for f in (h2d, d2h, both):
            f()
        r = [bw(h2d, n), bw(d2h, n // 3), bw(both, n + n // 3)]
        t = torch.tensor(r, device="cuda", dtype=torch.float64)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        info = "%s cpus=%d pages[%s]" % (mode, len(os.sched_getaffinity(0)), node_of(h_in))
        infos = [None] * world
        dist.all_gather_object(infos, info)
        if rank == 0:
            g = torch.stack(gathered).cpu()
            print("==", mode)
            for i in range(world):
                print("  rank %d  h2d %.1f  d2h %.1f  both %.1f GB/s   %s" % (i, g[i, 0], g[i, 1], g[i, 2], infos[i]))
            print("  aggregate h2d %.1f  d2h %.1f  both %.1f GB/s" % tuple(g.sum(0).tolist()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
