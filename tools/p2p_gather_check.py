#!/usr/bin/env python
"""2+ GPU check of parallel.PeerGather (fused unproject + all-gather over peer memory) against the NCCL all-gather
path and a single-GPU unproject of all views. Launch with torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_gather_check.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from prometheus_b200 import ops, parallel, scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
V_in, S_in = 4 * world, int(sys.argv[1]) if len(sys.argv) > 1 else 128
params_all = torch.stack([scenes.make_params(1, S_in, 77 + v)[0] for v in range(V_in)]).to(dev)
cams_all = scenes.make_cameras(V_in, S_in, 77).to(dev)
v0, v1 = parallel.shard_range(V_in, world, rank)
params, cams = params_all[v0:v1].contiguous(), cams_all[v0:v1].contiguous()
muls = torch.ones(14, device=dev)
n_local = (v1 - v0) * S_in * S_in
cfg = ("mvsplat", scenes.Z_NEAR, scenes.Z_FAR, scenes.S_MIN, scenes.S_MAX)
want = ops.unproject_fwd(params_all, cams_all, muls, *cfg)
nccl = parallel.gather_gaussians(lambda slots: ops.unproject_fwd(params, cams, muls, *cfg, out=slots), n_local, dev)
gathers = {}
for name, mc in (("peer stores", False), ("NVSwitch multicast", True)):
    pg = parallel.PeerGather(n_local, dev, multicast=mc, force_multicast=mc)
    if mc and not pg.multicast:
        continue  # no multicast support on this box
    gathers[name] = pg
    for it in range(3):  # repeated use of the same buffer
        got = pg.gather(lambda n, ptrs: ops.unproject_fwd_multi(params, cams, muls, *cfg, n, ptrs))
        torch.cuda.synchronize()
        for a, b, c in zip(got, want, nccl):
            assert torch.equal(a, b) and torch.equal(c, b), "mismatch on rank %d (%s)" % (rank, name)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / n * 1e3


times = {name: timeit(lambda pg=pg: pg.gather(lambda n, ptrs: ops.unproject_fwd_multi(params, cams, muls, *cfg, n, ptrs)))
         for name, pg in gathers.items()}
t_nccl = timeit(lambda: parallel.gather_gaussians(lambda slots: ops.unproject_fwd(params, cams, muls, *cfg, out=slots), n_local, dev))
if rank == 0:
    print("p2p gather ok on %d GPUs: %d Gaussians/rank; fused unproject+gather %s; unproject + NCCL all-gather %.3f ms" % (
        world, n_local, ", ".join("%s %.3f ms" % kv for kv in times.items()), t_nccl))
dist.destroy_process_group()
