#!/usr/bin/env python
"""Per-tile statistics of one config (developer tool; GPU): intersections per tile, how much of each
tile's list the fused renderer sorted, and how much the compositor needed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prometheus_b200 import GaussianConverter, ops, scenes  # noqa: E402
from prometheus_b200.rasterization import _rasterize_core  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda", 0)
sc = scenes.make_scene(cfg)
S = sc.size_out
conv = GaussianConverter("mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX).to(dev)
with torch.no_grad():
    g = conv(sc.params.to(dev), sc.cameras_in.to(dev))
    vm, K = ops.camera_prep(sc.cameras_out[0].to(dev), S, S)
    V = vm.shape[0]
    r, a, meta = _rasterize_core(g[0][0], g[4][0], g[3][0], g[2][0, :, 0], g[1][0, :, 0], vm, K, S, S,
                                 backgrounds=torch.ones(V, 3, device=dev), render_mode="RGB+ED", packed=False, fused=True)
off = meta["isect_offsets"].flatten().long()
n = int(meta["flatten_ids"].numel())
cnt = torch.cat([off[1:], torch.tensor([n], device=dev)]) - off
srt = meta["sorted_counts"].flatten().long()
q = torch.tensor([0.0, 0.1, 0.5, 0.9, 0.99, 0.999, 1.0], device=dev)
print("tiles", cnt.numel(), "isects", n)
print("count  quantiles", [int(x) for x in torch.quantile(cnt.float(), q)])
print("sorted quantiles", [int(x) for x in torch.quantile(srt.float(), q)], "sum", int(srt.sum()))
print("alpha min/mean", float(a.min()), float(a.mean()))
top = torch.argsort(srt, descending=True)[:12]
for t in top.tolist():
    print("tile", t, "cam", t // (meta["tile_width"] * meta["tile_height"]), "xy", (t % (meta["tile_width"] * meta["tile_height"])) % meta["tile_width"],
          (t % (meta["tile_width"] * meta["tile_height"])) // meta["tile_width"], "count", int(cnt[t]), "sorted", int(srt[t]))
# cost model: groups * count (collect passes) per tile
groups = (srt + 1023) // 1024
print("sum groups*count", int((groups * cnt).sum()), "max groups*count", int((groups * cnt).max()), "mean", float((groups * cnt).float().mean()))
