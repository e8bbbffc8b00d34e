"""GPU probe: time the binning/sort stages on the C2 scene (dev tool, not part of the product)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prometheus_b200 import GaussianConverter, ops, scenes

dev = "cuda"
sc = scenes.make_scene("C2")
conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(dev)
with torch.no_grad():
    xyz, feat, op, scl, rot = conv(sc.params.to(dev), sc.cameras_in.to(dev))
vm, K = ops.camera_prep(sc.cameras_out[0].to(dev), 512, 512)
means, quats, scales = xyz[0].contiguous(), rot[0].contiguous(), scl[0].contiguous()


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("project no hist   %.3f ms" % t(lambda: ops.project_fwd(means, quats, scales, vm, K, 512, 512)))
radii, means2d, depths, conics, tpg = ops.project_fwd(means, quats, scales, vm, K, 512, 512)
print("bin_count %.3f ms" % t(lambda: ops.bin_count(means2d, radii, 512, 512)))
tc = ops.bin_count(means2d, radii, 512, 512)
offsets, stats = ops.tile_scan(tc)
n, mx = stats.tolist()
print("n_isects", n, "max tile count", mx, "mean", n / 8192, "tiles>8192:", int((tc > 8192).sum()), "tpg max", int(tpg.max()), "radii max", int(radii.max()))
print("tile_scan %.3f ms" % t(lambda: ops.tile_scan(tc)))
print("bin_scatter %.3f ms" % t(lambda: ops.bin_scatter(means2d, radii, depths, offsets, 512, 512, n)))
rec = ops.bin_scatter(means2d, radii, depths, offsets, 512, 512, n)
print("tile_sort %.3f ms" % t(lambda: ops.tile_sort(rec, offsets, 8, 512, 512, max_tile_count=mx)))
ids, flat, flag = ops.tile_sort(rec, offsets, 8, 512, 512, max_tile_count=mx)
print("overflow flag", int(flag.item()))
q = torch.quantile(tc.float().flatten(), torch.tensor([0.5, 0.9, 0.99, 1.0], device=dev))
print("tile count quantiles", q.tolist())
import importlib
rz = importlib.import_module('prometheus_b200.rasterization')
def full(): 
    return rz.isect_tiles(means2d, radii, depths, tpg, 512, 512, 8)
print("isect_tiles binned total %.3f ms" % t(full))
a = full()
b = rz.isect_tiles(means2d, radii, depths, tpg, 512, 512, 8, force_generic=True)
print("binned == generic:", all(torch.equal(x, y) for x, y in zip(a, b)))
