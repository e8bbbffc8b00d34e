#!/usr/bin/env python
"""How often does the speculative sizing of the fast path fail along a camera trajectory? (developer tool; GPU)
C5 on one GPU: 120 frames in chunks of 8 cameras; per step: time, two-pass (exact / re-rendered) chunks, single-pass chunks."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prometheus_b200 import GaussianRenderer, ops, scenes  # noqa: E402

dev = torch.device("cuda", 0)
cfg = scenes.CONFIGS["C5"]
V_in, S_in, V_out, S = cfg["views_in"], cfg["size_in"], cfg["views_out"], cfg["size_out"]
params = torch.stack([scenes.make_params(1, S_in, 4321 + v)[0] for v in range(V_in)]).to(dev)
cams_in = scenes.make_cameras(V_in, S_in, 4321).to(dev)
cams_out = scenes.make_cameras(V_out, S, 4322, first_identity=False).to(dev)
rend = GaussianRenderer(background=[1, 1, 1]).to(dev)
if len(sys.argv) > 1:
    rend.tile_capacity_slack = float(sys.argv[1])
muls = torch.ones(14, device=dev)
with torch.no_grad():
    xyz, rgb, opac, scl, quat = ops.unproject_fwd(params, cams_in, muls, "mvsplat", scenes.Z_NEAR, scenes.Z_FAR, scenes.S_MIN, scenes.S_MAX)
    g = (xyz[None], rgb[None, :, None, :], opac[None, :, None], scl[None], quat[None])
    for step in range(5):
        ops.PROFILE = []
        torch.cuda.synchronize()
        t0 = time.time()
        rend(cams_out[None], g, bg_color=None, h=S, w=S)
        torch.cuda.synchronize()
        dt = (time.time() - t0) * 1e3
        names = [n for n, _, _ in ops.PROFILE]
        ops.PROFILE = None
        print("step %d: %.1f ms, two-pass chunks %d, single-pass chunks %d, tile-cap hint %s" % (
            step, dt, names.count("bin_scatter"), names.count("project_scatter_fwd"), list(rend._tile_cap_hint.values())))
