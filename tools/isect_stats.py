#!/usr/bin/env python
"""How many of gsplat's (Gaussian, tile) intersections can the compositor actually use? (developer tool)
For every visible pair: tiles of gsplat's AABB (3 sigma of the larger axis, square) vs tiles whose pixel-centre span
overlaps the per-axis extent of {alpha >= 1/255} (intersected with that AABB)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prometheus_b200 import GaussianConverter, ops, scenes  # noqa: E402

dev = torch.device("cuda", 0)
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc = scenes.make_scene(cfg)
S = sc.size_out
conv = GaussianConverter("mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX).to(dev)
with torch.no_grad():
    g = conv(sc.params.to(dev), sc.cameras_in.to(dev))
    cams = sc.cameras_out[0, :8].to(dev)
    vm, K = ops.camera_prep(cams, S, S)
    radii, means2d, depths, conics, tpg = ops.project_fwd(g[0][0], g[4][0], g[3][0], vm, K, S, S)
    op = g[2][0, :, 0][None].expand_as(radii)
    vis = radii > 0
    q0, q1, q2 = conics[..., 0], conics[..., 1], conics[..., 2]
    det = q0 * q2 - q1 * q1
    cxx, cyy = q2 / det, q0 / det          # covariance = conic^-1
    tau = torch.log((255.0 * op).clamp(min=1.0))
    hx = torch.sqrt(2 * tau * cxx) * 1.005 + 0.05
    hy = torch.sqrt(2 * tau * cyy) * 1.005 + 0.05
    T = S // 16
    mx, my, r = means2d[..., 0], means2d[..., 1], radii.float()

    def tiles(lo, hi):  # pixel-centre span of tile t: [16 t + 0.5, 16 t + 15.5]
        t0 = torch.ceil((lo - 15.5) / 16).clamp(0, T)
        t1 = (torch.floor((hi - 0.5) / 16) + 1).clamp(0, T)
        return t0, t1

    gx0 = torch.floor((mx - r) / 16).clamp(0, T); gx1 = torch.ceil((mx + r) / 16).clamp(0, T)
    gy0 = torch.floor((my - r) / 16).clamp(0, T); gy1 = torch.ceil((my + r) / 16).clamp(0, T)
    ax0, ax1 = tiles(mx - hx, mx + hx)
    ay0, ay1 = tiles(my - hy, my + hy)
    nx = (torch.minimum(gx1, ax1) - torch.maximum(gx0, ax0)).clamp(min=0)
    ny = (torch.minimum(gy1, ay1) - torch.maximum(gy0, ay0)).clamp(min=0)
    n_gsplat = ((gx1 - gx0) * (gy1 - gy0))[vis].sum().item()
    n_tight = (nx * ny)[vis & (op * 255 > 1)].sum().item()
    print("%s: visible pairs %d, gsplat intersections %d (%.2f per pair, tiles_per_gauss sum %d), alpha-extent intersections %d (%.1f %% fewer)"
          % (cfg, int(vis.sum()), n_gsplat, n_gsplat / vis.sum().item(), int(tpg.sum()), n_tight, 100 * (1 - n_tight / n_gsplat)))
    print("radius mean %.2f, hx mean %.2f, hy mean %.2f, opacity mean %.3f" % (r[vis].mean(), hx[vis].mean(), hy[vis].mean(), op[vis].mean()))
