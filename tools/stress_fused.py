#!/usr/bin/env python
"""Randomised stress test of the fused fast path against the separate-stage path (developer tool; GPU):
random sizes, camera counts, splat sizes, opacities and depth distributions (incl. thin dense shells and
exactly repeated depths that exercise the second-level buckets / radix fallback). Bit-for-bit equality of
render / alpha is required; a raised overflow flag is reported, not an error."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib  # noqa: E402

import torch  # noqa: E402

rz = importlib.import_module("prometheus_b200.rasterization")
from prometheus_b200 import ops  # noqa: E402

DEV = "cuda"
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
for case in range(n_cases):
    g = torch.Generator().manual_seed(1000 + case)
    r = lambda *s: torch.rand(*s, generator=g)
    W = int(16 + r(1) * 240)
    H = int(16 + r(1) * 240)
    C = int(1 + r(1) * 5)
    n = int(200 + r(1) ** 2 * 60000)
    mode = case % 5
    means = torch.zeros(n, 3)
    means[:, :2] = (r(n, 2) - 0.5) * (1.5 if mode != 3 else 0.2)
    if mode == 0:      # volume
        means[:, 2] = 1.5 + r(n) * 3
    elif mode == 1:    # thin shells
        means[:, 2] = 2.0 + torch.randint(0, 6, (n,), generator=g) * 0.3 + r(n) * 1e-4
    elif mode == 2:    # exactly repeated depths
        means[:, 2] = 2.0 + torch.randint(0, 40, (n,), generator=g) * 0.05
    elif mode == 3:    # everything on a few tiles, huge lists
        means[:, 2] = 2.0 + r(n) * 0.5
    else:              # one shell + outliers
        means[:, 2] = torch.where(r(n) < 0.97, 2.5 + r(n) * 1e-5, 1.0 + r(n) * 8)
    quats = torch.randn(n, 4, generator=g)
    scales = 0.002 + r(n, 3) * (0.03 if mode != 3 else 0.01)
    opac = r(n) ** (1 if case % 2 else 3) * (0.15 if mode in (2, 3) else 1.0)
    colors = r(n, 3)
    vm = torch.eye(4).repeat(C, 1, 1)
    vm[:, 0, 3] = torch.linspace(-0.2, 0.2, C) if C > 1 else 0.0
    f = 0.9 * max(W, H)
    K = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]]).repeat(C, 1, 1)
    args = [t.to(DEV) for t in (means, quats, scales, opac, colors, vm, K)] + [W, H]
    bg = r(C, 3).to(DEV)
    r1, a1, m1 = rz._rasterize_core(*args, backgrounds=bg, render_mode="RGB+ED", packed=False, fused=True)
    r2, a2, m2 = rz._rasterize_core(*args, backgrounds=bg, render_mode="RGB+ED", packed=False, fused=False)
    fell_back = m1["isect_ids"] is not None
    ok = torch.equal(r1, r2) and torch.equal(a1, a2)
    # the packed module path on the same inputs
    im, de, ma, info = rz.render_views_packed(*args[:7], bg, W, H)
    ok2 = True
    if int(info["overflow_flag"].item()) == 0:
        ok2 = torch.equal(im, (r2[..., :3] * 2 - 1).permute(0, 3, 1, 2)) and torch.equal(ma[:, 0], a2[..., 0].clamp(0, 1)) and \
            torch.equal(de[:, 0], r2[..., 3])
    n_is = int(m2["flatten_ids"].numel())
    cnt_max = int(torch.diff(torch.cat([m2["isect_offsets"].flatten(), torch.tensor([n_is], device=DEV, dtype=torch.int32)])).max()) if n_is else 0
    print("case %2d mode %d %3dx%3d C=%d n=%6d isects=%8d max tile=%6d fallback=%s  %s" % (
        case, mode, W, H, C, n, n_is, cnt_max, fell_back, "ok" if ok and ok2 else "MISMATCH fused=%s packed=%s" % (ok, ok2)))
    bad += not (ok and ok2)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
