#!/usr/bin/env python
"""Randomised stress test of the fused fast path against the separate-stage path (developer tool; GPU):
random sizes, camera counts, splat sizes, opacities and depth distributions (incl. thin dense shells and
exactly repeated depths that exercise the second-level buckets / radix fallback). Bit-for-bit equality of
render / alpha is required; a raised overflow flag is reported, not an error. Also per case: single-pass binning of the
usable tiles (same pixels), the operator's lazy meta (same arrays), and — every third case — the backward with the
forward's hit map against the backward with its own culling."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib  # noqa: E402

import torch  # noqa: E402

rz = importlib.import_module("prometheus_b200.rasterization")
from prometheus_b200 import ops  # noqa: E402

DEV = "cuda"


def run_case(case: int) -> bool:
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
    if case >= 1000:
        # adversarial splats for the usable-tile binning and the culling extents: needle- and disc-shaped Gaussians (axis
        # ratios up to 5000), a few that cover the whole image, opacities on both sides of the 1 / 255 threshold and at 1
        scales = torch.exp(torch.log(torch.tensor(1e-4)) + r(n, 3) * (torch.log(torch.tensor(0.5)) - torch.log(torch.tensor(1e-4))))
        scales[: max(1, n // 500)] = 2.0
        pick = r(n)
        opac = torch.where(pick < 0.25, 1.0 / 255.0 + (r(n) - 0.5) * 2e-4, torch.where(pick < 0.35, torch.ones(n), r(n)))
        if n > 20000:  # keep the lists of these image-sized splats bounded
            keep = torch.arange(n) % (n // 20000 + 1) == 0
            means, quats, scales, opac = means[keep], quats[keep], scales[keep], opac[keep]
            n = means.shape[0]
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
    # single-pass binning of the usable tiles only (the module's steady state): same pixels from fewer records
    ok3, n_u = True, n_is
    if n_is and int(info["overflow_flag"].item()) == 0:
        im3, de3, ma3, info3 = rz.render_views_packed(*args[:7], bg, W, H, capacity=n_is + 5, tile_capacity=cnt_max + 1)
        n_u = int(info3["stats"][0].item())
        ok3 = info3["usable_tiles"] and n_u <= n_is and int(info3["overflow_flag"].item()) == 0 and \
            torch.equal(im3, im) and torch.equal(de3, de) and torch.equal(ma3, ma)
    # the operator with its lazy meta: fused render, gsplat's sorted arrays when read
    r4, a4, m4 = rz.rasterization(*args, backgrounds=bg, render_mode="RGB+ED", packed=False)
    ok4 = torch.equal(r4, r2) and torch.equal(a4, a2) and torch.equal(m4["flatten_ids"], m2["flatten_ids"]) and \
        torch.equal(m4["isect_ids"], m2["isect_ids"]) and torch.equal(m4["tiles_per_gauss"], m2["tiles_per_gauss"])
    # backward: the forward's hit map vs the backward's own culling (same gradients up to the order of the atomics)
    ok5 = True
    if n_is and int(info["overflow_flag"].item()) == 0 and case % 3 == 0:
        grads = []
        for use_map in (False, True):
            ops.HIT_MAP = use_map
            leaves = [t.clone().requires_grad_(True) for t in args[:5]]
            im5, de5, ma5, _ = rz.render_views_packed(*leaves, *args[5:7], bg, W, H, capacity=n_is + 5, tile_capacity=cnt_max + 1)
            (im5.square().mean() + de5.mean() + ma5.mean()).backward()
            grads.append([t.grad for t in leaves])
        ops.HIT_MAP = True
        for ga, gb in zip(*grads):
            tol = 1e-5 * float(gb.abs().max()) + 1e-7
            ok5 = ok5 and bool(((ga - gb).abs() <= tol + 1e-3 * gb.abs()).all())
    print("case %2d mode %d %3dx%3d C=%d n=%6d isects=%8d (usable %8d) max tile=%6d fallback=%s  %s" % (
        case, mode, W, H, C, n, n_is, n_u, cnt_max, fell_back,
        "ok" if ok and ok2 and ok3 and ok4 and ok5 else "MISMATCH fused=%s packed=%s usable=%s lazy=%s hitmap=%s" % (ok, ok2, ok3, ok4, ok5)))
    return bool(ok and ok2 and ok3 and ok4 and ok5)


def run(n_cases: int = 40, first: int = 0) -> int:
    """-> number of mismatching cases among first .. first + n_cases - 1"""
    return sum(not run_case(c) for c in range(first, first + n_cases))


if __name__ == "__main__":  # stress_fused.py [n_cases [first_case]]   (first_case >= 1000: adversarial splats)
    bad = run(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print("mismatches:", bad)
    sys.exit(1 if bad else 0)
