"""Synthetic pixel-aligned Gaussian scenes of the shapes BASELINE.json names
(generator specified in SURVEY.md §8d). Host-side, deterministic (seeded torch CPU
generator); returns CPU tensors — callers move them to the GPU (bench.py does so
from pinned memory inside the timed e2e region).

Scene scale follows the reference's dataset normalisation (first camera identity,
baseline <~ 1, prometheus/datasets/base_dataset.py:587-603) and its default focal
(207.3 px at 256^2, prometheus/systems/gs_decoder_system.py:93); converter limits
are configurations/algorithm/base_algo.yaml:57-62.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn.functional as F

Z_NEAR, Z_FAR, S_MIN, S_MAX = 0.01, 100.0, 0.001, 1.0

# name -> (input views, input size, target views, target size, scenes)
CONFIGS = {
    "C1": dict(views_in=4, size_in=256, views_out=4, size_out=256, scenes=1),
    "C2": dict(views_in=8, size_in=512, views_out=8, size_out=512, scenes=1),
    "C3": dict(views_in=8, size_in=256, views_out=8, size_out=256, scenes=8),
    "C4": dict(views_in=8, size_in=512, views_out=8, size_out=512, scenes=64),
    "C5": dict(views_in=32, size_in=512, views_out=120, size_out=1024, scenes=1),
}


def focal_for(size: int) -> float:
    return 207.3 * size / 256.0


def make_cameras(n_views: int, size: int, seed: int, first_identity: bool = True) -> torch.Tensor:
    """[n_views, 18] OpenGL c2w (3x4 row-major) + fx, fy, cx, cy, H, W. Cameras orbit the
    point (0,0,-2) on a +-0.3 rad arc and look at it; view 0 is the identity."""
    g = torch.Generator().manual_seed(seed)
    theta = (torch.rand(n_views, generator=g) * 2 - 1) * 0.3
    if first_identity:
        theta[0] = 0.0
    f = focal_for(size)
    cams = torch.zeros(n_views, 18)
    for v in range(n_views):
        th = float(theta[v])
        s, c = math.sin(th), math.cos(th)
        R = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])  # columns: right, up, back
        t = torch.tensor([2.0 * s, 0.0, -2.0 + 2.0 * c])
        c2w = torch.cat([R, t[:, None]], dim=1)
        cams[v, :12] = c2w.reshape(-1)
        cams[v, 12:] = torch.tensor([f, f, size / 2.0, size / 2.0, float(size), float(size)])
    return cams


def make_params(n_views: int, size: int, seed: int) -> torch.Tensor:
    """Raw decoder maps [n_views, 14, size, size] (rgb3, uv2, depth1, opacity1, scale3, quat4)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    p = torch.empty(n_views, 14, size, size)
    p[:, 0:3] = rn(n_views, 3, size, size)
    p[:, 3:5] = 0.004 * rn(n_views, 2, size, size)
    lf = F.interpolate(rn(n_views, 1, 8, 8), size=(size, size), mode="bicubic", align_corners=False)
    p[:, 5:6] = -3.9 + 0.3 * lf
    p[:, 6:7] = 2.0 + 1.5 * rn(n_views, 1, size, size)
    x = (2.0 / focal_for(size) - S_MIN) / (S_MAX - S_MIN)
    p[:, 7:10] = math.log(x / (1.0 - x)) + 0.5 * rn(n_views, 3, size, size)
    p[:, 10:14] = rn(n_views, 4, size, size)
    return p


@dataclass
class Scene:
    params: torch.Tensor       # [B, N_in, 14, s, s]
    cameras_in: torch.Tensor   # [B, N_in, 18]
    cameras_out: torch.Tensor  # [B, N_out, 18]
    size_out: int
    name: str

    @property
    def n_gaussians(self) -> int:
        return self.params.shape[1] * self.params.shape[3] * self.params.shape[4]


def make_scene(name: str = "C2", seed: int = 1234, scenes: int | None = None, views_in: int | None = None,
               size_in: int | None = None, views_out: int | None = None, size_out: int | None = None) -> Scene:
    cfg = dict(CONFIGS[name])
    for k, v in dict(scenes=scenes, views_in=views_in, size_in=size_in, views_out=views_out, size_out=size_out).items():
        if v is not None:
            cfg[k] = v
    B = cfg["scenes"]
    params = torch.stack([make_params(cfg["views_in"], cfg["size_in"], seed + 17 * b) for b in range(B)])
    cams_in = torch.stack([make_cameras(cfg["views_in"], cfg["size_in"], seed + 17 * b) for b in range(B)])
    cams_out = torch.stack([make_cameras(cfg["views_out"], cfg["size_out"], seed + 17 * b + 1, first_identity=False)
                            for b in range(B)])
    return Scene(params, cams_in, cams_out, cfg["size_out"], name)


def render_loss(images: torch.Tensor, depths: torch.Tensor, masks: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Synthetic stand-in for the stage-1 loss (gs_decoder_system.py:214-239; LPIPS weights are
    not available offline): every one of the three outputs receives a gradient."""
    return ((images - target) ** 2).mean() + (1.0 / depths.clamp(0.1, 100.0)).mean() + (1.0 - masks).mean()


def render_loss_fused(images: torch.Tensor, depths: torch.Tensor, masks: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """`render_loss` with its per-pixel glue done by losses.render_loss_terms (one kernel each way instead of a dozen
    ATen passes); same value and gradients up to the summation order."""
    from .losses import render_loss_terms

    mse, disparity, _valid, entropy = render_loss_terms(images, depths, masks, target)
    return mse + disparity.mean() + entropy
