"""Shared helpers for the test-suite (scene builders on top of the oracle)."""
import numpy as np
import torch

from prometheus_b200 import scenes


def small_scene(oracle, views_in=2, size_in=32, views_out=2, size_out=48, seed=5, scenes_n=1):
    """Synthetic pixel-aligned scene -> oracle Gaussians + reference-path cameras (numpy)."""
    sc = scenes.make_scene("C1", seed=seed, views_in=views_in, size_in=size_in, views_out=views_out,
                           size_out=size_out, scenes=scenes_n)
    g = oracle.unproject(sc.params.numpy(), sc.cameras_in.numpy(), mode="mvsplat", z_near=scenes.Z_NEAR,
                         z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX)
    vm, K = oracle.camera_prep(sc.cameras_out[0].numpy(), size_out, size_out)
    return dict(scene=sc, means=g["xyz"][0], quats=g["rotations"][0], scales=g["scales"][0],
                opacities=g["opacity"][0, :, 0], colors=g["features"][0, :, 0], viewmats=vm, Ks=K,
                width=size_out, height=size_out)


def random_blobs(n=3000, C=2, W=64, H=48, seed=0, big=False):
    """Unstructured random Gaussians in front of C cameras (exercises culls and big radii)."""
    g = torch.Generator().manual_seed(seed)
    means = torch.randn(n, 3, generator=g) * 0.6 + torch.tensor([0.0, 0.0, 3.0])
    quats = torch.randn(n, 4, generator=g)
    scales = torch.rand(n, 3, generator=g) * (0.3 if big else 0.05) + 0.005
    opac = torch.rand(n, generator=g)
    colors = torch.rand(n, 3, generator=g)
    vm = torch.eye(4)[None].repeat(C, 1, 1)
    for c in range(C):
        vm[c, 0, 3] = 0.15 * c
        vm[c, 2, 3] = 0.1 * c
    K = torch.tensor([[90.0, 0, W / 2 + 0.5], [0, 95.0, H / 2 - 1.0], [0, 0, 1]])[None].repeat(C, 1, 1)
    return dict(means=means.numpy(), quats=quats.numpy(), scales=scales.numpy(), opacities=opac.numpy(),
                colors=colors.numpy(), viewmats=vm.numpy(), Ks=K.numpy(), width=W, height=H)


def allclose(a, b, atol=1e-5, rtol=1e-4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= atol + rtol * np.abs(b)))


def max_violation(a, b, atol=1e-5, rtol=1e-4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) - (atol + rtol * np.abs(b)))) if a.size else 0.0


def violations(a, b, atol=1e-5, rtol=1e-4):
    """Element-wise |a - b| <= atol + rtol |b| (the north_star's 1e-5 abs / 1e-4 rel): returns
    (share of violating entries, largest |a - b| / (atol + rtol |b|))."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0, 0.0
    ratio = np.abs(a - b) / (atol + rtol * np.abs(b))
    return float((ratio > 1.0).mean()), float(ratio.max())
