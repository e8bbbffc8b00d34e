"""Level-1 drop-in, measured the way an unedited Prometheus checkout would use it.

The reference's `GaussianRenderer.forward` (prometheus/modules/renderers/gaussians_renderer.py:475-512) renders a batch
as a host loop over (scene, view): per view it pulls the camera to the CPU (`cameras.cpu()`, :455 — a device sync),
rescales the intrinsics and inverts the pose with numpy (:453-472, MiniCam :139-169), builds K on the host with the
principal point at the image centre (:404-412), calls `rasterization(...)` for ONE camera with render_mode="RGB+ED",
packed=False (:413-427), keeps `info["means2d"]` / `info["radii"]` (:428-445), and at the end stacks the views, maps the
images to [-1, 1] and calls torch.cuda.synchronize() (:507-510).

`reference_style_forward` restates that control flow — same host round trips, same per-view C = 1 operator calls, same
final synchronize — on top of `prometheus_b200.rasterization`, i.e. what a user gets from `install_as_gsplat()` WITHOUT
switching `gs_decoder.py:62-63` to the batched `prometheus_b200.GaussianRenderer`. /root/reference does not exist on the
GPU box, so the loop is restated here instead of imported; tests/test_gpu_parity.py checks it against the batched module.
Used by bench.py (`operator_dropin`) and by that test only.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def host_camera(camera: torch.Tensor, h: int, w: int):
    """One camera [18] on the device -> (viewmat [4,4], K [3,3]) on the device, through the host like the reference:
    .cpu() (sync), intrinsics rescaled to the render size, fov = 2 atan(0.5 size / f) with the reference's h <-> w
    pairing (:469-470), OpenGL -> OpenCV flip, numpy inverse, K rebuilt from the fovs with a centred principal point."""
    dev = camera.device
    cam = camera.detach().cpu()
    c2w = torch.eye(4)
    c2w[:3, :] = cam[:12].reshape(3, 4)
    fx, fy, _cx, _cy, H, W = [cam[12 + i:13 + i] for i in range(6)]
    fx, fy = fx * h / H, fy * w / W
    fovy = (2 * torch.atan(0.5 * w / fy)).numpy()
    fovx = (2 * torch.atan(0.5 * h / fx)).numpy()
    m = c2w.numpy()
    m[:3, 1:3] *= -1
    w2c = np.linalg.inv(m)
    viewmat = torch.tensor(w2c).to(dev)
    fl_x = w / (2 * math.tan(float(fovx[0]) * 0.5))
    fl_y = h / (2 * math.tan(float(fovy[0]) * 0.5))
    K = torch.tensor([[fl_x, 0, w / 2.0], [0, fl_y, h / 2.0], [0, 0, 1]]).to(dev)
    return viewmat, K


def reference_style_forward(rasterization, cameras: torch.Tensor, gaussians, bg: torch.Tensor, h: int, w: int):
    """cameras [B,N,18], gaussians = the converter's 5-tuple, bg [3] on the device ->
    (images [B,N,3,h,w] in [-1,1], depths [B,N,1,h,w], masks [B,N,1,h,w], radii list, viewspace_points list)."""
    B, N = cameras.shape[:2]
    xyz, features, opacity, scales, rotations = gaussians
    radii, viewspace, images, depths, masks = [], [], [], [], []
    for i in range(B):
        for j in range(N):
            viewmat, K = host_camera(cameras[i, j], h, w)
            colors, alphas, info = rasterization(
                means=xyz[i], quats=rotations[i], scales=scales[i], opacities=opacity[i].squeeze(-1),
                colors=features[i][:, 0], viewmats=viewmat[None], Ks=K[None], backgrounds=bg[None],
                width=int(w), height=int(h), render_mode="RGB+ED", packed=False)
            try:
                info["means2d"].retain_grad()
            except Exception:  # noqa: BLE001  (no graph under no_grad, as in the reference's bare except)
                pass
            viewspace.append(info["means2d"])
            radii.append(info["radii"].squeeze(0))
            images.append(colors[0, :, :, :3].permute(2, 0, 1))
            depths.append(colors[0, :, :, 3:].permute(2, 0, 1))
            masks.append(alphas[0].permute(2, 0, 1).clamp(0, 1))
    images = torch.stack(images, dim=0).unflatten(0, (B, N)) * 2 - 1
    depths = torch.stack(depths, dim=0).unflatten(0, (B, N))
    masks = torch.stack(masks, dim=0).unflatten(0, (B, N))
    torch.cuda.synchronize()
    return images, depths, masks, radii, viewspace
