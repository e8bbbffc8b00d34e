"""Loss-side glue of the stage-1 rendering loss, fused.

`render_loss_terms` replaces the elementwise / reduction ATen kernels the reference runs on the renderer's outputs
before its loss networks (prometheus/systems/gs_decoder_system.py:214-239) and their autograd mirror images with one
kernel each way; its backward emits the packed cotangents the compositing backward consumes, so between the loss
networks and the Gaussians there is no full-image ATen pass left. The loss networks themselves (LPIPS, the
scale-and-shift-invariant depth loss) stay the reference's: they take `images`, `disparity` and `valid` as before.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops


class _LossTerms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, depths, masks, target):
        ctx.set_materialize_grads(False)  # undefined cotangents arrive as None, not as zero tensors the size of the outputs
        shape = images.shape
        im, tg = images.reshape(-1, 3, *shape[-2:]), target.reshape(-1, 3, *shape[-2:])
        de, ma = depths.reshape(-1, 1, *shape[-2:]), masks.reshape(-1, 1, *shape[-2:])
        disparity, valid, sums = ops.loss_terms_fwd(im, tg, de, ma)
        ctx.save_for_backward(im, tg, de)
        ctx.shapes = (images.shape, depths.shape, masks.shape)
        n_pix = de.numel()
        mse = (sums[0] / (3 * n_pix)).float()
        entropy = (sums[1] / n_pix).float()
        ctx.mark_non_differentiable(valid)
        return mse, disparity.reshape(depths.shape), valid.reshape(depths.shape), entropy

    @staticmethod
    def backward(ctx, g_mse, v_disparity, _v_valid, g_entropy):
        im, tg, de = ctx.saved_tensors
        v_images, v_depths, v_masks = ops.loss_terms_bwd(
            im, tg, de, g_mse, g_entropy, None if v_disparity is None else v_disparity.reshape(de.shape))
        s_im, s_de, s_ma = ctx.shapes
        return v_images.reshape(s_im), v_depths.reshape(s_de), v_masks.reshape(s_ma), None


def render_loss_terms(images: torch.Tensor, depths: torch.Tensor, masks: torch.Tensor, target: torch.Tensor
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(images [..,3,H,W], depths [..,1,H,W], masks [..,1,H,W]) as GaussianRenderer returns them, target like images ->
        mse       = F.mse_loss(images, target)                       (gs_decoder_system.py:221)
        disparity = 1 / depths.clip(0.1, 100)                        (:218)
        valid     = ones_like(depths) with depths <= 0.1 zeroed      (:216-217; no gradient)
        entropy   = (1 - masks).mean()                               (:233)
    Differentiable w.r.t. images, depths and masks."""
    return _LossTerms.apply(images, depths, masks, target)
