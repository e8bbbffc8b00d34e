"""`rasterization(...)` — the operator the reference binds to.

Same signature, argument meaning, return tuple and error behaviour as
`gsplat.rasterization` (gsplat 1.4.0) for the configuration Prometheus uses
(reference call site prometheus/modules/renderers/gaussians_renderer.py:413-427:
packed=False, sh_degree=None, pinhole, classic mode), computed by the sm_100a
kernels behind include/prometheus_gs.h. `meta` exposes every intermediate
upstream exposes (radii, means2d, depths, conics, opacities, tiles_per_gauss,
isect_ids, flatten_ids, isect_offsets, tile_width, ...) — they are the bit-exact
parity surface — and `meta["means2d"]` is a graph tensor that accepts
`retain_grad()` (used at gaussians_renderer.py:428-435 and by the SDS refiner).

Autograd is wired as two nodes like upstream (projection -> compositing), so
gradients reach means/quats/scales/opacities/colors/backgrounds.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import ops

_RENDER_MODES = ("RGB", "D", "ED", "RGB+D", "RGB+ED")

# test/bench hook: route every call through the generic emit + onesweep path
FORCE_GENERIC_SORT = False
# `rasterization(...)` renders with the fused sort-as-you-go kernel (same bits) and computes the fully sorted
# intersection arrays of `meta` only when one of them is read (_LazyMeta). "0": always run the full per-tile sort
# + composite_fwd (developer knob for A/B runs; bench.py's `operator_path.full_sort` figure).
OPERATOR_FUSED = __import__("os").environ.get("PM_OPERATOR_FUSED", "1") != "0"


class _LazyMeta(dict):
    """gsplat's `meta` dict. The reference reads `radii` and `means2d` from it (gaussians_renderer.py:428-445); the
    sorted intersection arrays `tiles_per_gauss`, `isect_ids`, `flatten_ids` — which the fused renderer never needs as
    a whole — are produced by the separate-stage kernels (projection with per-Gaussian tile counts, binning, full
    per-tile sort) the first time any of them is read through the mapping interface, and are then ordinary entries.
    Until then the keys exist (membership, len, keys() behave as usual)."""

    LAZY = frozenset(("tiles_per_gauss", "isect_ids", "flatten_ids"))

    def __init__(self, eager, thunk):
        super().__init__(eager)
        for k in self.LAZY:
            dict.__setitem__(self, k, None)
        self._thunk = thunk

    @property
    def materialized(self) -> bool:
        return self._thunk is None

    def materialize(self):
        thunk, self._thunk = self._thunk, None
        if thunk is not None:
            dict.update(self, thunk())
        return self

    def __getitem__(self, k):
        if k in self.LAZY:
            self.materialize()
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        if k in self.LAZY:
            self.materialize()
        return dict.get(self, k, default)

    def pop(self, k, *default):
        if k in self.LAZY:
            self.materialize()
        return dict.pop(self, k, *default)

    def __setitem__(self, k, v):
        if k in self.LAZY:
            self.materialize()
        dict.__setitem__(self, k, v)

    def __iter__(self):  # (also makes dict(meta) / {**meta} go through keys() + __getitem__ instead of a raw copy)
        return dict.__iter__(self)

    def items(self):
        return dict.items(self.materialize())

    def values(self):
        return dict.values(self.materialize())

    def copy(self):
        return dict(self.materialize())

    def __repr__(self):
        return dict.__repr__(self) if self.materialized else "_LazyMeta(sorted arrays pending; keys %s)" % sorted(self.keys())


def _new_touched(link, means2d):
    """Zero-filled pair map for the compositing backward, only when a projection node will consume it."""
    if link is None:
        return None
    C, N = means2d.shape[:2]
    return torch.zeros(C, N, dtype=torch.uint8, device=means2d.device)


def _leave_touched(link, touched, v_means2d, v_conics, v_depths):
    if link is not None and touched is not None:
        # valid only for exactly these gradient tensors (autograd may hand the projection node a SUM with other
        # contributions, e.g. a loss on meta["means2d"]; then the cotangents themselves are scanned)
        link["touched"] = (touched, v_means2d.data_ptr(), v_conics.data_ptr(), None if v_depths is None else v_depths.data_ptr())


def _take_touched(link, v_means2d, v_depths, v_conics):
    if not link or "touched" not in link:
        return None
    touched, p_m, p_c, p_d = link.pop("touched")
    same = (v_means2d is not None and v_means2d.data_ptr() == p_m and v_conics is not None and v_conics.data_ptr() == p_c
            and (None if v_depths is None else v_depths.data_ptr()) == p_d)
    return touched if same else None


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                radius_clip, fused_count=False, views_per_scene=0, link=None, tile_capacity=0, bin_opacities=None):
        """fused_count (module fast path): skip tiles_per_gauss, return the per-tile histogram instead.
        tile_capacity > 0 (with fused_count): single-pass binning — the intersection records are written as well, into
        fixed-capacity tile regions, and returned after the histogram. bin_opacities (not differentiated): bin only the
        tiles a splat can be seen in (ops.project_fwd).
        link: dict shared with the compositing node of the same render — its backward leaves the map of pairs that
        received a gradient there (`touched`), so that this node's backward reads 1 byte per pair to find its work."""
        ctx.set_materialize_grads(False)  # undefined cotangents arrive as None, not as zero tensors the size of the outputs
        ctx.link = link
        out = ops.project_fwd(means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                              radius_clip, count_tiles=not fused_count, tile_histogram=fused_count,
                              views_per_scene=views_per_scene, tile_capacity=tile_capacity if fused_count else 0,
                              opacities=bin_opacities if fused_count else None)
        ctx.views_per_scene = views_per_scene
        radii, means2d, depths, conics, tpg = out[:5]
        aux = out[5] if fused_count else tpg
        ctx.save_for_backward(means, quats, scales, viewmats, Ks, radii, conics)
        ctx.dims = (width, height, eps2d)
        if fused_count and tile_capacity:
            ctx.mark_non_differentiable(radii, aux, out[6])
            return means2d, depths, conics, radii, aux, out[6]
        ctx.mark_non_differentiable(radii, aux)
        return means2d, depths, conics, radii, aux

    @staticmethod
    def backward(ctx, v_means2d, v_depths, v_conics, _v_radii, _v_aux, _v_records=None):
        means, quats, scales, viewmats, Ks, radii, conics = ctx.saved_tensors
        width, height, eps2d = ctx.dims
        touched = _take_touched(ctx.link, v_means2d, v_depths, v_conics)
        v_means, v_quats, v_scales = ops.project_bwd(means, quats, scales, viewmats, Ks, width, height, radii,
                                                     conics, v_means2d, v_depths, v_conics, eps2d,
                                                     views_per_scene=ctx.views_per_scene, touched=touched)
        return (v_means, v_quats, v_scales) + (None,) * 13


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids, width,
                height, d_n, link=None):
        ctx.set_materialize_grads(False)  # undefined cotangents arrive as None, not as zero tensors the size of the outputs
        ctx.link = link
        means2d, conics = means2d.contiguous(), conics.contiguous()
        opacities = opacities.contiguous()
        colors = None if colors is None else colors.contiguous()
        depths = None if depths is None else depths.contiguous()
        backgrounds = None if backgrounds is None else backgrounds.contiguous().float()
        need_bwd = any(ctx.needs_input_grad[:6])  # then the forward also writes the splat stream the backward replays
        out = ops.composite_fwd(means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids, width,
                                height, d_n, want_stream=need_bwd)
        render, alphas, last_ids, trans = out[:4]
        ctx.save_for_backward(means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids,
                              alphas, last_ids, d_n, trans, out[4] if need_bwd else None)
        ctx.dims = (width, height)
        ctx.last_ids = last_ids
        return render, alphas.unsqueeze(-1)

    @staticmethod
    def backward(ctx, v_render, v_alphas):
        (means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids, alphas, last_ids,
         d_n, trans, stream) = ctx.saved_tensors
        width, height = ctx.dims
        v_alphas2 = None if v_alphas is None else v_alphas.squeeze(-1)
        touched = _new_touched(ctx.link, means2d)
        v_means2d, v_conics, v_colors, v_opac, v_depths = ops.composite_bwd(
            means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids, width, height, alphas,
            last_ids, v_render, v_alphas2, d_n, transmittance=trans, stream=stream, touched=touched)
        _leave_touched(ctx.link, touched, v_means2d, v_conics, v_depths)
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[5] and v_render is not None:
            D = backgrounds.shape[-1]
            v_bg = (v_render[..., :D] * trans.unsqueeze(-1)).sum(dim=(1, 2))
        return v_means2d, v_conics, v_opac, v_colors, v_depths, v_bg, None, None, None, None, None, None


class _TileRender(torch.autograd.Function):
    """Fused sort-as-you-go forward (ops.tile_render_fwd) + the ordinary compositing backward."""

    @staticmethod
    def forward(ctx, means2d, conics, opacities, colors, depths, backgrounds, records, offsets, width, height, link=None,
                tile_order=None):
        ctx.set_materialize_grads(False)  # undefined cotangents arrive as None, not as zero tensors the size of the outputs
        ctx.link = link
        means2d, conics = means2d.contiguous(), conics.contiguous()
        opacities = opacities.contiguous()
        colors = None if colors is None else colors.contiguous()
        depths = None if depths is None else depths.contiguous()
        backgrounds = None if backgrounds is None else backgrounds.contiguous().float()
        need_bwd = any(ctx.needs_input_grad[:6])
        out = ops.tile_render_fwd(records, offsets, means2d, conics, opacities, colors, depths, backgrounds, width, height,
                                  tile_order=tile_order, want_stream=need_bwd)
        render, alphas, last_ids, flat, sorted_counts, flag, trans = out[:7]
        ctx.save_for_backward(means2d, conics, opacities, colors, depths, backgrounds, offsets, flat, alphas, last_ids,
                              trans, out[7] if need_bwd else None)
        ctx.dims = (width, height)
        ctx.mark_non_differentiable(flag, flat, sorted_counts)
        return render, alphas.unsqueeze(-1), flag, flat, sorted_counts

    @staticmethod
    def backward(ctx, v_render, v_alphas, _vf, _vflat, _vsc):
        (means2d, conics, opacities, colors, depths, backgrounds, offsets, flat, alphas, last_ids,
         trans, stream) = ctx.saved_tensors
        width, height = ctx.dims
        v_alphas2 = None if v_alphas is None else v_alphas.squeeze(-1)
        touched = _new_touched(ctx.link, means2d)
        v_means2d, v_conics, v_colors, v_opac, v_depths = ops.composite_bwd(
            means2d, conics, opacities, colors, depths, backgrounds, offsets, flat, width, height, alphas, last_ids,
            v_render, v_alphas2, None, transmittance=trans, stream=stream, touched=touched)
        _leave_touched(ctx.link, touched, v_means2d, v_conics, v_depths)
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[5] and v_render is not None:
            D = backgrounds.shape[-1]
            v_bg = (v_render[..., :D] * trans.unsqueeze(-1)).sum(dim=(1, 2))
        return v_means2d, v_conics, v_opac, v_colors, v_depths, v_bg, None, None, None, None, None, None


class _TileRenderPacked(torch.autograd.Function):
    """Module fast path: fused sort-as-you-go forward with the packed epilogue (images = 2 rgb - 1 in CHW,
    expected depth, clamped alpha) + the compositing backward that takes the packed cotangents directly."""

    @staticmethod
    def forward(ctx, means2d, conics, opacities, colors, depths, backgrounds, records, offsets, stats, width, height,
                grad_sink=None, tile_order=None, views_per_scene=0, link=None, record_stride=0, capacity=None):
        ctx.set_materialize_grads(False)  # undefined cotangents arrive as None, not as zero tensors the size of the outputs
        ctx.link = link
        ctx.views_per_scene = views_per_scene
        ctx.grad_sink = grad_sink  # dict that receives the contiguous [C,N,2] means2d gradient (densification statistics)
        means2d, conics = means2d.contiguous(), conics.contiguous()
        opacities, colors, depths = opacities.contiguous(), colors.contiguous(), depths.contiguous()
        backgrounds = None if backgrounds is None else backgrounds.contiguous().float()
        # the exact final transmittance (4 B/pixel) and the splat stream the backward replays (48 B per staged
        # Gaussian) are written only when a backward will consume them
        need_bwd = any(ctx.needs_input_grad[:6])
        images, depth_out, masks, last_ids, flat, sorted_counts, flag, trans, stream, hit_words = ops.tile_render_packed_fwd(
            records, offsets, means2d, conics, opacities, colors, depths, backgrounds, width, height, d_n=stats,
            tile_order=tile_order, views_per_scene=views_per_scene, for_backward=need_bwd, record_stride=record_stride,
            capacity=capacity)
        ctx.save_for_backward(means2d, conics, opacities, colors, depths, backgrounds, offsets, flat, masks, depth_out,
                              last_ids, stats, sorted_counts, trans, stream, hit_words)
        ctx.dims = (width, height)
        ctx.mark_non_differentiable(flag, flat, sorted_counts)
        return images, depth_out, masks, flag, flat, sorted_counts

    @staticmethod
    def backward(ctx, v_images, v_depth_out, v_masks, _vf, _vflat, _vsc):
        (means2d, conics, opacities, colors, depths, backgrounds, offsets, flat, masks, depth_out, last_ids,
         stats, sorted_counts, trans, stream, hit_words) = ctx.saved_tensors
        width, height = ctx.dims
        # the backward replays what the forward sorted: tiles with the longest sorted prefix go first
        order = ops.tile_scan(sorted_counts, want_order=True)[2]
        touched = _new_touched(ctx.link, means2d)
        v_means2d, v_conics, v_colors, v_opac, v_depths = ops.composite_bwd_packed(
            means2d, conics, opacities, colors, depths, backgrounds, offsets, flat, width, height, masks, depth_out,
            last_ids, v_images, v_depth_out, v_masks, d_n=stats, tile_order=order, views_per_scene=ctx.views_per_scene,
            transmittance=trans, stream=stream, touched=touched, hit_words=hit_words)
        _leave_touched(ctx.link, touched, v_means2d, v_conics, v_depths)
        if ctx.grad_sink is not None:
            ctx.grad_sink["v_means2d"] = v_means2d
            for j, view in enumerate(ctx.grad_sink.get("views", ())):  # the reference's viewspace_points[j].grad
                g = v_means2d[j:j + 1]
                view.grad = g if view.grad is None else view.grad + g
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[5] and v_images is not None:
            v_bg = (2.0 * v_images * (trans.unsqueeze(1) if trans is not None else 1.0 - masks)).sum(dim=(2, 3))
        return (v_means2d, v_conics, v_opac, v_colors, v_depths, v_bg) + (None,) * 11


def render_views_packed(means, quats, scales, opacities, colors, viewmats, Ks, backgrounds, width, height,
                        capacity=None, near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, views_per_scene=0,
                        tile_capacity=None, usable_tiles=True):
    """Fast path behind GaussianRenderer: all C views of one scene, RGB + expected depth, outputs already in the
    module's layout. -> images [C,3,H,W] in [-1,1], depths [C,1,H,W], masks [C,1,H,W], info.

    `capacity` = speculative size of the intersection buffers. None: read the exact count from the device
    (one host sync between the tile scan and the scatter, as gsplat does). An int: no host sync at all — the
    count stays on the device (info["stats"] = {n_isects, max tile count}); the caller must check
    n_isects <= capacity afterwards and re-render otherwise (GaussianRenderer does).

    `tile_capacity` = T (only together with `capacity`): SINGLE-PASS BINNING — the projection kernel writes every
    intersection record straight into its tile's region of T records (pm_project_scatter_fwd), so the second pass over
    the pairs (pm_bin_scatter) disappears. Speculative like `capacity`: the caller must also check
    info["stats"][1] (the largest tile) <= T afterwards and re-render without it otherwise.
    `usable_tiles`: bin a splat only into the tiles of gsplat's AABB where it can reach
    alpha >= 1/255 (project.cu: usable_tiles) — same images / gradients bit for bit; the private lists
    (info["stats"], "isect_offsets", "flatten_ids", "sorted_counts") then describe a SUBSET of gsplat's intersections,
    in the same order. False: gsplat's intersections exactly.

    `views_per_scene` = V > 0: SCENE BATCHING — the Gaussian tensors hold C / V scenes back to back (N Gaussians
    each, N % 256 == 0) and camera c renders scene c / V: one launch set for a whole training batch."""
    C = viewmats.shape[0]
    if not means.is_cuda:
        raise RuntimeError("prometheus_b200 needs CUDA tensors: there is no CPU fallback")
    f32 = lambda t: (t if t.dtype == torch.float32 else t.float()).contiguous()
    means, quats, scales, opacities, colors, viewmats, Ks = map(f32, (means, quats, scales, opacities, colors, viewmats, Ks))
    link = {}  # compositing backward -> projection backward: the map of pairs that received a gradient
    single_pass = bool(tile_capacity) and capacity is not None
    proj = _Project.apply(means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
                          True, views_per_scene, link, int(tile_capacity) if single_pass else 0,
                          opacities.detach() if usable_tiles else None)
    means2d, depths, conics, radii, counts = proj[:5]
    # The reference keeps one [1,N,2] screen-space tensor per rendered view and reads its .grad after backward
    # (`viewspace_points`, gaussians_renderer.py:428-445). Routing autograd through per-view tensors costs a cat and a
    # stack of the whole [C,N,2] array each way; instead the per-view tensors are leaf views of the projected means
    # and the compositing backward hands each its slice of the contiguous gradient (no copy).
    grad_sink = {}
    if means2d.requires_grad:
        base = means2d.detach()
        means2d_views = [base[j:j + 1].requires_grad_(True) for j in range(C)]
        grad_sink["views"] = means2d_views
    else:
        means2d_views = [v.unsqueeze(0) for v in means2d.unbind(0)]
    isect_offsets, stats, tile_order = ops.tile_scan(counts, want_order=True)
    exact_counts = None
    if capacity is None:
        # the one host read of the exact-size path brings both counts: the caller can size its NEXT call of this
        # shape (e.g. the next chunk of a trajectory) at once instead of waiting for this call's deferred check
        exact_counts = tuple(int(v) for v in stats.tolist())
        capacity = exact_counts[0]
    if single_pass:
        records = proj[5]
    else:
        records = ops.bin_scatter(means2d.detach(), radii, depths.detach(), isect_offsets, width, height, int(capacity),
                                  conics=conics.detach() if usable_tiles else None,
                                  opacities=opacities.detach() if usable_tiles else None, views_per_scene=views_per_scene)
    images, depth_out, masks, flag, flatten_ids, sorted_counts = _TileRenderPacked.apply(
        means2d, conics, opacities, colors, depths, backgrounds, records, isect_offsets, stats, width, height, grad_sink,
        tile_order, views_per_scene, link, int(tile_capacity) if single_pass else 0, int(capacity))
    info = {"radii": radii, "means2d": means2d, "means2d_views": means2d_views, "depths": depths, "conics": conics,
            "isect_offsets": isect_offsets, "flatten_ids": flatten_ids, "sorted_counts": sorted_counts,
            "stats": stats, "overflow_flag": flag, "capacity": int(capacity), "means2d_grad_sink": grad_sink,
            "tile_capacity": int(tile_capacity) if single_pass else 0, "usable_tiles": bool(usable_tiles),
            "exact_counts": exact_counts}
    return images, depth_out, masks, info


# Largest per-tile group the fused kernel sorts at once (tile_render.cu: kTrCap)
FUSED_GROUP_CAP = 256  # tiles at or below this size can never overflow a sort group (kTrCap = 768): no flag read-back


def isect_tiles(means2d, radii, depths, tiles_per_gauss, width, height, C, force_generic=False):
    """Tile binning + depth sort (no gradient). Returns sorted isect_ids, flatten_ids and
    isect_offsets. One host sync: n_isects sizes the output arrays, as upstream.

    Main path (binned): per-tile histogram (from the projection) -> offsets -> scatter into tile
    segments -> per-tile shared-memory sort. Generic path (any segment larger than the shared-memory
    capacity, or no histogram): emit + hand-written onesweep radix sort + offset encode. Both produce
    the same arrays bit for bit."""
    if not force_generic:
        offsets, stats = ops.tile_scan(ops.bin_count(means2d, radii, width, height))
        n_isects, max_count = stats.tolist()
        records = ops.bin_scatter(means2d, radii, depths, offsets, width, height, n_isects)
        ids, flat, flag = ops.tile_sort(records, offsets, C, width, height, max_tile_count=max_count)
        # only tiles above the one-pass capacity can overflow (a depth sub-bucket that does not fit)
        if max_count <= ops.tile_sort_capacity() or int(flag.item()) == 0:
            return ids, flat, offsets
    cum, n_dev = ops.scan_tiles(tiles_per_gauss)
    n_isects = int(n_dev.item())
    tw, th = ops.tile_grid(width, height)
    ids, flat = ops.bin_emit(means2d, radii, depths, cum, width, height, n_isects)
    tb, cb = ops.key_bits(tw, th, C)
    ids, flat, ws = ops.sort_pairs(ids, flat, 32 + tb + cb, return_workspace=True)
    offsets = ops.tile_offsets(ids, C, width, height)
    # the onesweep look-back gives up after a bounded spin and raises a flag instead of hanging; a timed-out sort
    # must never be returned as if it were sorted (this generic path already synchronises on n_isects)
    if ops.sort_error_flag(ws):
        raise RuntimeError("pm_sort_pairs: decoupled look-back timed out; the sorted arrays are invalid")
    return ids, flat, offsets


def rasterization(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]
    scales: Tensor,  # [N, 3]
    opacities: Tensor,  # [N]
    colors: Tensor,  # [(C,) N, D]
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: str = "pinhole",
    covars: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    return _rasterize_core(means, quats, scales, opacities, colors, viewmats, Ks, width, height, near_plane,
                           far_plane, radius_clip, eps2d, sh_degree, packed, tile_size, backgrounds, render_mode,
                           sparse_grad, absgrad, rasterize_mode, channel_chunk, distributed, camera_model,
                           covars, per_view_means2d=False, fused=OPERATOR_FUSED, lazy_meta=True)


def _rasterize_core(means, quats, scales, opacities, colors, viewmats, Ks, width, height, near_plane=0.01,
                    far_plane=1e10, radius_clip=0.0, eps2d=0.3, sh_degree=None, packed=True, tile_size=16,
                    backgrounds=None, render_mode="RGB", sparse_grad=False, absgrad=False,
                    rasterize_mode="classic", channel_chunk=32, distributed=False, camera_model="pinhole",
                    covars=None, per_view_means2d=False, fused=False, lazy_meta=False):
    """Body of `rasterization`. `per_view_means2d=True` (used by renderer.GaussianRenderer)
    additionally returns meta["means2d_views"]: one [1,N,2] graph tensor per camera with
    retain_grad() set, the batched equivalent of gaussians_renderer.py:428-445.
    `fused=True` (module-level fast path) renders with the sort-as-you-go kernel: images are
    bit-identical, but meta["isect_ids"] is None and meta["flatten_ids"] is sorted only up to
    meta["sorted_counts"] entries per tile — unless `lazy_meta=True` (what `rasterization` passes): then `meta` is a
    _LazyMeta whose tiles_per_gauss / isect_ids / flatten_ids are gsplat's fully sorted arrays, computed on first read."""
    N = means.shape[0]
    C = viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    assert render_mode in _RENDER_MODES, render_mode
    assert (colors.dim() == 2 and colors.shape[0] == N) or (colors.dim() == 3 and colors.shape[:2] == (C, N)), colors.shape
    if backgrounds is not None:
        assert backgrounds.shape == (C, colors.shape[-1]), backgrounds.shape
    # what this drop-in covers is exactly the reference's configuration of the path
    if packed:
        raise NotImplementedError("prometheus_b200.rasterization: packed=True is not implemented "
                                  "(the reference calls with packed=False, gaussians_renderer.py:425)")
    for name, bad in (("sh_degree", sh_degree is not None), ("sparse_grad", sparse_grad), ("absgrad", absgrad),
                      ("distributed", distributed), ("covars", covars is not None),
                      ("rasterize_mode", rasterize_mode != "classic"), ("camera_model", camera_model != "pinhole"),
                      ("tile_size", tile_size != 16)):
        if bad:
            raise NotImplementedError(f"prometheus_b200.rasterization: {name} is outside the reference's path")
    if not means.is_cuda:
        raise RuntimeError("prometheus_b200.rasterization needs CUDA tensors: there is no CPU fallback")

    f32 = lambda t: t if t.dtype == torch.float32 else t.float()
    means, quats, scales, opacities, colors = map(f32, (means, quats, scales, opacities, colors))
    viewmats, Ks = f32(viewmats).contiguous(), f32(Ks).contiguous()

    use_fused = fused and not FORCE_GENERIC_SORT
    # compositing backward -> projection backward: the map of pairs that received a gradient. Not with per-view
    # means2d tensors (their gradients reach the projection node through a cat, as new tensors).
    link = None if per_view_means2d else {}
    means2d, depths, conics, radii, aux = _Project.apply(
        means.contiguous(), quats.contiguous(), scales.contiguous(), viewmats, Ks, width, height, eps2d,
        near_plane, far_plane, radius_clip, use_fused, 0, link)
    tpg = None if use_fused else aux
    with_depth = render_mode in ("D", "ED", "RGB+D", "RGB+ED")
    only_depth = render_mode in ("D", "ED")
    D = colors.shape[-1]
    if (0 if only_depth else D) + (1 if with_depth else 0) > 4:
        raise NotImplementedError("prometheus_b200.rasterization: at most 4 render channels (RGB + depth)")

    means2d_views = None
    means2d_in = means2d
    if per_view_means2d:
        means2d_views = [v.unsqueeze(0) for v in means2d.unbind(0)]
        if means2d.requires_grad:
            for v in means2d_views:
                v.retain_grad()
            means2d_in = torch.cat(means2d_views, dim=0)
    comp_args = (means2d_in, conics, opacities, None if only_depth else colors, depths if with_depth else None,
                 None if only_depth else backgrounds)

    render = None
    sorted_counts = None
    if use_fused:
        m2d, dep = means2d.detach(), depths.detach()
        # aux = per-tile histogram, counted inside the projection; tiles are rendered heaviest first
        isect_offsets, stats, tile_order = ops.tile_scan(aux, want_order=True)
        n_isects, max_count = stats.tolist()
        records = ops.bin_scatter(m2d, radii, dep, isect_offsets, width, height, n_isects)
        render, alphas, flag, flatten_ids, sorted_counts = _TileRender.apply(*comp_args, records, isect_offsets, width, height, link,
                                                                             tile_order)
        isect_ids = None
        # a depth sub-bucket larger than the group capacity can only exist if some tile is larger than it
        if max_count > FUSED_GROUP_CAP and int(flag.item()) != 0:
            render = None
    if render is None:
        if tpg is None:  # the fused attempt overflowed a sort group: the separate stages need the per-Gaussian counts
            with torch.no_grad():
                tpg = ops.project_fwd(means.detach(), quats.detach(), scales.detach(), viewmats, Ks, width, height, eps2d,
                                      near_plane, far_plane, radius_clip, count_tiles=True)[4]
        isect_ids, flatten_ids, isect_offsets = isect_tiles(means2d.detach(), radii, depths.detach(), tpg, width,
                                                            height, C, force_generic=FORCE_GENERIC_SORT)
        sorted_counts = None
        render, alphas = _Composite.apply(*comp_args, isect_offsets, flatten_ids, width, height, None, link)
    if render_mode in ("ED", "RGB+ED"):
        render = torch.cat([render[..., :-1], render[..., -1:] / alphas.clamp(min=1e-10)], dim=-1)

    tw, th = ops.tile_grid(width, height)
    meta = {
        "camera_ids": None, "gaussian_ids": None, "radii": radii, "means2d": means2d, "depths": depths,
        "conics": conics, "opacities": opacities.unsqueeze(0).expand(C, N), "tile_width": tw, "tile_height": th,
        "tiles_per_gauss": tpg, "isect_ids": isect_ids, "flatten_ids": flatten_ids,
        "isect_offsets": isect_offsets, "width": width, "height": height, "tile_size": tile_size,
        "n_cameras": C,
    }
    if means2d_views is not None:
        meta["means2d_views"] = means2d_views
    if sorted_counts is not None and lazy_meta:
        # the fused renderer produced the images; gsplat's sorted arrays follow on demand from the same projection
        proj_args = (means.detach(), quats.detach(), scales.detach(), viewmats, Ks, width, height, eps2d, near_plane,
                     far_plane, radius_clip)
        m2d, rad, dep = means2d.detach(), radii, depths.detach()

        def sorted_arrays():
            with torch.no_grad():
                tiles_per_gauss = ops.project_fwd(*proj_args, count_tiles=True)[4]
                ids, flat, _ = isect_tiles(m2d, rad, dep, tiles_per_gauss, width, height, C, force_generic=FORCE_GENERIC_SORT)
            return {"tiles_per_gauss": tiles_per_gauss, "isect_ids": ids, "flatten_ids": flat}

        return render, alphas, _LazyMeta(meta, sorted_arrays)
    if sorted_counts is not None:
        meta["sorted_counts"] = sorted_counts
    return render, alphas, meta
