"""Module-level mirror of the reference's renderer module
(prometheus/modules/renderers/gaussians_renderer.py): `GaussianConverter` and
`GaussianRenderer` with the same constructor arguments, forward signatures,
buffer names (`converter.muls`, `renderer.bg_color` — checkpoints are loaded with
strict=True in paper/stage1_eval.py:167), return tuples and side attributes
(`.radii`, `.viewspace_points`), computed B200-natively:

  * GaussianConverter.forward = ONE fused unproject kernel (fwd) / ONE (bwd)
    instead of ~40 ATen launches (gaussians_renderer.py:210-287);
  * GaussianRenderer.forward renders all N views of a scene in one batched launch
    set (C = N cameras), with the camera conversion done on the device — none of
    the reference's per-view `.cpu()` round trips, numpy inverses, host-built K
    uploads or the trailing torch.cuda.synchronize() (:453-472, :404-412, :510).
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import ops
from .rasterization import FUSED_GROUP_CAP, _rasterize_core, rasterization, render_views_packed


class _Unproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, cameras, muls, mode, z_near, z_far, s_min, s_max):
        ctx.set_materialize_grads(False)  # undefined cotangents arrive as None, not as zero tensors the size of the outputs
        # fp16 / bf16 maps (an autocast decoder) are read in their storage type; the reference widens them to fp32
        # by `* self.muls` (:219), and so does the kernel, in registers
        params = params.contiguous() if params.dtype in ops.PARAM_DTYPES else params.float().contiguous()
        cameras = cameras.float().contiguous()
        out = ops.unproject_fwd(params, cameras, muls, mode, z_near, z_far, s_min, s_max)
        ctx.save_for_backward(params, cameras, muls)
        ctx.cfg = (mode, z_near, z_far, s_min, s_max)
        return out

    @staticmethod
    def backward(ctx, v_xyz, v_rgb, v_opacity, v_scales, v_quats):
        params, cameras, muls = ctx.saved_tensors
        v_params = ops.unproject_bwd(params, cameras, muls, *ctx.cfg, v_xyz, v_rgb, v_opacity, v_scales, v_quats)
        return v_params, None, None, None, None, None, None, None


class GaussianConverter(nn.Module):
    """Drop-in for the reference class of the same name (gaussians_renderer.py:173-287)."""

    def __init__(self, gs_convert_mode: str = "mvsplat", z_near: float = 0.001, z_far: float = 100,
                 s_min: float = 0.001, s_max: float = 1.0):
        super().__init__()
        if gs_convert_mode not in ops.MODES:
            raise ValueError("unsupport decode type")
        self.gs_convert_mode = gs_convert_mode
        self.z_near, self.z_far = z_near, z_far
        self.s_min, self.s_max = s_min, s_max
        self.gaussian_channels = [3, 2, 1, 1, 3, 4]
        if gs_convert_mode == "director3d":
            muls = torch.Tensor([0.01] * 3 + [0.01] * 2 + [0.05] * 1 + [0.05] * 1 + [0.005] * 3 + [0.005] * 4).reshape(1, -1)
            self.register_buffer("muls", muls)
            self.muls = self.muls / self.muls.max()
        else:
            self.register_buffer("muls", torch.ones(1, 14))

    def forward(self, local_gaussian_params, cameras):
        B, N, C, h, w = local_gaussian_params.shape
        assert C == sum(self.gaussian_channels)
        xyz, rgb, opacity, scales, quats = _Unproject.apply(
            local_gaussian_params.reshape(B * N, C, h, w), cameras.reshape(B * N, 18), self.muls,
            self.gs_convert_mode, float(self.z_near), float(self.z_far), float(self.s_min), float(self.s_max))
        return (xyz.reshape(B, -1, 3), rgb.reshape(B, -1, 1, 3), opacity.reshape(B, -1, 1),
                scales.reshape(B, -1, 3), quats.reshape(B, -1, 4))


class GaussianRenderer(nn.Module):
    """Drop-in for the reference class of the same name (gaussians_renderer.py:289-512)."""

    # cameras rendered per batched launch set: bounds the per-pair / per-intersection buffers
    # (32 B per (camera, Gaussian) pair) when a long trajectory is rendered from a big scene
    view_chunk = 8
    # render with the fused sort-as-you-go kernel (tile_render.cu) and its packed epilogue: bit-identical
    # images, but only the part of each tile's list that the compositor consumes gets sorted, and the module's
    # output layout comes straight out of the kernel
    fused = True
    # fused path: size the intersection buffers from the previous call of the same shape (x slack) instead of
    # reading the count back between two kernels; the count is checked once all launches are queued
    speculative_capacity = True
    capacity_slack = 1.3
    # fused path, once a shape's counts are known: SINGLE-PASS BINNING — the projection kernel writes each intersection
    # record into its tile's fixed-capacity region (largest tile of the previous call x slack), so the second pass over
    # the pairs (bin_scatter) disappears. Checked like the capacity (largest tile <= region size); used only while the
    # regions stay within `single_pass_max_ratio` x the compact record array (uniformly loaded tiles)
    single_pass_binning = os.environ.get("PM_SINGLE_PASS_BINNING", "1") != "0"  # (env: developer knob for A/B runs)
    tile_capacity_slack = 1.25
    single_pass_max_ratio = 4.0
    # single-pass binning bins a splat only into the tiles where it can reach alpha >= 1/255 (project.cu: usable_tiles):
    # ~19 % fewer intersections on the benchmark scenes, images and gradients unchanged bit for bit
    usable_tiles = os.environ.get("PM_USABLE_TILES", "1") != "0"  # (env: developer knob for A/B runs)
    # False: forward() returns only after its checks passed (like the reference, which ends with
    # torch.cuda.synchronize(), :510). True: forward() never blocks; call validate() before trusting the images
    # (DecodeRenderPipeline does, when it hands the host copies out).
    defer_checks = False
    # several scenes per forward (a training batch): render them in ONE launch set when every scene's views fit a
    # chunk and the per-scene Gaussian count is a multiple of 256 (always, for pixel-aligned maps)
    scene_batch = True
    scene_chunk = 8  # scenes per launch set
    STATS_RING = 4096  # pinned {n_isects, max tile count, overflow} rows for the deferred checks

    def __init__(self, sh_degree=0, background=[1, 1, 1]):
        super().__init__()
        self.sh_degree = sh_degree
        self.register_buffer("bg_color", torch.tensor(background).float())
        self.radii: List[torch.Tensor] = []
        self.viewspace_points: List[torch.Tensor] = []
        self._cap_hint = {}
        self._tile_cap_hint = {}
        self.last_stats = None
        self._pending = []
        self._static_stats = None
        self._stats_ring = None
        self._stats_next = 0
        self._stat_sources = []
        self._stat_scenes = 0

    def _backgrounds(self, bg_color, n_views: int, device) -> torch.Tensor:
        # reference :381-386 — 'random' draws a fresh colour per rendered view
        if isinstance(bg_color, str) and bg_color == "random":
            return torch.rand(n_views, 3, device=device)
        if bg_color is None:
            return self.bg_color.to(device=device, dtype=torch.float32).expand(n_views, 3).contiguous()
        if torch.is_tensor(bg_color):
            return bg_color.to(device=device, dtype=torch.float32).reshape(1, 3).expand(n_views, 3).contiguous()
        # a tuple / list: materialised on the device ONCE per colour — a pageable host->device copy per call would
        # synchronise, and is illegal inside a CUDA-graph capture (graphed.py replays this path)
        key = (tuple(float(v) for v in bg_color), str(device))
        cache = self.__dict__.setdefault("_bg_cache", {})
        if key not in cache:
            cache[key] = torch.tensor(key[0], dtype=torch.float32, device=device).reshape(1, 3)
        return cache[key].expand(n_views, 3).contiguous()

    # ---------------------------------------------------------------- separate-stage path (fused = False)
    def _render_chunk_unfused(self, xyz, rot, sc, opac, feat, viewmats, Ks, bgs, h, w):
        render, alphas, info = _rasterize_core(
            means=xyz, quats=rot, scales=sc, opacities=opac, colors=feat, viewmats=viewmats, Ks=Ks, backgrounds=bgs,
            width=int(w), height=int(h), render_mode="RGB+ED", packed=False, per_view_means2d=True, fused=False)
        images = render[..., :3].permute(0, 3, 1, 2) * 2 - 1
        return images, render[..., 3:].permute(0, 3, 1, 2), alphas.permute(0, 3, 1, 2).clamp(0, 1), info

    # ---------------------------------------------------------------- fused path
    def _render_chunk_fused(self, args, h, w):
        xyz, rot, sc, opac, feat, viewmats, Ks, bgs = args
        key = (xyz.shape[0], viewmats.shape[0], int(h), int(w))
        cap = self._cap_hint.get(key) if self.speculative_capacity else None
        images, depths, masks, info = render_views_packed(xyz, rot, sc, opac, feat, viewmats, Ks, bgs, int(w), int(h),
                                                          capacity=cap, tile_capacity=self._tile_cap_for(key, cap, h, w),
                                                          usable_tiles=self.usable_tiles)
        self._hint_from_exact(key, cap, info)
        return images, depths, masks, info, key

    def _hint_from_exact(self, key, cap, info) -> None:
        """A first call of a shape ran with exact sizes and read {n_isects, max tile count} on the way: set the hints
        now, so that the following chunks of the same forward (a trajectory: 15 chunks of one shape) already run
        speculatively — single-pass binning, no host read — instead of all taking the exact path."""
        if cap is None and self.speculative_capacity and info.get("exact_counts") and key not in self._cap_hint:
            self._update_hint(key, *info["exact_counts"])

    def _tile_cap_for(self, key, cap, h, w):
        """Region size for single-pass binning, or None (two passes): needs both hints of the shape."""
        tcap = self._tile_cap_hint.get(key) if (self.single_pass_binning and cap is not None) else None
        if tcap is None:
            return None
        tiles = key[1] * ((int(h) + 15) // 16) * ((int(w) + 15) // 16)
        if tiles * tcap > self.single_pass_max_ratio * cap or tcap >= 2 ** 31:
            return None
        return tcap

    def _update_hint(self, key, n_isects: int, max_count: int = 0) -> None:
        """Capacity for the next call of this shape: the latest count with slack, but never much below the
        previous hint — along a camera trajectory the count varies from chunk to chunk and must not oscillate."""
        want = int(n_isects * self.capacity_slack) + 4096
        self._cap_hint[key] = max(want, int(self._cap_hint.get(key, 0) * 0.97))
        want_t = int(max_count * self.tile_capacity_slack) + 64
        self._tile_cap_hint[key] = max(want_t, int(self._tile_cap_hint.get(key, 0) * 0.97))

    def _check(self, job) -> bool:
        """True if the fused render of `job` is valid; updates the capacity hint."""
        key = job["key"]
        if job["event"] is not None:
            job["event"].synchronize()  # fires when the render kernel is done; later work (a backward pass) keeps running
        n_isects, max_count, overflow = (int(v) for v in job["stats_host"])
        self.last_stats = (n_isects, max_count, overflow)  # of the fast path's OWN lists (usable tiles: a subset of gsplat's)
        self._update_hint(key, n_isects, max_count)
        if n_isects > job["capacity"]:
            return False
        if job.get("tile_capacity") and max_count > job["tile_capacity"]:
            return False  # single-pass binning dropped records of an over-full tile
        # a depth sub-bucket larger than the group capacity can only exist if some tile is larger than it
        return max_count <= FUSED_GROUP_CAP or overflow == 0

    def take_pending(self) -> list:
        """Deferred-check mode: hand the unchecked jobs of the forward() calls so far to the caller
        (see check_jobs)."""
        jobs, self._pending = self._pending, []
        return jobs

    def check_jobs(self, jobs) -> bool:
        return all([self._check(j) for j in jobs])

    def validate(self) -> bool:
        """Deferred-check mode: True if every fused render since the last call was valid (otherwise re-run
        forward() with defer_checks = False)."""
        return self.check_jobs(self.take_pending())

    def _job_for(self, args, out, info, key):
        """Queue the device-side counts of a fused render for the later check: {n_isects, max tile count, overflow
        flag} travel to pinned memory behind the render kernel."""
        if self._static_stats is not None:
            # under CUDA-graph capture (graphed.py): a preallocated pinned row per job, no event — the two copies
            # become nodes of the graph and the caller synchronises on its own event before looking at the row
            rows, used = self._static_stats
            stats_host, ev = rows[used[0]], None
            used[0] += 1
        else:
            # a ring of pinned rows instead of one pinned allocation per job: the host allocator would otherwise
            # fall back to cudaHostAlloc (milliseconds, and it serialises with the device) whenever its free blocks
            # are still waiting for their events
            if self._stats_ring is None:
                self._stats_ring = torch.zeros(self.STATS_RING, 3, dtype=torch.int32).pin_memory()
            if len(self._pending) < self.STATS_RING - 64:
                stats_host = self._stats_ring[self._stats_next]
                self._stats_next = (self._stats_next + 1) % self.STATS_RING
            else:  # thousands of unchecked jobs outstanding: never overwrite a row that has not been looked at
                stats_host = torch.empty(3, dtype=torch.int32, pin_memory=True)
            ev = torch.cuda.Event()
        stats_host[:2].copy_(info["stats"], non_blocking=True)
        stats_host[2:].copy_(info["overflow_flag"], non_blocking=True)
        if ev is not None:
            ev.record()
        return dict(args=args, out=out, info=info, key=key, stats_host=stats_host, event=ev, capacity=info["capacity"],
                    tile_capacity=info.get("tile_capacity", 0))

    def _forward_batched(self, viewmats, Ks, gaussians, scaling_modifier, bg_color, h, w):
        """All B scenes of a training batch in ONE launch set (scene batching, `views_per_scene`): the B x N cameras
        become one camera list, camera c sees the Gaussians of scene c // N. Returns None if the deferred-free
        check fails (the caller then renders scene by scene)."""
        B, N = viewmats.shape[:2]
        xyz, features, opacity, scales, rotations = gaussians
        G = xyz.shape[1]
        sc = scales  # scaling_modifier: see forward()
        args = (xyz.reshape(B * G, 3), rotations.reshape(B * G, 4), sc.reshape(B * G, 3), opacity.reshape(B * G),
                features[:, :, 0].reshape(B * G, 3), viewmats.reshape(B * N, 4, 4), Ks.reshape(B * N, 3, 3),
                self._backgrounds(bg_color, B * N, xyz.device))
        key = (G, B * N, int(h), int(w), B)
        cap = self._cap_hint.get(key) if self.speculative_capacity else None
        im, de, ma, info = render_views_packed(*args, int(w), int(h), capacity=cap, views_per_scene=N,
                                               tile_capacity=self._tile_cap_for(key, cap, h, w),
                                               usable_tiles=self.usable_tiles)
        self._hint_from_exact(key, cap, info)
        job = self._job_for(args, (im, de, ma), info, key)
        if self.defer_checks:
            self._pending.append(dict(key=key, capacity=job["capacity"], stats_host=job["stats_host"], event=job["event"],
                                      tile_capacity=job["tile_capacity"]))
        elif not self._check(job):
            return None
        self._stat_sources = [(info["radii"], info.get("means2d_grad_sink"), info["means2d_views"])]
        self._stat_scenes = B
        for j in range(B * N):
            self.radii.append(info["radii"][j])
            self.viewspace_points.append(info["means2d_views"][j])
        return (im.reshape(B, N, 3, h, w), de.reshape(B, N, 1, h, w), ma.reshape(B, N, 1, h, w), None, None)

    def forward(self, cameras, gaussians, scaling_modifier=1.0, bg_color=None, h=256, w=256):
        """`scaling_modifier` is accepted for signature parity and IGNORED, exactly like the reference's gsplat
        branch (gaussians_renderer.py:377-440 never reads it; only the diff_gaussian_rasterization fallback
        branch, :337, which is out of scope, does)."""
        B, N = cameras.shape[:2]
        xyz, features, opacity, scales, rotations = gaussians
        device = xyz.device
        self.radii = []
        self.viewspace_points = []
        viewmats, Ks = ops.camera_prep(cameras.reshape(B * N, 18), h, w)
        viewmats, Ks = viewmats.reshape(B, N, 4, 4), Ks.reshape(B, N, 3, 3)
        if self.fused and self.scene_batch and B > 1 and N <= self.view_chunk and xyz.shape[1] % 256 == 0:
            outs, sources = [], []
            for b0 in range(0, B, self.scene_chunk):  # bounded per-pair / per-intersection buffers for big batches
                b1 = min(B, b0 + self.scene_chunk)
                out = self._forward_batched(viewmats[b0:b1], Ks[b0:b1], tuple(t[b0:b1] for t in gaussians),
                                            scaling_modifier, bg_color, h, w)
                if out is None:
                    break
                outs.append(out)
                sources.extend(self._stat_sources)
            if len(outs) * self.scene_chunk >= B:
                self._stat_sources, self._stat_scenes = sources, B
                if len(outs) == 1:
                    return outs[0]
                return tuple(torch.cat([o[k] for o in outs], 0) for k in range(3)) + (None, None)
            self.radii, self.viewspace_points = [], []  # a check failed: scene by scene below (exact sizes, fallbacks)
        jobs = []
        G = xyz.shape[1]
        for i in range(B):
            bgs = self._backgrounds(bg_color, N, device)
            if B == 1:
                # views, not selects: the backward of t[i] materialises a zero tensor and copies the gradient into it
                # (two kernels per Gaussian tensor and step), the backward of a reshape is free
                gs_i = (xyz.reshape(G, 3), rotations.reshape(G, 4), scales.reshape(G, 3), opacity.reshape(G),
                        features.reshape(G, -1, 3)[:, 0] if features.shape[2] != 1 else features.reshape(G, 3))
            else:
                gs_i = (xyz[i], rotations[i], scales[i], opacity[i].squeeze(-1), features[i][:, 0])
            for c0 in range(0, N, self.view_chunk):
                c1 = min(N, c0 + self.view_chunk)
                args = gs_i + (viewmats[i, c0:c1], Ks[i, c0:c1], bgs[c0:c1])
                if self.fused:
                    im, de, ma, info, key = self._render_chunk_fused(args, h, w)
                    jobs.append(self._job_for(args, (im, de, ma), info, key))
                else:
                    im, de, ma, info = self._render_chunk_unfused(*args, h, w)
                    jobs.append(dict(args=args, out=(im, de, ma), info=info, key=None))
        if self.fused and self.defer_checks:
            # only what the check needs: holding the jobs' tensors would pin every step's buffers
            self._pending.extend(dict(key=j["key"], capacity=j["capacity"], stats_host=j["stats_host"], event=j["event"],
                                      tile_capacity=j["tile_capacity"]) for j in jobs)
        elif self.fused:
            # every launch is queued; now look at the counts (the GPU is still busy rendering)
            for job in jobs:
                if not self._check(job):
                    a = job["args"]
                    n_isects, max_count = int(job["stats_host"][0]), int(job["stats_host"][1])
                    if n_isects > job["capacity"] or (job["tile_capacity"] and max_count > job["tile_capacity"]):
                        # speculative buffers were too small: exact sizes (two binning passes) this time
                        self._cap_hint.pop(job["key"], None)
                        self._tile_cap_hint.pop(job["key"], None)
                        im, de, ma, info, _ = self._render_chunk_fused(a, h, w)
                        ok = int(info["stats"][1].item()) <= FUSED_GROUP_CAP or int(info["overflow_flag"].item()) == 0
                        self._update_hint(job["key"], n_isects, max_count)
                    else:
                        ok = False
                    if not ok:  # a depth sub-bucket beyond the fused kernel's group capacity: separate stages
                        im, de, ma, info = self._render_chunk_unfused(*a, h, w)
                    job["out"], job["info"] = (im, de, ma), info
        # (radii [C,N], sink of the contiguous means2d gradient) per rendered chunk, for update_densification_stats
        self._stat_sources = [(j["info"]["radii"], j["info"].get("means2d_grad_sink"), j["info"]["means2d_views"]) for j in jobs]
        self._stat_scenes = B
        images, depths, masks = [], [], []
        per_scene = len(jobs) // max(B, 1)
        for i in range(B):
            sj = jobs[i * per_scene:(i + 1) * per_scene]
            for job in sj:
                nv = job["out"][0].shape[0]
                for j in range(nv):
                    self.radii.append(job["info"]["radii"][j])
                    self.viewspace_points.append(job["info"]["means2d_views"][j])
            for lst, k in ((images, 0), (depths, 1), (masks, 2)):
                lst.append(sj[0]["out"][k] if len(sj) == 1 else torch.cat([j["out"][k] for j in sj], 0))
        pack = (lambda l: l[0].unsqueeze(0)) if B == 1 else (lambda l: torch.stack(l, dim=0))
        return pack(images), pack(depths), pack(masks), None, None


def _update_densification_stats(self, is_visible, max_radii2D, xyz_gradient_accum, denom) -> None:
    """After `loss.backward()`: the SDS++ refiner's per-view bookkeeping (prometheus/models/sdspp_refiner.py:275-280
    with GaussianModel.add_densification_stats, prometheus/utils/gs_utils.py:247-252) for every view of the last
    forward(), in one kernel per rendered chunk instead of ~10 boolean-mask kernels per view. The four tensors are
    the refiner's own state (`new_gaussians.is_visible`, `.max_radii2D`, `.xyz_gradient_accum`, `.denom`), updated
    in place. One scene per forward (the refiner's case)."""
    assert self._stat_scenes == 1, "densification statistics are kept per Gaussian set: render one scene per forward()"
    for radii, sink, views in self._stat_sources:
        grad = sink.get("v_means2d") if sink else None
        if grad is None:  # separate-stage path: gather the per-view gradients autograd left on the side tensors
            grad = torch.cat([v.grad if v.grad is not None else torch.zeros_like(v) for v in views], dim=0)
        ops.densify_stats(radii, grad, is_visible, max_radii2D, xyz_gradient_accum, denom)


GaussianRenderer.update_densification_stats = _update_densification_stats


def install_as_gsplat() -> None:
    """Make `from gsplat import rasterization` (gaussians_renderer.py:11) resolve to
    this package. Call before `prometheus.modules` is imported."""
    import sys
    import types

    mod = types.ModuleType("gsplat")
    mod.rasterization = rasterization
    mod.__version__ = "1.4.0+prometheus_b200"
    mod.__doc__ = "prometheus_b200 drop-in for the gsplat.rasterization operator"
    sys.modules["gsplat"] = mod
