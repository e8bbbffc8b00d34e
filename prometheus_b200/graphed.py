"""Decode + render of fixed-shape scenes as ONE CUDA-graph replay.

The fused path never talks to the host between its kernels (speculative intersection capacity, counts checked
afterwards), so the whole launch set — unproject, camera preparation, project + count, tile scan, scatter, fused
sort-and-composite, the memsets between them and the two tiny copies of the device-side counts — is static for a
given input shape and can be captured once and replayed. That removes the per-step Python / launch cost, which is
what bounds small scenes (config C1: 0.43 ms of kernels in a 0.53 ms eager step).

Inference only (the captured work runs under no_grad); training steps stay on the eager modules.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops
from .renderer import GaussianConverter, GaussianRenderer


class GraphedDecodeRender:
    """`run(params, cameras_in, cameras_out)` = `renderer(cameras_out, converter(params, cameras_in))[:3]` with the
    launch set replayed from a CUDA graph.

    Inputs may live on the device or in (pinned) host memory; they are copied into the graph's static input
    buffers on the current stream. The returned (images, depths, masks) are the graph's static output buffers:
    they are overwritten by the next run() of the same shape — clone them (or copy them out, as
    DecodeRenderPipeline does) before that. A new input shape or dtype captures a new graph.

    The capacity / sort-overflow checks of the fused path are deferred exactly as in
    GaussianRenderer.defer_checks: run(check=True) looks at them before returning (one event synchronise);
    run(check=False) returns at once and `validate()` does it later. A failed check (the scene produced more
    intersections than the capture was sized for) re-renders eagerly and re-captures with the larger size."""

    MAX_JOBS = 256  # fused launch sets per forward (scenes x view chunks) a capture may hold

    def __init__(self, converter: GaussianConverter, renderer: GaussianRenderer, h: int, w: int, bg_color=None):
        if isinstance(bg_color, str):
            raise ValueError("a random background is drawn on the host per call: not capturable, pass a colour")
        if not renderer.fused or not renderer.speculative_capacity:
            raise ValueError("CUDA-graph capture needs the fused path with speculative capacity (no mid-step host sync)")
        self.conv, self.rend = converter, renderer
        self.h, self.w, self.bg = int(h), int(w), bg_color
        self._plans = {}
        self._last = None
        self.n_captures = 0     # telemetry: graphs captured (first use of a shape / slot, and re-captures)
        self.n_failed_checks = 0  # replays whose deferred checks failed (each is re-rendered eagerly by the caller)

    # ------------------------------------------------------------------ capture
    def _eager(self, params, cin, cout):
        prev, self.rend.defer_checks = self.rend.defer_checks, False
        try:
            with torch.no_grad():
                return self.rend(cout, self.conv(params, cin), bg_color=self.bg, h=self.h, w=self.w)[:3]
        finally:
            self.rend.defer_checks = prev

    def _capture(self, params, cin, cout, static_in=None) -> dict:
        dev = params.device if params.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if static_in is None:
            static_in = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in (params, cin, cout))
            for s, t in zip(static_in, (params, cin, cout)):
                s.copy_(t, non_blocking=True)
        # one eager pass with blocking checks: sets the capacity hints the captured buffers are sized from (and
        # takes the slow path now if this scene needs it)
        self._eager(*static_in)
        self.n_captures += 1
        rows = torch.zeros(self.MAX_JOBS, 3, dtype=torch.int32).pin_memory()
        used = [0]
        rend = self.rend
        prev = (rend.defer_checks, rend._static_stats, rend._pending)
        rend.defer_checks, rend._static_stats, rend._pending = True, (rows, used), []
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES
        try:
            with torch.cuda.graph(graph), torch.no_grad():
                outs = rend(static_in[2], self.conv(static_in[0], static_in[1]), bg_color=self.bg, h=self.h, w=self.w)[:3]
            jobs = rend._pending
        finally:
            rend.defer_checks, rend._static_stats, rend._pending = prev
        return dict(graph=graph, static_in=static_in, outs=tuple(outs), jobs=jobs, rows=rows, stale=False,
                    launches=ops.LAUNCHES - n0, done=torch.cuda.Event(), radii=list(rend.radii))

    @staticmethod
    def _key(params, cin, cout):
        return (tuple(params.shape), params.dtype, tuple(cin.shape), tuple(cout.shape))

    # ------------------------------------------------------------------ plans: capture / load / replay / check
    def plan_for(self, params, cameras_in, cameras_out, slot: int = 0) -> dict:
        """The captured plan for inputs of this shape (captured now if there is none). `slot` selects one of several
        independent plans of the same shape, each with its own static inputs and outputs (DecodeRenderPipeline
        double-buffers over two)."""
        key = self._key(params, cameras_in, cameras_out) + (slot,)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self._capture(params, cameras_in, cameras_out)
            plan["key"] = key
        return plan

    @staticmethod
    def load(plan: dict, params, cameras_in, cameras_out) -> None:
        """Copy one scene's inputs (device or pinned host) into the plan's static inputs, on the current stream."""
        for s, t in zip(plan["static_in"], (params, cameras_in, cameras_out)):
            s.copy_(t, non_blocking=True)

    def replay(self, plan: dict) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Launch the plan on the current stream. A plan marked stale (a scene of its shape overflowed the captured
        intersection buffers) is captured again first, with the capacity the renderer has learnt since."""
        if plan["stale"]:
            plan.update(self._capture(*plan["static_in"], static_in=plan["static_in"]))
        plan["graph"].replay()
        ops.LAUNCHES += plan["launches"]
        plan["done"].record()
        return plan["outs"]

    def check(self, plan: dict) -> bool:
        """True if the plan's last replay passed its deferred checks (waits for that replay to finish). On failure
        every plan of the shape is marked stale; the caller re-renders the scene (`render_eager`)."""
        plan["done"].synchronize()
        ok = self.rend.check_jobs(plan["jobs"])
        if not ok:
            self.n_failed_checks += 1
            for k, p in self._plans.items():
                if k[:-1] == plan["key"][:-1]:
                    p["stale"] = True
        return ok

    def render_eager(self, params, cameras_in, cameras_out):
        """The eager modules with blocking checks (exact buffer sizes, fallbacks): the recovery path."""
        dev = torch.device("cuda", torch.cuda.current_device())
        return tuple(self._eager(*(t.to(dev, non_blocking=True) for t in (params, cameras_in, cameras_out))))

    # ------------------------------------------------------------------ one-call interface
    def run(self, params: torch.Tensor, cameras_in: torch.Tensor, cameras_out: torch.Tensor,
            check: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        plan = self.plan_for(params, cameras_in, cameras_out)
        self.load(plan, params, cameras_in, cameras_out)
        self.replay(plan)
        self._last = plan
        if check and not self.validate():
            return self.recover()
        return plan["outs"]

    __call__ = run

    def validate(self) -> bool:
        """True if the last run()'s deferred checks passed (waits for that run to finish on the device)."""
        return True if self._last is None else self.check(self._last)

    def recover(self):
        """After validate() returned False: the correct (images, depths, masks) of that run, rendered eagerly (this
        also teaches the renderer the larger capacity; the next run() of the shape captures again)."""
        plan, self._last = self._last, None
        return tuple(self._eager(*plan["static_in"]))
