"""Streaming decode+render of host-resident scenes (the text-to-3D batch case, BASELINE config C4:
many scenes' decoder maps arrive from the host, images go back to the host).

`DecodeRenderPipeline` keeps three CUDA streams busy: while scene i is unprojected / projected /
sorted / composited on the compute stream, scene i+1's maps are already moving host->device and
scene i-1's images device->host. Inputs must be pinned host tensors; device buffers are double
buffered. This is the public end-to-end API that bench.py's `e2e` number goes through.

With `graph=True` the per-scene launch set is replayed from CUDA graphs (graphed.GraphedDecodeRender): two plans
with their own static inputs and outputs alternate, so the same three-stream overlap holds while the host only
issues a handful of calls per scene — what small scenes need, whose step is launch-bound rather than PCIe-bound.
In both modes a scene's pinned inputs must stay untouched until the scene after the next one has been submitted (or
flush() has returned): the copies are asynchronous, and a scene whose deferred check fails is rendered again from them.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .graphed import GraphedDecodeRender
from .renderer import GaussianConverter, GaussianRenderer


class DecodeRenderPipeline:
    max_in_flight = 2  # scenes whose deferred checks have not been looked at yet (eager mode)

    def __init__(self, converter: GaussianConverter, renderer: GaussianRenderer, h: int, w: int, device=None,
                 bg_color=None, graph: bool = False, compact_outputs: bool = False):
        """compact_outputs: hand the host uint8 frames [B,M,H,W,3] (the conversion of export_video, io.py:181, done on
        the device by pm_pack_frames_u8) and fp16 depths / masks [B,M,1,H,W] instead of three fp32 tensors — 7 instead
        of 20 bytes per pixel over PCIe, for consumers that write videos / previews. `out_h` must then hold pinned
        tensors of those dtypes and shapes."""
        self.compact = bool(compact_outputs)
        self.conv, self.rend = converter, renderer
        self.graphed = GraphedDecodeRender(converter, renderer, h, w, bg_color) if graph else None
        self._n_staged = 0
        self.h, self.w, self.bg = h, w, bg_color
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._staged = None      # (device inputs, ready event) of the next scene
        self._pending_out = []   # events of D2H copies in flight
        self._dev_out_keepalive = []
        self._unchecked = []     # (deferred renderer checks, device inputs, host outputs) per scene in flight

    def _stage(self, params_h, cams_in_h, cams_out_h):
        assert params_h.is_pinned() and cams_in_h.is_pinned() and cams_out_h.is_pinned(), "pin the host buffers"
        if self.graphed is not None:
            return self._stage_graph(params_h, cams_in_h, cams_out_h)
        with torch.cuda.stream(self.s_in):
            dev = tuple(t.to(self.device, non_blocking=True) for t in (params_h, cams_in_h, cams_out_h))
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return dev, ev

    # ------------------------------------------------------------------ graph mode
    def _stage_graph(self, params_h, cams_in_h, cams_out_h):
        """Host->device copy straight into the static inputs of the plan (slot = scene parity) this scene will
        replay. The slot's previous scene must have been rendered (its inputs consumed) first."""
        gr = self.graphed
        with torch.cuda.device(self.device):
            plan = gr.plan_for(params_h, cams_in_h, cams_out_h, slot=self._n_staged % 2)  # first use: captures (blocking)
        self._n_staged += 1
        self._settle_graph(plan)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(plan["done"])
            gr.load(plan, params_h, cams_in_h, cams_out_h)
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return (plan, (params_h, cams_in_h, cams_out_h)), ev

    def _settle_graph(self, plan) -> None:
        """Before a plan is reused: look at the deferred checks of the scene it rendered last (they live in the plan's
        pinned rows, which the next replay overwrites). A failed check re-renders that scene eagerly from its host
        inputs and replaces its host copy; the plans of the shape are captured again at their next replay."""
        entry = plan.pop("in_flight", None)
        if entry is None:
            return
        inputs_h, out_h = entry
        if self.graphed.check(plan):
            return
        with torch.cuda.device(self.device):
            outs = self.graphed.render_eager(*inputs_h)
        self._copy_out(out_h, outs, torch.cuda.current_stream(self.device))

    def _run_graph(self, out_h, next_inputs_h) -> None:
        (plan, inputs_h), ready = self._staged
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)
        if "copied" in plan:
            cur.wait_event(plan["copied"])  # the slot's previous images have left its static outputs
        outs = self.graphed.replay(plan)
        plan["copied"] = self._copy_out(out_h, outs, cur, record_stream=False)
        plan["in_flight"] = (inputs_h, out_h)
        # staging the next scene reuses the OTHER slot and first waits (on the host) for that slot's previous scene
        # to have finished — this scene is already queued behind it, so the device does not idle meanwhile
        self._staged = self._stage(*next_inputs_h) if next_inputs_h is not None else None

    def prefetch(self, params_h, cams_in_h, cams_out_h) -> None:
        """Start the host->device copy of the NEXT scene (call before `run` of the current one)."""
        self._staged = self._stage(params_h, cams_in_h, cams_out_h)

    def run(self, out_h: Sequence[torch.Tensor], next_inputs_h: Optional[Tuple[torch.Tensor, ...]] = None) -> None:
        """Decode + render the staged scene and start the device->host copy of (images, depths, masks)
        into the pinned tensors `out_h`. `next_inputs_h` = (params, cameras_in, cameras_out) of the NEXT
        scene: its host->device copy is enqueued first, so it overlaps this scene's compute."""
        assert self._staged is not None, "call prefetch() for the first scene"
        if self.graphed is not None:
            return self._run_graph(out_h, next_inputs_h)
        (params, cin, cout), ready = self._staged
        self._staged = self._stage(*next_inputs_h) if next_inputs_h is not None else None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)
        for t in (params, cin, cout):
            t.record_stream(cur)
        outs = self._render(params, cin, cout, deferred=True)
        jobs = self.rend.take_pending()
        self._copy_out(out_h, outs, cur)
        # keep what a re-render would need until the scene's deferred checks have been looked at
        self._unchecked.append((jobs, (params, cin, cout), out_h))
        # bounded run-ahead: with one scene executing and `max_in_flight` queued behind it the device never idles,
        # while a host that runs many scenes ahead only multiplies the buffers in flight (and the allocator's
        # cudaMalloc calls, tens of milliseconds each, until it has enough of them)
        while len(self._unchecked) > self.max_in_flight:
            self._settle(self._unchecked.pop(0))

    def run_batch(self, inputs_h: Tuple[torch.Tensor, ...], out_h: Sequence[torch.Tensor], chunk: Optional[int] = None) -> None:
        """A large host-resident batch of scenes (config C4: 64 scenes) as a stream of `chunk`-scene slices (default:
        the renderer's scene_chunk, or a quarter of the batch if that is smaller): the host->device copy of slice k+1 and the device->host copy of slice k-1 overlap
        the render of slice k, instead of the whole batch being copied, then rendered, then copied back.
        inputs_h = (params [B,N,14,h,w], cameras_in [B,N,18], cameras_out [B,M,18]) pinned; out_h = pinned
        (images [B,M,3,H,W], depths [B,M,1,H,W], masks [B,M,1,H,W]). Call flush() before reading out_h."""
        B = inputs_h[0].shape[0]
        # default slice: the renderer's scene_chunk, but at least four slices per batch — with one or two slices the
        # copies cannot overlap the render (8 scenes per rank at N = 8 went host->device, render, device->host in series)
        c = int(chunk or max(1, min(self.rend.scene_chunk, -(-B // 4))))
        slices = [slice(b, min(B, b + c)) for b in range(0, B, c)]
        self.prefetch(*(t[slices[0]] for t in inputs_h))
        for i, sl in enumerate(slices):
            nxt = tuple(t[slices[i + 1]] for t in inputs_h) if i + 1 < len(slices) else None
            self.run([o[sl] for o in out_h], next_inputs_h=nxt)

    def _render(self, params, cin, cout, deferred: bool):
        prev = self.rend.defer_checks
        self.rend.defer_checks = deferred
        try:
            with torch.no_grad():
                g = self.conv(params, cin)
                return self.rend(cout, g, bg_color=self.bg, h=self.h, w=self.w)[:3]
        finally:
            self.rend.defer_checks = prev

    def _copy_out(self, out_h, outs, cur, record_stream: bool = True):
        if self.compact:
            from . import ops

            images, depths, masks = outs[:3]
            B, M, _, H, W = images.shape
            outs = (ops.pack_frames_u8(images.reshape(B * M, 3, H, W)).reshape(B, M, H, W, 3), depths.half(), masks.half())
            record_stream = True  # fresh allocator blocks, also in graph mode
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            for dst, src in zip(out_h, outs):
                if record_stream:  # not for a graph's static outputs: they are never returned to the allocator
                    src.record_stream(self.s_out)
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_out)
        self._pending_out.append(ev)
        if len(self._pending_out) > 4:
            self._pending_out.pop(0).synchronize()
        return ev

    def _settle(self, entry) -> None:
        """Look at a scene's deferred checks (intersection capacity, sort-group overflow); in the rare case
        one failed, render the scene again with blocking checks and replace its host copy."""
        jobs, (params, cin, cout), out_h = entry
        if self.rend.check_jobs(jobs):
            return
        cur = torch.cuda.current_stream(self.device)
        outs = self._render(params, cin, cout, deferred=False)
        self._copy_out(out_h, outs, cur)

    def flush(self) -> None:
        """Wait — on the HOST — for every outstanding render and device->host copy: when flush() returns, every
        `out_h` handed to run() / run_batch() holds its final content and may be read on the CPU. (The current
        stream is ordered behind the copies as well.)"""
        cur = torch.cuda.current_stream(self.device)
        while self._unchecked:
            self._settle(self._unchecked.pop(0))
        if self.graphed is not None:
            for plan in list(self.graphed._plans.values()):
                self._settle_graph(plan)
        for ev in self._pending_out:
            cur.wait_event(ev)
            ev.synchronize()  # the host reads out_h next: ordering streams does not block the CPU
        self._pending_out = []
        cur.wait_stream(self.s_in)
