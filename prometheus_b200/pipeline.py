"""Streaming decode+render of host-resident scenes (the text-to-3D batch case, BASELINE config C4:
many scenes' decoder maps arrive from the host, images go back to the host).

`DecodeRenderPipeline` keeps three CUDA streams busy: while scene i is unprojected / projected /
sorted / composited on the compute stream, scene i+1's maps are already moving host->device and
scene i-1's images device->host. Inputs must be pinned host tensors; device buffers are double
buffered. This is the public end-to-end API that bench.py's `e2e` number goes through.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .renderer import GaussianConverter, GaussianRenderer


class DecodeRenderPipeline:
    def __init__(self, converter: GaussianConverter, renderer: GaussianRenderer, h: int, w: int, device=None,
                 bg_color=None):
        self.conv, self.rend = converter, renderer
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
        with torch.cuda.stream(self.s_in):
            dev = tuple(t.to(self.device, non_blocking=True) for t in (params_h, cams_in_h, cams_out_h))
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return dev, ev

    def prefetch(self, params_h, cams_in_h, cams_out_h) -> None:
        """Start the host->device copy of the NEXT scene (call before `run` of the current one)."""
        self._staged = self._stage(params_h, cams_in_h, cams_out_h)

    def run(self, out_h: Sequence[torch.Tensor], next_inputs_h: Optional[Tuple[torch.Tensor, ...]] = None) -> None:
        """Decode + render the staged scene and start the device->host copy of (images, depths, masks)
        into the pinned tensors `out_h`. `next_inputs_h` = (params, cameras_in, cameras_out) of the NEXT
        scene: its host->device copy is enqueued first, so it overlaps this scene's compute."""
        assert self._staged is not None, "call prefetch() for the first scene"
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
        while len(self._unchecked) > 4:
            self._settle(self._unchecked.pop(0))

    def _render(self, params, cin, cout, deferred: bool):
        prev = self.rend.defer_checks
        self.rend.defer_checks = deferred
        try:
            with torch.no_grad():
                g = self.conv(params, cin)
                return self.rend(cout, g, bg_color=self.bg, h=self.h, w=self.w)[:3]
        finally:
            self.rend.defer_checks = prev

    def _copy_out(self, out_h, outs, cur) -> None:
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            for dst, src in zip(out_h, outs):
                src.record_stream(self.s_out)
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_out)
        self._pending_out.append(ev)
        if len(self._pending_out) > 4:
            self._pending_out.pop(0).synchronize()

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
        """Make the current stream wait for every outstanding copy (call before reading `out_h`)."""
        cur = torch.cuda.current_stream(self.device)
        while self._unchecked:
            self._settle(self._unchecked.pop(0))
        for ev in self._pending_out:
            cur.wait_event(ev)
        self._pending_out = []
        cur.wait_stream(self.s_in)
