"""Novel-view trajectory rendering — the outer loop of `inference.py` / config C5.

Reference: `export_video` (prometheus/utils/io.py:167-190) renders `num_frames` single views in a host loop: per
frame one `sample_from_dense_cameras` call (prometheus/utils/camera.py:134-217: ~60 small ATen kernels plus
`matrix_to_quaternion`'s Python loop over all path poses, prometheus/utils/math.py:46-91), one single-view render, and a
blocking `.cpu()` of the fp32 image and depth.

Here: every pose of the trajectory is sampled by ONE kernel, frames are rendered `chunk` cameras at a time through
`GaussianRenderer` (deferred checks, no host sync), converted to uint8 HWC on the device and copied to pinned host
memory on a side stream while the next chunk renders.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import ops
from .renderer import GaussianRenderer


def sample_from_dense_cameras(dense_cameras: torch.Tensor, t: torch.Tensor, noise_strengths: Sequence[float] = (0, 0, 0, 0)):
    """Drop-in for prometheus.utils.camera.sample_from_dense_cameras: dense_cameras [B,N,18], t [B,M] -> [B,M,18].
    Equals the reference where the reference is well defined (B = 1, one sample per call) and is the per-sample
    stacking of that for any B, M. Position noise (noise_strengths[0]) is applied like the reference; rotation
    noise is not supported."""
    if not dense_cameras.is_cuda:
        raise RuntimeError("prometheus_b200.sample_from_dense_cameras needs CUDA tensors: there is no CPU fallback")
    if len(noise_strengths) > 1 and noise_strengths[1] > 0:
        raise NotImplementedError("rotation noise (noise_strengths[1]) is outside the inference path")
    out = ops.sample_cameras(dense_cameras, t.to(dense_cameras.device))
    if noise_strengths[0] > 0:
        pose = out[..., :12].view(*out.shape[:2], 3, 4)
        pose[..., :3, 3] += torch.randn_like(pose[..., :3, 3]) * noise_strengths[0]
    return out


def _video_schedule(num_frames: int, device) -> torch.Tensor:
    """t_i = i / num_frames exactly as export_video forms it (io.py:175: a Python double rounded to fp32; a
    device-side fp32 `arange / n` multiplies by the reciprocal and is off by one ulp for some i)."""
    return torch.tensor([[i / num_frames for i in range(num_frames)]], dtype=torch.float32).to(device)


class TrajectoryRenderer:
    """Render a camera path through one scene's Gaussians into host frames.

        tr = TrajectoryRenderer(renderer)
        frames_u8, depths = tr.render(gaussians, dense_cameras, num_frames=120, render_size=1024)

    frames_u8 [F,H,W,3] uint8 and depths [F,H,W,1] fp32 are pinned host tensors laid out like the lists
    export_video builds (io.py:181-182)."""

    def __init__(self, renderer: GaussianRenderer, chunk: Optional[int] = None):
        self.rend = renderer
        self.chunk = int(chunk or renderer.view_chunk)
        self._side = None

    def cameras(self, dense_cameras: torch.Tensor, num_frames: int) -> torch.Tensor:
        """Poses of the video schedule t_i = i / num_frames (io.py:175) -> [num_frames, 18] on the device."""
        dense = dense_cameras.reshape(1, -1, 18) if dense_cameras.dim() == 2 else dense_cameras[:1]
        return ops.sample_cameras(dense, _video_schedule(num_frames, dense.device))[0]

    def render(self, gaussians, dense_cameras: torch.Tensor, num_frames: int = 720, render_size: int = 512,
               bg_color=None, frames_out: Optional[torch.Tensor] = None, depths_out: Optional[torch.Tensor] = None,
               frame_range: Optional[Tuple[int, int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """`frame_range = (a, b)`: render only frames a..b-1 of the `num_frames` schedule (one rank's share when the
        frames of a video are sharded over GPUs); the outputs then hold b - a frames."""
        dev = gaussians[0].device
        H = W = int(render_size)
        fa, fb = (0, num_frames) if frame_range is None else (int(frame_range[0]), int(frame_range[1]))
        assert 0 <= fa <= fb <= num_frames
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
        if frames_out is None:
            frames_out = torch.empty(fb - fa, H, W, 3, dtype=torch.uint8, pin_memory=True)
        if depths_out is None:
            depths_out = torch.empty(fb - fa, H, W, 1, dtype=torch.float32, pin_memory=True)
        assert frames_out.shape[0] == fb - fa and depths_out.shape[0] == fb - fa
        cams = self.cameras(dense_cameras.to(dev), num_frames)
        cur = torch.cuda.current_stream(dev)
        prev_defer = self.rend.defer_checks
        self.rend.defer_checks = True
        pending = []
        try:
            with torch.no_grad():
                for f0 in range(fa, fb, self.chunk):
                    f1 = min(fb, f0 + self.chunk)
                    self._render_chunk(gaussians, cams, f0, f1, H, W, bg_color, frames_out, depths_out, cur, fa)
                    pending.append((f0, f1, self.rend.take_pending()))
                # the renders are queued: look at their deferred checks, redo a chunk in the rare failing case
                self.rend.defer_checks = False
                for f0, f1, jobs in pending:
                    if not self.rend.check_jobs(jobs):
                        self._render_chunk(gaussians, cams, f0, f1, H, W, bg_color, frames_out, depths_out, cur, fa)
        finally:
            self.rend.defer_checks = prev_defer
        cur.wait_stream(self._side)
        # the outputs are pinned HOST tensors a caller reads on the CPU: block the host until the side-stream
        # device->host copies have landed (ordering the current stream behind them is not enough for that)
        self._side.synchronize()
        return frames_out, depths_out

    def _render_chunk(self, gaussians, cams, f0, f1, H, W, bg_color, frames_out, depths_out, cur, base=0):
        images, depths, _, _, _ = self.rend(cams[None, f0:f1], gaussians, bg_color=bg_color, h=H, w=W)
        u8 = ops.pack_frames_u8(images[0])
        d = depths[0]  # [n,1,H,W] is bit-for-bit [n,H,W,1]
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self._side):
            self._side.wait_event(done)
            u8.record_stream(self._side)
            d.record_stream(self._side)
            frames_out[f0 - base:f1 - base].copy_(u8, non_blocking=True)
            depths_out[f0 - base:f1 - base].view(f1 - f0, 1, H, W).copy_(d, non_blocking=True)


def export_video(render_fn, save_path, name, dense_cameras, fps=60, num_frames=720, render_size=512, device="cuda:0"):
    """Signature of prometheus.utils.io.export_video. `render_fn` is either a `(renderer, gaussians)` pair — the
    batched fast path — or the reference's callable `(camera [1,1,18], h, w) -> (image, depth)`, which is then
    driven frame by frame but with all poses sampled in one launch. Returns (frames uint8 [F,H,W,3], depths
    [F,H,W,1]) as numpy arrays; writes `<name>.mp4` only if `save_path` is given and imageio is installed (video
    encoding and the depth colour map are host-side and outside this path)."""
    dense_cameras = dense_cameras.to(device)
    if isinstance(render_fn, tuple):
        renderer, gaussians = render_fn
        frames, depths = TrajectoryRenderer(renderer).render(gaussians, dense_cameras, num_frames, render_size)
        torch.cuda.current_stream().synchronize()
        frames, depths = frames.numpy(), depths.numpy()
    else:
        cams = ops.sample_cameras(dense_cameras.reshape(1, -1, 18), _video_schedule(num_frames, device))
        fr, de = [], []
        with torch.no_grad():
            for i in range(num_frames):
                image, depth = render_fn(cams[:, i:i + 1], render_size, render_size)
                fr.append(ops.pack_frames_u8(image.reshape(1, 3, render_size, render_size)))
                de.append(depth.reshape(1, render_size, render_size, 1))
        frames, depths = torch.cat(fr).cpu().numpy(), torch.cat(de).cpu().numpy()
    if save_path is not None:
        try:
            import imageio
        except ImportError as e:  # noqa: F841
            raise RuntimeError("imageio is not installed: pass save_path=None and encode the returned frames yourself")
        import os

        imageio.mimwrite(os.path.join(save_path, f"{name}.mp4"), list(frames), fps=fps, quality=8, macro_block_size=1)
    return frames, depths


def export_ply_for_gaussians(path, gaussians, opacity_threshold=0.00, bbox=(-2, 2), compatible=True):
    """Drop-in for prometheus.utils.io.export_ply_for_gaussians (io.py:25-125): writes `path + '.splat'` — like the
    reference (its `.ply` write is commented out) — and returns {'pts', 'color', 'alpha'}. The reference sorts with
    np.argsort and builds every 32-byte record in a Python loop; here the sort keys, the radix sort and the packing
    run on the device (ops.splat_pack) and one device->host copy feeds the file write. `path=None` skips the file."""
    xyz, features, opacity, scales, rotations = gaussians
    means3d, opac = xyz[0].contiguous().float(), opacity[0].contiguous().float()
    scl, rot, shs = scales[0].contiguous().float(), rotations[0].contiguous().float(), features[0].contiguous().float()
    if not means3d.is_cuda:
        raise RuntimeError("prometheus_b200.export_ply_for_gaussians needs CUDA tensors: there is no CPU fallback")
    mask = opac[..., 0] >= opacity_threshold
    means3d, opac, scl, rot, shs = means3d[mask], opac[mask], scl[mask], rot[mask], shs[mask]
    rgb = shs.detach().transpose(1, 2).flatten(start_dim=1).contiguous()
    splat = ops.splat_pack(means3d.detach(), rgb, opac[:, 0].detach().contiguous(), scl.detach(), rot.detach(), compatible)
    data = splat.cpu().numpy()
    if path is not None:
        with open(path + ".splat", "wb") as f:
            f.write(data.tobytes())
    return {"pts": means3d.detach().cpu().numpy(), "color": ((rgb + 1) / 2).cpu().numpy(), "alpha": opac.detach().cpu().numpy(),
            "splat": data}
