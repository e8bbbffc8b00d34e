#!/usr/bin/env python
"""Does the forward gain from running two independent steps on two streams (developer tool)? tile_render_fwd leaves ~29 %
of the issue slots idle (barrier waits); project_fwd<2> of another step could use them if the two kernels co-reside.
Prints ms per step for one stream and for two streams (same total number of steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prometheus_b200 import GaussianConverter, GaussianRenderer, scenes  # noqa: E402

dev = torch.device("cuda", 0)
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc = scenes.make_scene(cfg)
S = sc.size_out
conv = GaussianConverter("mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX).to(dev)
rends = [GaussianRenderer(background=[1, 1, 1]).to(dev) for _ in range(2)]
params, cin, cout = sc.params.to(dev), sc.cameras_in.to(dev), sc.cameras_out.to(dev)
streams = [torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)]


def step(r):
    with torch.no_grad():
        return r(cout, conv(params, cin), bg_color=None, h=S, w=S)


for r in rends:
    for _ in range(3):
        step(r)
    r.defer_checks = True
torch.cuda.synchronize()
K = 20


def run(two):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        for j in range(2):
            if two:
                streams[j].wait_stream(torch.cuda.current_stream()) if i == 0 else None
                with torch.cuda.stream(streams[j]):
                    step(rends[j])
            else:
                step(rends[j])
    if two:
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * K)


for _ in range(2):
    print("one stream %.4f ms/step   two streams %.4f ms/step" % (run(False), run(True)))
assert all(r.validate() for r in rends)
