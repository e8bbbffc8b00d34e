#!/usr/bin/env python
"""Per-stage timing of the module fast path on one config (developer tool; GPU only).
    python tools/stage_bench.py [--config C2] [--steps 20] [--bwd]
    PM_GS_LIB=/path/to/variant.so python tools/stage_bench.py     (A/B an alternative build)
Prints one JSON line: total ms/step (CUDA events, L2 flushed between steps) and ms per C-ABI call."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prometheus_b200 import GaussianConverter, GaussianRenderer, ops, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--two-pass", action="store_true", help="disable single-pass binning (project_count + bin_scatter)")
    ap.add_argument("--tag", default=os.environ.get("PM_GS_LIB", "default"))
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    sc = scenes.make_scene(a.config)
    S = sc.size_out
    conv = GaussianConverter("mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX).to(dev)
    rend = GaussianRenderer(background=[1, 1, 1]).to(dev)
    rend.fused = not a.unfused
    rend.single_pass_binning = not a.two_pass
    params, cin, cout = sc.params.to(dev), sc.cameras_in.to(dev), sc.cameras_out.to(dev)
    target = torch.rand(*cout.shape[:2], 3, S, S, device=dev) * 2 - 1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        if not a.bwd:
            with torch.no_grad():
                return rend(cout, conv(params, cin), bg_color=None, h=S, w=S)
        p = params.detach().requires_grad_(True)
        im, de, ma, _, _ = rend(cout, conv(p, cin), bg_color=None, h=S, w=S)
        scenes.render_loss(im, de, ma, target).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ops.PROFILE = []
    evs = []
    for _ in range(a.steps):
        if not a.no_flush:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    st = {}
    for n, e0, e1 in prof:
        st[n] = st.get(n, 0.0) + e0.elapsed_time(e1) / a.steps
    tot = sum(x.elapsed_time(y) for x, y in evs) / a.steps
    print(json.dumps({"tag": a.tag, "config": a.config, "bwd": a.bwd, "ms_per_step": round(tot, 4),
                      "kernels_ms": round(sum(st.values()), 4), "stages_ms": {k: round(v, 4) for k, v in st.items()}}))


if __name__ == "__main__":
    main()
