#!/usr/bin/env python
"""Where does the end-to-end step go? H2D alone, D2H alone, compute alone, the streaming pipeline (developer tool)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prometheus_b200 import DecodeRenderPipeline, GaussianConverter, GaussianRenderer, scenes  # noqa: E402

dev = torch.device("cuda", 0)
sc = scenes.make_scene(sys.argv[1] if len(sys.argv) > 1 else "C2")
S = sc.size_out
conv = GaussianConverter("mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX).to(dev)
rend = GaussianRenderer(background=[1, 1, 1]).to(dev)
ph, cih, coh = sc.params.pin_memory(), sc.cameras_in.pin_memory(), sc.cameras_out.pin_memory()
B, V = sc.cameras_out.shape[:2]
out_h = [torch.empty(B, V, c, S, S).pin_memory() for c in (3, 1, 1)]
out_d = [torch.empty(B, V, c, S, S, device=dev) for c in (3, 1, 1)]
n = 30


def timeit(fn, n=n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3), round(host, 3)


pd = ph.to(dev)
print("h2d 117MB (gpu ms, host enqueue ms):", timeit(lambda: pd.copy_(ph, non_blocking=True)))
print("d2h 42MB:", timeit(lambda: [h.copy_(d, non_blocking=True) for h, d in zip(out_h, out_d)]))
cid, cod = cih.to(dev), coh.to(dev)


def compute():
    with torch.no_grad():
        rend(cod, conv(pd, cid), h=S, w=S)


compute()
rend.defer_checks = True
print("compute (deferred checks):", timeit(compute))
rend.validate()
rend.defer_checks = False
print("compute (blocking checks):", timeit(compute))
pipe = DecodeRenderPipeline(conv, rend, S, S, device=dev)


def loop(k):
    pipe.prefetch(ph, cih, coh)
    for i in range(k):
        pipe.run(out_h, next_inputs_h=(ph, cih, coh) if i + 1 < k else None)
    pipe.flush()


loop(5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
loop(n)
e1.record()
host = (time.perf_counter() - t0) / n * 1e3
torch.cuda.synchronize()
print("pipeline per step (gpu ms, host ms):", round(e0.elapsed_time(e1) / n, 3), round(host, 3))
