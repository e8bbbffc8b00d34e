#!/usr/bin/env python
"""Culling statistics of the forward compositor (needs the -DPM_STATS build: tools/build_variant.sh stats -DPM_STATS;
run with PM_GS_LIB=gpurun_variants/stats.so)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prometheus_b200 import GaussianConverter, GaussianRenderer, _lib, scenes  # noqa: E402

dev = torch.device("cuda", 0)
sc = scenes.make_scene(sys.argv[1] if len(sys.argv) > 1 else "C2")
S = sc.size_out
conv = GaussianConverter("mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX).to(dev)
rend = GaussianRenderer(background=[1, 1, 1]).to(dev)
L = _lib.load()
buf = (ctypes.c_ulonglong * 16)()
with torch.no_grad():
    g = conv(sc.params.to(dev), sc.cameras_in.to(dev))
    rend(sc.cameras_out.to(dev), g, h=S, w=S)
    L.pm_debug_stats(buf, 1)
    rend(sc.cameras_out.to(dev), g, h=S, w=S)
    L.pm_debug_stats(buf, 1)
v = list(buf)
print("warp-chunks with hits %d, warp hits %d (%.1f per chunk)" % (v[0], v[1], v[1] / max(v[0], 1)))
print("lane evals %d; above alpha threshold %d (%.1f%%); live+above %d (%.1f%%)" % (v[1] * 32, v[2], 100 * v[2] / (v[1] * 32), v[3], 100 * v[3] / (v[1] * 32)))
print("hits with >=1 lane above threshold %d (%.1f%%); with >=1 live lane above threshold %d (%.1f%%)" % (v[4], 100 * v[4] / v[1], v[5], 100 * v[5] / v[1]))
t = max(v[13], 1)
print("tiles %d, records/tile %.0f, passes/tile %.2f, groups/tile %.2f, level-2 refinements/tile %.3f, radix-sorted groups %d, staged/tile %.0f"
      % (v[13], v[14] / t, v[6] / t, v[7] / t, v[8] / t, v[9], v[10] / t))
print("footprint candidates %d (%.1f%% of staged x 8 warps), dense cull rounds %d (%.1f candidates each)"
      % (v[11], 100 * v[11] / max(v[10] * 8, 1), v[12], v[11] / max(v[12], 1)))
