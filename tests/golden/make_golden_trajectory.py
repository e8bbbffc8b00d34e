"""Golden fixture for the trajectory pose sampler: outputs of the REFERENCE's own
`sample_from_dense_cameras` (prometheus/utils/camera.py:134-217, with matrix_to_quaternion /
quaternion_to_matrix from prometheus/utils/math.py:18-91 and quaternion_slerp camera.py:81-109),
imported from /root/reference (never copied) and run on CPU in fp32.

    python tests/golden/make_golden_trajectory.py     ->  tests/golden/trajectory.npz

The reference function only broadcasts correctly for B = 1 trajectories and M = 1 sample per call (its slerp
multiplies a [B,M,1] fraction with a [B,M] angle) — which is how export_video (prometheus/utils/io.py:167-190) calls
it, once per frame. The fixture therefore calls it once per t and stacks the results; the batched B200 kernel
must reproduce exactly that.

Cases: orbit with large rotations (all four matrix_to_quaternion branches), a smooth path with random t, a path
with a repeated pose (zero-length segment), t = 0 / t = 1 end points (t = 1 falls through the reference's strict
`<` segment test and lands in segment 0), the 120-frame video schedule t = i / 120.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_UTILS = "/root/reference/prometheus/utils"


def load_reference_camera():
    """Import prometheus/utils/camera.py and math.py without executing prometheus/utils/__init__.py
    (which pulls in training-only dependencies)."""
    pkg = types.ModuleType("refutils")
    pkg.__path__ = [REF_UTILS]
    sys.modules["refutils"] = pkg
    mods = {}
    for name in ("math", "camera"):
        spec = importlib.util.spec_from_file_location("refutils." + name, os.path.join(REF_UTILS, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["refutils." + name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["camera"]


def rot(axis, ang):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def make_path(rng, n, big_rotations, repeat_at=None):
    cams = np.zeros((n, 18), np.float32)
    R = rot(rng.normal(size=3), rng.uniform(0, np.pi))
    p = rng.normal(size=3) * 0.5
    for i in range(n):
        if repeat_at is not None and i == repeat_at:
            pass  # same pose as the previous one: a zero-length segment
        else:
            step = rng.uniform(0.3, 1.2) if big_rotations else rng.uniform(0.02, 0.15)
            R = rot(rng.normal(size=3), step) @ R
            p = p + rng.normal(size=3) * 0.1
        cams[i, :12] = np.concatenate([R, p[:, None]], 1).reshape(-1)
        f = rng.uniform(400, 600)
        cams[i, 12:] = [f, f * rng.uniform(0.95, 1.05), 256 + rng.normal(), 256 + rng.normal(), 512, 512]
    return cams


def main():
    cam = load_reference_camera()
    rng = np.random.default_rng(7)
    out = {}
    cases = {
        "orbit_big": (np.stack([make_path(rng, 9, True)]), np.linspace(0, 1, 23, dtype=np.float32)[None]),
        "smooth": (np.stack([make_path(rng, 30, False)]), rng.uniform(0, 1, size=(1, 40)).astype(np.float32)),
        "repeat": (np.stack([make_path(rng, 6, False, repeat_at=3)]), np.linspace(0, 0.999, 17, dtype=np.float32)[None]),
        "video120": (np.stack([make_path(rng, 8, False)]), (np.arange(120, dtype=np.float32) / 120)[None]),
    }
    for name, (dense, t) in cases.items():
        with torch.no_grad():
            d = torch.from_numpy(dense.copy())
            got = torch.cat([cam.sample_from_dense_cameras(d, torch.tensor([[float(ti)]])) for ti in t[0]], dim=1)
        assert got.shape == (dense.shape[0], t.shape[1], 18) and torch.isfinite(got).all(), name
        out[name + "_dense"], out[name + "_t"], out[name + "_out"] = dense, t, got.numpy()
    np.savez_compressed(os.path.join(HERE, "trajectory.npz"), **out)
    print("wrote trajectory.npz:", {k: v.shape for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
