"""Golden fixture for the .splat exporter: the byte stream written by the REFERENCE's own
`export_ply_for_gaussians` (prometheus/utils/io.py:25-125), imported from /root/reference and run on CPU.

    python tests/golden/make_golden_splat.py     ->  tests/golden/splat.npz

`plyfile`, `imageio` and the depth colour map are not installed here; they are replaced by arithmetic-free stubs
(the exporter only uses plyfile as a wrapper around one numpy structured array: `el["scale_0"]` / `el[idx]`).
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_UTILS = "/root/reference/prometheus/utils"


def load_reference_io():
    ply = types.ModuleType("plyfile")

    class PlyElement:
        @staticmethod
        def describe(elements, name):
            return elements  # the structured array itself: supports el["field"] and el[idx]["field"]

    class PlyData:
        def __init__(self, elements):
            self._el = {"vertex": elements[0]}

        def __getitem__(self, k):
            return self._el[k]

    ply.PlyData, ply.PlyElement = PlyData, PlyElement
    sys.modules["plyfile"] = ply
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    pkg = types.ModuleType("refutils")
    pkg.__path__ = [REF_UTILS]
    sys.modules["refutils"] = pkg
    img = types.ModuleType("refutils.image_utils")
    img.colorize_depth_maps = lambda *a, **k: None
    sys.modules["refutils.image_utils"] = img
    mods = {}
    for name in ("math", "camera", "io"):
        spec = importlib.util.spec_from_file_location("refutils." + name, os.path.join(REF_UTILS, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["refutils." + name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["io"]


def main():
    io = load_reference_io()
    gen = torch.Generator().manual_seed(3)
    n = 700
    xyz = torch.randn(1, n, 3, generator=gen)
    features = torch.rand(1, n, 1, 3, generator=gen)                       # rgb in [0,1] (sigmoid outputs of the decoder)
    opacity = torch.sigmoid(torch.randn(1, n, 1, generator=gen) * 2.0)
    scales = 0.001 + 0.3 * torch.rand(1, n, 3, generator=gen) ** 2
    rotations = torch.randn(1, n, 4, generator=gen)
    rotations = rotations / rotations.norm(dim=-1, keepdim=True)
    out = dict(xyz=xyz.numpy(), features=features.numpy(), opacity=opacity.numpy(), scales=scales.numpy(), rotations=rotations.numpy())
    for tag, thr in (("all", 0.0), ("pruned", 0.3)):
        with tempfile.TemporaryDirectory() as d:
            ret = io.export_ply_for_gaussians(os.path.join(d, "g"), (xyz, features, opacity, scales, rotations), opacity_threshold=thr)
            out["splat_" + tag] = np.frombuffer(open(os.path.join(d, "g.splat"), "rb").read(), np.uint8).copy()
            for k, v in ret.items():
                out["ret_%s_%s" % (tag, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "splat.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
