"""Golden fixture for half-precision decoder maps: the REFERENCE GaussianConverter
(gaussians_renderer.py:210-287, read from /root/reference, never copied) run on fp16 / bf16 maps,
forward and autograd backward (the gradient arrives in the maps' dtype).

Reuses the fp32 inputs of converter_<mode>.npz, rounded to the half type. Run in the build container only:
    python tests/golden/make_golden_halfmaps.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference_module  # noqa: E402


def main():
    ref = load_reference_module()
    out = {}
    for mode, dt, name in (("mvsplat", torch.bfloat16, "bf16"), ("mvsplat", torch.float16, "f16"),
                           ("director3d", torch.bfloat16, "bf16"), ("gslrm", torch.float16, "f16")):
        g = np.load(os.path.join(HERE, f"converter_{mode}.npz"))
        conv = ref.GaussianConverter(gs_convert_mode=mode, z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0)
        params = torch.from_numpy(g["params"]).to(dt).requires_grad_(True)
        outs = conv(params, torch.from_numpy(g["cameras"]))
        keys = ("xyz", "features", "opacity", "scales", "rotations")
        assert all(o.dtype == torch.float32 for o in outs)
        loss = sum((o * torch.from_numpy(g["w_" + k])).sum() for o, k in zip(outs, keys))
        loss.backward()
        assert params.grad.dtype == dt
        pre = f"{mode}_{name}_"
        for o, k in zip(outs, keys):
            out[pre + k] = o.detach().numpy()
        # half tensors stored widened to fp32 (exact); the test narrows them again
        out[pre + "params"] = params.detach().float().numpy()
        out[pre + "v_params"] = params.grad.float().numpy()
    np.savez_compressed(os.path.join(HERE, "converter_half.npz"), **out)
    print("converter_half.npz", os.path.getsize(os.path.join(HERE, "converter_half.npz")))


if __name__ == "__main__":
    main()
