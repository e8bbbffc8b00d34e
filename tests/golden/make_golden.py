"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own
Python code (read from /root/reference; never copied into the repo).

What is pinned:
  converter_<mode>.npz  GaussianConverter.forward (gaussians_renderer.py:210-287) on
                        small seeded inputs, all three modes, plus its autograd
                        gradients for a fixed upstream gradient;
  cameras.npz           convert_camera_parameters_into_viewpoint_cameras + MiniCam +
                        the K build (:453-472, :139-169, :404-412): viewmat / K per camera;
  renderer_e2e.npz      the reference GaussianRenderer.forward (:475-512) driven end to
                        end with `gsplat.rasterization` bound to the oracle restatement
                        (gsplat itself is not installable here => stages 2-4 stay
                        "parity unpinned"; this fixture pins everything around them:
                        camera handling, argument marshalling, output packing).

Run in the build container only (it needs /root/reference):
    python tests/golden/make_golden.py
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/prometheus/modules/renderers/gaussians_renderer.py"

from oracle import gs_oracle  # noqa: E402


def load_reference_module():
    stub = types.ModuleType("gsplat")
    stub.rasterization = gs_oracle.rasterization
    sys.modules["gsplat"] = stub
    torch.cuda.synchronize = lambda *a, **k: None  # gaussians_renderer.py:510 on a CPU-only box
    spec = importlib.util.spec_from_file_location("ref_gaussians_renderer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.RASTERIZER_BACKEND == "gsplat"
    return mod


def rand_cameras(gen, n, size, H=None):
    """Random rigid OpenGL c2w + intrinsics; H/W may differ from the render size."""
    H = H or size
    cams = torch.zeros(n, 18)
    for i in range(n):
        ax = torch.randn(3, generator=gen)
        ax = ax / ax.norm()
        ang = float(torch.rand(1, generator=gen)) * 0.5
        Kx = torch.tensor([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = torch.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * (Kx @ Kx)
        t = torch.randn(3, generator=gen) * 0.3
        cams[i, :12] = torch.cat([R, t[:, None]], 1).reshape(-1)
        f = 0.8 * H + 0.2 * H * float(torch.rand(1, generator=gen))
        cams[i, 12:] = torch.tensor([f, f * 1.01, H / 2 + 1.5, H / 2 - 2.0, float(H), float(H)])
    return cams


def main():
    ref = load_reference_module()
    gen = torch.Generator().manual_seed(20240607)

    # ---- converter, three modes, forward + backward
    B, N, h, w = 2, 2, 12, 20
    for mode in ("mvsplat", "gslrm", "director3d"):
        conv = ref.GaussianConverter(gs_convert_mode=mode, z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0)
        params = torch.randn(B, N, 14, h, w, generator=gen)
        params[:, :, 5] = params[:, :, 5] * 0.3 - 3.9
        params[:, :, 3:5] *= 0.01
        cams = rand_cameras(gen, B * N, 64, H=64).reshape(B, N, 18)
        params.requires_grad_(True)
        outs = conv(params, cams)
        ws = [torch.randn(o.shape, generator=gen) for o in outs]
        loss = sum((o * w_).sum() for o, w_ in zip(outs, ws))
        loss.backward()
        np.savez_compressed(
            os.path.join(HERE, f"converter_{mode}.npz"), params=params.detach().numpy(), cameras=cams.numpy(),
            muls=conv.muls.numpy(), xyz=outs[0].detach().numpy(), features=outs[1].detach().numpy(),
            opacity=outs[2].detach().numpy(), scales=outs[3].detach().numpy(), rotations=outs[4].detach().numpy(),
            w_xyz=ws[0].numpy(), w_features=ws[1].numpy(), w_opacity=ws[2].numpy(), w_scales=ws[3].numpy(),
            w_rotations=ws[4].numpy(), v_params=params.grad.numpy())

    # ---- camera preparation
    rend = ref.GaussianRenderer(sh_degree=0, background=[1, 1, 1])
    cams = rand_cameras(gen, 6, 96, H=128)
    vms, Ks = [], []
    for i in range(6):
        cam = rend.convert_camera_parameters_into_viewpoint_cameras(cams[i], h=96, w=96)
        vms.append(cam.world_view_transform.transpose(0, 1).numpy())
        # the K build of gaussians_renderer.py:405-412, verbatim arithmetic (math.tan on the fp32 fov)
        tanfovx, tanfovy = math.tan((cam.FoVx * 0.5).item()), math.tan((cam.FoVy * 0.5).item())
        Ks.append(torch.tensor([[cam.image_width / (2 * tanfovx), 0, cam.image_width / 2.0],
                                [0, cam.image_height / (2 * tanfovy), cam.image_height / 2.0], [0, 0, 1]]).numpy())
    np.savez_compressed(os.path.join(HERE, "cameras.npz"), cameras=cams.numpy(), h=96, w=96,
                        viewmats=np.stack(vms).astype(np.float32), Ks=np.stack(Ks))

    # ---- reference renderer end to end over the oracle operator
    sys.path.insert(0, ROOT)
    from prometheus_b200 import scenes

    sc = scenes.make_scene("C1", seed=77, views_in=2, size_in=24, views_out=3, size_out=40, scenes=2)
    conv = ref.GaussianConverter(gs_convert_mode="mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0)
    params = sc.params.clone().requires_grad_(True)
    gaussians = conv(params, sc.cameras_in)
    images, depths, masks, _, _ = rend(sc.cameras_out, gaussians, bg_color=(0.2, 0.5, 0.9), h=40, w=40)
    tgt = torch.rand(images.shape, generator=gen) * 2 - 1
    loss = scenes.render_loss(images, depths, masks, tgt) * images.numel()
    loss.backward()
    np.savez_compressed(
        os.path.join(HERE, "renderer_e2e.npz"), params=sc.params.numpy(), cameras_in=sc.cameras_in.numpy(),
        cameras_out=sc.cameras_out.numpy(), bg=np.array([0.2, 0.5, 0.9], np.float32), h=40, w=40,
        images=images.detach().numpy(), depths=depths.detach().numpy(), masks=masks.detach().numpy(),
        target=tgt.numpy(), v_params=params.grad.numpy(),
        radii=np.stack([r.numpy() for r in rend.radii]),
        means2d_grad=np.stack([p.grad.numpy() for p in rend.viewspace_points]))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
