"""Non-square camera fixture: the reference's camera preparation with h != w AND H != W.

convert_camera_parameters_into_viewpoint_cameras (gaussians_renderer.py:453-472) rescales `fx, cx` by h / H and
`fy, cy` by w / W and pairs fovy with w, fovx with h (:469-470); MiniCam / the K build (:139-169, :404-412) then
turn them back into focal lengths with the OTHER pairing. For square images the swaps cancel; this fixture pins what
the reference really computes when they do not. Runs the reference's own code (needs /root/reference):
    python tests/golden/make_golden_cameras_nonsquare.py
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference_module, rand_cameras  # noqa: E402


def main():
    ref = load_reference_module()
    gen = torch.Generator().manual_seed(20261018)
    rend = ref.GaussianRenderer(sh_degree=0, background=[1, 1, 1])
    cams = rand_cameras(gen, 6, 96, H=128)
    cams[:, 17] = 160.0                     # W != H in the camera record
    cams[:, 13] = cams[:, 12] * 1.2         # fy != fx
    cams[:, 14] = 81.0                      # cx, cy off-centre (the K build ignores them)
    h, w = 72, 104                          # render size: h != w
    vms, Ks = [], []
    for i in range(6):
        cam = rend.convert_camera_parameters_into_viewpoint_cameras(cams[i].clone(), h=h, w=w)
        vms.append(cam.world_view_transform.transpose(0, 1).numpy())
        tanfovx, tanfovy = math.tan((cam.FoVx * 0.5).item()), math.tan((cam.FoVy * 0.5).item())
        Ks.append(torch.tensor([[cam.image_width / (2 * tanfovx), 0, cam.image_width / 2.0],
                                [0, cam.image_height / (2 * tanfovy), cam.image_height / 2.0], [0, 0, 1]]).numpy())
        assert (cam.image_width, cam.image_height) == (w, h)
    np.savez_compressed(os.path.join(HERE, "cameras_nonsquare.npz"), cameras=cams.numpy(), h=h, w=w,
                        viewmats=np.stack(vms).astype(np.float32), Ks=np.stack(Ks))
    print(np.stack(Ks)[0])


if __name__ == "__main__":
    main()
