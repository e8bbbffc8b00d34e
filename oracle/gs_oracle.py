"""ctypes/numpy binding of the CPU oracle (oracle/gs_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of gs_oracle.c. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. PARITY UNPINNED for stages 2-4 (gsplat 1.4.0 is not
available); stage 1 and the camera preparation are pinned by tests/golden/.

The high-level entry point `rasterization(...)` mirrors the keyword contract of
gsplat.rasterization as called at
prometheus/modules/renderers/gaussians_renderer.py:413-427 and returns torch CPU
tensors wired into autograd, so the *unmodified* reference module can be run on
top of it (tests/golden/make_golden.py does exactly that).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from ctypes import POINTER, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgs_oracle.so")
_lib = None

MODES = {"mvsplat": 0, "gslrm": 1, "director3d": 2}


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "gs_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "all"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_bin_emit.restype = c_int64
        _lib.orc_num_threads.restype = c_int
        _lib.orc_tile_n_bits.restype = c_int
        _lib.orc_cam_n_bits.restype = c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(c_int(int(n)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ty=c_float):
    if a is None:
        return None
    return a.ctypes.data_as(POINTER(ty))


# --------------------------------------------------------------------------- stage 1
def unproject(params, cameras, muls=None, mode="mvsplat", z_near=0.01, z_far=100.0,
              s_min=0.001, s_max=1.0):
    """params [B,N,14,h,w], cameras [B,N,18] -> dict of per-scene arrays
    (restates GaussianConverter.forward, gaussians_renderer.py:210-287)."""
    params = _f32(params)
    cameras = _f32(cameras)
    B, N, Cc, h, w = params.shape
    assert Cc == 14
    V = B * N
    G = V * h * w
    if muls is None:
        muls = converter_muls(mode)
    muls = _f32(muls).reshape(14)
    xyz = np.empty((G, 3), np.float32)
    rgb = np.empty((G, 3), np.float32)
    opacity = np.empty((G,), np.float32)
    scales = np.empty((G, 3), np.float32)
    quats = np.empty((G, 4), np.float32)
    lib().orc_unproject(_p(params), _p(cameras), _p(muls), c_int(MODES[mode]),
                        c_float(z_near), c_float(z_far), c_float(s_min), c_float(s_max),
                        c_int(V), c_int(h), c_int(w),
                        _p(xyz), _p(rgb), _p(opacity), _p(scales), _p(quats))
    g = N * h * w
    return dict(xyz=xyz.reshape(B, g, 3), features=rgb.reshape(B, g, 1, 3),
                opacity=opacity.reshape(B, g, 1), scales=scales.reshape(B, g, 3),
                rotations=quats.reshape(B, g, 4))


def converter_muls(mode="mvsplat"):
    """`GaussianConverter.muls` buffer (gaussians_renderer.py:188-207)."""
    if mode == "director3d":
        m = np.array([0.01] * 3 + [0.01] * 2 + [0.05] + [0.05] + [0.005] * 3 + [0.005] * 4, np.float32)
        return (m / m.max()).astype(np.float32)
    return np.ones(14, np.float32)


def camera_prep(cameras, h, w):
    """cameras [n,18] -> (viewmats [n,4,4], Ks [n,3,3]) following the reference's
    HOST path op for op: convert_camera_parameters_into_viewpoint_cameras
    (gaussians_renderer.py:453-472), MiniCam (:139-169), K build (:404-412)."""
    cameras = _f32(cameras).reshape(-1, 18)
    n = cameras.shape[0]
    viewmats = np.zeros((n, 4, 4), np.float32)
    Ks = np.zeros((n, 3, 3), np.float32)
    for i in range(n):
        cam = cameras[i]
        c2w = np.eye(4, dtype=np.float32)
        c2w[:3, :] = cam[:12].reshape(3, 4)
        fx, fy, cx, cy, H, W = [np.float32(v) for v in cam[12:]]
        fx = np.float32(np.float32(fx * np.float32(h)) / H)
        fy = np.float32(np.float32(fy * np.float32(w)) / W)
        fovy = np.float32(2) * np.arctan(np.float32(np.float32(0.5 * w) / fy), dtype=np.float32)
        fovx = np.float32(2) * np.arctan(np.float32(np.float32(0.5 * h) / fx), dtype=np.float32)
        c2w[:3, 1:3] *= -1
        w2c = np.linalg.inv(c2w)
        viewmats[i] = w2c
        tanfovx = math.tan(float(np.float32(fovx * np.float32(0.5))))
        tanfovy = math.tan(float(np.float32(fovy * np.float32(0.5))))
        Ks[i] = np.array([[w / (2 * tanfovx), 0, w / 2.0], [0, h / (2 * tanfovy), h / 2.0], [0, 0, 1]],
                         dtype=np.float32)
    return viewmats, Ks


# --------------------------------------------------------------------------- stages 2-4
def project(means, quats, scales, viewmats, Ks, width, height, tile_size=16, eps2d=0.3,
            near_plane=0.01, far_plane=1e10, radius_clip=0.0):
    means, quats, scales = _f32(means), _f32(quats), _f32(scales)
    viewmats, Ks = _f32(viewmats), _f32(Ks)
    N, C = means.shape[0], viewmats.shape[0]
    radii = np.zeros((C, N), np.int32)
    means2d = np.zeros((C, N, 2), np.float32)
    depths = np.zeros((C, N), np.float32)
    conics = np.zeros((C, N, 3), np.float32)
    tpg = np.zeros((C, N), np.int32)
    lib().orc_project(_p(means), _p(quats), _p(scales), _p(viewmats), _p(Ks), c_int(N), c_int(C),
                      c_int(width), c_int(height), c_int(tile_size), c_float(eps2d),
                      c_float(near_plane), c_float(far_plane), c_float(radius_clip),
                      _p(radii, c_int32), _p(means2d), _p(depths), _p(conics), _p(tpg, c_int32))
    return dict(radii=radii, means2d=means2d, depths=depths, conics=conics, tiles_per_gauss=tpg)


def tile_bits(tile_width, tile_height, C):
    L = lib()
    return int(L.orc_tile_n_bits(c_int(tile_width * tile_height))), int(L.orc_cam_n_bits(c_int(C)))


def bin_emit(means2d, radii, depths, tile_size, tile_width, tile_height):
    means2d, depths = _f32(means2d), _f32(depths)
    radii = np.ascontiguousarray(radii, np.int32)
    C, N = radii.shape
    args = (_p(means2d), _p(radii, c_int32), _p(depths), c_int(N), c_int(C), c_int(tile_size),
            c_int(tile_width), c_int(tile_height))
    n = int(lib().orc_bin_emit(*args, None, None))
    ids = np.zeros((n,), np.int64)
    flat = np.zeros((n,), np.int32)
    if n:
        lib().orc_bin_emit(*args, _p(ids, c_int64), _p(flat, c_int32))
    return ids, flat


def sort_pairs(keys, vals, end_bit):
    keys = np.array(keys, dtype=np.int64, copy=True)
    vals = np.array(vals, dtype=np.int32, copy=True)
    lib().orc_sort_pairs(_p(keys, c_int64), _p(vals, c_int32), c_int64(keys.shape[0]), c_int(end_bit))
    return keys, vals


def tile_offsets(sorted_ids, C, tile_width, tile_height):
    sorted_ids = np.ascontiguousarray(sorted_ids, np.int64)
    off = np.zeros((C, tile_height, tile_width), np.int32)
    lib().orc_tile_offsets(_p(sorted_ids, c_int64), c_int64(sorted_ids.shape[0]), c_int(C),
                           c_int(tile_width), c_int(tile_height), _p(off, c_int32))
    return off


def _color_args(colors, C, N):
    colors = _f32(colors)
    per_cam = 1 if colors.ndim == 3 else 0
    D = colors.shape[-1]
    if per_cam:
        assert colors.shape[:2] == (C, N)
    else:
        assert colors.shape[0] == N
    return colors, per_cam, D


def composite_fwd(means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids,
                  width, height, tile_size=16, want_fragile=True):
    means2d, conics, opacities = _f32(means2d), _f32(conics), _f32(opacities)
    C, N = means2d.shape[:2]
    colors, per_cam, D = _color_args(colors, C, N)
    depths = None if depths is None else _f32(depths)
    backgrounds = None if backgrounds is None else _f32(backgrounds)
    offsets = np.ascontiguousarray(offsets, np.int32)
    flatten_ids = np.ascontiguousarray(flatten_ids, np.int32)
    CD = D + (0 if depths is None else 1)
    render = np.zeros((C, height, width, CD), np.float32)
    alphas = np.zeros((C, height, width), np.float32)
    last_ids = np.zeros((C, height, width), np.int32)
    fragile = np.zeros((C, height, width), np.uint8) if want_fragile else None
    trans = np.zeros((C, height, width), np.float32)
    lib().orc_composite_fwd(_p(means2d), _p(conics), _p(opacities), _p(colors), c_int(per_cam), c_int(D),
                            _p(depths), _p(backgrounds), _p(offsets, c_int32), _p(flatten_ids, c_int32),
                            c_int64(flatten_ids.shape[0]), c_int(N), c_int(C), c_int(width), c_int(height),
                            c_int(tile_size), _p(render), _p(alphas), _p(last_ids, c_int32),
                            _p(fragile, c_uint8), _p(trans))
    composite_fwd.transmittance = trans  # the exact final T per pixel of the last call (see composite_bwd)
    return render, alphas, last_ids, fragile


def composite_bwd(means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids,
                  width, height, alphas, last_ids, v_render, v_alphas, tile_size=16, transmittance=None):
    """transmittance [C,H,W] = the forward's exact final T (what the CUDA path keeps); None = upstream's
    T_final = 1 - alphas, which rounds away up to ~3e-4 relative of it on opaque pixels."""
    means2d, conics, opacities = _f32(means2d), _f32(conics), _f32(opacities)
    C, N = means2d.shape[:2]
    colors, per_cam, D = _color_args(colors, C, N)
    depths = None if depths is None else _f32(depths)
    backgrounds = None if backgrounds is None else _f32(backgrounds)
    offsets = np.ascontiguousarray(offsets, np.int32)
    flatten_ids = np.ascontiguousarray(flatten_ids, np.int32)
    alphas = _f32(alphas)
    last_ids = np.ascontiguousarray(last_ids, np.int32)
    v_render = None if v_render is None else _f32(v_render)
    v_alphas = None if v_alphas is None else _f32(v_alphas)
    transmittance = None if transmittance is None else _f32(transmittance)
    v_means2d = np.zeros((C, N, 2), np.float64)
    v_conics = np.zeros((C, N, 3), np.float64)
    v_colors = np.zeros(colors.shape, np.float64)
    v_opac = np.zeros((N,), np.float64)
    v_depths = np.zeros((C, N), np.float64) if depths is not None else None
    lib().orc_composite_bwd(_p(means2d), _p(conics), _p(opacities), _p(colors), c_int(per_cam), c_int(D),
                            _p(depths), _p(backgrounds), _p(offsets, c_int32), _p(flatten_ids, c_int32),
                            c_int64(flatten_ids.shape[0]), c_int(N), c_int(C), c_int(width), c_int(height),
                            c_int(tile_size), _p(alphas), _p(last_ids, c_int32), _p(v_render), _p(v_alphas),
                            _p(v_means2d, c_double), _p(v_conics, c_double), _p(v_colors, c_double),
                            _p(v_opac, c_double), _p(v_depths, c_double), _p(transmittance))
    return dict(v_means2d=v_means2d, v_conics=v_conics, v_colors=v_colors, v_opacities=v_opac,
                v_depths=v_depths)


def project_bwd(means, quats, scales, viewmats, Ks, width, height, radii, conics,
                v_means2d, v_depths, v_conics, eps2d=0.3):
    means, quats, scales = _f32(means), _f32(quats), _f32(scales)
    viewmats, Ks = _f32(viewmats), _f32(Ks)
    N, C = means.shape[0], viewmats.shape[0]
    radii = np.ascontiguousarray(radii, np.int32)
    conics = _f32(conics)
    v_means2d = None if v_means2d is None else _f32(v_means2d)
    v_depths = None if v_depths is None else _f32(v_depths)
    v_conics = None if v_conics is None else _f32(v_conics)
    v_means = np.zeros((N, 3), np.float64)
    v_quats = np.zeros((N, 4), np.float64)
    v_scales = np.zeros((N, 3), np.float64)
    lib().orc_project_bwd(_p(means), _p(quats), _p(scales), _p(viewmats), _p(Ks), c_int(N), c_int(C),
                          c_int(width), c_int(height), c_float(eps2d), _p(radii, c_int32), _p(conics),
                          _p(v_means2d), _p(v_depths), _p(v_conics),
                          _p(v_means, c_double), _p(v_quats, c_double), _p(v_scales, c_double))
    return dict(v_means=v_means, v_quats=v_quats, v_scales=v_scales)


# --------------------------------------------------------------------------- pipeline
def forward_pipeline(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                     near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, tile_size=16,
                     backgrounds=None, with_depth=True, want_fragile=True):
    """All forward stages on numpy arrays; returns a dict with every intermediate
    (the bit-exact parity surface) plus render / alphas / last_ids / fragile."""
    C = np.asarray(viewmats).shape[0]
    tw = (width + tile_size - 1) // tile_size
    th = (height + tile_size - 1) // tile_size
    pr = project(means, quats, scales, viewmats, Ks, width, height, tile_size, eps2d, near_plane,
                 far_plane, radius_clip)
    ids, flat = bin_emit(pr["means2d"], pr["radii"], pr["depths"], tile_size, tw, th)
    tb, cb = tile_bits(tw, th, C)
    ids_s, flat_s = sort_pairs(ids, flat, 32 + tb + cb)
    off = tile_offsets(ids_s, C, tw, th)
    render, alphas, last_ids, fragile = composite_fwd(
        pr["means2d"], pr["conics"], opacities, colors, pr["depths"] if with_depth else None,
        backgrounds, off, flat_s, width, height, tile_size, want_fragile)
    out = dict(pr)
    out.update(isect_ids_unsorted=ids, flatten_ids_unsorted=flat, isect_ids=ids_s, flatten_ids=flat_s,
               isect_offsets=off, render=render, alphas=alphas, last_ids=last_ids, fragile=fragile,
               transmittance=composite_fwd.transmittance,
               tile_width=tw, tile_height=th, tile_n_bits=tb, cam_n_bits=cb)
    return out


def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                  near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, sh_degree=None,
                  packed=False, tile_size=16, backgrounds=None, render_mode="RGB", exact_transmittance=True,
                  **_unused):
    """gsplat.rasterization stand-in on torch CPU tensors, with autograd wired as
    two nodes (projection -> compositing) like upstream, so `meta["means2d"]`
    accepts retain_grad() (reference call site gaussians_renderer.py:413-435).
    exact_transmittance=False makes the backward rebuild T_final = 1 - alpha as
    upstream does (see composite_bwd). Oracle use only."""
    import torch

    assert sh_degree is None and not packed
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED")
    with_depth = render_mode in ("D", "ED", "RGB+D", "RGB+ED")
    only_depth = render_mode in ("D", "ED")
    vm_np = viewmats.detach().numpy().astype(np.float32)
    K_np = Ks.detach().numpy().astype(np.float32)
    C = vm_np.shape[0]
    tw = (width + tile_size - 1) // tile_size
    th = (height + tile_size - 1) // tile_size
    st = {}

    class _Project(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means, quats, scales):
            m, q, s = (t.detach().contiguous().numpy() for t in (means, quats, scales))
            pr = project(m, q, s, vm_np, K_np, width, height, tile_size, eps2d, near_plane, far_plane,
                         radius_clip)
            st.update(pr)
            ctx.inputs = (m, q, s)
            return (torch.from_numpy(pr["means2d"].copy()), torch.from_numpy(pr["depths"].copy()),
                    torch.from_numpy(pr["conics"].copy()))

        @staticmethod
        def backward(ctx, v_means2d, v_depths, v_conics):
            m, q, s = ctx.inputs
            tonp = lambda t: None if t is None else t.contiguous().numpy()
            pg = project_bwd(m, q, s, vm_np, K_np, width, height, st["radii"], st["conics"],
                             tonp(v_means2d), tonp(v_depths), tonp(v_conics), eps2d)
            f32 = lambda a: torch.from_numpy(a.astype(np.float32))
            return f32(pg["v_means"]), f32(pg["v_quats"]), f32(pg["v_scales"])

    class _Composite(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means2d, conics, depths, opacities, colors):
            m2, con, dep, o = (t.detach().contiguous().numpy() for t in (means2d, conics, depths, opacities))
            col = np.zeros((o.shape[0], 0), np.float32) if only_depth else colors.detach().contiguous().numpy()
            bg = None if backgrounds is None or only_depth else backgrounds.detach().numpy().astype(np.float32)
            ids, flat = bin_emit(m2, st["radii"], dep, tile_size, tw, th)
            tb, cb = tile_bits(tw, th, C)
            ids_s, flat_s = sort_pairs(ids, flat, 32 + tb + cb)
            off = tile_offsets(ids_s, C, tw, th)
            render, alphas, last_ids, fragile = composite_fwd(
                m2, con, o, col, dep if with_depth else None, bg, off, flat_s, width, height, tile_size, True)
            st.update(isect_ids=ids_s, flatten_ids=flat_s, isect_offsets=off, last_ids=last_ids,
                      fragile=fragile, alphas=alphas, tile_n_bits=tb, cam_n_bits=cb,
                      transmittance=composite_fwd.transmittance)
            ctx.saved = (m2, con, dep, o, col, bg)
            return torch.from_numpy(render), torch.from_numpy(alphas)[..., None]

        @staticmethod
        def backward(ctx, v_render, v_alphas):
            m2, con, dep, o, col, bg = ctx.saved
            vr = None if v_render is None else v_render.contiguous().numpy()
            va = None if v_alphas is None else v_alphas.contiguous().numpy()[..., 0]
            g = composite_bwd(m2, con, o, col, dep if with_depth else None, bg, st["isect_offsets"],
                              st["flatten_ids"], width, height, st["alphas"], st["last_ids"], vr, va, tile_size,
                              transmittance=st["transmittance"] if exact_transmittance else None)
            f32 = lambda a: None if a is None else torch.from_numpy(a.astype(np.float32))
            return (f32(g["v_means2d"]), f32(g["v_conics"]), f32(g["v_depths"]), f32(g["v_opacities"]),
                    None if only_depth else f32(g["v_colors"]))

    means2d, depths, conics = _Project.apply(means, quats, scales)
    render, alphas = _Composite.apply(means2d, conics, depths, opacities, colors)
    if render_mode in ("ED", "RGB+ED"):
        render = torch.cat([render[..., :-1], render[..., -1:] / alphas.clamp(min=1e-10)], dim=-1)
    meta = {"means2d": means2d, "depths": depths, "conics": conics, "width": width, "height": height,
            "tile_size": tile_size, "n_cameras": C, "tile_width": tw, "tile_height": th,
            "opacities": opacities[None].expand(C, -1)}
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets", "fragile", "last_ids"):
        meta[k] = torch.from_numpy(np.asarray(st[k]))
    return render, alphas, meta
