"""Pure-PyTorch (CPU, fp32) restatement of the rasterisation path — the analogue
of gsplat's `_torch_impl` reference that BASELINE.json names as the CPU baseline.

TEST INFRASTRUCTURE ONLY (same rules as gs_oracle.c). PARITY UNPINNED: gsplat
1.4.0 is not available, so this restates its published algorithm (SURVEY.md §8c
steps 1-11), anchored on the reference call site
prometheus/modules/renderers/gaussians_renderer.py:413-427.

Two jobs:
  * an INDEPENDENT second restatement: its integer stage outputs must equal the C
    oracle's bit for bit (tests/test_oracle.py), and `torch.autograd` through it
    checks the C oracle's hand-written backward;
  * the "pure-PyTorch on host cores" CPU baseline timed by bench.py.

Every arithmetic step of the projection is written elementwise in the same
order as gs_oracle.c (torch CPU kernels round each op once, no FMA contraction),
so radii / tile counts / keys agree exactly. The compositing is vectorised per
tile with sequential semantics (cumprod), in chunks with early termination.
"""
from __future__ import annotations

import torch


def _sum3(a, b, c):
    return (a + b) + c


def _sqrt(x):
    """Correctly rounded fp32 sqrt. torch.sqrt on CPU fp32 is vectorised with a
    <=1-ulp routine (measured: 0.7% of inputs differ from IEEE), which would break
    the bit-exact agreement with gs_oracle.c; sqrt in fp64 rounded to fp32 is exact."""
    return torch.sqrt(x.double()).float()


def quat_to_rotmat(q):
    w, x, y, z = q.unbind(-1)
    inv = 1.0 / _sqrt(((x * x + y * y) + z * z) + w * w)
    x, y, z, w = x * inv, y * inv, z * inv, w * inv
    x2, y2, z2 = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    R = [1.0 - 2.0 * (y2 + z2), 2.0 * (xy - wz), 2.0 * (xz + wy),
         2.0 * (xy + wz), 1.0 - 2.0 * (x2 + z2), 2.0 * (yz - wx),
         2.0 * (xz - wy), 2.0 * (yz + wx), 1.0 - 2.0 * (x2 + y2)]
    return R  # row-major list of 9 tensors [N]


def project(means, quats, scales, viewmat, K, width, height, eps2d=0.3, near_plane=0.01,
            far_plane=1e10, radius_clip=0.0):
    """One camera. Returns radii [N] i32, means2d [N,2], depths [N], conics [N,3]."""
    R = [viewmat[r, k] for r in range(3) for k in range(3)]
    t = [viewmat[r, 3] for r in range(3)]
    mx, my, mz = means.unbind(-1)
    px = _sum3(R[0] * mx, R[1] * my, R[2] * mz) + t[0]
    py = _sum3(R[3] * mx, R[4] * my, R[5] * mz) + t[1]
    pz = _sum3(R[6] * mx, R[7] * my, R[8] * mz) + t[2]
    valid = ~((pz < near_plane) | (pz > far_plane))

    Rq = quat_to_rotmat(quats)
    s = scales.unbind(-1)
    M = [Rq[r * 3 + k] * s[k] for r in range(3) for k in range(3)]
    S = {}
    for (i, j) in ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)):
        S[(i, j)] = _sum3(M[i * 3] * M[j * 3], M[i * 3 + 1] * M[j * 3 + 1], M[i * 3 + 2] * M[j * 3 + 2])
        S[(j, i)] = S[(i, j)]
    T = [[_sum3(R[r * 3] * S[(0, k)], R[r * 3 + 1] * S[(1, k)], R[r * 3 + 2] * S[(2, k)]) for k in range(3)]
         for r in range(3)]

    def c_(i, j):
        return _sum3(T[i][0] * R[j * 3], T[i][1] * R[j * 3 + 1], T[i][2] * R[j * 3 + 2])

    c00, c01, c02, c11, c12, c22 = c_(0, 0), c_(0, 1), c_(0, 2), c_(1, 1), c_(1, 2), c_(2, 2)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    Wf = torch.tensor(float(width))
    Hf = torch.tensor(float(height))
    tan_fovx = 0.5 * Wf / fx
    tan_fovy = 0.5 * Hf / fy
    lim_x_pos = (Wf - cx) / fx + 0.3 * tan_fovx
    lim_x_neg = cx / fx + 0.3 * tan_fovx
    lim_y_pos = (Hf - cy) / fy + 0.3 * tan_fovy
    lim_y_neg = cy / fy + 0.3 * tan_fovy
    rz = 1.0 / pz
    rz2 = rz * rz
    tx = pz * torch.minimum(lim_x_pos, torch.maximum(-lim_x_neg, px * rz))
    ty = pz * torch.minimum(lim_y_pos, torch.maximum(-lim_y_neg, py * rz))
    J00, J02 = fx * rz, -((fx * tx) * rz2)
    J11, J12 = fy * rz, -((fy * ty) * rz2)
    A00 = J00 * c00 + J02 * c02
    A01 = J00 * c01 + J02 * c12
    A02 = J00 * c02 + J02 * c22
    A11 = J11 * c11 + J12 * c12
    A12 = J11 * c12 + J12 * c22
    v00 = (A00 * J00 + A02 * J02) + eps2d
    v01 = A01 * J11 + A02 * J12
    v11 = (A11 * J11 + A12 * J12) + eps2d
    m2x = (fx * px) * rz + cx
    m2y = (fy * py) * rz + cy
    det = v00 * v11 - v01 * v01
    valid = valid & (det > 0)
    inv_det = 1.0 / det
    b = 0.5 * (v00 + v11)
    v1 = b + _sqrt(torch.clamp(b * b - det, min=0.01))
    radius = torch.ceil(3.0 * _sqrt(v1))
    valid = valid & ~(radius <= radius_clip)
    valid = valid & ~((m2x + radius <= 0) | (m2x - radius >= width) | (m2y + radius <= 0) | (m2y - radius >= height))
    zero = torch.zeros_like(m2x)
    radii = torch.where(valid, radius.detach(), zero).to(torch.int32)
    means2d = torch.stack([torch.where(valid, m2x, zero), torch.where(valid, m2y, zero)], -1)
    depths = torch.where(valid, pz, zero)
    conics = torch.stack([torch.where(valid, v11 * inv_det, zero), torch.where(valid, -v01 * inv_det, zero),
                          torch.where(valid, v00 * inv_det, zero)], -1)
    return radii, means2d, depths, conics


def n_bits(n: int) -> int:
    return int(n).bit_length()


def isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height):
    """means2d [C,N,2], radii [C,N], depths [C,N] -> tiles_per_gauss, sorted isect_ids,
    flatten_ids, offsets [C,th,tw]  (steps 6-9)."""
    C, N = radii.shape
    ts = float(tile_size)
    m = means2d.detach()
    tr = radii.to(torch.float32) / ts
    txc, tyc = m[..., 0] / ts, m[..., 1] / ts
    x0 = torch.clamp(torch.floor(txc - tr), 0, tile_width).to(torch.int64)
    x1 = torch.clamp(torch.ceil(txc + tr), 0, tile_width).to(torch.int64)
    y0 = torch.clamp(torch.floor(tyc - tr), 0, tile_height).to(torch.int64)
    y1 = torch.clamp(torch.ceil(tyc + tr), 0, tile_height).to(torch.int64)
    nx, ny = x1 - x0, y1 - y0
    tpg = torch.where(radii > 0, nx * ny, torch.zeros_like(nx))
    flat_tpg = tpg.flatten()
    n = int(flat_tpg.sum())
    f = torch.repeat_interleave(torch.arange(C * N), flat_tpg)          # flatten id of each isect
    start = torch.cumsum(flat_tpg, 0) - flat_tpg
    k = torch.arange(n) - start[f]                                      # index within the pair's AABB
    nxf = nx.flatten()[f]
    ty = y0.flatten()[f] + torch.div(k, nxf, rounding_mode="floor")
    tx = x0.flatten()[f] + k % nxf
    tile_n_bits = n_bits(tile_width * tile_height)
    dbits = depths.detach().flatten()[f].contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    cid = torch.div(f, N, rounding_mode="floor")
    ids = (cid << (32 + tile_n_bits)) | ((ty * tile_width + tx) << 32) | dbits
    ids_sorted, order = torch.sort(ids, stable=True)
    flat_sorted = f[order].to(torch.int32)
    tile_key = ids_sorted >> 32
    n_tiles = tile_width * tile_height
    lin = (tile_key >> tile_n_bits) * n_tiles + (tile_key & ((1 << tile_n_bits) - 1))
    offsets = torch.searchsorted(lin, torch.arange(C * n_tiles)).to(torch.int32).reshape(C, tile_height, tile_width)
    return tpg.to(torch.int32), ids_sorted, flat_sorted, offsets


def composite(means2d, conics, opacities, colors, backgrounds, offsets, flatten_ids, width, height,
              tile_size=16, chunk=512):
    """means2d [C,N,2], conics [C,N,3], opacities [N], colors [C,N,CD] (depth already
    appended), backgrounds [C,CD] or None -> render [C,H,W,CD], alphas [C,H,W,1] (step 10)."""
    C, N = means2d.shape[:2]
    CD = colors.shape[-1]
    th, tw = offsets.shape[1:]
    n_isects = flatten_ids.shape[0]
    m2 = means2d.reshape(C * N, 2)
    con = conics.reshape(C * N, 3)
    col = colors.reshape(C * N, CD)
    off = offsets.flatten().tolist() + [n_isects]
    render = torch.zeros(C, height, width, CD)
    alphas = torch.zeros(C, height, width, 1)
    tiles_r, tiles_a = [], []
    for ct in range(C * th * tw):
        c, tile = divmod(ct, th * tw)
        tyi, txi = divmod(tile, tw)
        ys = torch.arange(tyi * tile_size, min((tyi + 1) * tile_size, height))
        xs = torch.arange(txi * tile_size, min((txi + 1) * tile_size, width))
        py = (ys.to(torch.float32) + 0.5)[:, None].expand(len(ys), len(xs)).reshape(-1)
        px = (xs.to(torch.float32) + 0.5)[None, :].expand(len(ys), len(xs)).reshape(-1)
        npx = px.shape[0]
        T = torch.ones(npx)
        done = torch.zeros(npx, dtype=torch.bool)
        out = torch.zeros(npx, CD)
        for s in range(off[ct], off[ct + 1], chunk):
            if bool(done.all()):
                break
            f = flatten_ids[s:min(s + chunk, off[ct + 1])].long()
            g = f % N
            dx = m2[f, 0][:, None] - px[None]
            dy = m2[f, 1][:, None] - py[None]
            cq = con[f]
            sigma = 0.5 * (cq[:, 0:1] * dx * dx + cq[:, 2:3] * dy * dy) + cq[:, 1:2] * dx * dy
            alpha = torch.clamp(opacities[g][:, None] * torch.exp(-sigma), max=0.999)
            skip = (sigma < 0) | (alpha < 1.0 / 255.0)
            a = torch.where(skip, torch.zeros_like(alpha), alpha)
            next_T = T[None] * torch.cumprod(1.0 - a, dim=0)             # T after each Gaussian
            T_before = torch.cat([T[None], next_T[:-1]], 0)
            stop = (~skip) & (next_T <= 1e-4)
            stopped = (torch.cumsum(stop.to(torch.int32), 0) > 0) | done[None]   # at or after the stop
            w = torch.where(stopped, torch.zeros_like(a), a * T_before)
            out = out + w.t() @ col[f]
            # transmittance after the chunk: value just before the first stop, else the last
            any_stop = stop.any(0) & ~done
            first = torch.argmax(stop.to(torch.int8), 0)
            T_new = torch.where(any_stop, T_before.gather(0, first[None])[0], next_T[-1])
            T = torch.where(done, T, T_new)
            done = done | any_stop
        r = out if backgrounds is None else out + T[:, None] * backgrounds[c][None]
        render[c, ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1] = r.reshape(len(ys), len(xs), CD)
        alphas[c, ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1, 0] = (1.0 - T).reshape(len(ys), len(xs))
    return render, alphas


def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                  near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, tile_size=16,
                  backgrounds=None, render_mode="RGB", **_unused):
    """gsplat.rasterization contract (packed=False) in plain torch; differentiable."""
    C = viewmats.shape[0]
    N = means.shape[0]
    tw = (width + tile_size - 1) // tile_size
    th = (height + tile_size - 1) // tile_size
    pr = [project(means, quats, scales, viewmats[c], Ks[c], width, height, eps2d, near_plane, far_plane,
                  radius_clip) for c in range(C)]
    radii = torch.stack([p[0] for p in pr])
    means2d = torch.stack([p[1] for p in pr])
    depths = torch.stack([p[2] for p in pr])
    conics = torch.stack([p[3] for p in pr])
    tpg, ids, flat, offsets = isect_tiles(means2d, radii, depths, tile_size, tw, th)
    col = colors if colors.dim() == 3 else colors[None].expand(C, -1, -1)
    bg = backgrounds
    if render_mode in ("D", "ED"):
        col = depths[..., None]
        bg = None
    elif render_mode in ("RGB+D", "RGB+ED"):
        col = torch.cat([col, depths[..., None]], -1)
        if bg is not None:
            bg = torch.cat([bg, torch.zeros(C, 1)], -1)
    render, alphas = composite(means2d, conics, opacities, col, bg, offsets, flat, width, height, tile_size)
    if render_mode in ("ED", "RGB+ED"):
        render = torch.cat([render[..., :-1], render[..., -1:] / alphas.clamp(min=1e-10)], -1)
    meta = dict(radii=radii, means2d=means2d, depths=depths, conics=conics, tiles_per_gauss=tpg,
                isect_ids=ids, flatten_ids=flat, isect_offsets=offsets, tile_width=tw, tile_height=th,
                width=width, height=height, tile_size=tile_size, n_cameras=C,
                opacities=opacities[None].expand(C, N))
    return render, alphas, meta
