"""CPU restatement (numpy, fp32) of the reference's trajectory pose sampler — TEST INFRASTRUCTURE ONLY
(imported by tests/ and bench.py's CPU legs; nothing under prometheus_b200/ touches it).

Follows prometheus/utils/camera.py:134-217 (sample_from_dense_cameras) with matrix_to_quaternion
(prometheus/utils/math.py:46-91), quaternion_slerp (camera.py:81-109) and quaternion_to_matrix (math.py:18-44),
for the only call pattern the reference's broadcasting supports and export_video uses (io.py:167-190): one
trajectory, one sample per call; `sample_from_dense_cameras(dense [N,18], t [M])` returns the M calls stacked.
PINNED: tests/golden/trajectory.npz holds the reference function's own outputs (tests/golden/make_golden_trajectory.py).
Also restates the frame packing of export_video (io.py:181): uint8(clamp(x/2 + 1/2, 0, 1) * 255), HWC.
"""
import numpy as np

F = np.float32


def matrix_to_quaternion(M):
    """math.py:46-91, per matrix, fp32, wxyz."""
    M = np.asarray(M, F).reshape(-1, 3, 3)
    out = np.zeros((M.shape[0], 4), F)
    for i, m in enumerate(M):
        tr = F(F(F(F(1) + m[0, 0]) + m[1, 1]) + m[2, 2])
        if tr > 0:
            r = F(np.sqrt(tr) / F(2.0))
            x = F((m[2, 1] - m[1, 2]) / F(F(4) * r))
            y = F((m[0, 2] - m[2, 0]) / F(F(4) * r))
            z = F((m[1, 0] - m[0, 1]) / F(F(4) * r))
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            S = F(np.sqrt(F(F(F(F(1.0) + m[0, 0]) - m[1, 1]) - m[2, 2])) * F(2))
            r = F((m[2, 1] - m[1, 2]) / S)
            x = F(F(0.25) * S)
            y = F((m[0, 1] + m[1, 0]) / S)
            z = F((m[0, 2] + m[2, 0]) / S)
        elif m[1, 1] > m[2, 2]:
            S = F(np.sqrt(F(F(F(F(1.0) + m[1, 1]) - m[0, 0]) - m[2, 2])) * F(2))
            r = F((m[0, 2] - m[2, 0]) / S)
            x = F((m[0, 1] + m[1, 0]) / S)
            y = F(F(0.25) * S)
            z = F((m[1, 2] + m[2, 1]) / S)
        else:
            S = F(np.sqrt(F(F(F(F(1.0) + m[2, 2]) - m[0, 0]) - m[1, 1])) * F(2))
            r = F((m[1, 0] - m[0, 1]) / S)
            x = F((m[0, 2] + m[2, 0]) / S)
            y = F((m[1, 2] + m[2, 1]) / S)
            z = F(F(0.25) * S)
        out[i] = (r, x, y, z)
    return out


def quaternion_to_matrix(q):
    """math.py:18-44."""
    q = np.asarray(q, F)
    r, i, j, k = q
    two_s = F(F(2.0) / F(np.sum(q * q, dtype=F)))
    return np.array([
        [1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r)],
        [two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r)],
        [two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)],
    ], F)


def quaternion_slerp(q0, q1, fraction):
    """camera.py:81-109 (shortest path, spin 0)."""
    q0, q1 = np.asarray(q0, F), np.array(q1, F)
    d = F(np.sum(q0 * q1, dtype=F))
    if d < 0:
        # camera.py:94-95 flips d FIRST (`d[d < 0] = -d[d < 0]`), so the mask of the next statement
        # (`q1[d < 0] = -q1[d < 0]`) is already all-false: q1 is never negated. Reproduced as is.
        d = F(-d)
    d = F(min(max(d, F(-1)), F(1)))
    angle = F(np.arccos(d))
    isin = F(F(1.0) / F(np.sin(angle) + F(1e-10)))
    q0_ = q0 * F(np.sin(F(F(1.0) - fraction) * angle)) * isin
    q1_ = q1 * F(np.sin(F(fraction * angle))) * isin
    q = (q0_ + q1_).astype(F)
    return q0 if angle < F(1e-5) else q


def sample_from_dense_cameras(dense, t):
    """dense [N,18], t [M] -> [M,18]: arc-length (position + rotation angle) parametrised pose interpolation."""
    dense = np.asarray(dense, F)
    t = np.asarray(t, F).reshape(-1)
    N = dense.shape[0]
    poses = dense[:, :12].reshape(N, 3, 4)
    pos, rots = poses[:, :, 3], poses[:, :, :3]
    quats = matrix_to_quaternion(rots)
    d_pos = np.sqrt(np.sum((pos[1:] - pos[:-1]) ** 2, axis=-1, dtype=F)).astype(F)
    qn = (quats / np.sqrt(np.sum(quats * quats, axis=-1, keepdims=True, dtype=F))).astype(F)
    dot = np.sum(qn[:-1] * qn[1:], axis=-1, dtype=F)
    angle = (F(2) * np.arccos(np.clip(dot, F(-1), F(1)))).astype(F)
    d = np.sqrt(d_pos ** 2 + angle ** 2).astype(F)
    cum = np.zeros(N, F)
    for i in range(1, N):
        cum[i] = F(cum[i - 1] + d[i - 1])  # sequential fp32 sum, as torch.cumsum on the CPU
    total = cum[-1]
    out = np.zeros((t.shape[0], dense.shape[1]), F)
    for m, tm in enumerate(t):
        desired = F(tm * total)
        mask = (desired >= cum[:-1]) & (desired < cum[1:])
        seg = int(np.argmax(mask))  # 0 when nothing matches (t = 1)
        lower, upper = cum[seg], cum[seg + 1]
        with np.errstate(invalid="ignore", divide="ignore"):
            frac = F(F(desired - lower) / F(upper - lower))
        frac = F(min(max(frac, F(0)), F(1))) if not np.isnan(frac) else frac
        p = (F(1) - frac) * pos[seg] + frac * pos[seg + 1]
        q = quaternion_slerp(quats[seg], quats[seg + 1], frac)
        q = (q / np.sqrt(np.sum(q * q, dtype=F))).astype(F)
        R = quaternion_to_matrix(q)
        out[m, :12] = np.concatenate([R, p[:, None].astype(F)], axis=1).reshape(-1)
        out[m, 12:] = (F(1) - frac) * dense[seg, 12:] + frac * dense[seg + 1, 12:]
    return out


def pack_frames_u8(images):
    """images [F,3,H,W] in [-1,1] -> uint8 [F,H,W,3] as export_video does per frame (io.py:181):
    mul(1/2).add(1/2).clamp(0,1).mul(255) then numpy's float -> uint8 truncation."""
    x = np.asarray(images, F)
    y = np.clip(x * F(0.5) + F(0.5), F(0), F(1)) * F(255)
    return np.transpose(y, (0, 2, 3, 1)).astype(np.uint8)


def pack_splat(xyz, rgb, opacity, scales, rots, compatible=True):
    """The .splat byte stream of export_ply_for_gaussians (prometheus/utils/io.py:40-121) for activated Gaussians
    xyz [n,3], rgb [n,3], opacity [n], scales [n,3], rots [n,4]: [n,32] uint8, plus the sort keys (ascending).
    PINNED: tests/golden/splat.npz holds the reference's own output (tests/golden/make_golden_splat.py)."""
    xyz, rgb, opacity, scales, rots = (np.asarray(a, F) for a in (xyz, rgb, opacity, scales, rots))
    if compatible:  # io.py:46-48: inverse_sigmoid (math.py:7) and log(scale + 1e-8)
        op = np.log(opacity / (F(1) - opacity)).astype(F)
        ls = np.log(scales + F(1e-8)).astype(F)
    else:
        op, ls = opacity, scales
    keys = (-np.exp(ls[:, 0] + ls[:, 1] + ls[:, 2]) / (F(1) + np.exp(-op))).astype(F)
    order = np.argsort(keys, kind="stable")
    out = np.zeros((len(order), 32), np.uint8)
    out[:, 0:12] = xyz[order].view(np.uint8).reshape(-1, 12)
    out[:, 12:24] = np.exp(ls[order]).astype(F).view(np.uint8).reshape(-1, 12)
    color = np.concatenate([rgb[order], (F(1) / (F(1) + np.exp(-op[order])))[:, None]], axis=1).astype(F)
    out[:, 24:28] = (color * F(255)).clip(0, 255).astype(np.uint8)
    r = rots[order]
    rn = np.sqrt(np.sum(r * r, axis=1, keepdims=True, dtype=F)).astype(F)
    out[:, 28:32] = ((r / rn) * F(128) + F(128)).clip(0, 255).astype(np.uint8)
    return out, keys[order]
