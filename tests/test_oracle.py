"""CPU tests of the oracle itself: known-answer cases, the reference's own golden
outputs (tests/golden/, produced by tests/golden/make_golden.py from the reference
module), and agreement of the two independent restatements (C and torch)."""
import math
import os

import numpy as np
import pytest
import torch

from tests.util import allclose, random_blobs, small_scene

EYE = np.eye(4, dtype=np.float32)[None]


def K_of(f, W, H):
    return np.array([[[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]]], np.float32)


# --------------------------------------------------------------------------- golden: the reference itself
@pytest.mark.parametrize("mode", ["mvsplat", "gslrm", "director3d"])
def test_unproject_matches_reference_converter(oracle, golden_dir, mode):
    g = np.load(os.path.join(golden_dir, f"converter_{mode}.npz"))
    o = oracle.unproject(g["params"], g["cameras"], g["muls"], mode, 0.01, 100.0, 0.001, 1.0)
    for k in ("xyz", "features", "opacity", "scales", "rotations"):
        assert o[k].shape == g[k].shape
        assert allclose(o[k], g[k], atol=1e-5, rtol=1e-4), k


@pytest.mark.parametrize("mode,name", [("mvsplat", "bf16"), ("mvsplat", "f16"), ("director3d", "bf16"), ("gslrm", "f16")])
def test_unproject_on_half_precision_maps_matches_reference_converter(oracle, golden_dir, mode, name):
    """fp16 / bf16 maps: the reference widens them by `* muls` (gaussians_renderer.py:219), i.e. the restatement
    applied to the widened maps (stored exactly, as fp32, in the fixture)."""
    g = np.load(os.path.join(golden_dir, "converter_half.npz"))
    base = np.load(os.path.join(golden_dir, f"converter_{mode}.npz"))
    pre = f"{mode}_{name}_"
    o = oracle.unproject(g[pre + "params"], base["cameras"], base["muls"], mode, 0.01, 100.0, 0.001, 1.0)
    for k in ("xyz", "features", "opacity", "scales", "rotations"):
        assert allclose(o[k], g[pre + k], atol=1e-5, rtol=1e-4), k


def test_camera_prep_matches_reference(oracle, golden_dir):
    d = np.load(os.path.join(golden_dir, "cameras.npz"))
    vm, K = oracle.camera_prep(d["cameras"], int(d["h"]), int(d["w"]))
    assert np.array_equal(vm, d["viewmats"])          # same np.linalg.inv on the same floats
    assert allclose(K, d["Ks"], atol=0, rtol=1e-6)    # torch.atan vs numpy arctan: <= 1 ulp


def test_camera_prep_matches_reference_non_square(oracle, golden_dir):
    """h != w and H != W: the reference pairs fovx with h and fovy with w (gaussians_renderer.py:469-470) and the K
    build un-pairs them again (:405-412); for square images the swaps cancel, here they do not."""
    d = np.load(os.path.join(golden_dir, "cameras_nonsquare.npz"))
    h, w = int(d["h"]), int(d["w"])
    assert h != w and d["cameras"][0, 16] != d["cameras"][0, 17]
    vm, K = oracle.camera_prep(d["cameras"], h, w)
    assert np.array_equal(vm, d["viewmats"])
    assert allclose(K, d["Ks"], atol=0, rtol=1e-6)
    # what the two swaps amount to: fx' = w fx / H, fy' = h fy / W, principal point at the image centre
    cam = d["cameras"][0]
    assert allclose(K[0, 0, 0], w * cam[12] / cam[16], atol=0, rtol=1e-5) and allclose(K[0, 1, 1], h * cam[13] / cam[17], atol=0, rtol=1e-5)
    assert K[0, 0, 2] == w / 2 and K[0, 1, 2] == h / 2


# --------------------------------------------------------------------------- known answers
def test_single_isotropic_gaussian_closed_form(oracle):
    """One isotropic Gaussian on the optical axis: cov2d = (f s / z)^2 + 0.3, closed-form alpha(r)."""
    f, W, H, s, z, op = 100.0, 64, 64, 0.02, 2.0, 0.8
    means = np.array([[0, 0, z]], np.float32)
    quats = np.array([[1, 0, 0, 0]], np.float32)
    scales = np.full((1, 3), s, np.float32)
    st = oracle.forward_pipeline(means, quats, scales, np.array([op], np.float32), np.array([[1.0, 0.5, 0.25]], np.float32),
                                 EYE, K_of(f, W, H), W, H, backgrounds=None, with_depth=True)
    var = (f * s / z) ** 2 + 0.3
    assert st["radii"][0, 0] == math.ceil(3 * math.sqrt(var + 0.1))  # v1 = b + sqrt(max(0.01, 0))
    np.testing.assert_allclose(st["means2d"][0, 0], [32, 32], atol=1e-5)
    np.testing.assert_allclose(st["conics"][0, 0], [1 / var, 0, 1 / var], rtol=1e-5, atol=1e-7)
    assert st["depths"][0, 0] == np.float32(z)
    ys, xs = np.mgrid[0:H, 0:W]
    r2 = (xs + 0.5 - 32) ** 2 + (ys + 0.5 - 32) ** 2
    alpha = np.minimum(0.999, op * np.exp(-0.5 * r2 / var))
    alpha[alpha < 1 / 255] = 0
    tiles = np.zeros((H, W), bool)  # only tiles inside the radius AABB are rasterised
    r = st["radii"][0, 0]
    x0, x1 = int(math.floor((32 - r) / 16)), int(math.ceil((32 + r) / 16))
    tiles[x0 * 16:x1 * 16, x0 * 16:x1 * 16] = True
    alpha[~tiles] = 0
    np.testing.assert_allclose(st["alphas"][0], alpha, atol=2e-6)
    np.testing.assert_allclose(st["render"][0, ..., 0], alpha, atol=2e-6)
    np.testing.assert_allclose(st["render"][0, ..., 3], alpha * z, atol=5e-6)
    assert st["tiles_per_gauss"][0, 0] == (x1 - x0) ** 2 == len(st["isect_ids"])


def test_two_gaussians_transmittance_and_depth_order(oracle):
    """Two coincident splats at different depths: the nearer one is composited first whatever the input order."""
    f, W, H = 120.0, 32, 32
    for order in ((0, 1), (1, 0)):
        z = np.array([1.5, 3.0], np.float32)[list(order)]
        op = np.array([0.6, 0.9], np.float32)[list(order)]
        col = np.array([[1, 0, 0], [0, 1, 0]], np.float32)[list(order)]
        means = np.stack([np.zeros(2), np.zeros(2), z], 1).astype(np.float32)
        quats = np.tile(np.array([[1, 0, 0, 0]], np.float32), (2, 1))
        scales = np.stack([np.full(3, 0.02 * zi / 1.5, np.float32) for zi in z])
        st = oracle.forward_pipeline(means, quats, scales, op, col, EYE, K_of(f, W, H), W, H,
                                     backgrounds=np.array([[0, 0, 1]], np.float32), with_depth=False)
        var = (f * 0.02 / 1.5) ** 2 + 0.3
        d2 = 0.5 * (0.5 ** 2 + 0.5 ** 2) / var
        a_near, a_far = 0.6 * math.exp(-d2), 0.9 * math.exp(-d2)
        px = st["render"][0, 16, 16]
        exp = np.array([a_near, (1 - a_near) * a_far, (1 - a_near) * (1 - a_far)])
        np.testing.assert_allclose(px, exp, atol=2e-6)
        near_id = int(np.argmin(z))
        assert st["flatten_ids"][0] == near_id  # front-most first in every tile
        np.testing.assert_allclose(st["alphas"][0, 16, 16], 1 - (1 - a_near) * (1 - a_far), atol=2e-6)


def test_transmittance_cutoff(oracle):
    """Opaque stack: compositing stops (without adding) once T*(1-alpha) <= 1e-4; alpha clamps at 0.999."""
    n, f, W, H = 6, 100.0, 16, 16
    means = np.stack([np.zeros(n), np.zeros(n), 1.0 + 0.1 * np.arange(n)], 1).astype(np.float32)
    quats = np.tile(np.array([[1, 0, 0, 0]], np.float32), (n, 1))
    scales = np.full((n, 3), 50.0, np.float32)  # huge splats: sigma ~ 0, alpha clamps at 0.999 everywhere
    st = oracle.forward_pipeline(means, quats, scales, np.ones(n, np.float32), np.ones((n, 3), np.float32), EYE,
                                 K_of(f, W, H), W, H, backgrounds=None, with_depth=False)
    # alpha = 0.999 each: T = 1 -> 1e-3 -> (1e-6 <= 1e-4: stop before adding the 2nd)
    np.testing.assert_allclose(st["alphas"][0], 0.999, atol=1e-6)
    np.testing.assert_allclose(st["render"][0, ..., 0], 0.999, atol=1e-6)
    assert (st["last_ids"][0] == 0).all()


def test_cull_cases(oracle):
    f, W, H = 100.0, 64, 64
    means = np.array([[0, 0, 0.005],     # in front of the near plane
                      [0, 0, -1.0],      # behind the camera
                      [50.0, 0, 1.0],    # far off screen
                      [0, 0, 2.0],       # visible
                      [0, 0, 2.0]], np.float32)
    quats = np.tile(np.array([[1, 0, 0, 0]], np.float32), (5, 1))
    scales = np.full((5, 3), 0.01, np.float32)
    pr = oracle.project(means, quats, scales, EYE, K_of(f, W, H), W, H)
    assert list(pr["radii"][0] > 0) == [False, False, False, True, True]
    assert (pr["tiles_per_gauss"][0][:3] == 0).all()
    pr2 = oracle.project(means, quats, scales, EYE, K_of(f, W, H), W, H, radius_clip=100.0)
    assert (pr2["radii"] == 0).all()
    pr3 = oracle.project(means, quats, scales, EYE, K_of(f, W, H), W, H, far_plane=1.5)
    assert (pr3["radii"][0][3:] == 0).all()
    # opacity below 1/255 never contributes
    st = oracle.forward_pipeline(means[3:4], quats[3:4], scales[3:4], np.array([0.003], np.float32),
                                 np.ones((1, 3), np.float32), EYE, K_of(f, W, H), W, H, backgrounds=None)
    assert st["alphas"].max() == 0.0


def test_tile_edges(oracle):
    """mean +- radius exactly on multiples of 16: floor/ceil of the AABB."""
    f, W, H = 64.0, 64, 64
    # put the projected mean exactly at pixel (32, 32): x = y = 0
    means = np.array([[0, 0, 1.0]], np.float32)
    quats = np.array([[1, 0, 0, 0]], np.float32)
    scales = np.full((1, 3), 0.0825, np.float32)
    pr = oracle.project(means, quats, scales, EYE, K_of(f, W, H), W, H)
    r = int(pr["radii"][0, 0])
    assert r == 16  # ceil(3*sqrt((64*0.0825)^2 + 0.3 + 0.1)) = 16 -> AABB [16,48] exactly on tile borders
    assert pr["tiles_per_gauss"][0, 0] == 4  # tiles 1..2 in x and y: floor(1.0)=1, ceil(3.0)=3
    ids, flat = oracle.bin_emit(pr["means2d"], pr["radii"], pr["depths"], 16, 4, 4)
    tb, _ = oracle.tile_bits(4, 4, 1)
    tiles = sorted(int((i >> 32) & ((1 << tb) - 1)) for i in ids)
    assert tiles == [5, 6, 9, 10]
    depth_bits = np.float32(1.0).view(np.int32)
    assert all(int(i & 0xFFFFFFFF) == int(depth_bits) for i in ids)


def test_key_layout_sort_and_offsets(oracle):
    sc = random_blobs(n=1500, C=3, W=80, H=48, seed=3)
    st = oracle.forward_pipeline(**sc, backgrounds=None)
    ids, flat = st["isect_ids"], st["flatten_ids"]
    assert st["tile_n_bits"] == 4 and st["cam_n_bits"] == 2  # 15 tiles, 3 cameras
    assert (np.diff(ids) >= 0).all()
    assert st["tiles_per_gauss"].sum() == len(ids)
    # stability: equal keys keep emission (= flatten id) order
    same = np.diff(ids) == 0
    assert (np.diff(flat.astype(np.int64))[same] > 0).all()
    # offsets = lower bounds
    off = st["isect_offsets"].reshape(-1)
    tb = st["tile_n_bits"]
    lin = (ids >> (32 + tb)) * 15 + ((ids >> 32) & ((1 << tb) - 1))
    assert np.array_equal(off, np.searchsorted(lin, np.arange(3 * 15), side="left"))
    # camera of every key equals camera of its flatten id; depth bits round-trip
    N = sc["means"].shape[0]
    assert np.array_equal(ids >> (32 + tb), flat // N)
    assert np.array_equal((ids & 0xFFFFFFFF).astype(np.uint32), st["depths"].reshape(-1)[flat].view(np.uint32))


def test_empty_inputs(oracle):
    z = lambda *s: np.zeros(s, np.float32)
    st = oracle.forward_pipeline(z(0, 3), z(0, 4), z(0, 3), z(0), z(0, 3), EYE, K_of(50, 32, 32), 32, 32,
                                 backgrounds=np.ones((1, 3), np.float32))
    assert len(st["isect_ids"]) == 0 and (st["isect_offsets"] == 0).all()
    assert (st["alphas"] == 0).all() and (st["render"][..., :3] == 1).all() and (st["render"][..., 3] == 0).all()


# --------------------------------------------------------------------------- two independent restatements agree
@pytest.mark.parametrize("big", [False, True])
def test_c_oracle_equals_torch_restatement(oracle, big):
    from oracle import torch_ref

    sc = random_blobs(n=2000, C=2, W=64, H=48, seed=11, big=big)
    bg = torch.rand(2, 3, generator=torch.Generator().manual_seed(1))
    outs = []
    for fn in (oracle.rasterization, torch_ref.rasterization):
        ins = [torch.from_numpy(sc[k]).clone().requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")]
        r, a, meta = fn(*ins, torch.from_numpy(sc["viewmats"]), torch.from_numpy(sc["Ks"]), sc["width"], sc["height"],
                        backgrounds=bg, render_mode="RGB+ED", packed=False)
        meta["means2d"].retain_grad()
        gen = torch.Generator().manual_seed(2)
        wr, wa = torch.rand(r.shape, generator=gen), torch.rand(a.shape, generator=gen)
        if "fragile" in meta:
            keep = (~meta["fragile"].bool()).float()[..., None]
            outs_keep = keep
        ((r * wr * outs_keep).sum() + (a * wa * outs_keep).sum()).backward()
        outs.append((r.detach(), a.detach(), meta, [t.grad for t in ins], meta["means2d"].grad))
    (r1, a1, m1, g1, gm1), (r2, a2, m2, g2, gm2) = outs
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert torch.equal(m1[k].long(), m2[k].long()), k
    for k in ("means2d", "depths", "conics"):
        assert torch.equal(m1[k].detach(), m2[k].detach()), k
    keep = ~m1["fragile"].bool()
    assert keep.float().mean() > 0.98
    assert allclose(r1[keep], r2[keep], atol=1e-5, rtol=1e-4)
    assert allclose(a1[keep], a2[keep], atol=1e-5, rtol=1e-4)
    for ga, gb in zip(g1, g2):  # hand-written backward (C) vs torch.autograd through the restatement
        scale = max(1.0, float(gb.abs().max()))
        assert allclose(ga / scale, gb / scale, atol=2e-5, rtol=1e-3)
    assert allclose(gm1, gm2, atol=2e-4, rtol=1e-3)


def test_pixel_aligned_scene_statistics(oracle):
    """The synthetic generator reproduces the scene statistics SURVEY §8d quotes."""
    s = small_scene(oracle, views_in=4, size_in=64, views_out=1, size_out=64)
    st = oracle.forward_pipeline(s["means"], s["quats"], s["scales"], s["opacities"], s["colors"], s["viewmats"],
                                 s["Ks"], 64, 64, backgrounds=np.ones((1, 3), np.float32))
    vis = (st["radii"] > 0).mean()
    assert vis > 0.95
    ratio = len(st["isect_ids"]) / (st["radii"] > 0).sum()
    assert 2.0 < ratio < 4.5
    assert st["alphas"].mean() > 0.99
    assert st["fragile"].mean() < 0.02


# --------------------------------------------------------------------------- trajectory row
@pytest.mark.parametrize("name", ["orbit_big", "smooth", "repeat", "video120"])
def test_trajectory_oracle_matches_reference_golden(name):
    """oracle/trajectory_oracle.py against the reference's own sample_from_dense_cameras outputs
    (tests/golden/trajectory.npz, generated by tests/golden/make_golden_trajectory.py from /root/reference)."""
    import os

    from oracle import trajectory_oracle as T

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "trajectory.npz"))
    got = T.sample_from_dense_cameras(g[name + "_dense"][0], g[name + "_t"][0])
    want = g[name + "_out"][0]
    assert np.allclose(got[:, :12], want[:, :12], atol=1e-5, rtol=0)
    assert np.allclose(got[:, 12:], want[:, 12:], atol=0, rtol=2e-5)
    # t = 0 returns the first pose; t = 1 falls into segment 0 at fraction 1 (the reference's strict `<` test)
    if name == "orbit_big":
        d = g[name + "_dense"][0]
        assert np.allclose(want[0, 12:], d[0, 12:]) and np.allclose(want[-1, 12:], d[1, 12:])


def test_pack_frames_oracle_is_the_reference_chain():
    """export_video's per-frame conversion (io.py:181) restated with torch CPU ops vs the numpy oracle."""
    import torch

    from oracle import trajectory_oracle as T

    x = torch.randn(3, 3, 9, 7) * 0.8
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 0.999999])
    want = torch.stack([f.permute(1, 2, 0).mul(1 / 2).add(1 / 2).clamp(0, 1).mul(255) for f in x]).numpy().astype(np.uint8)
    assert np.array_equal(T.pack_frames_u8(x.numpy()), want)


def _splat_records(buf):
    """[n,32] uint8 -> (xyz f32 [n,3], scales f32 [n,3], rgba u8 [n,4], rot u8 [n,4])"""
    buf = np.ascontiguousarray(buf).reshape(-1, 32)
    return (buf[:, 0:12].copy().view(np.float32), buf[:, 12:24].copy().view(np.float32), buf[:, 24:28], buf[:, 28:32])


def check_splat_against_reference(got, want, n_expected):
    """Record-level parity with the reference's byte stream: same Gaussians (positions bit-exact), scales to fp32
    round-off of the exp(log(.)) round trip, 8-bit fields within one code (libm vs numpy exp / log differ by an ulp),
    and the same volume x opacity ordering up to swaps of near-equal keys."""
    gx, gs, gc, gr = _splat_records(got)
    wx, ws, wc, wr = _splat_records(want)
    assert len(gx) == len(wx) == n_expected
    gi = np.lexsort(gx.T[::-1])
    wi = np.lexsort(wx.T[::-1])
    assert np.array_equal(gx[gi], wx[wi])
    assert np.allclose(gs[gi], ws[wi], rtol=2e-6, atol=0)
    assert np.abs(gc[gi].astype(int) - wc[wi].astype(int)).max() <= 1
    assert np.abs(gr[gi].astype(int) - wr[wi].astype(int)).max() <= 1
    # ordering: rank of every Gaussian in both streams differs only where keys are (nearly) tied
    rank_g = np.empty(len(gi), int); rank_g[gi] = np.arange(len(gi))
    rank_w = np.empty(len(wi), int); rank_w[wi] = np.arange(len(wi))
    assert np.abs(np.argsort(rank_g) - np.argsort(rank_w)).max() <= 2 or np.mean(np.argsort(rank_g) == np.argsort(rank_w)) > 0.99


@pytest.mark.parametrize("tag,thr", [("all", 0.0), ("pruned", 0.3)])
def test_splat_oracle_matches_reference_golden(tag, thr):
    """oracle pack_splat vs the reference exporter's own bytes (tests/golden/splat.npz)."""
    import os

    from oracle import trajectory_oracle as T

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "splat.npz"))
    keep = g["opacity"][0, :, 0] >= thr
    got, keys = T.pack_splat(g["xyz"][0][keep], g["features"][0][keep][:, 0], g["opacity"][0][keep][:, 0], g["scales"][0][keep],
                             g["rotations"][0][keep])
    assert np.all(np.diff(keys) >= 0)
    check_splat_against_reference(got, g["splat_" + tag], int(keep.sum()))


# --------------------------------------------------------------------------- size-independent properties of the oracle
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_compositing_properties(oracle, seed):
    """Properties any correct restatement of the path has, whatever the scene: compositing is linear in the colours
    (no background), the result does not depend on the order the Gaussians are listed in (distinct depths), alpha
    and transmittance stay in range, expected depth lies between the nearest and farthest visible splat, and every
    pixel's last contributing entry lies inside its tile's range."""
    s = random_blobs(n=1500, C=2, W=72, H=40, seed=30 + seed)
    a = oracle.forward_pipeline(**s, want_fragile=False)
    # ranges
    assert a["alphas"].min() >= 0.0 and a["alphas"].max() <= 1.0
    vis_depth = a["depths"][a["radii"] > 0]
    hit = a["alphas"] > 0.5
    ed = a["render"][..., 3] / np.maximum(a["alphas"], 1e-10)
    assert ed[hit].min() >= vis_depth.min() * (1 - 1e-5) and ed[hit].max() <= vis_depth.max() * (1 + 1e-5)
    # last_ids inside the tile's range of the sorted list
    off = a["isect_offsets"].reshape(-1)
    C, H, W = a["last_ids"].shape
    tw = a["tile_width"]
    ends = np.append(off[1:], a["flatten_ids"].shape[0])
    for c in range(C):
        for (y, x) in ((0, 0), (H - 1, W - 1), (H // 2, W // 3)):
            t = (c * a["tile_height"] + y // 16) * tw + x // 16
            last = a["last_ids"][c, y, x]
            assert (off[t] <= last < ends[t]) or a["alphas"][c, y, x] == 0.0
    # linearity in the colours
    rng = np.random.default_rng(seed)
    col2 = rng.random(s["colors"].shape, dtype=np.float32)
    s2 = dict(s, colors=col2)
    s3 = dict(s, colors=(0.25 * s["colors"] + 0.5 * col2).astype(np.float32))
    b, c3 = oracle.forward_pipeline(**s2, want_fragile=False), oracle.forward_pipeline(**s3, want_fragile=False)
    assert np.array_equal(a["alphas"], b["alphas"]) and np.array_equal(a["last_ids"], b["last_ids"])
    assert allclose(c3["render"][..., :3], 0.25 * a["render"][..., :3] + 0.5 * b["render"][..., :3], atol=2e-6, rtol=1e-5)
    # permutation invariance (depths are distinct with probability 1; ties would be ordered by index)
    d0 = a["depths"][0][a["radii"][0] > 0]
    assert np.unique(d0).size == d0.size
    perm = rng.permutation(s["means"].shape[0])
    sp = dict(s, **{k: s[k][perm] for k in ("means", "quats", "scales", "opacities", "colors")})
    p = oracle.forward_pipeline(**sp, want_fragile=False)
    assert np.array_equal(p["render"], a["render"]) and np.array_equal(p["alphas"], a["alphas"])
    assert np.array_equal(p["radii"], a["radii"][:, perm])


def test_oracle_background_and_opacity_limits(oracle):
    """Zero opacity renders pure background with alpha 0; the background enters as T * bg exactly."""
    s = random_blobs(n=400, C=1, W=48, H=32, seed=77)
    bg = np.array([[0.2, 0.5, 0.9]], np.float32)
    z = oracle.forward_pipeline(**dict(s, opacities=np.zeros_like(s["opacities"])), backgrounds=bg, want_fragile=False)
    assert np.all(z["alphas"] == 0.0) and np.all(z["render"][..., :3] == bg[0]) and np.all(z["render"][..., 3] == 0.0)
    a = oracle.forward_pipeline(**s, want_fragile=False)
    b = oracle.forward_pipeline(**s, backgrounds=bg, want_fragile=False)
    T = 1.0 - a["alphas"]
    assert allclose(b["render"][..., :3], a["render"][..., :3] + T[..., None] * bg[0], atol=1e-6, rtol=1e-6)
    assert np.array_equal(b["render"][..., 3], a["render"][..., 3])  # the depth channel has no background
