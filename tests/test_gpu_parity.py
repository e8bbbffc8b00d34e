"""GPU parity tests (run with -m gpu on the B200 box): every stage of the CUDA path,
called through the C-ABI (prometheus_b200.ops -> libprometheus_gs.so), against the CPU
oracle on the same seeded inputs, against the committed golden fixtures, and — at
BASELINE.json's full size — through size-independent properties.

Bars (BASELINE.json north_star): bit-exact radii / tile counts / sort keys / order /
per-tile ranges; images, depth, alpha and gradients within 1e-5 abs / 1e-4 rel in fp32
(gradients: element-wise, see `grad_close`), asserted on pixels whose decisions are not within
a few ulp of a threshold (oracle's fragile mask).
"""
import os

import numpy as np
import pytest
import torch

from tests.util import allclose, max_violation, random_blobs, small_scene, violations

pytestmark = pytest.mark.gpu

DEV = "cuda"


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def grad_close(a, b, atol=1e-5, rtol=1e-4, max_share=0.0, what=""):
    """Gradient comparison at the north_star's bar, ELEMENT-WISE: |a - b| <= atol + rtol |b| for every entry
    (the fp64-accumulating oracle is b). `max_share` = share of entries allowed to violate it (0 unless the
    test says why); the measured share and the worst ratio are printed so that `-s` shows the margin."""
    share, worst = violations(a, b, atol, rtol)
    print("grad_close %s: violating share %.3g (allowed %.3g), worst |a-b| / (atol + rtol |b|) = %.3g" % (what, share, max_share, worst))
    return share <= max_share


# Allowed share of gradient entries outside |a - b| <= 1e-5 + 1e-4 |b| (measured shares are printed by grad_close):
PROJECT_BWD_SHARE = 1e-3    # fp32 VJP chains vs the oracle's fp64 ones on strongly anisotropic splats
OPERATOR_GRAD_SHARE = 2e-3  # two correct fp32 forwards differ in the last bit of some alphas (see the test); measured: <= 3 of 2331 entries (ragged), <= 1.3e-4 at C2 size
ORDER_SHARE = 1e-3          # same kernels, different fp32 atomic summation order


def scenes_for_parity(oracle):
    return {
        "pixel_aligned": small_scene(oracle, views_in=3, size_in=48, views_out=2, size_out=64),
        "blobs": random_blobs(n=4000, C=3, W=80, H=56, seed=7),
        "big_blobs": random_blobs(n=1500, C=2, W=96, H=64, seed=9, big=True),
        "ragged": random_blobs(n=777, C=1, W=50, H=37, seed=4),  # image not a multiple of the tile
    }


@pytest.fixture(scope="module")
def cases(oracle):
    out = {}
    for name, s in scenes_for_parity(oracle).items():
        C = s["viewmats"].shape[0]
        bg = np.random.default_rng(3).random((C, 3)).astype(np.float32)
        st = oracle.forward_pipeline(s["means"], s["quats"], s["scales"], s["opacities"], s["colors"], s["viewmats"],
                                     s["Ks"], s["width"], s["height"], backgrounds=bg, with_depth=True)
        out[name] = (s, bg, st)
    return out


CASE_NAMES = ["pixel_aligned", "blobs", "big_blobs", "ragged"]


# --------------------------------------------------------------------------- stage 1
@pytest.mark.parametrize("mode", ["mvsplat", "gslrm", "director3d"])
def test_unproject_fwd_bwd_vs_reference_golden(golden_dir, mode):
    from prometheus_b200 import GaussianConverter

    g = np.load(os.path.join(golden_dir, f"converter_{mode}.npz"))
    conv = GaussianConverter(gs_convert_mode=mode, z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    assert allclose(conv.muls.cpu().numpy(), g["muls"], atol=0, rtol=1e-7)
    params = dev(g["params"]).requires_grad_(True)
    outs = conv(params, dev(g["cameras"]))
    for o, k in zip(outs, ("xyz", "features", "opacity", "scales", "rotations")):
        assert o.shape == g[k].shape, k
        assert allclose(o.detach().cpu().numpy(), g[k], atol=1e-5, rtol=1e-4), (k, max_violation(o.detach().cpu().numpy(), g[k]))
    loss = sum((o * dev(g["w_" + k])).sum() for o, k in zip(outs, ("xyz", "features", "opacity", "scales", "rotations")))
    loss.backward()
    assert grad_close(params.grad.cpu().numpy(), g["v_params"], atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("mode,name", [("mvsplat", "bf16"), ("mvsplat", "f16"), ("director3d", "bf16"), ("gslrm", "f16")])
def test_unproject_half_precision_maps_vs_reference_golden(golden_dir, mode, name):
    """Decoder maps that arrive in fp16 / bf16 (autocast decoder) are read natively: same Gaussians as the
    reference converter on those maps (it widens by `* muls`, gaussians_renderer.py:219), the gradient comes
    back in the maps' dtype, and both equal the fp32 entry points on the widened maps bit for bit."""
    from prometheus_b200 import GaussianConverter, ops

    dt = {"bf16": torch.bfloat16, "f16": torch.float16}[name]
    g = np.load(os.path.join(golden_dir, "converter_half.npz"))
    base = np.load(os.path.join(golden_dir, f"converter_{mode}.npz"))
    pre = f"{mode}_{name}_"
    keys = ("xyz", "features", "opacity", "scales", "rotations")
    conv = GaussianConverter(gs_convert_mode=mode, z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    params = dev(g[pre + "params"]).to(dt)
    assert torch.equal(params.float().cpu(), torch.from_numpy(g[pre + "params"]))  # the fixture is exact in dt
    params.requires_grad_(True)
    cams = dev(base["cameras"])
    outs = conv(params, cams)
    for o, k in zip(outs, keys):
        assert o.dtype == torch.float32 and o.shape == g[pre + k].shape, k
        assert allclose(o.detach().cpu().numpy(), g[pre + k], atol=1e-5, rtol=1e-4), k
    loss = sum((o * dev(base["w_" + k])).sum() for o, k in zip(outs, keys))
    loss.backward()
    assert params.grad.dtype == dt
    # one ulp of the half type on top of the fp32 tolerance: a last-bit fp32 difference may flip the rounding
    ulp = 2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10
    got, want = params.grad.float().cpu().numpy(), g[pre + "v_params"]
    assert np.all(np.abs(got - want) <= 1e-5 + (ulp + 1e-4) * np.abs(want)), np.abs(got - want).max()

    # bit-exact against the fp32 kernels on the widened maps
    wide = params.detach().float()
    B, N = wide.shape[:2]
    p32, p16 = wide.reshape(B * N, 14, *wide.shape[3:]), params.detach().reshape(B * N, 14, *wide.shape[3:])
    args = (cams.reshape(B * N, 18), conv.muls.reshape(-1), mode, 0.01, 100.0, 0.001, 1.0)
    o32, o16 = ops.unproject_fwd(p32, *args), ops.unproject_fwd(p16, *args)
    for a, b in zip(o32, o16):
        assert torch.equal(a, b)
    vs = [dev(base["w_" + k]).reshape(o.shape).contiguous() for o, k in zip(o32, keys)]
    v32, v16 = ops.unproject_bwd(p32, *args, *vs), ops.unproject_bwd(p16, *args, *vs)
    assert v16.dtype == dt and torch.equal(v32.to(dt), v16)


def test_unproject_vs_oracle_larger(oracle):
    from prometheus_b200 import ops, scenes

    sc = scenes.make_scene("C1", seed=3, views_in=4, size_in=96)
    o = oracle.unproject(sc.params.numpy(), sc.cameras_in.numpy(), mode="mvsplat", z_near=0.01, z_far=100.0)
    muls = torch.ones(14, device=DEV)
    xyz, rgb, op, scl, q = ops.unproject_fwd(sc.params[0].to(DEV), sc.cameras_in[0].to(DEV), muls, "mvsplat", 0.01, 100.0, 0.001, 1.0)
    for a, b in ((xyz, o["xyz"][0]), (rgb, o["features"][0, :, 0]), (op, o["opacity"][0, :, 0]), (scl, o["scales"][0]), (q, o["rotations"][0])):
        assert allclose(a.cpu().numpy(), b, atol=1e-5, rtol=1e-4)


def test_camera_prep_vs_reference_golden(golden_dir):
    from prometheus_b200 import ops

    d = np.load(os.path.join(golden_dir, "cameras.npz"))
    vm, K = ops.camera_prep(dev(d["cameras"]), int(d["h"]), int(d["w"]))
    assert allclose(vm.cpu().numpy(), d["viewmats"], atol=2e-6, rtol=1e-5)
    assert allclose(K.cpu().numpy(), d["Ks"], atol=0, rtol=1e-5)


def test_camera_prep_non_square_vs_reference_golden(golden_dir):
    """h != w, H != W, fx != fy (tests/golden/make_golden_cameras_nonsquare.py ran the reference's own host path)."""
    from prometheus_b200 import ops

    d = np.load(os.path.join(golden_dir, "cameras_nonsquare.npz"))
    vm, K = ops.camera_prep(dev(d["cameras"]), int(d["h"]), int(d["w"]))
    assert allclose(vm.cpu().numpy(), d["viewmats"], atol=2e-6, rtol=1e-5)
    assert allclose(K.cpu().numpy(), d["Ks"], atol=0, rtol=1e-5)


def test_level1_dropin_loop_matches_the_batched_module(golden_dir):
    """The reference's own control flow (per-view .cpu(), numpy inverse, host K, C = 1 operator calls, final
    synchronize — tools/dropin_loop.py restates gaussians_renderer.py:475-512) over prometheus_b200.rasterization,
    against the batched module fast path: non-square render size and camera record, forward and gradients."""
    import prometheus_b200 as pb
    from prometheus_b200 import GaussianConverter, GaussianRenderer, scenes
    from tools.dropin_loop import reference_style_forward

    sc = scenes.make_scene("C1", seed=29, views_in=3, size_in=48, views_out=3, size_out=64, scenes=2)
    nsq = np.load(os.path.join(golden_dir, "cameras_nonsquare.npz"))
    cams_out = sc.cameras_out.clone()
    cams_out[..., 12:] = torch.from_numpy(nsq["cameras"][:3, 12:]) * torch.tensor([0.6, 0.6, 1, 1, 1, 1])  # H != W, fx != fy
    h, w = 56, 88
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    rend = GaussianRenderer(background=[0.2, 0.5, 0.9]).to(DEV)
    res = []
    for loop in (True, False):
        params = sc.params.to(DEV).requires_grad_(True)
        g = conv(params, sc.cameras_in.to(DEV))
        if loop:
            im, de, ma, radii, vsp = reference_style_forward(pb.rasterization, cams_out.to(DEV), g, rend.bg_color, h, w)
        else:
            im, de, ma, _, _ = rend(cams_out.to(DEV), g, bg_color=None, h=h, w=w)
            radii, vsp = rend.radii, rend.viewspace_points
        (scenes.render_loss(im, de, ma, torch.zeros_like(im)) * im.numel()).backward()
        res.append((im.detach(), de.detach(), ma.detach(), torch.stack(radii), params.grad.clone(), torch.stack([v.grad for v in vsp])))
    a, b = res
    assert a[0].shape == (2, 3, 3, h, w) and float(a[2].mean()) > 0.3
    # the host path inverts the pose with LAPACK, the device path analytically: the cameras differ by an ulp, so a
    # handful of borderline Gaussians / threshold pixels may differ — everything else agrees to fp32 round-off
    assert (a[3] != b[3]).float().mean() < 1e-3
    for x, y, tol in ((a[0], b[0], 2e-4), (a[1], b[1], 5e-4), (a[2], b[2], 2e-4)):
        assert float(torch.quantile((x - y).abs().flatten()[:2_000_000], 0.999)) < tol
    assert float((a[4] - b[4]).abs().sum() / b[4].abs().sum()) < 2e-3
    assert float((a[5] - b[5]).abs().sum() / b[5].abs().sum()) < 2e-3


def test_modules_cast_to_bf16_like_eval_py(golden_dir):
    """eval.py:91 casts the whole decoder system — and with it `converter.muls` and `renderer.bg_color` — to bf16 and
    feeds bf16 decoder maps: the bf16-cast modules must give what the fp32 modules give on the same (bf16) maps."""
    from prometheus_b200 import GaussianConverter, GaussianRenderer, scenes

    sc = scenes.make_scene("C1", seed=33, views_in=2, size_in=40, views_out=2, size_out=48)
    maps = sc.params.to(torch.bfloat16).to(DEV)
    outs = []
    for cast in (True, False):
        conv = GaussianConverter("director3d", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
        rend = GaussianRenderer(background=[1, 1, 1]).to(DEV)
        if cast:
            conv, rend = conv.to(torch.bfloat16), rend.to(torch.bfloat16)
            assert conv.muls.dtype == torch.bfloat16 and rend.bg_color.dtype == torch.bfloat16
        else:
            conv.muls.copy_(conv.muls.to(torch.bfloat16).float())  # the same multipliers, held in fp32
        with torch.no_grad():
            g = conv(maps, sc.cameras_in.to(DEV))
            outs.append(g + rend(sc.cameras_out.to(DEV), g, h=48, w=48)[:3])
    for x, y in zip(*outs):
        assert x.dtype == torch.float32 and torch.equal(x, y)


# --------------------------------------------------------------------------- stage 2
@pytest.mark.parametrize("name", CASE_NAMES)
def test_project_bit_exact(cases, name):
    from prometheus_b200 import ops

    s, bg, st = cases[name]
    radii, means2d, depths, conics, tpg = ops.project_fwd(dev(s["means"]), dev(s["quats"]), dev(s["scales"]),
                                                          dev(s["viewmats"]), dev(s["Ks"]), s["width"], s["height"])
    assert np.array_equal(radii.cpu().numpy(), st["radii"])
    assert np.array_equal(tpg.cpu().numpy(), st["tiles_per_gauss"])
    # the floats that feed integer outputs are reproduced bit for bit, not merely closely
    assert np.array_equal(means2d.cpu().numpy().view(np.uint32), st["means2d"].view(np.uint32))
    assert np.array_equal(depths.cpu().numpy().view(np.uint32), st["depths"].view(np.uint32))
    assert np.array_equal(conics.cpu().numpy().view(np.uint32), st["conics"].view(np.uint32))


def test_project_cull_parameters(oracle):
    from prometheus_b200 import ops

    s = random_blobs(n=2000, C=2, W=64, H=64, seed=21)
    for kw in (dict(near_plane=2.9), dict(far_plane=3.1), dict(radius_clip=4.0), dict(eps2d=0.1)):
        ref = oracle.project(s["means"], s["quats"], s["scales"], s["viewmats"], s["Ks"], 64, 64, **kw)
        radii, means2d, depths, conics, tpg = ops.project_fwd(dev(s["means"]), dev(s["quats"]), dev(s["scales"]),
                                                              dev(s["viewmats"]), dev(s["Ks"]), 64, 64, **kw)
        assert np.array_equal(radii.cpu().numpy(), ref["radii"]), kw
        assert np.array_equal(tpg.cpu().numpy(), ref["tiles_per_gauss"]), kw
        assert np.array_equal(conics.cpu().numpy().view(np.uint32), ref["conics"].view(np.uint32)), kw


# --------------------------------------------------------------------------- stage 3
@pytest.mark.parametrize("name", CASE_NAMES)
def test_binning_sort_offsets_bit_exact(cases, name):
    from prometheus_b200 import ops

    s, bg, st = cases[name]
    C = s["viewmats"].shape[0]
    cum, n_dev = ops.scan_tiles(dev(st["tiles_per_gauss"]))
    assert int(n_dev.item()) == len(st["isect_ids"])
    assert np.array_equal(cum.cpu().numpy(), np.cumsum(st["tiles_per_gauss"].reshape(-1)).astype(np.int32))
    ids, flat = ops.bin_emit(dev(st["means2d"]), dev(st["radii"]), dev(st["depths"]), cum, s["width"], s["height"], int(n_dev.item()))
    assert np.array_equal(ids.cpu().numpy(), st["isect_ids_unsorted"])
    assert np.array_equal(flat.cpu().numpy(), st["flatten_ids_unsorted"])
    tb, cb = ops.key_bits(st["tile_width"], st["tile_height"], C)
    assert (tb, cb) == (st["tile_n_bits"], st["cam_n_bits"])
    ids_s, flat_s, ws = ops.sort_pairs(ids, flat, 32 + tb + cb, return_workspace=True)
    assert ops.sort_error_flag(ws) == 0
    assert np.array_equal(ids_s.cpu().numpy(), st["isect_ids"])
    assert np.array_equal(flat_s.cpu().numpy(), st["flatten_ids"])
    off = ops.tile_offsets(ids_s, C, s["width"], s["height"])
    assert np.array_equal(off.cpu().numpy(), st["isect_offsets"])


@pytest.mark.parametrize("n,end_bit", [(1, 44), (31, 40), (4096, 47), (4097, 47), (100_003, 47), (1_000_000, 44), (3_000_001, 64), (50_000, 9)])
def test_radix_sort_against_stable_reference(n, end_bit):
    """Stand-alone check of the onesweep sort: random keys with many duplicates, ragged sizes."""
    from prometheus_b200 import ops

    g = torch.Generator().manual_seed(n)
    hi = torch.randint(0, 1 << 12, (n,), generator=g, dtype=torch.int64)          # few distinct tiles -> duplicates
    lo = torch.randint(0, 1 << 20, (n,), generator=g, dtype=torch.int64) << 11    # depth-like bits
    keys = ((hi << 32) | lo)
    if end_bit < 64:
        keys = keys & ((1 << end_bit) - 1)
    vals = torch.arange(n, dtype=torch.int32)
    ref_keys, order = torch.sort(keys, stable=True)
    k, v, ws = ops.sort_pairs(keys.to(DEV), vals.to(DEV), end_bit, return_workspace=True)
    assert ops.sort_error_flag(ws) == 0
    assert torch.equal(k.cpu(), ref_keys)
    assert torch.equal(v.cpu(), vals[order])


def test_sort_with_device_side_count():
    """No-host-sync mode: the element count is read from device memory; the tail is ignored."""
    from prometheus_b200 import ops

    n, cap = 70_001, 90_000
    g = torch.Generator().manual_seed(5)
    keys = torch.randint(0, 1 << 44, (cap,), generator=g, dtype=torch.int64)
    vals = torch.arange(cap, dtype=torch.int32)
    ref_keys, order = torch.sort(keys[:n], stable=True)
    d_n = torch.tensor([n], dtype=torch.int32, device=DEV)
    k, v = ops.sort_pairs(keys.to(DEV), vals.to(DEV), 44, d_n=d_n)
    assert torch.equal(k[:n].cpu(), ref_keys) and torch.equal(v[:n].cpu(), vals[:n][order])


def _binned(ops, st, s, C):
    """Run the binned path (histogram -> scan -> scatter -> per-tile sort) on the oracle's projection."""
    _, _, _, _, tpg = None, None, None, None, None
    radii, means2d, depths, conics, tpg = ops.project_fwd(dev(s["means"]), dev(s["quats"]), dev(s["scales"]), dev(s["viewmats"]),
                                                          dev(s["Ks"]), s["width"], s["height"])
    tile_counts = ops.bin_count(means2d, radii, s["width"], s["height"])
    offsets, stats = ops.tile_scan(tile_counts)
    n, max_count = stats.tolist()
    records = ops.bin_scatter(means2d, radii, depths, offsets, s["width"], s["height"], n)
    ids, flat, flag = ops.tile_sort(records, offsets, C, s["width"], s["height"], max_tile_count=max_count)
    return tile_counts, offsets, n, max_count, ids, flat, flag


@pytest.mark.parametrize("name", CASE_NAMES)
def test_binned_sort_bit_exact(cases, name):
    """The B200-first sort (tile histogram + scatter + shared-memory sort) gives the same sorted keys, order
    and per-tile ranges as the stable 64-bit radix sort, bit for bit."""
    from prometheus_b200 import ops

    s, bg, st = cases[name]
    C = s["viewmats"].shape[0]
    tile_counts, offsets, n, max_count, ids, flat, flag = _binned(ops, st, s, C)
    ref_off = st["isect_offsets"].reshape(-1)
    ref_cnt = np.diff(np.append(ref_off, len(st["isect_ids"])))
    assert n == len(st["isect_ids"]) and max_count == ref_cnt.max()
    assert np.array_equal(tile_counts.cpu().numpy().reshape(-1), ref_cnt)
    assert np.array_equal(offsets.cpu().numpy(), st["isect_offsets"])
    assert int(flag.item()) == 0
    assert np.array_equal(ids.cpu().numpy(), st["isect_ids"])
    assert np.array_equal(flat.cpu().numpy(), st["flatten_ids"])


def test_binned_sort_depth_ties_and_overflow_fallback(oracle):
    """(a) exact depth ties (duplicated Gaussians) are ordered by flatten id = stable-sort order;
    (b) a tile holding more intersections than the shared-memory capacity raises the overflow flag and the
    operator falls back to the generic onesweep path — same result either way."""
    import prometheus_b200 as pb
    from prometheus_b200 import ops

    base = random_blobs(n=600, C=2, W=64, H=48, seed=31)
    dup = {k: (np.concatenate([v, v, v]) if k in ("means", "quats", "scales", "opacities", "colors") else v) for k, v in base.items()}
    st = oracle.forward_pipeline(dup["means"], dup["quats"], dup["scales"], dup["opacities"], dup["colors"], dup["viewmats"],
                                 dup["Ks"], 64, 48, backgrounds=None)
    assert (np.diff(st["isect_ids"]) == 0).sum() > 100  # the scene really has ties
    _, offsets, n, _, ids, flat, flag = _binned(ops, st, dup, 2)
    assert int(flag.item()) == 0
    assert np.array_equal(ids.cpu().numpy(), st["isect_ids"]) and np.array_equal(flat.cpu().numpy(), st["flatten_ids"])

    # (b) one tile far above the one-pass shared-memory capacity (8192 entries)
    def stack_on_one_pixel(n_big, same_depth):
        g = torch.Generator().manual_seed(2)
        means = np.zeros((n_big, 3), np.float32)
        means[:, 2] = 2.5 if same_depth else 2.0 + torch.rand(n_big, generator=g).numpy()
        return dict(means=means, quats=np.tile(np.array([[1, 0, 0, 0]], np.float32), (n_big, 1)),
                    scales=np.full((n_big, 3), 0.01, np.float32), opacities=np.full(n_big, 0.01, np.float32),
                    colors=torch.rand(n_big, 3, generator=g).numpy(), viewmats=np.eye(4, dtype=np.float32)[None],
                    Ks=np.array([[[100.0, 0, 24.0], [0, 100.0, 24.0], [0, 0, 1]]], np.float32), width=64, height=64)

    # (b1) 17000 splats with distinct depths: split by depth through the scratch buffer, then sorted group by group
    big = stack_on_one_pixel(17000, same_depth=False)
    st = oracle.forward_pipeline(big["means"], big["quats"], big["scales"], big["opacities"], big["colors"], big["viewmats"],
                                 big["Ks"], 64, 64, backgrounds=None)
    _, _, _, max_count, ids, flat, flag = _binned(ops, st, big, 1)
    assert max_count == 17000 > ops.tile_sort_capacity() and int(flag.item()) == 0
    assert np.array_equal(ids.cpu().numpy(), st["isect_ids"]) and np.array_equal(flat.cpu().numpy(), st["flatten_ids"])
    # (b2) 9000 splats with IDENTICAL depth bits: one depth sub-bucket larger than the capacity -> overflow flag,
    # and the operator falls back to the generic onesweep path
    big = stack_on_one_pixel(9000, same_depth=True)
    st = oracle.forward_pipeline(big["means"], big["quats"], big["scales"], big["opacities"], big["colors"], big["viewmats"],
                                 big["Ks"], 64, 64, backgrounds=None)
    _, _, _, max_count, _, _, flag = _binned(ops, st, big, 1)
    assert max_count == 9000 and int(flag.item()) == 1
    r, a, m = pb.rasterization(*[dev(big[k]) for k in ("means", "quats", "scales", "opacities", "colors")], dev(big["viewmats"]),
                               dev(big["Ks"]), 64, 64, packed=False)
    assert np.array_equal(m["isect_ids"].cpu().numpy(), st["isect_ids"]) and np.array_equal(m["flatten_ids"].cpu().numpy(), st["flatten_ids"])
    assert np.array_equal(m["isect_offsets"].cpu().numpy(), st["isect_offsets"])


def test_binned_sort_hashed_privatisation_on_a_large_image(oracle):
    """Scattered splats on a 1024x800 image: a CTA's tiles no longer fit the dense shared-memory window
    (64x50 tiles > 2048 cells), so bin_count / bin_scatter take the hash-table path."""
    from prometheus_b200 import ops

    g = torch.Generator().manual_seed(12)
    n = 6000
    means = torch.rand(n, 3, generator=g) * torch.tensor([5.0, 4.0, 2.0]) + torch.tensor([-2.5, -2.0, 2.0])
    s = dict(means=means.numpy(), quats=torch.randn(n, 4, generator=g).numpy(),
             scales=(torch.rand(n, 3, generator=g) * 0.05 + 0.005).numpy(), opacities=torch.rand(n, generator=g).numpy(),
             colors=torch.rand(n, 3, generator=g).numpy(), viewmats=np.eye(4, dtype=np.float32)[None],
             Ks=np.array([[[400.0, 0, 512.0], [0, 400.0, 400.0], [0, 0, 1]]], np.float32), width=1024, height=800)
    st = oracle.forward_pipeline(s["means"], s["quats"], s["scales"], s["opacities"], s["colors"], s["viewmats"], s["Ks"],
                                 1024, 800, backgrounds=None, want_fragile=False)
    assert len(st["isect_ids"]) > 10000
    tile_counts, offsets, n_is, max_count, ids, flat, flag = _binned(ops, st, s, 1)
    assert n_is == len(st["isect_ids"]) and int(flag.item()) == 0
    assert np.array_equal(offsets.cpu().numpy(), st["isect_offsets"])
    assert np.array_equal(ids.cpu().numpy(), st["isect_ids"]) and np.array_equal(flat.cpu().numpy(), st["flatten_ids"])


def test_generic_and_binned_paths_agree_through_the_operator(cases):
    import importlib

    import prometheus_b200 as pb

    rz = importlib.import_module("prometheus_b200.rasterization")  # the module (the package attribute is the function)
    assert hasattr(rz, "isect_tiles")

    s, bg, st = cases["blobs"]
    args = [dev(s[k]) for k in ("means", "quats", "scales", "opacities", "colors")] + [dev(s["viewmats"]), dev(s["Ks"]), s["width"], s["height"]]
    r1, a1, m1 = pb.rasterization(*args, backgrounds=dev(bg), render_mode="RGB+ED", packed=False)
    rz.FORCE_GENERIC_SORT = True
    try:
        r2, a2, m2 = pb.rasterization(*args, backgrounds=dev(bg), render_mode="RGB+ED", packed=False)
    finally:
        rz.FORCE_GENERIC_SORT = False
    for k in ("isect_ids", "flatten_ids", "isect_offsets"):
        assert torch.equal(m1[k], m2[k]), k
    assert torch.equal(r1, r2) and torch.equal(a1, a2)


@pytest.mark.parametrize("mode", ["RGB+ED", "RGB", "D"])
def test_operator_lazy_meta_equals_the_eager_one(cases, mode):
    """`rasterization(...)` renders with the fused kernel and computes gsplat's sorted meta arrays when they are read:
    no sort kernel runs before that, the arrays equal the separate-stage path's, the images are the same bits and so
    are the gradients' inputs (same backward kernel on the same saved state)."""
    import importlib

    import prometheus_b200 as pb
    from prometheus_b200 import ops

    rz = importlib.import_module("prometheus_b200.rasterization")
    s, bg, st = cases["blobs"]
    args = [dev(s[k]) for k in ("means", "quats", "scales", "opacities", "colors")] + [dev(s["viewmats"]), dev(s["Ks"]), s["width"], s["height"]]
    kw = dict(backgrounds=None if mode == "D" else dev(bg), render_mode=mode, packed=False)
    prof, ops.PROFILE = ops.PROFILE, []
    try:
        r1, a1, m1 = pb.rasterization(*args, **kw)
        names = [n for n, _, _ in ops.PROFILE]
    finally:
        ops.PROFILE = prof
    assert isinstance(m1, rz._LazyMeta) and not m1.materialized
    assert "tile_render_fwd" in names and "tile_sort" not in names and "composite_fwd" not in names
    assert "isect_ids" in m1 and set(m1.keys()) >= {"radii", "means2d", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"}
    r2, a2, m2 = rz._rasterize_core(*args, **kw)           # separate stages, eager meta
    assert not isinstance(m2, rz._LazyMeta)
    assert torch.equal(r1, r2) and torch.equal(a1, a2)
    for k in ("radii", "means2d", "depths", "conics", "isect_offsets"):
        assert torch.equal(m1[k], m2[k]), k
    assert not m1.materialized
    for k in ("tiles_per_gauss", "isect_ids", "flatten_ids"):
        assert torch.equal(m1[k], m2[k]), k
    assert m1.materialized and dict(m1).keys() == m1.keys()
    old = rz.OPERATOR_FUSED
    try:  # the knob: every call through the full sort
        rz.OPERATOR_FUSED = False
        r3, a3, m3 = pb.rasterization(*args, **kw)
        assert not isinstance(m3, rz._LazyMeta) and torch.equal(r3, r1) and torch.equal(m3["flatten_ids"], m1["flatten_ids"])
    finally:
        rz.OPERATOR_FUSED = old


# --------------------------------------------------------------------------- stage 4
@pytest.mark.parametrize("name", CASE_NAMES)
def test_composite_fwd(cases, name):
    from prometheus_b200 import ops

    s, bg, st = cases[name]
    render, alphas, last_ids, trans = ops.composite_fwd(dev(st["means2d"]), dev(st["conics"]), dev(s["opacities"]), dev(s["colors"]),
                                                        dev(st["depths"]), dev(bg), dev(st["isect_offsets"]), dev(st["flatten_ids"]),
                                                        s["width"], s["height"])
    keep = st["fragile"] == 0
    assert keep.mean() > 0.97, keep.mean()
    r, a, l = render.cpu().numpy(), alphas.cpu().numpy(), last_ids.cpu().numpy()
    assert allclose(r[keep], st["render"][keep], atol=1e-5, rtol=1e-4), max_violation(r[keep], st["render"][keep])
    assert allclose(a[keep], st["alphas"][keep], atol=1e-5, rtol=1e-4)
    assert np.array_equal(l[keep], st["last_ids"][keep])
    # the exact final transmittance (what the backward consumes instead of 1 - alpha): relative bar only, T ~ 1e-4..1
    t = trans.cpu().numpy()
    assert np.all(np.abs(t[keep] - st["transmittance"][keep]) <= 1e-9 + 1e-4 * st["transmittance"][keep])
    assert torch.equal(1.0 - trans, alphas)
    # fragile pixels may flip one decision, never more than one Gaussian's worth of weight
    assert np.abs(a - st["alphas"]).max() < 5e-3


@pytest.mark.parametrize("D,with_depth,with_bg", [(3, False, True), (1, True, False), (0, True, False), (2, False, False), (3, True, False)])
def test_composite_channel_layouts(oracle, D, with_depth, with_bg):
    from prometheus_b200 import ops

    s = random_blobs(n=2500, C=2, W=64, H=48, seed=13)
    colors = np.random.default_rng(1).random((2, 2500, D)).astype(np.float32)  # per-camera colours [C,N,D]
    bg = np.random.default_rng(2).random((2, D)).astype(np.float32) if with_bg else None
    st = oracle.forward_pipeline(s["means"], s["quats"], s["scales"], s["opacities"], colors, s["viewmats"], s["Ks"],
                                 64, 48, backgrounds=bg, with_depth=with_depth)
    render, alphas, _, _ = ops.composite_fwd(dev(st["means2d"]), dev(st["conics"]), dev(s["opacities"]),
                                          dev(colors) if D else None, dev(st["depths"]) if with_depth else None,
                                          dev(bg) if with_bg else None, dev(st["isect_offsets"]), dev(st["flatten_ids"]), 64, 48)
    keep = st["fragile"] == 0
    assert render.shape[-1] == D + int(with_depth)
    assert allclose(render.cpu().numpy()[keep], st["render"][keep], atol=1e-5, rtol=1e-4)
    assert allclose(alphas.cpu().numpy()[keep], st["alphas"][keep], atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_composite_bwd(oracle, cases, name):
    from prometheus_b200 import ops

    s, bg, st = cases[name]
    rng = np.random.default_rng(17)
    keep = (st["fragile"] == 0).astype(np.float32)
    v_render = (rng.standard_normal(st["render"].shape).astype(np.float32)) * keep[..., None]
    v_alphas = (rng.standard_normal(st["alphas"].shape).astype(np.float32)) * keep
    # feed the oracle's forward state so this test isolates the backward kernel; both forms of T_final:
    # the exact transmittance the forward kept (the product path) and upstream's 1 - alpha
    for trans in (st["transmittance"], None):
        ref = oracle.composite_bwd(st["means2d"], st["conics"], s["opacities"], s["colors"], st["depths"], bg,
                                   st["isect_offsets"], st["flatten_ids"], s["width"], s["height"], st["alphas"],
                                   st["last_ids"], v_render, v_alphas, transmittance=trans)
        out = ops.composite_bwd(dev(st["means2d"]), dev(st["conics"]), dev(s["opacities"]), dev(s["colors"]), dev(st["depths"]),
                                dev(bg), dev(st["isect_offsets"]), dev(st["flatten_ids"]), s["width"], s["height"],
                                dev(st["alphas"]), dev(st["last_ids"]), dev(v_render), dev(v_alphas),
                                transmittance=None if trans is None else dev(trans))
        for got, key in zip(out, ("v_means2d", "v_conics", "v_colors", "v_opacities", "v_depths")):
            # fp32 reductions in arbitrary order vs fp64 accumulation: a handful of heavily cancelling sums miss the
            # bar (measured: <= 5 of the 9000 v_conics entries of big_blobs, whose splats cover hundreds of pixels
            # with |dx|^2 up to 2500; 0 .. 1e-4 elsewhere; <= 6.4e-6 on the BASELINE configs, test_gpu_parity_full.py)
            assert grad_close(got.cpu().numpy(), ref[key], max_share=1e-3, what="%s/%s" % (name, key)), key


@pytest.mark.parametrize("name", ["pixel_aligned", "blobs", "big_blobs"])
def test_project_bwd(oracle, cases, name):
    from prometheus_b200 import ops

    s, bg, st = cases[name]
    rng = np.random.default_rng(23)
    vis = (st["radii"] > 0).astype(np.float32)
    v_m2 = rng.standard_normal(st["means2d"].shape).astype(np.float32) * vis[..., None]
    v_d = rng.standard_normal(st["depths"].shape).astype(np.float32) * vis
    v_c = rng.standard_normal(st["conics"].shape).astype(np.float32) * vis[..., None]
    ref = oracle.project_bwd(s["means"], s["quats"], s["scales"], s["viewmats"], s["Ks"], s["width"], s["height"],
                             st["radii"], st["conics"], v_m2, v_d, v_c)
    got = ops.project_bwd(dev(s["means"]), dev(s["quats"]), dev(s["scales"]), dev(s["viewmats"]), dev(s["Ks"]),
                          s["width"], s["height"], dev(st["radii"]), dev(st["conics"]), dev(v_m2), dev(v_d), dev(v_c))
    for g_, key in zip(got, ("v_means", "v_quats", "v_scales")):
        assert grad_close(g_.cpu().numpy(), ref[key], max_share=PROJECT_BWD_SHARE, what="%s/%s" % (name, key)), key


# --------------------------------------------------------------------------- operator level (the drop-in)
@pytest.mark.parametrize("name", CASE_NAMES)
@pytest.mark.parametrize("render_mode", ["RGB+ED", "RGB"])
def test_rasterization_operator_fwd_bwd(oracle, cases, name, render_mode):
    """`rasterization(...)` end to end (gsplat contract) against the oracle's, forward and gradients."""
    import prometheus_b200 as pb

    s, bg, st = cases[name]
    keys = ("means", "quats", "scales", "opacities", "colors")
    ins_ref = [torch.from_numpy(s[k]).clone().requires_grad_(True) for k in keys]
    r_ref, a_ref, m_ref = oracle.rasterization(*ins_ref, torch.from_numpy(s["viewmats"]), torch.from_numpy(s["Ks"]),
                                               s["width"], s["height"], backgrounds=torch.from_numpy(bg),
                                               render_mode=render_mode, packed=False)
    ins = [dev(s[k]).requires_grad_(True) for k in keys]
    r, a, m = pb.rasterization(*ins, dev(s["viewmats"]), dev(s["Ks"]), s["width"], s["height"], backgrounds=dev(bg),
                               render_mode=render_mode, packed=False)
    assert r.shape == r_ref.shape and a.shape == a_ref.shape
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert torch.equal(m[k].cpu().long(), m_ref[k].long()), k
    assert m["means2d"].shape == m_ref["means2d"].shape and m["opacities"].shape == (s["viewmats"].shape[0], s["means"].shape[0])
    for k in ("tile_width", "tile_height", "width", "height", "tile_size", "n_cameras"):
        assert m[k] == m_ref[k], k
    keep = ~m_ref["fragile"].bool()
    assert allclose(r.detach().cpu()[keep], r_ref.detach()[keep], atol=1e-5, rtol=1e-4)
    assert allclose(a.detach().cpu()[keep], a_ref.detach()[keep], atol=1e-5, rtol=1e-4)
    # backward with O(1) upstream gradients on the non-fragile pixels
    gen = torch.Generator().manual_seed(3)
    wr = torch.randn(r_ref.shape, generator=gen) * keep[..., None]
    wa = torch.randn(a_ref.shape, generator=gen) * keep[..., None]
    m_ref["means2d"].retain_grad()
    m["means2d"].retain_grad()
    ((r_ref * wr).sum() + (a_ref * wa).sum()).backward()
    ((r * wr.to(DEV)).sum() + (a * wa.to(DEV)).sum()).backward()
    # End to end at the north_star's bar, element-wise. Both sides carry the exact final transmittance from
    # their own forward (round 1 rebuilt it as 1 - alpha like upstream and needed 1e-4 / 1e-3). What is left is
    # the difference between two correct fp32 forwards (expf vs ex2.approx): a 1-ulp change of an alpha near
    # 0.99 is a 6e-6 relative change of (1 - alpha) and of everything behind that splat; the share of entries
    # that miss the bar for that reason is bounded by OPERATOR_GRAD_SHARE and printed.
    for t, t_ref, k in zip(ins, ins_ref, keys):
        assert t.grad is not None, k
        assert grad_close(t.grad.cpu().numpy(), t_ref.grad.numpy(), max_share=OPERATOR_GRAD_SHARE, what="%s/%s/%s" % (name, render_mode, k)), k
    gm, gm_ref = m["means2d"].grad.cpu().numpy(), m_ref["means2d"].grad.numpy()  # the refiner side-contract
    assert grad_close(gm, gm_ref, max_share=OPERATOR_GRAD_SHARE, what="%s/%s/means2d" % (name, render_mode))


def test_reference_module_runs_unmodified_on_the_operator(golden_dir):
    """Module-level drop-in: our GaussianConverter/GaussianRenderer reproduce what the REFERENCE's
    modules produced (fixture made by running them over the oracle operator), incl. gradients,
    `.radii` and `.viewspace_points[k].grad`."""
    from prometheus_b200 import GaussianConverter, GaussianRenderer, scenes

    d = np.load(os.path.join(golden_dir, "renderer_e2e.npz"))
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    rend = GaussianRenderer(sh_degree=0, background=[1, 1, 1]).to(DEV)
    params = dev(d["params"]).requires_grad_(True)
    gaussians = conv(params, dev(d["cameras_in"]))
    h, w = int(d["h"]), int(d["w"])
    images, depths, masks, n1, n2 = rend(dev(d["cameras_out"]), gaussians, bg_color=tuple(d["bg"].tolist()), h=h, w=w)
    assert n1 is None and n2 is None
    assert images.shape == d["images"].shape and depths.shape == d["depths"].shape and masks.shape == d["masks"].shape
    B, N = d["cameras_out"].shape[:2]
    assert len(rend.radii) == B * N and len(rend.viewspace_points) == B * N
    # camera matrices differ from the reference's host path by an ulp (analytic inverse vs LAPACK),
    # so integer outputs may differ for a handful of borderline Gaussians and pixels near thresholds
    radii = torch.stack(rend.radii).cpu().numpy()
    assert (radii != d["radii"]).mean() < 1e-3
    for got, key, tol in ((images, "images", 2e-4), (depths, "depths", 5e-4), (masks, "masks", 2e-4)):
        diff = np.abs(got.detach().cpu().numpy() - d[key])
        assert np.quantile(diff, 0.999) < tol, (key, float(np.quantile(diff, 0.999)))
    loss = scenes.render_loss(images, depths, masks, dev(d["target"])) * images.numel()
    loss.backward()
    gp, gr = params.grad.cpu().numpy(), d["v_params"]
    rel = np.abs(gp - gr).sum() / np.abs(gr).sum()
    assert rel < 2e-3, rel
    assert rend.viewspace_points[0].shape == (1, gaussians[0].shape[1], 2)
    g2 = np.stack([p.grad.cpu().numpy() for p in rend.viewspace_points])
    assert np.abs(g2 - d["means2d_grad"]).sum() / np.abs(d["means2d_grad"]).sum() < 2e-3


def test_install_as_gsplat_shim():
    import sys

    import prometheus_b200 as pb

    pb.install_as_gsplat()
    from gsplat import rasterization  # the reference's import line, gaussians_renderer.py:11

    assert rasterization is pb.rasterization
    del sys.modules["gsplat"]


# --------------------------------------------------------------------------- full size (C2) properties
def test_full_size_properties_c2():
    """BASELINE configs[1]: 8 x 512^2 -> 2.1M Gaussians -> 8 views 512^2. Too big for the oracle in a test,
    so check size-independent properties of every stage."""
    from prometheus_b200 import GaussianConverter, ops, scenes
    from prometheus_b200.rasterization import rasterization

    sc = scenes.make_scene("C2")
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    xyz, feat, op, scl, rot = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
    G = sc.n_gaussians
    assert xyz.shape == (1, G, 3) and G == 8 * 512 * 512
    assert torch.isfinite(xyz).all() and (op > 0).all() and (op < 1).all()
    assert allclose(rot[0].norm(dim=-1).cpu().numpy(), 1.0, atol=1e-5, rtol=0)
    vm, K = ops.camera_prep(sc.cameras_out[0].to(DEV), 512, 512)
    bg = torch.ones(8, 3, device=DEV)
    r, a, m = rasterization(xyz[0], rot[0], scl[0], op[0, :, 0], feat[0, :, 0], vm, K, 512, 512, backgrounds=bg,
                            render_mode="RGB+ED", packed=False)
    ids, flat, off, tpg, radii = m["isect_ids"], m["flatten_ids"], m["isect_offsets"], m["tiles_per_gauss"], m["radii"]
    n = ids.numel()
    assert n == int(tpg.sum().item()) and n > 2 * G            # ~3 tiles per visible Gaussian, 8 cameras
    assert bool((ids[1:] >= ids[:-1]).all())                     # sortedness
    same = ids[1:] == ids[:-1]
    assert bool((flat[1:][same] > flat[:-1][same]).all())       # stability on equal keys
    assert torch.equal(torch.sort(flat.long())[0], torch.repeat_interleave(torch.arange(8 * G, device=DEV), tpg.flatten().long()))
    tb, cb = ops.key_bits(32, 32, 8)
    lin = (ids >> (32 + tb)) * 1024 + ((ids >> 32) & ((1 << tb) - 1))
    assert torch.equal(off.flatten().long(), torch.searchsorted(lin, torch.arange(8 * 1024, device=DEV)))
    assert torch.equal(ids >> (32 + tb), (flat // G).long())    # camera bits agree with the pair index
    assert torch.equal((ids & 0xFFFFFFFF).int(), m["depths"].flatten()[flat.long()].view(torch.int32))
    assert ((radii > 0).float().mean() > 0.9)
    assert torch.isfinite(r).all() and float(a.min()) >= 0.0 and float(a.max()) <= 1.0 + 1e-6
    assert float(a.mean()) > 0.99
    # idempotence / determinism of the forward path
    r2, a2, m2 = rasterization(xyz[0], rot[0], scl[0], op[0, :, 0], feat[0, :, 0], vm, K, 512, 512, backgrounds=bg,
                               render_mode="RGB+ED", packed=False)
    assert torch.equal(r, r2) and torch.equal(a, a2) and torch.equal(m2["isect_ids"], ids)
    # linearity in colour: compositing is linear in (colours, background)
    r3, _, _ = rasterization(xyz[0], rot[0], scl[0], op[0, :, 0], feat[0, :, 0] * 0.5, vm, K, 512, 512,
                             backgrounds=bg * 0.5, render_mode="RGB", packed=False)
    assert allclose(r3.cpu().numpy(), 0.5 * r[..., :3].cpu().numpy(), atol=1e-6, rtol=1e-5)


def test_gather_gaussians_single_rank_matches_plain_unproject():
    """world_size 1: the gather path (unproject writes into the rank's slot of the packed buffer) returns the
    same Gaussians, in the same order, as the plain converter."""
    from prometheus_b200 import GaussianConverter, ops, parallel, scenes

    sc = scenes.make_scene("C1", seed=21, views_in=3, size_in=24)
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    ref = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
    muls = torch.ones(14, device=DEV)
    n_local = sc.n_gaussians
    full = parallel.gather_gaussians(
        lambda slots: ops.unproject_fwd(sc.params[0].to(DEV), sc.cameras_in[0].to(DEV), muls, "mvsplat", 0.01, 100.0,
                                        0.001, 1.0, out=slots), n_local, torch.device(DEV))
    for got, want in zip(full, (ref[0][0], ref[1][0, :, 0], ref[2][0, :, 0], ref[3][0], ref[4][0])):
        assert torch.equal(got, want)


# --------------------------------------------------------------------------- fused sort-as-you-go fast path
@pytest.mark.parametrize("name", CASE_NAMES)
def test_fused_tile_render_is_bit_identical_to_separate_stages(cases, name):
    """tile_render_fwd (sorts only the prefix of each tile the compositor consumes) must give the same bits as
    tile_sort + composite_fwd: render, alphas, last_ids, and the sorted prefix of flatten_ids."""
    from prometheus_b200 import ops

    s, bg, st = cases[name]
    C = s["viewmats"].shape[0]
    args = (dev(st["means2d"]), dev(st["conics"]), dev(s["opacities"]), dev(s["colors"]), dev(st["depths"]), dev(bg))
    ref_r, ref_a, ref_l, ref_t = ops.composite_fwd(*args, dev(st["isect_offsets"]), dev(st["flatten_ids"]), s["width"], s["height"])
    counts = ops.bin_count(dev(st["means2d"]), dev(st["radii"]), s["width"], s["height"])
    offsets, stats = ops.tile_scan(counts)
    n, _ = stats.tolist()
    records = ops.bin_scatter(dev(st["means2d"]), dev(st["radii"]), dev(st["depths"]), offsets, s["width"], s["height"], n)
    r, a, l, flat, sorted_counts, flag, t = ops.tile_render_fwd(records, offsets, *args, s["width"], s["height"])
    assert int(flag.item()) == 0
    assert torch.equal(r, ref_r) and torch.equal(a, ref_a) and torch.equal(l, ref_l) and torch.equal(t, ref_t)
    off = st["isect_offsets"].reshape(-1)
    sc = sorted_counts.cpu().numpy().reshape(-1)
    cnt = np.diff(np.append(off, len(st["flatten_ids"])))
    assert (sc <= cnt).all() and (sc >= 0).all()
    flat_np = flat.cpu().numpy()
    for t in np.nonzero(sc)[0][:200]:
        assert np.array_equal(flat_np[off[t]:off[t] + sc[t]], st["flatten_ids"][off[t]:off[t] + sc[t]])
    # every pixel's last contributor lies inside the sorted prefix of its tile
    lmax = l.cpu().numpy().reshape(C, -1).max()
    assert lmax < len(st["flatten_ids"])


def test_fused_module_path_matches_unfused_including_gradients():
    """GaussianRenderer(fused=True) vs (fused=False): identical images; gradients agree to summation order."""
    from prometheus_b200 import GaussianConverter, GaussianRenderer, scenes

    sc = scenes.make_scene("C1", seed=17, views_in=3, size_in=64, views_out=3, size_out=80)
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    outs = []
    for fused in (True, False):
        rend = GaussianRenderer().to(DEV)
        rend.fused = fused
        params = sc.params.to(DEV).requires_grad_(True)
        images, depths, masks, _, _ = rend(sc.cameras_out.to(DEV), conv(params, sc.cameras_in.to(DEV)), bg_color=(0.1, 0.6, 0.3), h=80, w=80)
        (scenes.render_loss(images, depths, masks, torch.zeros_like(images)) * images.numel()).backward()
        outs.append((images.detach(), depths.detach(), masks.detach(), params.grad.clone(), torch.stack(rend.radii)))
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[4], b[4])
    assert grad_close(a[3].cpu().numpy(), b[3].cpu().numpy(), max_share=ORDER_SHARE, what="fused vs unfused")


def test_fused_overflow_falls_back():
    """3000 splats with identical depth on one tile: even a second-level depth sub-bucket exceeds the fused group capacity (768),
    the overflow flag is raised and the module falls back to the separate stages — same image either way."""
    import importlib

    from prometheus_b200 import ops

    rz = importlib.import_module("prometheus_b200.rasterization")
    n = 3000
    g = torch.Generator().manual_seed(4)
    means = torch.zeros(n, 3)
    means[:, 2] = 2.5
    args = [means.to(DEV), torch.tensor([[1.0, 0, 0, 0]]).repeat(n, 1).to(DEV), torch.full((n, 3), 0.01).to(DEV),
            torch.full((n,), 0.002).to(DEV), torch.rand(n, 3, generator=g).to(DEV), torch.eye(4)[None].to(DEV),
            torch.tensor([[[100.0, 0, 24.0], [0, 100.0, 24.0], [0, 0, 1]]]).to(DEV), 64, 64]
    r1, a1, m1 = rz._rasterize_core(*args, packed=False, fused=True)
    r2, a2, m2 = rz._rasterize_core(*args, packed=False, fused=False)
    assert m1["isect_ids"] is not None  # fell back to the separate-stage path
    assert torch.equal(r1, r2) and torch.equal(a1, a2)


@pytest.mark.parametrize("n_same", [90, 400, 700])
def test_fused_equal_depth_runs(n_same):
    """Runs of identical depth inside one tile (ties ordered by flatten id): short runs go through the fused
    kernel's all-pairs ranking, runs above its limit (128) through the in-kernel radix sort, and neither may
    raise the overflow flag while the run still fits one group (1024). Mixed with 700 splats at spread depths."""
    import importlib

    rz = importlib.import_module("prometheus_b200.rasterization")
    g = torch.Generator().manual_seed(9)
    n_far = 700
    n = n_same + n_far
    means = torch.zeros(n, 3)
    means[:, :2] = (torch.rand(n, 2, generator=g) - 0.5) * 0.08
    means[:n_same, 2] = 2.5
    means[n_same:, 2] = 2.0 + torch.rand(n_far, generator=g) * 2.0
    perm = torch.randperm(n, generator=g)
    means = means[perm]
    args = [means.to(DEV), torch.tensor([[1.0, 0, 0, 0]]).repeat(n, 1).to(DEV), torch.full((n, 3), 0.01).to(DEV),
            torch.full((n,), 0.01).to(DEV), torch.rand(n, 3, generator=g).to(DEV), torch.eye(4)[None].to(DEV),
            torch.tensor([[[100.0, 0, 24.0], [0, 100.0, 24.0], [0, 0, 1]]]).to(DEV), 64, 64]
    r1, a1, m1 = rz._rasterize_core(*args, packed=False, fused=True)
    r2, a2, m2 = rz._rasterize_core(*args, packed=False, fused=False)
    assert m1["isect_ids"] is None  # stayed on the fused path
    assert torch.equal(r1, r2) and torch.equal(a1, a2)
    off = m2["isect_offsets"].reshape(-1).cpu().numpy()
    sc = m1["sorted_counts"].reshape(-1).cpu().numpy()
    f1, f2 = m1["flatten_ids"].cpu().numpy(), m2["flatten_ids"].cpu().numpy()
    assert sc.max() > n_same  # low opacity: the whole list is consumed, including the equal-depth run
    for t in np.nonzero(sc)[0]:
        assert np.array_equal(f1[off[t]:off[t] + sc[t]], f2[off[t]:off[t] + sc[t]])


@pytest.mark.parametrize("name", CASE_NAMES)
def test_project_with_fused_tile_histogram(cases, name):
    """pm_project_count_fwd = pm_project_fwd + pm_bin_count in one pass: identical projection outputs and
    exactly the per-tile histogram the oracle's offsets imply."""
    from prometheus_b200 import ops

    s, _, st = cases[name]
    proj_args = [dev(s[k]) for k in ("means", "quats", "scales", "viewmats", "Ks")] + [s["width"], s["height"]]
    a = ops.project_fwd(*proj_args)
    b = ops.project_fwd(*proj_args, count_tiles=False, tile_histogram=True)
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x, y)
    assert b[4] is None
    off = st["isect_offsets"].reshape(-1)
    want = np.diff(np.append(off, len(st["flatten_ids"])))
    assert np.array_equal(b[5].cpu().numpy().reshape(-1), want)
    assert torch.equal(b[5], ops.bin_count(a[1], a[0], s["width"], s["height"]))


def test_project_fused_histogram_scattered_large_image():
    """Hash-table path of the fused histogram (window of touched tiles > 2048 cells) and > 32 cameras
    (two launches of the camera-looping kernel)."""
    from prometheus_b200 import ops

    g = torch.Generator().manual_seed(21)
    n, C, W, H = 3000, 35, 1024, 800
    means = torch.cat([(torch.rand(n, 2, generator=g) - 0.5) * 3.0, 2.0 + torch.rand(n, 1, generator=g)], 1)
    quats = torch.randn(n, 4, generator=g)
    scales = 0.005 + 0.05 * torch.rand(n, 3, generator=g)
    viewmats = torch.eye(4).repeat(C, 1, 1)
    viewmats[:, 0, 3] = torch.linspace(-0.3, 0.3, C)
    Ks = torch.tensor([[600.0, 0, W / 2], [0, 600.0, H / 2], [0, 0, 1]]).repeat(C, 1, 1)
    args = [dev(means), dev(quats), dev(scales), dev(viewmats), dev(Ks), W, H]
    a = ops.project_fwd(*args)
    b = ops.project_fwd(*args, count_tiles=True, tile_histogram=True)
    for x, y in zip(a[:5], b[:5]):
        assert torch.equal(x, y)
    assert torch.equal(b[5], ops.bin_count(a[1], a[0], W, H))
    assert int(b[5].sum()) == int(a[4].sum()) > 0


# --------------------------------------------------------------------------- packed module path, no-sync capacity
def _module_scene():
    from prometheus_b200 import GaussianConverter, scenes

    sc = scenes.make_scene("C1", seed=23, views_in=3, size_in=64, views_out=3, size_out=80)
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    return sc, conv


def test_speculative_capacity_too_small_is_detected_and_rerendered():
    """The fused module path sizes its intersection buffers from the previous call instead of reading the count
    back mid-pipeline. A hint that is far too small must be caught by the end-of-forward check and re-rendered
    with the exact size; a generous hint must give the same bits as the exact size."""
    from prometheus_b200 import GaussianRenderer

    sc, conv = _module_scene()
    with torch.no_grad():
        g = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
        ref = GaussianRenderer().to(DEV)
        ref.fused = False
        want = ref(sc.cameras_out.to(DEV), g, h=80, w=80)[:3]
        rend = GaussianRenderer().to(DEV)
        first = rend(sc.cameras_out.to(DEV), g, h=80, w=80)[:3]          # no hint yet: exact count (one sync)
        (key, hint), = rend._cap_hint.items()
        assert hint > 4096
        second = rend(sc.cameras_out.to(DEV), g, h=80, w=80)[:3]         # speculative, generous
        rend._cap_hint[key] = 100
        third = rend(sc.cameras_out.to(DEV), g, h=80, w=80)[:3]          # speculative, far too small
        assert rend._cap_hint[key] == hint
    for got in (first, second, third):
        for a, b in zip(got, want):
            assert torch.equal(a, b)


def test_single_pass_binning_same_bits_and_overflow_recovery():
    """Once a shape's counts are known the fused path bins in ONE pass (pm_project_scatter_fwd: fixed-capacity tile
    regions, no bin_scatter). Same images bit for bit as two passes; a region size that is too small for the largest
    tile must be detected (max tile count > tile capacity) and the scene rendered again."""
    from prometheus_b200 import GaussianRenderer, ops

    sc, conv = _module_scene()
    cams = sc.cameras_out.to(DEV)
    with torch.no_grad():
        g = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
        two = GaussianRenderer().to(DEV)
        two.single_pass_binning = False
        two(cams, g, h=80, w=80)
        want = two(cams, g, h=80, w=80)[:3]
        rend = GaussianRenderer().to(DEV)
        first = rend(cams, g, h=80, w=80)[:3]                  # exact sizes, two passes; sets both hints
        (key, tcap), = rend._tile_cap_hint.items()
        assert tcap > 64 and rend._tile_cap_for(key, rend._cap_hint[key], 80, 80) == tcap
        prof, ops.PROFILE = ops.PROFILE, []
        second = rend(cams, g, h=80, w=80)[:3]                 # single pass
        names = [n for n, _, _ in ops.PROFILE]
        ops.PROFILE = prof
        assert "project_scatter_fwd" in names and "bin_scatter" not in names
        rend._tile_cap_hint[key] = 16                          # far too small: records of full tiles are dropped
        third = rend(cams, g, h=80, w=80)[:3]
        assert 64 < rend._tile_cap_hint[key] <= tcap           # re-rendered, hint restored (from the usable-tile counts)
        rend.defer_checks = True
        rend._tile_cap_hint[key] = 16
        rend(cams, g, h=80, w=80)
        assert not rend.validate()
        rend(cams, g, h=80, w=80)
        assert rend.validate()
    for got in (first, second, third):
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    # fwd + bwd through the single-pass path: same gradients as two passes (same stream, same kernels)
    grads = []
    for single in (False, True):
        r = GaussianRenderer().to(DEV)
        r.single_pass_binning = single
        for _ in range(2):
            p = sc.params.to(DEV).clone().requires_grad_(True)
            im, de, ma, _, _ = r(cams, conv(p, sc.cameras_in.to(DEV)), h=80, w=80)
            (im.square().mean() + de.mean() + ma.mean()).backward()
        grads.append(p.grad.clone())
    torch.testing.assert_close(grads[0], grads[1], rtol=1e-4, atol=1e-7)


def test_hit_map_backward_equals_the_culling_backward():
    """The fused forward leaves a map of the list entries each warp footprint evaluated; composite_bwd replays exactly
    those instead of culling every chunk itself. Same gradients as the culling form (up to the order of the atomics),
    on the two-pass and the single-pass layouts, and the map is not empty."""
    from prometheus_b200 import GaussianRenderer, ops

    sc, conv = _module_scene()
    cams = sc.cameras_out.to(DEV)
    grads = {}
    old = ops.HIT_MAP
    try:
        for use_map in (False, True):
            ops.HIT_MAP = use_map
            r = GaussianRenderer().to(DEV)
            for call in range(2):  # first call: two binning passes; second: single pass, usable tiles
                p = sc.params.to(DEV).clone().requires_grad_(True)
                im, de, ma, _, _ = r(cams, conv(p, sc.cameras_in.to(DEV)), h=80, w=80)
                (im.square().mean() + de.mean() + ma.mean()).backward()
                grads[(use_map, call)] = p.grad.clone()
    finally:
        ops.HIT_MAP = old
    for call in range(2):
        share, worst = violations(grads[(True, call)].cpu().numpy(), grads[(False, call)].cpu().numpy())
        assert share <= 1e-4, (call, share, worst)
    # the map itself: bits only inside the tiles' lists, at least one per non-empty sorted tile
    with torch.no_grad():
        g = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
        vm, K = ops.camera_prep(cams[0], 80, 80)
        radii, m2, dep, con, _, counts = ops.project_fwd(g[0][0], g[4][0], g[3][0], vm, K, 80, 80, count_tiles=False, tile_histogram=True)
        offs, stats, order = ops.tile_scan(counts, want_order=True)
        n = int(stats[0].item())
        rec = ops.bin_scatter(m2, radii, dep, offs, 80, 80, n)
        out = ops.tile_render_packed_fwd(rec, offs, m2, con, g[2][0, :, 0].contiguous(), g[1][0, :, 0].contiguous(), dep,
                                         torch.ones(vm.shape[0], 3, device=DEV), 80, 80, tile_order=order, for_backward=True)
    hit = out[9]
    assert hit is not None and hit.shape[0] == 8 and hit.shape[1] >= (n >> 5) + counts.numel() + 1
    words = hit.cpu().numpy().view(np.uint32)
    off = offs.reshape(-1).cpu().numpy().astype(np.int64)
    cnt = counts.reshape(-1).cpu().numpy().astype(np.int64)
    used = np.zeros(words.shape[1], dtype=bool)
    for t in range(len(off)):
        b = (off[t] >> 5) + t
        nw = (cnt[t] + 31) // 32
        used[b:b + nw] = True
        if cnt[t] % 32 and nw:
            assert not (words[:, b + nw - 1] >> np.uint32(cnt[t] % 32)).any()   # no bit beyond the tile's list
    assert not words[:, ~used].any() and words.any()


def test_single_pass_binning_fallback_paths():
    """Single-pass binning outside its comfort zone: unstructured Gaussians on a 66 x 66 tile grid (a CTA's bounding
    box exceeds the shared-memory window: records go to global memory one atomic each) with a few splats so large
    that their tile AABB does not fit the packed form (> 63 tiles on an axis). Same bits as two passes."""
    from prometheus_b200.rasterization import render_views_packed

    s = random_blobs(n=4096, C=2, W=1056, H=1056, seed=11)
    s["scales"][:6] = 4.0   # a handful of splats that cover the whole image
    s["Ks"][:, 0, 0] = s["Ks"][:, 1, 1] = 900.0
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    args = [dev(s[k]) for k in ("means", "quats", "scales", "opacities", "colors", "viewmats", "Ks")]
    bg = torch.ones(2, 3, device=DEV)
    with torch.no_grad():
        im0, de0, ma0, info0 = render_views_packed(*args, bg, 1056, 1056, usable_tiles=False)
        n_isects, max_count = (int(v) for v in info0["stats"].tolist())
        assert max_count > 6 and int(info0["overflow_flag"].item()) == 0
        im1, de1, ma1, info1 = render_views_packed(*args, bg, 1056, 1056, capacity=n_isects + 100, tile_capacity=max_count + 3,
                                                   usable_tiles=False)
        assert info1["tile_capacity"] == max_count + 3 and info1["stats"].tolist() == [n_isects, max_count]
        assert torch.equal(info0["isect_offsets"], info1["isect_offsets"])
        # the module's default: only the tiles a splat can be seen in — fewer records, same pixels
        im2, de2, ma2, info2 = render_views_packed(*args, bg, 1056, 1056, capacity=n_isects + 100, tile_capacity=max_count + 3)
        assert info2["usable_tiles"] and info2["stats"][0].item() < n_isects and info2["stats"][1].item() <= max_count
        # ... and the two-pass form of the same rule (first call of a shape): same counts, same pixels
        im3, de3, ma3, info3 = render_views_packed(*args, bg, 1056, 1056)
        assert info3["usable_tiles"] and torch.equal(info3["stats"], info2["stats"])
        assert torch.equal(info3["isect_offsets"], info2["isect_offsets"])
        assert torch.equal(im3, im0) and torch.equal(de3, de0) and torch.equal(ma3, ma0)
    for a, b, c in ((im0, im1, im2), (de0, de1, de2), (ma0, ma1, ma2)):
        assert torch.equal(a, b) and torch.equal(a, c)


@pytest.mark.parametrize("C", [11, 35])
def test_single_pass_binning_many_cameras(C):
    """Camera counts that do not fill the 8-camera window batches (11 = 8 + 3) or the 32-camera launches (35 = 32 + 3)."""
    from prometheus_b200.rasterization import render_views_packed

    s = random_blobs(n=3000, C=C, W=96, H=64, seed=5)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    args = [dev(s[k]) for k in ("means", "quats", "scales", "opacities", "colors", "viewmats", "Ks")]
    bg = torch.rand(C, 3, device=DEV)
    with torch.no_grad():
        im0, de0, ma0, info0 = render_views_packed(*args, bg, 96, 64, usable_tiles=False)
        n_isects, max_count = (int(v) for v in info0["stats"].tolist())
        im1, de1, ma1, info1 = render_views_packed(*args, bg, 96, 64, capacity=n_isects + 7, tile_capacity=max_count,
                                                   usable_tiles=False)
        assert info1["stats"].tolist() == [n_isects, max_count] and torch.equal(info0["isect_offsets"], info1["isect_offsets"])
        im2, de2, ma2, info2 = render_views_packed(*args, bg, 96, 64, capacity=n_isects + 7, tile_capacity=max_count)
        assert info2["stats"][0].item() <= n_isects and bool((info2["isect_offsets"] <= info0["isect_offsets"]).all())
        # histogram-only form of the same call (records = NULL): the counts the scatter call bins
        from prometheus_b200 import ops
        radii, m2, dep, con, _, counts = ops.project_fwd(*args[:3], *args[5:7], 96, 64, count_tiles=False, tile_histogram=True,
                                                         opacities=args[3])
        offs, stats = ops.tile_scan(counts)
        assert torch.equal(offs, info2["isect_offsets"])
        # the scatter pass of the two-pass form applies the same rule (tile_sort.cu vs project.cu: same tiles bit for
        # bit): every slot of every segment holds a record of a visible pair with that pair's depth
        n_u = int(stats[0].item())
        rec = ops.bin_scatter(m2, radii, dep, offs, 96, 64, n_u, conics=con, opacities=args[3])
        f = (rec >> 32) & 0xffffffff
        assert bool((f < radii.numel()).all()) and bool((radii.reshape(-1)[f] > 0).all())
        assert torch.equal((rec & 0xffffffff).to(torch.int32), dep.reshape(-1)[f].view(torch.int32))
        im3, de3, ma3, info3 = render_views_packed(*args, bg, 96, 64)   # two passes, usable tiles
        assert torch.equal(info3["stats"], info2["stats"]) and torch.equal(im3, im0) and torch.equal(de3, de0)
    for a, b, c in ((im0, im1, im2), (de0, de1, de2), (ma0, ma1, ma2)):
        assert torch.equal(a, b) and torch.equal(a, c)


def test_backward_with_undefined_cotangents():
    """The autograd nodes do not materialise undefined cotangents (they used to arrive as zero tensors the size of the
    outputs — hundreds of MB per step for the non-differentiable ones): a loss on the images alone hands the compositing
    backward None for depths / masks; same gradient as with explicit zero-weight terms."""
    from prometheus_b200 import GaussianRenderer

    sc, conv = _module_scene()
    cams = sc.cameras_out.to(DEV)
    grads = []
    for explicit in (False, True):
        rend = GaussianRenderer().to(DEV)
        p = sc.params.to(DEV).clone().requires_grad_(True)
        im, de, ma, _, _ = rend(cams, conv(p, sc.cameras_in.to(DEV)), h=80, w=80)
        loss = im.square().mean()
        if explicit:
            loss = loss + 0.0 * de.sum() + 0.0 * ma.sum()
        loss.backward()
        assert torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0
        grads.append(p.grad.clone())
    torch.testing.assert_close(grads[0], grads[1], rtol=1e-4, atol=1e-7)
    # and the other way round: only the depth output
    rend = GaussianRenderer().to(DEV)
    p = sc.params.to(DEV).clone().requires_grad_(True)
    im, de, ma, _, _ = rend(cams, conv(p, sc.cameras_in.to(DEV)), h=80, w=80)
    de.mean().backward()
    assert torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0


def test_deferred_checks_and_pipeline_recover_from_a_bad_hint():
    """defer_checks: forward() never blocks and validate() reports the bad hint; DecodeRenderPipeline re-renders
    the affected scene before handing out the host copies."""
    from prometheus_b200 import DecodeRenderPipeline, GaussianRenderer

    sc, conv = _module_scene()
    cams = sc.cameras_out.to(DEV)
    with torch.no_grad():
        g = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
        rend = GaussianRenderer().to(DEV)
        want = rend(cams, g, h=80, w=80)[:3]
        (key, hint), = rend._cap_hint.items()
        rend.defer_checks = True
        rend(cams, g, h=80, w=80)
        assert rend.validate()
        rend._cap_hint[key] = 64
        rend(cams, g, h=80, w=80)
        assert not rend.validate()
        rend.defer_checks = False
    pipe = DecodeRenderPipeline(conv, rend, 80, 80, device=DEV)
    ins = tuple(t.pin_memory() for t in (sc.params, sc.cameras_in, sc.cameras_out))
    outs = [[torch.empty_like(w.cpu()).pin_memory() for w in want] for _ in range(3)]
    pipe.prefetch(*ins)
    for i in range(3):
        if i == 1:
            rend._cap_hint[key] = 64   # scene 1 renders into buffers that are too small
        pipe.run(outs[i], next_inputs_h=ins if i < 2 else None)
    pipe.flush()  # blocks the host until the device->host copies have landed: out_h is read straight away
    for o in outs:
        for a, b in zip(o, want):
            assert torch.equal(a, b.cpu())


def test_graphed_decode_render_replays_bit_identical_and_recovers_from_overflow():
    """GraphedDecodeRender: the whole decode+render launch set replayed from a CUDA graph gives the eager modules'
    bits for every scene of the captured shape (device or pinned-host inputs, several scenes per call), counts its
    launches, and a scene that overflows the captured intersection buffers is detected, re-rendered and the plan
    re-captured."""
    from prometheus_b200 import GaussianConverter, GaussianRenderer, GraphedDecodeRender, ops, scenes

    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    rend = GaussianRenderer().to(DEV)
    eager = GaussianRenderer().to(DEV)
    gr = GraphedDecodeRender(conv, rend, 80, 80)

    def want_for(sc):
        with torch.no_grad():
            return eager(sc.cameras_out.to(DEV), conv(sc.params.to(DEV), sc.cameras_in.to(DEV)), h=80, w=80)[:3]

    for B in (1, 2):  # one scene, and a scene-batched pair
        scs = [scenes.make_scene("C1", seed=s, scenes=B, views_in=3, size_in=64, views_out=3, size_out=80) for s in (23, 24, 25)]
        for i, sc in enumerate(scs):
            ins = (sc.params, sc.cameras_in, sc.cameras_out)
            ins = tuple(t.pin_memory() for t in ins) if i == 1 else tuple(t.to(DEV) for t in ins)
            n0 = ops.LAUNCHES
            got = gr.run(*ins)
            assert ops.LAUNCHES - n0 >= 5  # unproject, camera prep, project + scatter, scan, render (+ capture's eager pass)
            for a, b in zip(got, want_for(sc)):
                assert torch.equal(a, b)
        assert len(gr._plans) == B  # one capture per shape, replayed for the other scenes

    # a scene with ~4x larger splats needs more intersections than the capture was sized for
    sc = scenes.make_scene("C1", seed=23, scenes=1, views_in=3, size_in=64, views_out=3, size_out=80)
    big = sc.params.clone()
    big[:, :, 7:10] += 1.5
    key = gr._key(sc.params, sc.cameras_in, sc.cameras_out) + (0,)
    old_cap = gr._plans[key]["jobs"][0]["capacity"]
    got = gr.run(big.to(DEV), sc.cameras_in.to(DEV), sc.cameras_out.to(DEV), check=False)
    assert not gr.validate()
    got = gr.recover()
    with torch.no_grad():
        want = eager(sc.cameras_out.to(DEV), conv(big.to(DEV), sc.cameras_in.to(DEV)), h=80, w=80)[:3]
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert gr._plans[key]["stale"]
    got = gr.run(big.to(DEV), sc.cameras_in.to(DEV), sc.cameras_out.to(DEV))  # re-captured, larger buffers
    assert not gr._plans[key]["stale"] and gr._plans[key]["jobs"][0]["capacity"] > old_cap
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_pipeline_graph_mode_streams_scenes_and_recovers_from_overflow():
    """DecodeRenderPipeline(graph=True): pinned host scenes in, pinned host images out, every scene replayed from
    one of two alternating CUDA-graph plans; same bits as the eager modules, including for a scene in the middle of
    the stream that overflows the captured buffers (re-rendered, plans re-captured) and the scenes after it."""
    from prometheus_b200 import DecodeRenderPipeline, GaussianConverter, GaussianRenderer, scenes

    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    eager = GaussianRenderer().to(DEV)
    scs = [scenes.make_scene("C1", seed=40 + i, scenes=1, views_in=3, size_in=64, views_out=3, size_out=80) for i in range(7)]
    scs[3].params[:, :, 7:10] += 1.5  # ~4x larger splats: more intersections than the plans were captured for
    ins = [tuple(t.clone().pin_memory() for t in (sc.params, sc.cameras_in, sc.cameras_out)) for sc in scs]
    want = []
    with torch.no_grad():
        for sc in scs:
            want.append(eager(sc.cameras_out.to(DEV), conv(sc.params.to(DEV), sc.cameras_in.to(DEV)), h=80, w=80)[:3])
    pipe = DecodeRenderPipeline(conv, GaussianRenderer().to(DEV), 80, 80, device=DEV, graph=True)
    outs = [[torch.empty_like(w.cpu()).pin_memory() for w in ws] for ws in want]
    pipe.prefetch(*ins[0])
    for i in range(len(scs)):
        pipe.run(outs[i], next_inputs_h=ins[i + 1] if i + 1 < len(scs) else None)
    pipe.flush()
    for i, (o, ws) in enumerate(zip(outs, want)):
        for a, b in zip(o, ws):
            assert torch.equal(a, b.cpu()), i
    assert len(pipe.graphed._plans) == 2


@pytest.mark.parametrize("graph", [False, True])
def test_pipeline_run_batch_streams_a_large_batch_in_slices(graph):
    """run_batch: a host batch of 5 scenes in slices of 2 (2 + 2 + 1) equals the one-shot render of the batch."""
    from prometheus_b200 import DecodeRenderPipeline, GaussianConverter, GaussianRenderer, scenes

    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    sc = scenes.make_scene("C1", seed=61, scenes=5, views_in=2, size_in=64, views_out=3, size_out=72)
    with torch.no_grad():
        want = GaussianRenderer().to(DEV)(sc.cameras_out.to(DEV), conv(sc.params.to(DEV), sc.cameras_in.to(DEV)), h=72, w=72)[:3]
    pipe = DecodeRenderPipeline(conv, GaussianRenderer().to(DEV), 72, 72, device=DEV, graph=graph)
    ins = tuple(t.pin_memory() for t in (sc.params, sc.cameras_in, sc.cameras_out))
    outs = [torch.empty_like(w.cpu()).pin_memory() for w in want]
    pipe.run_batch(ins, outs, chunk=2)
    pipe.flush()
    for a, b in zip(outs, want):
        assert torch.equal(a, b.cpu())


@pytest.mark.parametrize("graph", [False, True])
def test_pipeline_compact_outputs(graph):
    """DecodeRenderPipeline(compact_outputs=True): uint8 frames (export_video's conversion, io.py:181) and fp16
    depth / mask instead of fp32 — the same pixels, converted on the device before the copy."""
    from prometheus_b200 import DecodeRenderPipeline, GaussianConverter, GaussianRenderer, scenes

    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    sc = scenes.make_scene("C1", seed=71, scenes=2, views_in=2, size_in=48, views_out=3, size_out=64)
    with torch.no_grad():
        im, de, ma = GaussianRenderer().to(DEV)(sc.cameras_out.to(DEV), conv(sc.params.to(DEV), sc.cameras_in.to(DEV)), h=64, w=64)[:3]
    pipe = DecodeRenderPipeline(conv, GaussianRenderer().to(DEV), 64, 64, device=DEV, graph=graph, compact_outputs=True)
    ins = tuple(t.pin_memory() for t in (sc.params, sc.cameras_in, sc.cameras_out))
    outs = [torch.empty(2, 3, 64, 64, 3, dtype=torch.uint8).pin_memory(), torch.empty(2, 3, 1, 64, 64, dtype=torch.float16).pin_memory(),
            torch.empty(2, 3, 1, 64, 64, dtype=torch.float16).pin_memory()]
    pipe.prefetch(*ins)
    pipe.run(outs)
    pipe.flush()
    want_u8 = im.cpu().permute(0, 1, 3, 4, 2).mul(1 / 2).add(1 / 2).clamp(0, 1).mul(255).to(torch.uint8)
    assert torch.equal(outs[0], want_u8) and float(outs[0].float().std()) > 5
    assert torch.equal(outs[1], de.half().cpu()) and torch.equal(outs[2], ma.half().cpu())


# --------------------------------------------------------------------------- trajectory row
@pytest.mark.parametrize("name", ["orbit_big", "smooth", "repeat", "video120"])
def test_sample_cameras_vs_reference_golden(golden_dir, name):
    """pm_sample_cameras against the reference's own sample_from_dense_cameras outputs (pinned golden), all
    samples of a trajectory in one launch."""
    import os

    from prometheus_b200 import sample_from_dense_cameras

    g = np.load(os.path.join(golden_dir, "trajectory.npz"))
    got = sample_from_dense_cameras(dev(g[name + "_dense"]), dev(g[name + "_t"])).cpu().numpy()
    want = g[name + "_out"]
    assert got.shape == want.shape
    assert np.allclose(got[..., :12], want[..., :12], atol=2e-5, rtol=0)      # float tolerance: sinf / acosf differ by ~1 ulp
    assert np.allclose(got[..., 12:], want[..., 12:], atol=0, rtol=5e-5)


def test_sample_cameras_batched_equals_per_trajectory(golden_dir, oracle):
    import os

    from oracle import trajectory_oracle as T
    from prometheus_b200 import ops

    g = np.load(os.path.join(golden_dir, "trajectory.npz"))
    a, b = g["orbit_big_dense"][0][:8], g["video120_dense"][0][:8]
    dense = np.stack([a, b])
    t = np.stack([np.linspace(0, 1, 50, dtype=np.float32), np.random.default_rng(3).uniform(0, 1, 50).astype(np.float32)])
    got = ops.sample_cameras(dev(dense), dev(t)).cpu().numpy()
    for i in range(2):
        want = T.sample_from_dense_cameras(dense[i], t[i])
        assert np.allclose(got[i, :, :12], want[:, :12], atol=2e-5, rtol=0)
        assert np.allclose(got[i, :, 12:], want[:, 12:], atol=0, rtol=5e-5)


@pytest.mark.parametrize("shape", [(2, 3, 16, 32), (3, 3, 9, 7), (1, 3, 5, 4)])
def test_pack_frames_u8_bit_exact(shape):
    """Device conversion = export_video's host chain (io.py:181), bit for bit, incl. widths not divisible by 4."""
    from prometheus_b200 import ops

    gen = torch.Generator().manual_seed(5)
    x = torch.randn(*shape, generator=gen) * 0.8
    x.view(-1)[:4] = torch.tensor([-1.0, 1.0, 0.0, 0.999999])
    want = torch.stack([f.permute(1, 2, 0).mul(1 / 2).add(1 / 2).clamp(0, 1).mul(255) for f in x]).numpy().astype(np.uint8)
    got = ops.pack_frames_u8(x.to(DEV)).cpu().numpy()
    assert np.array_equal(got, want)


def test_trajectory_renderer_matches_the_per_frame_loop(golden_dir):
    """TrajectoryRenderer (all poses in one launch, chunks of cameras, device-side uint8 packing, async copies)
    vs the reference's export_video loop restated frame by frame on top of the same renderer."""
    import os

    from prometheus_b200 import GaussianConverter, GaussianRenderer, TrajectoryRenderer, ops, scenes

    sc = scenes.make_scene("C1", seed=31, views_in=3, size_in=48, views_out=2, size_out=64)
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    rend = GaussianRenderer().to(DEV)
    rend.view_chunk = 4
    dense = sc.cameras_in[0].clone()          # the 3 input cameras as a path (OpenGL c2w + intrinsics)
    n_frames, S = 10, 64
    with torch.no_grad():
        g = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
        frames, depths = TrajectoryRenderer(rend).render(g, dense.to(DEV), num_frames=n_frames, render_size=S)
        first_read = frames.clone()  # render() has blocked the host on its copies: readable straight away
        torch.cuda.synchronize()
        assert torch.equal(first_read, frames)
        t = torch.tensor([[i / n_frames for i in range(n_frames)]], dtype=torch.float32)  # io.py:175
        cams = ops.sample_cameras(dense[None].to(DEV), t.to(DEV))
        ref = GaussianRenderer().to(DEV)
        for i in range(n_frames):
            im, de, _, _, _ = ref(cams[:, i:i + 1], g, h=S, w=S)
            want = im.reshape(3, S, S).permute(1, 2, 0).cpu().mul(1 / 2).add(1 / 2).clamp(0, 1).mul(255).numpy().astype(np.uint8)
            assert np.array_equal(frames[i].numpy(), want), i
            assert np.array_equal(depths[i].numpy(), de.reshape(1, S, S).permute(1, 2, 0).cpu().numpy()), i
    assert frames.shape == (n_frames, S, S, 3) and frames.numpy().std() > 5
    # one rank's share of the frames (view-sharded video): same bits as the corresponding slice
    with torch.no_grad():
        part_f, part_d = TrajectoryRenderer(rend).render(g, dense.to(DEV), num_frames=n_frames, render_size=S, frame_range=(3, 9))
    assert part_f.shape[0] == 6 and torch.equal(part_f, frames[3:9]) and torch.equal(part_d, depths[3:9])


# --------------------------------------------------------------------------- loss-side glue
@pytest.mark.parametrize("shape", [(2, 3, 40, 56), (1, 1, 17, 9)])
def test_render_loss_terms_match_the_reference_expressions(shape):
    """render_loss_terms vs the expressions of gs_decoder_system.py:214-239 written with ATen ops, values and
    gradients (depths straddle both clip bounds; validity mask bit-exact)."""
    from prometheus_b200 import render_loss_terms

    B, N, H, W = shape
    g = torch.Generator().manual_seed(5)
    images = (torch.rand(B, N, 3, H, W, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    target = (torch.rand(B, N, 3, H, W, generator=g) * 2 - 1).to(DEV)
    depths = torch.exp(torch.randn(B, N, 1, H, W, generator=g) * 3.0)
    depths.view(-1)[:7] = torch.tensor([0.1, 100.0, 0.0, 0.05, 150.0, 0.1000001, 99.99999])
    depths = depths.to(DEV).requires_grad_(True)
    masks = torch.rand(B, N, 1, H, W, generator=g).to(DEV).requires_grad_(True)
    w_disp = torch.randn(B, N, 1, H, W, generator=g).to(DEV)

    def reference(images, depths, masks):
        mask = torch.ones_like(depths)
        mask[depths <= 0.1] = 0
        disp = 1 / depths.clip(0.1, 100)
        return torch.nn.functional.mse_loss(images, target), disp, mask, (1 - masks).mean()

    want = reference(images, depths, masks)
    (0.7 * want[0] + (want[1] * w_disp).sum() + 1.3 * want[3]).backward()
    want_g = [t.grad.clone() for t in (images, depths, masks)]
    for t in (images, depths, masks):
        t.grad = None
    got = render_loss_terms(images, depths, masks, target)
    (0.7 * got[0] + (got[1] * w_disp).sum() + 1.3 * got[3]).backward()
    assert torch.allclose(got[0], want[0], rtol=1e-6, atol=0) and torch.allclose(got[3], want[3], rtol=1e-6, atol=0)
    assert torch.equal(got[1], want[1].detach()) and torch.equal(got[2], want[2])
    for a, b in zip((images.grad, depths.grad, masks.grad), want_g):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-12), (a - b).abs().max()


def test_fused_loss_glue_through_the_renderer():
    """The training-style step with render_loss_fused equals the one with the ATen loss: loss and map gradients."""
    from prometheus_b200 import GaussianRenderer, scenes

    sc, conv = _module_scene()
    rend = GaussianRenderer().to(DEV)
    target = (torch.rand(1, 3, 3, 80, 80, generator=torch.Generator().manual_seed(2)) * 2 - 1).to(DEV)
    res = []
    for fn in (scenes.render_loss, scenes.render_loss_fused):
        p = sc.params.to(DEV).requires_grad_(True)
        im, de, ma, _, _ = rend(sc.cameras_out.to(DEV), conv(p, sc.cameras_in.to(DEV)), h=80, w=80)
        loss = fn(im, de, ma, target) * im.numel()  # O(1) cotangents per pixel
        loss.backward()
        res.append((loss.detach(), p.grad.clone()))
    assert torch.allclose(res[0][0], res[1][0], rtol=1e-6)
    assert float(res[0][1].abs().max()) > 1.0
    assert grad_close(res[1][1].cpu().numpy(), res[0][1].cpu().numpy(), max_share=ORDER_SHARE, what="fused loss glue")


# --------------------------------------------------------------------------- densification statistics
@pytest.mark.parametrize("fused", [True, False])
def test_densification_stats_match_the_refiner_loop(fused):
    """GaussianRenderer.update_densification_stats vs the refiner's per-view loop restated literally
    (sdspp_refiner.py:275-280 + gs_utils.py:247-252) on the renderer's own side lists, two refinement steps."""
    from prometheus_b200 import GaussianConverter, GaussianRenderer, scenes

    sc = scenes.make_scene("C1", seed=41, views_in=3, size_in=48, views_out=4, size_out=64)
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    rend = GaussianRenderer().to(DEV)
    rend.fused = fused
    N = sc.n_gaussians
    ours = [torch.zeros(N, device=DEV), torch.zeros(N, device=DEV), torch.zeros(N, 1, device=DEV), torch.zeros(N, 1, device=DEV)]
    ref = [t.clone() for t in ours]
    for step in range(2):
        params = (sc.params + 0.01 * step).to(DEV).requires_grad_(True)
        images, depths, masks, _, _ = rend(sc.cameras_out.to(DEV), conv(params, sc.cameras_in.to(DEV)), bg_color="random", h=64, w=64)
        (images.square().sum() + masks.sum() + depths.sum()).backward()
        rend.update_densification_stats(*ours)
        is_visible, max_radii2D, grad_accum, denom = ref
        for radii, vp in zip(rend.radii, rend.viewspace_points):          # the reference's loop, literally
            vis = radii > 0
            is_visible[vis] = 1
            max_radii2D[vis] = torch.max(max_radii2D[vis], radii[vis])
            grad_accum[vis] += torch.norm(vp.grad[0, vis, :2], dim=-1, keepdim=True)
            denom[vis] += 1
    assert torch.equal(ours[0], ref[0]) and torch.equal(ours[1], ref[1]) and torch.equal(ours[3], ref[3])
    assert float(ref[3].max()) == 8.0 and float(ref[2].sum()) > 0
    assert torch.allclose(ours[2], ref[2], rtol=1e-6, atol=1e-12)


# --------------------------------------------------------------------------- .splat export
@pytest.mark.parametrize("tag,thr", [("all", 0.0), ("pruned", 0.3)])
def test_splat_export_vs_reference_golden(golden_dir, tag, thr, tmp_path):
    """Device-side .splat packing against the bytes the reference's export_ply_for_gaussians wrote (pinned golden)."""
    import os

    from prometheus_b200 import export_ply_for_gaussians
    from tests.test_oracle import check_splat_against_reference

    g = np.load(os.path.join(golden_dir, "splat.npz"))
    gauss = tuple(dev(g[k]) for k in ("xyz", "features", "opacity", "scales", "rotations"))
    ret = export_ply_for_gaussians(str(tmp_path / "g"), gauss, opacity_threshold=thr)
    data = np.frombuffer(open(str(tmp_path / "g.splat"), "rb").read(), np.uint8)
    n = int((g["opacity"][0, :, 0] >= thr).sum())
    check_splat_against_reference(data, g["splat_" + tag], n)
    assert np.array_equal(ret["pts"], g["ret_%s_pts" % tag])
    assert np.allclose(ret["color"], g["ret_%s_color" % tag], atol=1e-6)


def test_splat_pack_large_is_sorted_and_complete(oracle):
    """2^20 Gaussians: keys non-increasing in volume x opacity, every Gaussian present exactly once."""
    from oracle import trajectory_oracle as T
    from prometheus_b200 import ops

    gen = torch.Generator().manual_seed(8)
    n = 1 << 20
    xyz = torch.randn(n, 3, generator=gen)
    rgb = torch.rand(n, 3, generator=gen)
    opa = torch.sigmoid(torch.randn(n, generator=gen) * 2)
    scl = 0.001 + 0.3 * torch.rand(n, 3, generator=gen) ** 2
    rot = torch.randn(n, 4, generator=gen)
    out = ops.splat_pack(*(t.to(DEV) for t in (xyz, rgb, opa, scl, rot))).cpu().numpy()
    pos = out[:, :12].copy().view(np.float32)
    # every Gaussian exactly once: match the stream back to its source through the positions
    X = xyz.numpy()
    i_src, i_out = np.lexsort(X.T[::-1]), np.lexsort(pos.T[::-1])
    assert np.array_equal(X[i_src], pos[i_out])
    src = np.empty(n, np.int64)
    src[i_out] = i_src
    # ... in the reference's order: volume x opacity (io.py:76-79), evaluated here in fp64
    vol = (scl.numpy().astype(np.float64) + 1e-8).prod(axis=1) * opa.numpy().astype(np.float64)
    key = vol[src]
    assert np.all(key[1:] <= key[:-1] * (1 + 1e-5))


def test_tile_order_is_a_heaviest_first_permutation():
    """pm_tile_scan's launch order: a permutation of the tiles, counts non-increasing up to the 1024-class quantisation."""
    from prometheus_b200 import ops

    gen = torch.Generator().manual_seed(2)
    counts = (torch.rand(3, 20, 17, generator=gen) ** 3 * 9000).to(torch.int32).to(DEV)
    off, stats, order = ops.tile_scan(counts, want_order=True)
    o = order.cpu().numpy()
    c = counts.reshape(-1).cpu().numpy()
    assert np.array_equal(np.sort(o), np.arange(c.size))
    shift = max(0, int(c.max()).bit_length() - 10)
    cls = c[o] >> shift
    assert np.all(np.diff(cls) <= 0)
    assert int(stats[0]) == int(c.sum()) and int(stats[1]) == int(c.max())


# --------------------------------------------------------------------------- edge cases of the module fast path
def test_fast_path_nothing_visible_renders_background():
    """All Gaussians behind the cameras: zero intersections, images = 2 * background - 1, masks = 0, depth = 0."""
    from prometheus_b200 import GaussianRenderer

    n = 500
    g = torch.Generator().manual_seed(1)
    xyz = torch.randn(1, n, 3, generator=g)
    xyz[..., 2] = 5.0  # OpenGL cameras look down -z: +z is behind them
    gauss = (xyz, torch.rand(1, n, 1, 3, generator=g), torch.rand(1, n, 1, generator=g), torch.full((1, n, 3), 0.05),
             torch.tensor([1.0, 0, 0, 0]).repeat(1, n, 1))
    cams = torch.zeros(1, 2, 18)
    cams[..., :12] = torch.eye(4)[:3].reshape(-1)
    cams[..., 12:] = torch.tensor([60.0, 60.0, 35.0, 25.0, 50.0, 70.0])
    rend = GaussianRenderer(background=[0.2, 0.5, 0.9]).to(DEV)
    for _ in range(2):  # second call goes through the speculative-capacity path with a zero count
        images, depths, masks, _, _ = rend(cams.to(DEV), tuple(t.to(DEV) for t in gauss), h=50, w=70)
    want = 2 * torch.tensor([0.2, 0.5, 0.9], device=DEV).view(1, 1, 3, 1, 1) - 1
    assert images.shape == (1, 2, 3, 50, 70) and torch.allclose(images, want.expand_as(images), atol=1e-7)
    assert float(masks.abs().max()) == 0.0 and float(depths.abs().max()) == 0.0
    assert all(int((r > 0).sum()) == 0 for r in rend.radii)


@pytest.mark.parametrize("C,h,w", [(1, 50, 70), (33, 37, 41)])
def test_fast_path_odd_sizes_and_many_cameras(C, h, w):
    """Image sizes that are not multiples of 16 and more cameras than one projection launch handles (32):
    fused packed path == separate stages, bit for bit."""
    from prometheus_b200 import GaussianConverter, GaussianRenderer, scenes

    sc = scenes.make_scene("C1", seed=51, views_in=2, size_in=40, views_out=C, size_out=64)
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    outs = []
    with torch.no_grad():
        g = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
        for fused in (True, False):
            rend = GaussianRenderer().to(DEV)
            rend.fused, rend.view_chunk = fused, 40
            outs.append(rend(sc.cameras_out.to(DEV), g, bg_color=(0.3, 0.1, 0.7), h=h, w=w)[:3] + (torch.stack(rend.radii),))
    for a, b in zip(*outs):
        assert a.shape == b.shape and torch.equal(a, b)
    assert outs[0][0].shape == (1, C, 3, h, w) and float(outs[0][2].mean()) > 0.05


def test_scene_batched_forward_backward_equals_per_scene():
    """B scenes in one launch set (views_per_scene) vs the scene-by-scene loop: images / depths / masks / radii bit
    for bit, parameter gradients to summation order; 'random' backgrounds draw B*N colours either way."""
    from prometheus_b200 import GaussianConverter, GaussianRenderer, scenes

    sc = scenes.make_scene("C1", seed=61, views_in=4, size_in=32, views_out=3, size_out=56, scenes=3)   # G = 4096 per scene
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    outs = []
    for batched in (True, False):
        rend = GaussianRenderer().to(DEV)
        rend.scene_batch = batched
        params = sc.params.to(DEV).requires_grad_(True)
        images, depths, masks, _, _ = rend(sc.cameras_out.to(DEV), conv(params, sc.cameras_in.to(DEV)), bg_color=(0.2, 0.4, 0.9), h=56, w=56)
        (scenes.render_loss(images, depths, masks, torch.zeros_like(images)) * images.numel()).backward()
        outs.append((images.detach(), depths.detach(), masks.detach(), torch.stack(rend.radii), params.grad.clone(),
                     torch.cat([v.grad for v in rend.viewspace_points])))
    a, b = outs
    assert a[0].shape == (3, 3, 3, 56, 56)
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x, y)
    assert grad_close(a[4].cpu().numpy(), b[4].cpu().numpy(), max_share=ORDER_SHARE, what="scene batched params")
    assert grad_close(a[5].cpu().numpy(), b[5].cpu().numpy(), max_share=ORDER_SHARE, what="scene batched means2d")


def test_scene_batched_single_pass_binning_same_bits():
    """Scene batching + single-pass binning (second call of the shape): the records of scene b's cameras carry flatten
    ids of scene b's Gaussians; images bit for bit as the first (two-pass) call and as the unbatched renderer, also
    when the scene-batched region size is too small (the batch is then rendered scene by scene)."""
    from prometheus_b200 import GaussianConverter, GaussianRenderer, ops, scenes

    sc = scenes.make_scene("C1", seed=63, views_in=4, size_in=32, views_out=3, size_out=56, scenes=3)
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    cams = sc.cameras_out.to(DEV)
    with torch.no_grad():
        g = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
        ref = GaussianRenderer().to(DEV)
        ref.scene_batch = False
        ref.single_pass_binning = False
        want = ref(cams, g, h=56, w=56)[:3]
        rend = GaussianRenderer().to(DEV)
        first = rend(cams, g, h=56, w=56)[:3]
        key = [k for k in rend._tile_cap_hint if len(k) == 5][0]
        prof, ops.PROFILE = ops.PROFILE, []
        second = rend(cams, g, h=56, w=56)[:3]
        names = [n for n, _, _ in ops.PROFILE]
        ops.PROFILE = prof
        assert names.count("project_scatter_fwd") == 1 and "bin_scatter" not in names   # one launch set, single pass
        rend._tile_cap_hint[key] = 8
        third = rend(cams, g, h=56, w=56)[:3]
    for got in (first, second, third):
        for a, b in zip(got, want):
            assert torch.equal(a, b)


def test_scene_batched_bad_hint_falls_back_to_per_scene():
    """A too-small speculative capacity on the scene-batched launch set is caught and the batch re-rendered scene by
    scene: same images as the unbatched renderer."""
    from prometheus_b200 import GaussianConverter, GaussianRenderer, scenes

    sc = scenes.make_scene("C1", seed=62, views_in=4, size_in=32, views_out=2, size_out=48, scenes=2)
    conv = GaussianConverter("mvsplat", z_near=0.01, z_far=100.0, s_min=0.001, s_max=1.0).to(DEV)
    with torch.no_grad():
        g = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
        ref = GaussianRenderer().to(DEV)
        ref.scene_batch = False
        want = ref(sc.cameras_out.to(DEV), g, h=48, w=48)[:3]
        rend = GaussianRenderer().to(DEV)
        first = rend(sc.cameras_out.to(DEV), g, h=48, w=48)[:3]
        key = [k for k in rend._cap_hint if len(k) == 5][0]
        rend._cap_hint[key] = 128
        second = rend(sc.cameras_out.to(DEV), g, h=48, w=48)[:3]
        assert rend._cap_hint[key] > 128 and len(rend.radii) == 4
        third = rend(sc.cameras_out.to(DEV), g, h=48, w=48)[:3]
    for got in (first, second, third):
        for a, b in zip(got, want):
            assert torch.equal(a, b)


def test_randomised_fused_vs_separate_stages():
    """tools/stress_fused.py on 15 random configurations (volumes, thin shells, repeated depths, huge per-tile lists,
    shells + outliers): fused and packed fast paths == separate stages, bit for bit."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "stress_fused.py"), "15"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "mismatches: 0" in p.stdout


def test_peer_gather_on_two_gpus():
    """Fused unproject + all-gather over NVLink peer memory (parallel.PeerGather) == NCCL all-gather == single-GPU
    unproject, bit for bit. Needs two GPUs (skipped on the single-GPU test box; run with gpurun --gpus 2)."""
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29577", os.path.join(root, "tools", "p2p_gather_check.py"), "64"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "p2p gather ok" in p.stdout, p.stdout[-1500:] + p.stderr[-1500:]


def test_nvtx_ranges_do_not_change_results():
    """PM_NVTX=1 wraps every C-ABI call in an NVTX range (ops.NVTX); same outputs, launches still counted."""
    from prometheus_b200 import ops

    cams = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "golden", "cameras.npz"))["cameras"]).to(DEV).float()
    want = ops.camera_prep(cams, 96, 96)
    before, ops.NVTX = ops.LAUNCHES, True
    try:
        got = ops.camera_prep(cams, 96, 96)
    finally:
        ops.NVTX = False
    assert ops.LAUNCHES == before + 1
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_randomised_stress_of_the_fused_paths():
    """tools/stress_fused.py, 15 cases: random sizes (not multiples of 16), camera counts, anisotropic splats, thin depth
    shells, exactly repeated depths, everything on a few tiles — fused renderer vs separate stages (same bits), usable-tile
    single-pass binning (same pixels), the operator's lazy meta (same arrays), hit-map backward vs culling backward."""
    from tools import stress_fused

    assert stress_fused.run(15) == 0
    # adversarial splats: needles and discs with axis ratios up to 5000, image-sized splats, opacities around 1 / 255
    assert stress_fused.run(8, first=1000) == 0
