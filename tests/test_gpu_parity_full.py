"""Full-size parity on the BASELINE configurations themselves (run with -m gpu on the B200 box).

BASELINE.json north_star: "results must match the reference on the same inputs ... on every config". The small
scenes of test_gpu_parity.py cover each stage; this file runs the CUDA path at the sizes bench.py times — C1
(4 views x 256^2 -> 262 144 Gaussians -> 4 views 256^2), C2 (8 x 512^2 -> 2 097 152 Gaussians -> 8 views 512^2),
one scene of C3 (8 x 256^2 -> 524 288 Gaussians -> 8 views 256^2, forward + backward) and one 1024^2 frame of C5
(32 x 512^2 -> 8 388 608 Gaussians) — against oracle/gs_oracle.c on identical inputs (reference call site:
prometheus/modules/renderers/gaussians_renderer.py:413-440):

  * operator path (`prometheus_b200.rasterization`): radii, tiles_per_gauss, isect_ids, flatten_ids, isect_offsets and
    the bits of means2d / depths / conics EQUAL; render / alpha within 1e-5 abs + 1e-4 rel and last_ids equal on
    pixels the oracle does not flag fragile; fragile share < 1 %;
  * module fast path (fused sort-as-you-go renderer, what bench.py's headline times): bit-identical to it;
  * backward on identical forward state (the oracle's) against orc_composite_bwd (fp64 accumulation) and
    orc_project_bwd, element-wise at 1e-5 + 1e-4 |b| with the violating share printed and bounded.

This is where the multi-group walk, the level-2 sub-buckets, the sampled kmin / kmax, the hashed privatisation and the
speculative capacity of the fast path run at the list lengths (~5 800 per tile, 8 thin depth shells) they were built for.
"""
import numpy as np
import pytest
import torch

from tests.util import violations

pytestmark = pytest.mark.gpu

DEV = "cuda"
KEYS = ("means", "quats", "scales", "opacities", "colors")


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def decode(oracle, name, n_views_out=None, seed=1234, **over):
    """Scene of config `name` -> Gaussians from the CUDA unproject (checked against the oracle's at 1e-5 / 1e-4), device
    cameras from pm_camera_prep. Everything downstream gets these same arrays on both sides."""
    from prometheus_b200 import GaussianConverter, ops, scenes

    sc = scenes.make_scene(name, seed=seed, scenes=1, **over)
    conv = GaussianConverter("mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX).to(DEV)
    with torch.no_grad():
        g = conv(sc.params.to(DEV), sc.cameras_in.to(DEV))
    ref = oracle.unproject(sc.params.numpy(), sc.cameras_in.numpy(), mode="mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR,
                           s_min=scenes.S_MIN, s_max=scenes.S_MAX)
    for got, key in zip(g, ("xyz", "features", "opacity", "scales", "rotations")):
        share, worst = violations(got.cpu().numpy(), ref[key])
        assert share == 0.0, (name, key, share, worst)
    S = sc.size_out
    cams = sc.cameras_out[0] if n_views_out is None else sc.cameras_out[0, :n_views_out]
    vm, K = ops.camera_prep(cams.to(DEV), S, S)
    gauss = dict(means=g[0][0], quats=g[4][0], scales=g[3][0], opacities=g[2][0, :, 0], colors=g[1][0, :, 0])
    gauss = {k: v.contiguous() for k, v in gauss.items()}
    return sc, gauss, vm, K, S


def oracle_forward(oracle, gauss, vm, K, S, bg):
    return oracle.forward_pipeline(*[gauss[k].cpu().numpy() for k in KEYS], vm.cpu().numpy(), K.cpu().numpy(), S, S,
                                   backgrounds=bg, with_depth=True)


def check_operator_forward(name, gauss, vm, K, S, bg, st):
    """`rasterization(...)` (full sort + composite_fwd) against the oracle's forward state `st`."""
    import prometheus_b200 as pb

    C = vm.shape[0]
    with torch.no_grad():
        r, a, m = pb.rasterization(*[gauss[k] for k in KEYS], vm, K, S, S, backgrounds=dev(bg), render_mode="RGB+D", packed=False)
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert np.array_equal(m[k].cpu().numpy(), st[k]), (name, k)
    for k in ("means2d", "depths", "conics"):
        assert np.array_equal(m[k].cpu().numpy().view(np.uint32), st[k].view(np.uint32)), (name, k)
    keep = st["fragile"] == 0
    fragile_frac = 1.0 - keep.mean()
    assert fragile_frac < 0.01, (name, fragile_frac)
    rn, an = r.cpu().numpy(), a[..., 0].cpu().numpy()
    s_r, w_r = violations(rn[keep], st["render"][keep])
    s_a, w_a = violations(an[keep], st["alphas"][keep])
    print("%s operator fwd: %d Gaussians x %d views, %d intersections, fragile %.4f %%, worst ratio render %.3f alpha %.3f" % (
        name, gauss["means"].shape[0], C, len(st["isect_ids"]), 100 * fragile_frac, w_r, w_a))
    assert s_r == 0.0 and s_a == 0.0, (name, s_r, w_r, s_a, w_a)
    # a fragile pixel may flip one decision, never more than one Gaussian's worth of weight
    assert np.abs(an - st["alphas"]).max() < 5e-3
    return r, a, m


def check_fast_path(name, gauss, vm, K, S, bg, st, r_op, a_op):
    """Fused sort-as-you-go renderer (project + histogram, scan, scatter, tile_render_fwd): same bits as the operator path,
    last_ids equal to the oracle's on non-fragile pixels, sorted prefixes equal to the oracle's order (a subsequence of it
    when only the usable tiles are binned: check_usable_tiles)."""
    from prometheus_b200 import ops
    from prometheus_b200.rasterization import render_views_packed

    radii, means2d, depths, conics, _, counts = ops.project_fwd(gauss["means"], gauss["quats"], gauss["scales"], vm, K, S, S,
                                                                count_tiles=False, tile_histogram=True)
    offsets, stats, order = ops.tile_scan(counts, want_order=True)
    n, max_count = stats.tolist()
    assert n == len(st["isect_ids"]) and np.array_equal(offsets.cpu().numpy(), st["isect_offsets"])
    records = ops.bin_scatter(means2d, radii, depths, offsets, S, S, n)
    r, a, last, flat, sorted_counts, flag, trans = ops.tile_render_fwd(records, offsets, means2d, conics, gauss["opacities"],
                                                                        gauss["colors"], depths, dev(bg), S, S, tile_order=order)
    assert int(flag.item()) == 0
    assert torch.equal(r, r_op) and torch.equal(a, a_op[..., 0])
    keep = st["fragile"] == 0
    assert np.array_equal(last.cpu().numpy()[keep], st["last_ids"][keep])
    off = st["isect_offsets"].reshape(-1)
    cnt = np.diff(np.append(off, n))
    sc = sorted_counts.cpu().numpy().reshape(-1)
    assert (sc <= cnt).all() and (sc >= 0).all()
    # every tile's sorted prefix is the oracle's order (all tiles: one vectorised comparison)
    flat_np = flat.cpu().numpy()
    idx = np.repeat(off, sc) + (np.arange(sc.sum()) - np.repeat(np.cumsum(sc) - sc, sc))
    assert np.array_equal(flat_np[idx], st["flatten_ids"][idx])
    # every non-fragile pixel's last contributor lies inside its tile's sorted prefix
    C = vm.shape[0]
    tl = torch.from_numpy(st["last_ids"] * keep).reshape(C, S // 16, 16, S // 16, 16).amax(dim=(2, 4)).reshape(-1).numpy()
    has = cnt > 0
    assert np.all((tl - off + 1)[has] <= sc[has])
    i_need = int(np.maximum(tl - off + 1, 0)[has].sum())
    print("%s fast path: max tile %d, sorted %d of %d intersections (I_need %d)" % (name, max_count, sc.sum(), n, i_need))
    # the packed module epilogue = the same pixels in the module's layout
    with torch.no_grad():
        images, depth_out, masks, info = render_views_packed(*[gauss[k] for k in KEYS], vm, K, dev(bg), S, S)
    assert torch.equal(images, (2.0 * r[..., :3] - 1.0).permute(0, 3, 1, 2))
    assert torch.equal(masks[:, 0], a.clamp(0, 1))
    assert torch.equal(depth_out[:, 0], r[..., 3] / a.clamp(min=1e-10))
    # single-pass binning (what the module runs from the second call of a shape on): the projection kernel writes the
    # records into fixed-capacity tile regions — same counts, same images, same sorted prefixes, at size
    with torch.no_grad():
        tcap = (int(max_count * 1.25) + 65) & ~1
        im1, de1, ma1, info1 = render_views_packed(*[gauss[k] for k in KEYS], vm, K, dev(bg), S, S,
                                                   capacity=int(n * 1.3) + 4096, tile_capacity=tcap, usable_tiles=False)
    assert info1["tile_capacity"] == tcap and info1["stats"].tolist() == [n, max_count]
    assert torch.equal(info1["isect_offsets"], offsets) and int(info1["overflow_flag"].item()) == 0
    assert torch.equal(im1, images) and torch.equal(de1, depth_out) and torch.equal(ma1, masks)
    # (the group boundaries follow a strided SAMPLE of the records in their scatter order, so the length of a tile's
    # sorted prefix may differ between the two layouts; what is sorted is the oracle's order either way)
    sc1 = info1["sorted_counts"].cpu().numpy().reshape(-1)
    assert (sc1 <= cnt).all() and np.all((tl - off + 1)[has] <= sc1[has])
    idx1 = np.repeat(off, sc1) + (np.arange(sc1.sum()) - np.repeat(np.cumsum(sc1) - sc1, sc1))
    assert np.array_equal(info1["flatten_ids"].cpu().numpy()[idx1], st["flatten_ids"][idx1])
    check_usable_tiles(name, gauss, vm, K, S, bg, st, (images, depth_out, masks), n, tcap, off, cnt, tl, has)
    return trans


def check_usable_tiles(name, gauss, vm, K, S, bg, st, want, n, tcap, off, cnt, tl, has):
    """The module's default: single-pass binning of the USABLE tiles only (a splat is binned where it can reach
    alpha >= 1/255). Same images bit for bit; every tile's list is a SUBSET of the oracle's, its sorted prefix a
    SUBSEQUENCE of the oracle's order, and it still holds the last contributor of every non-fragile pixel."""
    from prometheus_b200.rasterization import render_views_packed

    with torch.no_grad():
        im, de, ma, info = render_views_packed(*[gauss[k] for k in KEYS], vm, K, dev(bg), S, S,
                                               capacity=int(n * 1.3) + 4096, tile_capacity=tcap)
    assert info["usable_tiles"] and int(info["overflow_flag"].item()) == 0
    for a, b in zip((im, de, ma), want):
        assert torch.equal(a, b), name
    n_u, max_u = info["stats"].tolist()
    off_u = info["isect_offsets"].reshape(-1).cpu().numpy().astype(np.int64)
    cnt_u = np.diff(np.append(off_u, n_u))
    assert n_u <= n and (cnt_u <= cnt).all() and (cnt_u >= 0).all()
    sc = info["sorted_counts"].cpu().numpy().reshape(-1).astype(np.int64)
    assert (sc <= cnt_u).all() and (sc >= 0).all()
    # position of every sorted fast-path entry in the oracle's list of the same tile (flatten ids are unique in a tile)
    pairs = int(vm.shape[0]) * int(gauss["means"].shape[0])
    tiles = torch.repeat_interleave(torch.arange(len(cnt), device=DEV), torch.from_numpy(cnt).to(DEV))
    okey, opos = torch.sort(tiles * pairs + torch.from_numpy(st["flatten_ids"].astype(np.int64)).to(DEV))
    del tiles
    idx = np.repeat(off_u, sc) + (np.arange(sc.sum()) - np.repeat(np.cumsum(sc) - sc, sc))
    ftile = torch.from_numpy(np.repeat(np.arange(len(sc)), sc)).to(DEV)
    fkey = ftile * pairs + info["flatten_ids"].long()[torch.from_numpy(idx).to(DEV)]
    at = torch.searchsorted(okey, fkey).clamp(max=len(okey) - 1)
    assert bool((okey[at] == fkey).all()), name                     # subset
    pos = opos[at]
    same_tile = ftile[1:] == ftile[:-1]
    assert bool((pos[1:] > pos[:-1])[same_tile].all()), name        # oracle's order
    # the last contributor of every non-fragile pixel (oracle position tl in its tile) is in the tile's sorted prefix
    need = np.nonzero(has & (tl - off >= 0) & (tl > 0))[0]  # (0 = the oracle's 'nothing blended')
    lkey = torch.from_numpy(need * pairs + st["flatten_ids"][tl[need]].astype(np.int64)).to(DEV)
    fsorted = torch.sort(fkey)[0]
    at = torch.searchsorted(fsorted, lkey).clamp(max=len(fsorted) - 1)
    assert bool((fsorted[at] == lkey).all()), name
    print("%s usable tiles: %d of %d intersections binned (%.1f %% fewer), max tile %d, sorted %d" % (
        name, n_u, n, 100.0 * (1 - n_u / n), max_u, sc.sum()))


def check_backward(oracle, name, gauss, vm, K, S, bg, st, max_share):
    """composite_bwd + project_bwd on the ORACLE's forward state, vs orc_composite_bwd (fp64 accumulation) and
    orc_project_bwd: element-wise 1e-5 + 1e-4 |b|."""
    from prometheus_b200 import ops

    rng = np.random.default_rng(17)
    keep = (st["fragile"] == 0).astype(np.float32)
    v_render = rng.standard_normal(st["render"].shape).astype(np.float32) * keep[..., None]
    v_alphas = rng.standard_normal(st["alphas"].shape).astype(np.float32) * keep
    g_np = {k: gauss[k].cpu().numpy() for k in KEYS}
    ref = oracle.composite_bwd(st["means2d"], st["conics"], g_np["opacities"], g_np["colors"], st["depths"], bg,
                               st["isect_offsets"], st["flatten_ids"], S, S, st["alphas"], st["last_ids"], v_render, v_alphas,
                               transmittance=st["transmittance"])
    out = ops.composite_bwd(dev(st["means2d"]), dev(st["conics"]), gauss["opacities"], gauss["colors"], dev(st["depths"]), dev(bg),
                            dev(st["isect_offsets"]), dev(st["flatten_ids"]), S, S, dev(st["alphas"]), dev(st["last_ids"]),
                            dev(v_render), dev(v_alphas), transmittance=dev(st["transmittance"]))
    worst_share = 0.0
    for got, key in zip(out, ("v_means2d", "v_conics", "v_colors", "v_opacities", "v_depths")):
        share, worst = violations(got.cpu().numpy(), ref[key])
        print("%s composite_bwd %-11s violating share %.3g, worst ratio %.3g, max |ref| %.3g" % (name, key, share, worst, np.abs(ref[key]).max()))
        worst_share = max(worst_share, share)
        assert share <= max_share, (name, key, share, worst)
    f32 = lambda k: ref[k].astype(np.float32)
    vm_np, K_np = vm.cpu().numpy(), K.cpu().numpy()
    pref = oracle.project_bwd(g_np["means"], g_np["quats"], g_np["scales"], vm_np, K_np, S, S, st["radii"], st["conics"],
                              f32("v_means2d"), f32("v_depths"), f32("v_conics"))
    pgot = ops.project_bwd(gauss["means"], gauss["quats"], gauss["scales"], vm, K, S, S, dev(st["radii"]), dev(st["conics"]),
                           dev(f32("v_means2d")), dev(f32("v_depths")), dev(f32("v_conics")))
    for got, key in zip(pgot, ("v_means", "v_quats", "v_scales")):
        share, worst = violations(got.cpu().numpy(), pref[key])
        print("%s project_bwd   %-11s violating share %.3g, worst ratio %.3g, max |ref| %.3g" % (name, key, share, worst, np.abs(pref[key]).max()))
        assert share <= max_share, (name, key, share, worst)
    return worst_share


# share of gradient entries allowed outside the bar at full size: fp32 atomics in arbitrary order (the composite
# gradients are sums of up to thousands of signed terms) and fp32 VJP chains vs the oracle's fp64 accumulation
FULL_GRAD_SHARE = 1e-4  # measured on the B200: <= 6.4e-6 (v_conics, C1); everything else 0 .. 3e-6


def test_c1_all_views_forward_and_backward(oracle):
    """BASELINE configs[0]: 4 x 256^2 -> 262 144 Gaussians -> 4 views 256^2, everything at once."""
    sc, gauss, vm, K, S = decode(oracle, "C1")
    bg = np.random.default_rng(3).random((vm.shape[0], 3)).astype(np.float32)
    st = oracle_forward(oracle, gauss, vm, K, S, bg)
    r, a, m = check_operator_forward("C1", gauss, vm, K, S, bg, st)
    check_fast_path("C1", gauss, vm, K, S, bg, st, r, a)
    check_backward(oracle, "C1", gauss, vm, K, S, bg, st, FULL_GRAD_SHARE)


def test_c2_all_views_forward_one_view_backward(oracle):
    """BASELINE configs[1] (the bench headline): 8 x 512^2 -> 2 097 152 Gaussians -> 8 views 512^2 forward (operator
    path and fast path), and the backward of one view."""
    sc, gauss, vm, K, S = decode(oracle, "C2")
    bg = np.ones((vm.shape[0], 3), np.float32)
    st = oracle_forward(oracle, gauss, vm, K, S, bg)
    r, a, m = check_operator_forward("C2", gauss, vm, K, S, bg, st)
    check_fast_path("C2", gauss, vm, K, S, bg, st, r, a)
    del st, r, a, m
    st1 = oracle_forward(oracle, gauss, vm[:1], K[:1], S, bg[:1])
    check_backward(oracle, "C2[view 0]", gauss, vm[:1], K[:1], S, bg[:1], st1, FULL_GRAD_SHARE)


def test_c3_one_scene_forward_and_backward(oracle):
    """BASELINE configs[2], one of its 8 scenes: 8 x 256^2 -> 524 288 Gaussians -> 8 views 256^2, per-view random
    backgrounds (training.render_bg_color = 'random'), forward + backward."""
    sc, gauss, vm, K, S = decode(oracle, "C3")
    bg = np.random.default_rng(5).random((vm.shape[0], 3)).astype(np.float32)
    st = oracle_forward(oracle, gauss, vm, K, S, bg)
    r, a, m = check_operator_forward("C3", gauss, vm, K, S, bg, st)
    check_fast_path("C3", gauss, vm, K, S, bg, st, r, a)
    check_backward(oracle, "C3", gauss, vm, K, S, bg, st, FULL_GRAD_SHARE)


def test_c5_one_frame(oracle):
    """BASELINE configs[4]: 32 x 512^2 -> 8 388 608 Gaussians, one 1024^2 frame of the trajectory."""
    sc, gauss, vm, K, S = decode(oracle, "C5", n_views_out=1, seed=4321)
    assert gauss["means"].shape[0] == 8388608 and S == 1024
    bg = np.ones((1, 3), np.float32)
    st = oracle_forward(oracle, gauss, vm, K, S, bg)
    r, a, m = check_operator_forward("C5[frame 0]", gauss, vm, K, S, bg, st)
    check_fast_path("C5[frame 0]", gauss, vm, K, S, bg, st, r, a)


def test_c2_module_training_step_vs_oracle_autograd(oracle):
    """The whole module path at size — GaussianConverter -> GaussianRenderer (fused fast path, packed epilogue) ->
    render loss -> backward to the decoder maps — against the oracle operator's autograd on one C2 view: forward within
    1e-5 / 1e-4 on non-fragile pixels, Gaussian gradients element-wise with the share bounded by OPERATOR_GRAD_SHARE."""
    from prometheus_b200 import GaussianRenderer, scenes

    sc, gauss, vm, K, S = decode(oracle, "C2", n_views_out=1)
    leaves = {k: v.clone().requires_grad_(True) for k, v in gauss.items()}
    g = (leaves["means"][None], leaves["colors"][None, :, None, :], leaves["opacities"][None, :, None], leaves["scales"][None],
         leaves["quats"][None])
    rend = GaussianRenderer(background=[1, 1, 1]).to(DEV)
    images, depths, masks, _, _ = rend(sc.cameras_out[:, :1].to(DEV), g, bg_color=None, h=S, w=S)
    cpu = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in gauss.items()}
    r_ref, a_ref, m_ref = oracle.rasterization(*[cpu[k] for k in KEYS], vm.cpu(), K.cpu(), S, S, backgrounds=torch.ones(1, 3),
                                               render_mode="RGB+ED", packed=False)
    keep = ~m_ref["fragile"].bool()
    assert keep.float().mean() > 0.99
    img_ref = (r_ref[..., :3] * 2 - 1).permute(0, 3, 1, 2)
    got = images[0].detach().cpu()
    k3 = keep[:, None].expand_as(got)
    share, worst = violations((got[k3].numpy() + 1) / 2, (img_ref.detach()[k3].numpy() + 1) / 2)
    assert share == 0.0, (share, worst)
    share, worst = violations(masks[0, :, 0].detach().cpu()[keep].numpy(), a_ref[..., 0].detach()[keep].numpy())
    assert share == 0.0, (share, worst)
    gen = torch.Generator().manual_seed(3)
    w_img = torch.randn(images.shape, generator=gen) * keep[None, :, None].float()
    w_dep = torch.randn(depths.shape, generator=gen) * keep[None, :, None].float()
    w_msk = torch.randn(masks.shape, generator=gen) * keep[None, :, None].float()
    ((images * w_img.to(DEV)).sum() + (depths * w_dep.to(DEV)).sum() + (masks * w_msk.to(DEV)).sum()).backward()
    dep_ref = r_ref[..., 3:].permute(0, 3, 1, 2)
    msk_ref = a_ref.permute(0, 3, 1, 2).clamp(0, 1)
    ((img_ref * w_img[0]).sum() + (dep_ref * w_dep[0]).sum() + (msk_ref * w_msk[0]).sum()).backward()
    for k in KEYS:
        share, worst = violations(leaves[k].grad.cpu().numpy(), cpu[k].grad.numpy())
        print("C2 module step grad %-9s violating share %.3g, worst ratio %.3g" % (k, share, worst))
        assert share <= 5e-4, (k, share, worst)  # measured: means 8e-5, scales 1.3e-4, others <= 1e-7
