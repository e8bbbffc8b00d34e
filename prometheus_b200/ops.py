"""Stage-level ops: torch tensors in, torch tensors out, computed by the CUDA
kernels behind the C-ABI (include/prometheus_gs.h). No autograd here — see
rasterization.py / renderer.py for the differentiable wrappers.

Every function requires CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

TILE = 16
MODES = {"mvsplat": 0, "gslrm": 1, "director3d": 2}

# ---- instrumentation (used by bench.py): how many of OUR kernels were launched, and, when
# `PROFILE` is a list, a (name, start_event, end_event) record per C-ABI call, taken with CUDA
# events on the launching stream.
LAUNCHES = 0
PROFILE = None
# PM_NVTX=1: every C-ABI call is wrapped in an NVTX range named after the stage ("pm:project_scatter_fwd", ...), so a
# timeline tool shows the path's stages inside the caller's own ranges (the reference has none)
NVTX = __import__("os").environ.get("PM_NVTX", "0") not in ("", "0")


def _instrument(name, n_launches):
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapper(*a, **k):
            global LAUNCHES
            n = n_launches(*a, **k) if callable(n_launches) else n_launches
            LAUNCHES += n
            if NVTX:
                torch.cuda.nvtx.range_push("pm:" + (name(*a, **k) if callable(name) else name))
                try:
                    return fn(*a, **k)
                finally:
                    torch.cuda.nvtx.range_pop()
            if PROFILE is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            PROFILE.append((name(*a, **k) if callable(name) else name, e0, e1))
            return out

        return wrapper

    return deco


def _sort_launches(keys, vals, end_bit, *a, **k):
    return 2 + (end_bit + 7) // 8 if keys.numel() > 0 else 0


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (prometheus_b200 has no CPU path)")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


PARAM_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}  # enum pm_dtype


def _req_params(t: torch.Tensor, name: str = "params") -> torch.Tensor:
    """Decoder maps: fp32, or fp16 / bf16 as they come out of an autocast decoder (read natively)."""
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (prometheus_b200 has no CPU path)")
    if t.dtype not in PARAM_DTYPES:
        t = t.float()
    return t.contiguous()


def _opt(t, dtype, name):
    return None if t is None else _req(t, dtype, name)


def tile_grid(width: int, height: int) -> Tuple[int, int]:
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def key_bits(tile_width: int, tile_height: int, C: int) -> Tuple[int, int]:
    """floor(log2(n))+1 bit widths gsplat packs tile and camera ids with."""
    return int(tile_width * tile_height).bit_length(), int(C).bit_length()


# ------------------------------------------------------------------ stage 1
@_instrument("unproject_fwd", 1)
def unproject_fwd(params, cameras, muls, mode, z_near, z_far, s_min, s_max, out=None):
    """params [V,14,h,w], cameras [V,18] -> (xyz [G,3], rgb [G,3], opacity [G], scales [G,3], quats [G,4]).
    `out` may hold five preallocated tensors (e.g. slots of an all-gather buffer)."""
    L = _lib.load()
    params = _req_params(params)
    cameras = _req(cameras, torch.float32, "cameras")
    muls = _req(muls, torch.float32, "muls").reshape(-1)
    V, Cc, h, w = params.shape
    assert Cc == 14 and cameras.shape == (V, 18) and muls.numel() == 14
    G = V * h * w
    dev = params.device
    if out is None:
        out = (torch.empty(G, 3, device=dev), torch.empty(G, 3, device=dev), torch.empty(G, device=dev),
               torch.empty(G, 3, device=dev), torch.empty(G, 4, device=dev))
    xyz, rgb, opacity, scales, quats = out
    for t, n in zip(out, (3 * G, 3 * G, G, 3 * G, 4 * G)):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    _lib.check(L.pm_unproject_fwd_dt(_ptr(params), PARAM_DTYPES[params.dtype], _ptr(cameras), _ptr(muls), MODES[mode],
                                     z_near, z_far, s_min, s_max, V, h, w, _ptr(xyz), _ptr(rgb), _ptr(opacity),
                                     _ptr(scales), _ptr(quats), _stream()), "pm_unproject_fwd_dt")
    return xyz, rgb, opacity, scales, quats


@_instrument("unproject_fwd", 1)
def unproject_fwd_multi(params, cameras, muls, mode, z_near, z_far, s_min, s_max, n_dst, dst_ptrs):
    """Unproject this rank's views into n_dst destination sets (pm_unproject_fwd_multi): `dst_ptrs` is a ctypes
    array of 5 * n_dst device pointers {xyz, rgb, opacity, scales, quats} — see parallel.PeerGather."""
    L = _lib.load()
    params = _req_params(params)
    cameras = _req(cameras, torch.float32, "cameras")
    muls = _req(muls, torch.float32, "muls").reshape(-1)
    V, Cc, h, w = params.shape
    assert Cc == 14 and cameras.shape == (V, 18) and muls.numel() == 14
    _lib.check(L.pm_unproject_fwd_multi(_ptr(params), PARAM_DTYPES[params.dtype], _ptr(cameras), _ptr(muls), MODES[mode], z_near, z_far, s_min, s_max,
                                        V, h, w, int(n_dst), dst_ptrs, _stream()), "pm_unproject_fwd_multi")


@_instrument("unproject_bwd", 1)
def unproject_bwd(params, cameras, muls, mode, z_near, z_far, s_min, s_max, v_xyz, v_rgb, v_opacity,
                  v_scales, v_quats):
    """v_params in the maps' own dtype (fp32 gradient rounded to nearest for fp16 / bf16 maps)."""
    L = _lib.load()
    params = _req_params(params)
    cameras = _req(cameras, torch.float32, "cameras")
    muls = _req(muls, torch.float32, "muls").reshape(-1)
    V, _, h, w = params.shape
    vs = [_opt(t, torch.float32, "grad") for t in (v_xyz, v_rgb, v_opacity, v_scales, v_quats)]
    v_params = torch.empty_like(params)
    _lib.check(L.pm_unproject_bwd_dt(_ptr(params), PARAM_DTYPES[params.dtype], _ptr(cameras), _ptr(muls), MODES[mode],
                                     z_near, z_far, s_min, s_max, V, h, w, *[_ptr(t) for t in vs], _ptr(v_params),
                                     _stream()), "pm_unproject_bwd_dt")
    return v_params


@_instrument("camera_prep", 1)
def camera_prep(cameras, h: int, w: int):
    """cameras [n,18] -> viewmats [n,4,4], Ks [n,3,3], all on the device, no host sync."""
    L = _lib.load()
    cameras = _req(cameras, torch.float32, "cameras").reshape(-1, 18)
    n = cameras.shape[0]
    viewmats = torch.empty(n, 4, 4, device=cameras.device)
    Ks = torch.empty(n, 3, 3, device=cameras.device)
    _lib.check(L.pm_camera_prep(_ptr(cameras), n, h, w, _ptr(viewmats), _ptr(Ks), _stream()), "pm_camera_prep")
    return viewmats, Ks


# ------------------------------------------------------------------ stage 2
@_instrument(lambda *a, **k: "project_scatter_fwd" if k.get("tile_capacity") else "project_fwd",
             lambda means, quats, scales, viewmats, *a, **k: (viewmats.shape[0] + 31) // 32)
def project_fwd(means, quats, scales, viewmats, Ks, width, height, eps2d=0.3, near_plane=0.01,
                far_plane=1e10, radius_clip=0.0, count_tiles=True, tile_histogram=False, views_per_scene=0,
                tile_capacity=0, opacities=None):
    """-> radii, means2d, depths, conics, tiles_per_gauss (None unless count_tiles)
    [, tile_counts [C,th,tw] int32 when tile_histogram: the fused pm_bin_count]
    [, records [C*th*tw, tile_capacity] int64 when tile_capacity > 0: single-pass binning, pm_project_scatter_fwd].
    opacities [n_scenes * N] (single-pass binning only): bin only the tiles a splat can be seen in (alpha >= 1/255
    somewhere) instead of gsplat's whole AABB — the counts and records are then a subset of gsplat's intersections.
    views_per_scene = V > 0 (tile_histogram only): scene batching — means / quats / scales hold C / V blocks of
    N Gaussians and camera c sees block c / V; the pair outputs are [C, N]."""
    L = _lib.load()
    means, quats, scales = (_req(t, torch.float32, n) for t, n in ((means, "means"), (quats, "quats"), (scales, "scales")))
    viewmats, Ks = _req(viewmats, torch.float32, "viewmats"), _req(Ks, torch.float32, "Ks")
    C = viewmats.shape[0]
    n_scenes = C // views_per_scene if views_per_scene else 1
    assert means.shape[0] % n_scenes == 0 and (not views_per_scene or tile_histogram)
    N = means.shape[0] // n_scenes
    dev = means.device
    radii = torch.empty(C, N, dtype=torch.int32, device=dev)
    means2d = torch.empty(C, N, 2, device=dev)
    depths = torch.empty(C, N, device=dev)
    conics = torch.empty(C, N, 3, device=dev)
    tpg = torch.empty(C, N, dtype=torch.int32, device=dev) if count_tiles else None
    if tile_histogram:
        tw, th = tile_grid(width, height)
        counts = torch.zeros(C, th, tw, dtype=torch.int32, device=dev)
        if tile_capacity:
            assert not count_tiles
            if opacities is not None:
                opacities = _req(opacities, torch.float32, "opacities")
                assert opacities.numel() == means.shape[0]
            records = torch.empty(C * th * tw, int(tile_capacity), dtype=torch.int64, device=dev)
            _lib.check(L.pm_project_scatter_fwd(_ptr(means), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks), N, C,
                                                int(views_per_scene), width, height, TILE, eps2d, near_plane, far_plane,
                                                radius_clip, _ptr(radii), _ptr(means2d), _ptr(depths), _ptr(conics),
                                                _ptr(counts), _ptr(records), int(tile_capacity), _ptr(opacities),
                                                _stream()),
                       "pm_project_scatter_fwd")
            return radii, means2d, depths, conics, None, counts, records
        if opacities is not None:  # histogram of the USABLE tiles only (what a scatter call with `opacities` will bin)
            assert not count_tiles
            opacities = _req(opacities, torch.float32, "opacities")
            assert opacities.numel() == means.shape[0]
            _lib.check(L.pm_project_scatter_fwd(_ptr(means), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks), N, C,
                                                int(views_per_scene), width, height, TILE, eps2d, near_plane, far_plane,
                                                radius_clip, _ptr(radii), _ptr(means2d), _ptr(depths), _ptr(conics),
                                                _ptr(counts), None, 0, _ptr(opacities), _stream()),
                       "pm_project_scatter_fwd")
            return radii, means2d, depths, conics, None, counts
        _lib.check(L.pm_project_count_fwd(_ptr(means), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks), N, C,
                                          int(views_per_scene), width, height, TILE, eps2d, near_plane, far_plane, radius_clip, _ptr(radii),
                                          _ptr(means2d), _ptr(depths), _ptr(conics), _ptr(tpg), _ptr(counts), _stream()),
                   "pm_project_count_fwd")
        return radii, means2d, depths, conics, tpg, counts
    _lib.check(L.pm_project_fwd(_ptr(means), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks), N, C, width,
                                height, TILE, eps2d, near_plane, far_plane, radius_clip, _ptr(radii),
                                _ptr(means2d), _ptr(depths), _ptr(conics), _ptr(tpg), _stream()),
               "pm_project_fwd")
    return radii, means2d, depths, conics, tpg


@_instrument("project_bwd", 1)
def project_bwd(means, quats, scales, viewmats, Ks, width, height, radii, conics, v_means2d, v_depths,
                v_conics, eps2d=0.3, views_per_scene=0, touched=None):
    """touched [C,N] uint8: the pair map composite_bwd* marked for exactly these cotangents (pairs with a zero byte
    are skipped unread); None = look at the cotangents themselves."""
    L = _lib.load()
    means, quats, scales = (_req(t, torch.float32, "in") for t in (means, quats, scales))
    viewmats, Ks = _req(viewmats, torch.float32, "viewmats"), _req(Ks, torch.float32, "Ks")
    C = viewmats.shape[0]
    N = means.shape[0] // (C // views_per_scene if views_per_scene else 1)
    v_means2d, v_depths, v_conics = (_opt(t, torch.float32, "grad") for t in (v_means2d, v_depths, v_conics))
    v_means, v_quats, v_scales = torch.empty_like(means), torch.empty_like(quats), torch.empty_like(scales)
    _lib.check(L.pm_project_bwd(_ptr(means), _ptr(quats), _ptr(scales), _ptr(viewmats), _ptr(Ks), N, C,
                                int(views_per_scene), width, height, eps2d, _ptr(radii), _ptr(conics), _ptr(v_means2d), _ptr(v_depths),
                                _ptr(v_conics), _ptr(v_means), _ptr(v_quats), _ptr(v_scales), _ptr(touched), _stream()),
               "pm_project_bwd")
    return v_means, v_quats, v_scales


# ------------------------------------------------------------------ stage 3
@_instrument("scan_tiles", 3)
def scan_tiles(tiles_per_gauss):
    """Inclusive prefix sum -> (cum int32 [P], n_isects device int32 [1])."""
    L = _lib.load()
    tpg = _req(tiles_per_gauss, torch.int32, "tiles_per_gauss").reshape(-1)
    P = tpg.numel()
    dev = tpg.device
    cum = torch.empty(P, dtype=torch.int32, device=dev)
    n_out = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = int(L.pm_scan_workspace_bytes(P))
    ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=dev)
    _lib.check(L.pm_scan_tiles(_ptr(tpg), P, _ptr(cum), _ptr(n_out), _ptr(ws), ws_bytes, _stream()), "pm_scan_tiles")
    return cum, n_out


@_instrument("bin_emit", 1)
def bin_emit(means2d, radii, depths, cum, width, height, capacity: int):
    L = _lib.load()
    C, N = radii.shape
    tw, th = tile_grid(width, height)
    dev = radii.device
    ids = torch.empty(capacity, dtype=torch.int64, device=dev)
    flat = torch.empty(capacity, dtype=torch.int32, device=dev)
    _lib.check(L.pm_bin_emit(_ptr(means2d), _ptr(radii), _ptr(depths), _ptr(cum), N, C, TILE, tw, th,
                             _ptr(ids), _ptr(flat), capacity, _stream()), "pm_bin_emit")
    return ids, flat


@_instrument("bin_count", 1)
def bin_count(means2d, radii, width, height):
    """Per-(camera,tile) intersection histogram [C,th,tw] int32."""
    L = _lib.load()
    C, N = radii.shape
    tw, th = tile_grid(width, height)
    counts = torch.zeros(C, th, tw, dtype=torch.int32, device=radii.device)
    _lib.check(L.pm_bin_count(_ptr(means2d), _ptr(radii), N, C, TILE, tw, th, _ptr(counts), _stream()), "pm_bin_count")
    return counts


@_instrument("tile_scan", 1)
def tile_scan(tile_counts, want_order=False):
    """Exclusive scan of the per-tile histogram -> (isect_offsets like tile_counts, stats int32[2] =
    {n_isects, max tile count}, both on the device) [, tile_order int32[T]: tile indices heaviest first]."""
    L = _lib.load()
    assert tile_counts.dtype == torch.int32 and tile_counts.is_contiguous() and tile_counts.is_cuda
    offsets = torch.empty_like(tile_counts)
    stats = torch.empty(2, dtype=torch.int32, device=tile_counts.device)
    order = torch.empty(tile_counts.numel(), dtype=torch.int32, device=tile_counts.device) if want_order else None
    _lib.check(L.pm_tile_scan(_ptr(tile_counts), tile_counts.numel(), _ptr(offsets), _ptr(stats), _ptr(order), _stream()),
               "pm_tile_scan")
    return (offsets, stats, order) if want_order else (offsets, stats)


@_instrument("bin_scatter", 1)
def bin_scatter(means2d, radii, depths, offsets, width, height, capacity: int, conics=None, opacities=None,
                views_per_scene=0):
    """Scatter intersections into their tile segments as packed {depth bits, flatten id} records.
    conics + opacities: only the tiles a splat can be seen in (`offsets` must then come from the histogram of
    project_fwd(tile_histogram=True, opacities=...))."""
    L = _lib.load()
    C, N = radii.shape
    tw, th = tile_grid(width, height)
    cursors = torch.zeros_like(offsets)
    records = torch.empty(capacity, dtype=torch.int64, device=radii.device)
    if opacities is not None:
        conics, opacities = _req(conics, torch.float32, "conics"), _req(opacities, torch.float32, "opacities")
        assert opacities.numel() == N * (C // views_per_scene if views_per_scene else 1)
    else:
        conics = None
    _lib.check(L.pm_bin_scatter(_ptr(means2d), _ptr(radii), _ptr(depths), _ptr(offsets), _ptr(cursors), N, C, TILE,
                                tw, th, _ptr(records), capacity, _ptr(conics), _ptr(opacities), int(views_per_scene),
                                _stream()), "pm_bin_scatter")
    return records


def tile_sort_capacity() -> int:
    return int(_lib.load().pm_tile_sort_capacity())


@_instrument("tile_sort", lambda *a, **k: 2 if (k.get("max_tile_count", -1) < 0 or k.get("max_tile_count", -1) > 8192) else 1)
def tile_sort(records, offsets, C: int, width: int, height: int, d_n=None, max_tile_count: int = -1):
    """Per-tile shared-memory depth sort -> (isect_ids int64 [n], flatten_ids int32 [n], overflow flag int32[1]).
    Tiles above the one-pass capacity are split by depth through a scratch buffer first."""
    L = _lib.load()
    tw, th = tile_grid(width, height)
    n = records.numel()
    ids = torch.empty(n, dtype=torch.int64, device=records.device)
    flat = torch.empty(n, dtype=torch.int32, device=records.device)
    flag = torch.zeros(1, dtype=torch.int32, device=records.device)
    need_big = max_tile_count < 0 or max_tile_count > tile_sort_capacity()
    scratch = torch.empty_like(records) if need_big else None
    _lib.check(L.pm_tile_sort(_ptr(records), _ptr(scratch), _ptr(offsets), n, _ptr(d_n), C, tw, th, int(max_tile_count),
                              _ptr(ids), _ptr(flat), _ptr(flag), _stream()), "pm_tile_sort")
    return ids, flat, flag


@_instrument("sort_pairs", _sort_launches)
def sort_pairs(keys, vals, end_bit: int, d_n=None, return_workspace=False):
    """Stable radix sort of (int64 key, int32 value) pairs on bits [0, end_bit).
    `keys`/`vals` are clobbered (ping-pong buffer 0). Returns the sorted pair."""
    L = _lib.load()
    assert keys.dtype == torch.int64 and vals.dtype == torch.int32 and keys.is_cuda and vals.is_cuda
    assert keys.is_contiguous() and vals.is_contiguous() and keys.numel() == vals.numel()
    n = keys.numel()
    keys1, vals1 = torch.empty_like(keys), torch.empty_like(vals)
    ws_bytes = int(L.pm_sort_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
    sel = ctypes.c_int(0)
    _lib.check(L.pm_sort_pairs(_ptr(keys), _ptr(vals), _ptr(keys1), _ptr(vals1), n, _ptr(d_n), end_bit,
                               _ptr(ws), ws_bytes, ctypes.byref(sel), _stream()), "pm_sort_pairs")
    out = (keys1, vals1) if sel.value == 1 else (keys, vals)
    if return_workspace:
        return out + (ws,)
    return out


def sort_error_flag(ws: torch.Tensor) -> int:
    """Look-back timeout flag of a finished sort (debug aid; synchronises)."""
    words = ws.view(torch.int32)
    return int(words[8 * 256 + 8].item())


@_instrument("tile_offsets", 1)
def tile_offsets(sorted_ids, C: int, width: int, height: int, d_n=None):
    L = _lib.load()
    tw, th = tile_grid(width, height)
    off = torch.empty(C, th, tw, dtype=torch.int32, device=sorted_ids.device)
    _lib.check(L.pm_tile_offsets(_ptr(sorted_ids), sorted_ids.numel(), _ptr(d_n), C, tw, th, _ptr(off), _stream()),
               "pm_tile_offsets")
    return off


# ------------------------------------------------------------------ stage 4
def _colors_layout(colors, C, N):
    if colors is None:
        return None, 0, 0
    if colors.dim() == 3:
        assert colors.shape[:2] == (C, N), f"colors {tuple(colors.shape)} vs C={C}, N={N}"
        return colors, 1, colors.shape[-1]
    assert colors.dim() == 2 and colors.shape[0] == N
    return colors, 0, colors.shape[-1]


@_instrument("composite_fwd", 1)
def composite_fwd(means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids, width,
                  height, d_n=None, want_stream=False):
    """-> render [C,H,W,CD], alphas [C,H,W], last_ids [C,H,W], transmittance [C,H,W] (see pm_composite_fwd)
    [, splat_stream [n,12] when want_stream: what composite_bwd replays]."""
    L = _lib.load()
    C, N = means2d.shape[:2]
    colors, per_cam, D = _colors_layout(colors, C, N)
    CD = D + (0 if depths is None else 1)
    dev = means2d.device
    render = torch.empty(C, height, width, CD, device=dev)
    alphas = torch.empty(C, height, width, device=dev)
    last_ids = torch.empty(C, height, width, dtype=torch.int32, device=dev)
    trans = torch.empty(C, height, width, device=dev)
    stream = torch.empty(flatten_ids.numel(), STREAM_FLOATS, device=dev) if want_stream else None
    _lib.check(L.pm_composite_fwd(_ptr(means2d), _ptr(conics), _ptr(opacities), _ptr(colors), per_cam, D,
                                  _ptr(depths), _ptr(backgrounds), _ptr(offsets), _ptr(flatten_ids),
                                  flatten_ids.numel(), _ptr(d_n), N, C, width, height, TILE, _ptr(render),
                                  _ptr(alphas), _ptr(last_ids), _ptr(trans), _ptr(stream), _stream()), "pm_composite_fwd")
    return (render, alphas, last_ids, trans, stream) if want_stream else (render, alphas, last_ids, trans)


STREAM_FLOATS = 12  # floats per splat-stream record (include/prometheus_gs.h)
# the fused forward leaves the backward a map of the list entries each warp footprint evaluated (1 bit per entry and
# footprint), so that composite_bwd replays those and skips its own culling ("0": developer knob for A/B runs)
HIT_MAP = __import__("os").environ.get("PM_HIT_MAP", "1") != "0"


@_instrument("splat_stream_build", 1)
def splat_stream_build(means2d, conics, opacities, colors, depths, flatten_ids, d_n=None, views_per_scene=0):
    """Gather the whole sorted list into a splat stream [n,12] (pm_splat_stream_build): for a backward on a forward
    state that did not come from our forward kernels."""
    L = _lib.load()
    C, N = means2d.shape[:2]
    n_scenes = C // views_per_scene if views_per_scene else 1
    colors, per_cam, D = _colors_layout(colors, C, N * n_scenes)
    n = flatten_ids.numel()
    stream = torch.empty(n, STREAM_FLOATS, device=means2d.device)
    _lib.check(L.pm_splat_stream_build(_ptr(means2d), _ptr(conics), _ptr(opacities), _ptr(colors), per_cam, D, _ptr(depths),
                                       _ptr(flatten_ids), n, _ptr(d_n), N, C, int(views_per_scene), _ptr(stream), _stream()),
               "pm_splat_stream_build")
    return stream


@_instrument("tile_render_fwd", 1)
def tile_render_fwd(records, offsets, means2d, conics, opacities, colors, depths, backgrounds, width, height, d_n=None,
                    tile_order=None, want_stream=False):
    """Fused sort-as-you-go forward -> render, alphas, last_ids, flatten_ids (sorted prefix per tile),
    sorted_counts [C,th,tw], overflow flag int32[1], transmittance [C,H,W]."""
    L = _lib.load()
    C, N = means2d.shape[:2]
    colors, per_cam, D = _colors_layout(colors, C, N)
    CD = D + (0 if depths is None else 1)
    dev = means2d.device
    n = records.numel()
    render = torch.empty(C, height, width, CD, device=dev)
    alphas = torch.empty(C, height, width, device=dev)
    last_ids = torch.empty(C, height, width, dtype=torch.int32, device=dev)
    flat = torch.empty(n, dtype=torch.int32, device=dev)
    sorted_counts = torch.empty_like(offsets)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    trans = torch.empty(C, height, width, device=dev)
    stream = torch.empty(n, STREAM_FLOATS, device=dev) if want_stream else None
    _lib.check(L.pm_tile_render_fwd(_ptr(records), _ptr(offsets), n, _ptr(d_n), _ptr(means2d), _ptr(conics),
                                    _ptr(opacities), _ptr(colors), per_cam, D, _ptr(depths), _ptr(backgrounds), N, C,
                                    width, height, TILE, _ptr(render), _ptr(alphas), _ptr(last_ids), _ptr(flat),
                                    _ptr(sorted_counts), _ptr(flag), _ptr(tile_order), _ptr(trans), _ptr(stream), _stream()),
               "pm_tile_render_fwd")
    out = (render, alphas, last_ids, flat, sorted_counts, flag, trans)
    return out + (stream,) if want_stream else out


@_instrument("composite_bwd", 1)
def composite_bwd(means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids, width,
                  height, alphas, last_ids, v_render, v_alphas, d_n=None, transmittance=None, stream=None, touched=None):
    """transmittance: the forward's exact final T per pixel; None = rebuild it as 1 - alphas (upstream's form).
    stream: the forward's splat stream; None = gather it from the sorted list first (pm_splat_stream_build).
    touched: zero-filled uint8 [C,N] that receives 1 for every pair with a gradient (for project_bwd), or None."""
    L = _lib.load()
    C, N = means2d.shape[:2]
    colors, per_cam, D = _colors_layout(colors, C, N)
    if stream is None:
        stream = splat_stream_build(means2d, conics, opacities, colors, depths, flatten_ids, d_n)
    v_render = _opt(v_render, torch.float32, "v_render")
    v_alphas = _opt(v_alphas, torch.float32, "v_alphas")
    v_means2d = torch.zeros_like(means2d)
    v_conics = torch.zeros_like(conics)
    v_colors = torch.zeros_like(colors) if colors is not None else None
    v_opac = torch.zeros_like(opacities)
    v_depths = torch.zeros_like(depths) if depths is not None else None
    _lib.check(L.pm_composite_bwd(_ptr(stream), per_cam, D, int(depths is not None), _ptr(backgrounds), _ptr(offsets),
                                  stream.shape[0], _ptr(d_n), N, C, width, height, TILE, _ptr(alphas),
                                  _ptr(last_ids), _ptr(_opt(transmittance, torch.float32, "transmittance")),
                                  _ptr(v_render), _ptr(v_alphas), _ptr(v_means2d),
                                  _ptr(v_conics), _ptr(v_colors), _ptr(v_opac), _ptr(v_depths), _ptr(touched), _stream()),
               "pm_composite_bwd")
    return v_means2d, v_conics, v_colors, v_opac, v_depths


# ------------------------------------------------------------------ module fast path (packed outputs)
@_instrument("tile_render_fwd", 1)
def tile_render_packed_fwd(records, offsets, means2d, conics, opacities, colors, depths, backgrounds, width, height,
                           d_n=None, tile_order=None, views_per_scene=0, for_backward=False, record_stride=0,
                           capacity=None):
    """Fused sort-as-you-go forward with the module epilogue: -> images [C,3,H,W] (2 rgb - 1),
    depth_out [C,1,H,W] (expected depth), masks [C,1,H,W], last_ids [C,H,W], flatten_ids (sorted prefix per
    tile), sorted_counts [C,th,tw], overflow flag int32[1], transmittance [C,H,W], splat_stream [n,12] and hit_words
    [8, W] int32 (all three None unless for_backward: the backward's exact T_final, the records it replays and the map of
    the entries each warp footprint evaluated, which replaces the backward's culling). `records` may be a speculative
    capacity with the true count in d_n[0] (see pm_tile_render_packed_fwd). record_stride > 0: `records` is the
    fixed-capacity tile layout of project_fwd(tile_capacity=...) and `capacity` sizes flatten_ids / the stream."""
    L = _lib.load()
    C, N = means2d.shape[:2]
    n_scenes = C // views_per_scene if views_per_scene else 1
    colors, per_cam, D = _colors_layout(colors, C, N * n_scenes)
    assert D == 3 and depths is not None and opacities.numel() == N * n_scenes
    dev = means2d.device
    n = int(capacity) if record_stride else records.numel()
    images = torch.empty(C, 3, height, width, device=dev)
    depth_out = torch.empty(C, 1, height, width, device=dev)
    masks = torch.empty(C, 1, height, width, device=dev)
    last_ids = torch.empty(C, height, width, dtype=torch.int32, device=dev)
    flat = torch.empty(n, dtype=torch.int32, device=dev)
    sorted_counts = torch.empty_like(offsets)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    trans = torch.empty(C, height, width, device=dev) if for_backward else None
    stream = torch.empty(n, STREAM_FLOATS, device=dev) if for_backward else None
    hit_words = None
    if for_backward and HIT_MAP:
        hit_words = torch.zeros(8, int(L.pm_hit_words_stride(n, C, width, height, TILE)), dtype=torch.int32, device=dev)
    _lib.check(L.pm_tile_render_packed_fwd(_ptr(records), int(record_stride), _ptr(offsets), n, _ptr(d_n), _ptr(means2d), _ptr(conics),
                                           _ptr(opacities), _ptr(colors), per_cam, _ptr(depths), _ptr(backgrounds), N, C,
                                           width, height, TILE, _ptr(images), _ptr(depth_out), _ptr(masks),
                                           _ptr(last_ids), _ptr(flat), _ptr(sorted_counts), _ptr(flag), _ptr(tile_order),
                                           int(views_per_scene), _ptr(trans), _ptr(stream), _ptr(hit_words),
                                           0 if hit_words is None else hit_words.shape[1], _stream()),
               "pm_tile_render_packed_fwd")
    return images, depth_out, masks, last_ids, flat, sorted_counts, flag, trans, stream, hit_words


@_instrument("composite_bwd", 1)
def composite_bwd_packed(means2d, conics, opacities, colors, depths, backgrounds, offsets, flatten_ids, width, height,
                         masks, depth_out, last_ids, v_images, v_depth_out, v_masks, d_n=None, tile_order=None,
                         views_per_scene=0, transmittance=None, stream=None, touched=None, hit_words=None):
    """hit_words: the forward's hit map (tile_render_packed_fwd), valid only together with ITS stream."""
    L = _lib.load()
    C, N = means2d.shape[:2]
    colors, per_cam, D = _colors_layout(colors, C, N * (C // views_per_scene if views_per_scene else 1))
    if stream is None:
        stream = splat_stream_build(means2d, conics, opacities, colors, depths, flatten_ids, d_n, views_per_scene)
        hit_words = None
    v_images, v_depth_out, v_masks = (_opt(t, torch.float32, "grad") for t in (v_images, v_depth_out, v_masks))
    # the five accumulation targets as views of ONE zero-filled allocation (one fill kernel instead of five)
    shapes = [means2d.shape, conics.shape, colors.shape, opacities.shape, depths.shape]
    sizes = [(int(torch.Size(sh).numel()) + 3) & ~3 for sh in shapes]  # 16-byte aligned views
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=means2d.device)
    views, o = [], 0
    for sh, n in zip(shapes, sizes):
        views.append(flat[o:o + int(torch.Size(sh).numel())].view(sh))
        o += n
    v_means2d, v_conics, v_colors, v_opac, v_depths = views
    _lib.check(L.pm_composite_bwd_packed(_ptr(stream), per_cam, _ptr(backgrounds), _ptr(offsets), stream.shape[0], _ptr(d_n),
                                         N, C, width, height, TILE, _ptr(masks), _ptr(depth_out), _ptr(last_ids),
                                         _ptr(transmittance), _ptr(v_images), _ptr(v_depth_out), _ptr(v_masks),
                                         _ptr(v_means2d), _ptr(v_conics), _ptr(v_colors), _ptr(v_opac), _ptr(v_depths),
                                         _ptr(tile_order), int(views_per_scene), _ptr(touched), _ptr(hit_words),
                                         0 if hit_words is None else hit_words.shape[1], _stream()),
               "pm_composite_bwd_packed")
    return v_means2d, v_conics, v_colors, v_opac, v_depths


# ------------------------------------------------------------------ trajectory row
@_instrument("sample_cameras", 1)
def sample_cameras(dense_cameras, t):
    """dense_cameras [B,N,18], t [B,M] -> [B,M,18] (pm_sample_cameras)."""
    L = _lib.load()
    dense_cameras = _req(dense_cameras, torch.float32, "dense_cameras")
    t = _req(t, torch.float32, "t")
    B, N, Cc = dense_cameras.shape
    assert Cc == 18 and t.dim() == 2 and t.shape[0] == B, (tuple(dense_cameras.shape), tuple(t.shape))
    M = t.shape[1]
    out = torch.empty(B, M, 18, device=dense_cameras.device)
    _lib.check(L.pm_sample_cameras(_ptr(dense_cameras), _ptr(t), B, N, M, _ptr(out), _stream()), "pm_sample_cameras")
    return out


@_instrument("pack_frames_u8", 1)
def pack_frames_u8(images, out=None):
    """images [F,3,H,W] fp32 in [-1,1] -> uint8 [F,H,W,3] (pm_pack_frames_u8)."""
    L = _lib.load()
    images = _req(images, torch.float32, "images")
    Fn, Cc, H, W = images.shape
    assert Cc == 3
    if out is None:
        out = torch.empty(Fn, H, W, 3, dtype=torch.uint8, device=images.device)
    assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.numel() == Fn * H * W * 3
    _lib.check(L.pm_pack_frames_u8(_ptr(images), Fn, H, W, _ptr(out), _stream()), "pm_pack_frames_u8")
    return out


# ------------------------------------------------------------------ densification statistics
@_instrument("densify_stats", 1)
def densify_stats(radii, v_means2d, is_visible, max_radii2d, xyz_gradient_accum, denom):
    """In-place update of the refiner's per-Gaussian state from C rendered views (pm_densify_stats).
    radii [C,N] int32, v_means2d [C,N,2] fp32; state tensors fp32 with N elements each ([N] or [N,1])."""
    L = _lib.load()
    radii = _req(radii, torch.int32, "radii")
    v_means2d = _req(v_means2d, torch.float32, "v_means2d")
    C, N = radii.shape
    assert v_means2d.shape == (C, N, 2)
    for t in (is_visible, max_radii2d, xyz_gradient_accum, denom):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == N
    _lib.check(L.pm_densify_stats(_ptr(radii), _ptr(v_means2d), C, N, _ptr(is_visible), _ptr(max_radii2d),
                                  _ptr(xyz_gradient_accum), _ptr(denom), _stream()), "pm_densify_stats")


# ------------------------------------------------------------------ .splat packing
@_instrument("splat_pack", lambda *a, **k: 8)
def splat_pack(xyz, rgb, opacity, scales, rots, compatible=True):
    """Activated Gaussians ([n,3], [n,3], [n], [n,3], [n,4]) -> the .splat byte stream [n,32] uint8 in the
    reference's order (volume x opacity, descending), all on the device."""
    L = _lib.load()
    xyz, rgb, opacity, scales, rots = (_req(t, torch.float32, "gaussians") for t in (xyz, rgb, opacity, scales, rots))
    n = xyz.shape[0]
    dev = xyz.device
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    vals = torch.empty(n, dtype=torch.int32, device=dev)
    out = torch.empty(n, 32, dtype=torch.uint8, device=dev)
    if n == 0:
        return out
    _lib.check(L.pm_splat_keys(_ptr(opacity), _ptr(scales), n, int(bool(compatible)), _ptr(keys), _ptr(vals), _stream()),
               "pm_splat_keys")
    _, order, ws = sort_pairs(keys, vals, 32, return_workspace=True)
    _lib.check(L.pm_splat_pack(_ptr(xyz), _ptr(rgb), _ptr(opacity), _ptr(scales), _ptr(rots), _ptr(order), n,
                               int(bool(compatible)), _ptr(out), _stream()), "pm_splat_pack")
    if sort_error_flag(ws):  # (the exporter copies `out` to the host next: this sync is not an extra one)
        raise RuntimeError("pm_sort_pairs: decoupled look-back timed out; the .splat order is invalid")
    return out


# --------------------------------------------------------------------------- loss-side glue
@_instrument("loss_terms_fwd", 1)
def loss_terms_fwd(images, target, depths, masks):
    """images / target [V,3,H,W], depths / masks [V,1,H,W] -> (disparity [V,1,H,W], valid [V,1,H,W],
    sums = {sum (images - target)^2, sum (1 - masks)} as 2 doubles on the device)."""
    L = _lib.load()
    images, target = _req(images, torch.float32, "images"), _req(target, torch.float32, "target")
    depths, masks = _req(depths, torch.float32, "depths"), _req(masks, torch.float32, "masks")
    V, three, H, W = images.shape
    assert three == 3 and target.shape == images.shape and depths.shape == (V, 1, H, W) and masks.shape == (V, 1, H, W)
    disparity, valid = torch.empty_like(depths), torch.empty_like(depths)
    sums = torch.empty(2, dtype=torch.float64, device=images.device)
    _lib.check(L.pm_loss_terms_fwd(_ptr(images), _ptr(target), _ptr(depths), _ptr(masks), V, H, W, _ptr(disparity),
                                   _ptr(valid), _ptr(sums), _stream()), "pm_loss_terms_fwd")
    return disparity, valid, sums


@_instrument("loss_terms_bwd", 1)
def loss_terms_bwd(images, target, depths, g_mse, g_entropy, v_disparity):
    """-> (v_images, v_depths, v_masks); g_mse / g_entropy: 0-dim device tensors or None, v_disparity [V,1,H,W] or None."""
    L = _lib.load()
    V, _, H, W = images.shape
    g_mse, g_entropy = _opt(g_mse, torch.float32, "g_mse"), _opt(g_entropy, torch.float32, "g_entropy")
    v_disparity = _opt(v_disparity, torch.float32, "v_disparity")
    v_images, v_depths, v_masks = torch.empty_like(images), torch.empty_like(depths), torch.empty_like(depths)
    _lib.check(L.pm_loss_terms_bwd(_ptr(images), _ptr(target), _ptr(depths), V, H, W, _ptr(g_mse), _ptr(g_entropy),
                                   _ptr(v_disparity), _ptr(v_images), _ptr(v_depths), _ptr(v_masks), _stream()),
               "pm_loss_terms_bwd")
    return v_images, v_depths, v_masks
