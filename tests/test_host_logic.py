"""CPU tests of the host-side logic: synthetic scene generator, key-bit / tile-grid helpers,
sharding arithmetic and the world_size-2 gather path over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from prometheus_b200 import ops, parallel, scenes


def test_scene_generator_is_deterministic_and_shaped():
    a = scenes.make_scene("C1", views_in=2, size_in=16, views_out=3, size_out=24)
    b = scenes.make_scene("C1", views_in=2, size_in=16, views_out=3, size_out=24)
    assert torch.equal(a.params, b.params) and torch.equal(a.cameras_out, b.cameras_out)
    assert a.params.shape == (1, 2, 14, 16, 16) and a.cameras_in.shape == (1, 2, 18) and a.cameras_out.shape == (1, 3, 18)
    assert a.n_gaussians == 2 * 16 * 16
    c2w0 = a.cameras_in[0, 0, :12].reshape(3, 4)
    assert torch.allclose(c2w0, torch.eye(4)[:3])                    # first input camera is the identity
    R = a.cameras_out[0, 1, :12].reshape(3, 4)[:, :3]
    assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-6)          # rigid
    assert float(a.cameras_in[0, 0, 12]) == pytest.approx(207.3 * 16 / 256)
    for name, cfg in scenes.CONFIGS.items():
        assert cfg["views_in"] * cfg["size_in"] ** 2 in (262144, 2097152, 524288, 8388608), name


def test_key_bits_and_tile_grid_match_gsplat_formulas():
    # floor(log2(n)) + 1
    assert ops.key_bits(32, 32, 8) == (11, 4)
    assert ops.key_bits(16, 16, 1) == (9, 1)
    assert ops.key_bits(64, 64, 120) == (13, 7)
    assert ops.tile_grid(512, 512) == (32, 32) and ops.tile_grid(50, 37) == (4, 3)


def test_oracle_bits_agree_with_host_helper(oracle):
    for tw, th, C in ((32, 32, 8), (4, 3, 1), (64, 64, 120), (1, 1, 2)):
        assert ops.key_bits(tw, th, C) == oracle.tile_bits(tw, th, C)


@pytest.mark.parametrize("n,world", [(8, 2), (120, 8), (7, 4), (3, 8), (0, 2)])
def test_shard_range_partitions(n, world):
    spans = [parallel.shard_range(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
        assert e0 == s1 and e0 >= s0
    sizes = [e - s for s, e in spans]
    assert max(sizes) - min(sizes) <= 1


def test_slot_views_layout():
    n = 5
    slot = torch.arange(14 * n, dtype=torch.float32)
    xyz, rgb, op, sc, q = parallel.slot_views(slot, n)
    assert xyz.shape == (n, 3) and rgb.shape == (n, 3) and op.shape == (n,) and sc.shape == (n, 3) and q.shape == (n, 4)
    assert float(rgb[0, 0]) == 3 * n and float(op[0]) == 6 * n and float(sc[0, 0]) == 7 * n and float(q[0, 0]) == 10 * n
    xyz[0, 0] = -1.0
    assert float(slot[0]) == -1.0  # views, not copies


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import gs_oracle as O

        sc = scenes.make_scene("C1", seed=9, views_in=4, size_in=8, views_out=6, size_out=16)
        v0, v1 = parallel.shard_range(4, world, rank)      # this rank decodes input views [v0, v1)
        n_local = (v1 - v0) * 8 * 8

        def unproject_into(slots):
            g = O.unproject(sc.params[:, v0:v1].numpy(), sc.cameras_in[:, v0:v1].numpy(), mode="mvsplat", z_near=0.01, z_far=100.0)
            for dst, key in zip(slots, ("xyz", "features", "opacity", "scales", "rotations")):
                dst.copy_(torch.from_numpy(g[key][0]).reshape(dst.shape))

        full = parallel.gather_gaussians(unproject_into, n_local, "cpu")
        (s, e), frames = parallel.render_sharded_views(lambda cams: (cams.clone(),), sc.cameras_out[0])
        ref = O.unproject(sc.params.numpy(), sc.cameras_in.numpy(), mode="mvsplat", z_near=0.01, z_far=100.0)
        ok = all(np.array_equal(t.numpy().reshape(-1), ref[k][0].reshape(-1))
                 for t, k in zip(full, ("xyz", "features", "opacity", "scales", "rotations")))
        ret[rank] = (ok, (s, e), tuple(frames[0].shape))
    finally:
        dist.destroy_process_group()


def test_view_sharded_gather_world2_gloo():
    """N>1 path on CPU: two ranks unproject their own input views into their gather slots, ONE all-gather,
    every rank ends up with the full scene in single-GPU order; target views are split without a collective."""
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == (0, 3) and ret[1][1] == (3, 6)
    assert ret[0][2] == (3, 18)


def test_lazy_meta_mapping_semantics():
    """`rasterization`'s meta dict: the sorted arrays are computed once, by the first read of any of them through the
    mapping interface; membership / keys / eager entries never trigger it."""
    from prometheus_b200.rasterization import _LazyMeta

    calls = []

    def thunk():
        calls.append(1)
        return {"tiles_per_gauss": "tpg", "isect_ids": "ids", "flatten_ids": "flat"}

    m = _LazyMeta({"radii": "r", "isect_offsets": "off"}, thunk)
    assert "isect_ids" in m and len(m) == 5 and set(m.keys()) == {"radii", "isect_offsets", "tiles_per_gauss", "isect_ids", "flatten_ids"}
    assert m["radii"] == "r" and m.get("isect_offsets") == "off" and not m.materialized and not calls
    assert "pending" in repr(m) and not calls
    assert m["flatten_ids"] == "flat" and m.materialized and calls == [1]
    assert m["isect_ids"] == "ids" and m.get("tiles_per_gauss") == "tpg" and calls == [1]
    for make in (dict, lambda d: {**d}, lambda d: d.copy(), lambda d: dict(d.items())):
        m2 = _LazyMeta({"radii": "r"}, thunk)
        assert make(m2)["isect_ids"] == "ids" and m2.materialized
    m3 = _LazyMeta({"radii": "r"}, thunk)
    assert list(m3.values()).count("ids") == 1
    m4 = _LazyMeta({"radii": "r"}, thunk)
    m4["isect_ids"] = "mine"                      # a caller's own value survives
    assert m4["isect_ids"] == "mine" and m4["flatten_ids"] == "flat"
    m5 = _LazyMeta({"radii": "r"}, thunk)
    assert m5.pop("tiles_per_gauss") == "tpg" and "tiles_per_gauss" not in m5
