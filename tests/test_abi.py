"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports
every symbol include/prometheus_gs.h declares; argument validation happens on the
host before anything is enqueued (so it is testable without a GPU)."""
import ctypes
import os
import re

import pytest

from prometheus_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "prometheus_gs.h")).read()
    return sorted(set(re.findall(r"PM_API[^;(]*?\b(pm_\w+)\s*\(", hdr)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for s in ("pm_unproject_fwd", "pm_unproject_bwd", "pm_camera_prep", "pm_project_fwd", "pm_project_bwd",
              "pm_scan_tiles", "pm_bin_emit", "pm_sort_pairs", "pm_tile_offsets", "pm_composite_fwd",
              "pm_tile_render_fwd", "pm_bin_count", "pm_tile_scan", "pm_bin_scatter", "pm_tile_sort", "pm_tile_sort_capacity",
              "pm_composite_bwd", "pm_abi_version"):
        assert s in syms
    assert set(syms) == set(_lib.SIGNATURES)  # the ctypes binding mirrors the header one to one


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert os.path.exists(_lib.library_path())
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert lib.pm_abi_version() == 6
    assert lib.pm_status_string(0) == b"ok"
    assert b"workspace" in lib.pm_status_string(-3)


def test_no_torch_or_oracle_in_the_abi():
    """The boundary is plain C: the shared library must not link torch or the oracle."""
    import subprocess

    out = subprocess.run(["ldd", _lib.library_path()], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "c10" not in out


def test_argument_errors_are_reported_before_launch():
    lib = _lib.load()
    null = ctypes.c_void_p(None)
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    assert lib.pm_unproject_fwd(null, null, null, 0, 0.01, 100.0, 0.001, 1.0, 1, 4, 4, null, null, null, null, null, null) == -1
    assert lib.pm_unproject_fwd(one, one, one, 7, 0.01, 100.0, 0.001, 1.0, 1, 4, 4, one, one, one, one, one, null) == -2
    assert lib.pm_project_fwd(one, one, one, one, one, 4, 1, 32, 32, 8, 0.3, 0.01, 1e10, 0.0, one, one, one, one, null, null, null) == -2  # tile_size 8
    assert lib.pm_project_fwd(one, one, one, one, one, -1, 1, 32, 32, 16, 0.3, 0.01, 1e10, 0.0, one, one, one, one, null, null, null) == -1
    assert lib.pm_project_scatter_fwd(one, one, one, one, one, 4, 1, 0, 32, 32, 16, 0.3, 0.01, 1e10, 0.0, one, one, one, one, one, null, 64, null, null) == -1  # no records but a capacity
    assert lib.pm_project_scatter_fwd(one, one, one, one, one, 4, 1, 0, 32, 32, 16, 0.3, 0.01, 1e10, 0.0, one, one, one, one, one, one, 0, null, null) == -1  # records but tile_cap 0
    sel = ctypes.c_int(-1)
    assert lib.pm_sort_pairs(one, one, one, one, 100, null, 40, one, 8, ctypes.byref(sel), null) == -3  # workspace too small
    assert lib.pm_sort_pairs(one, one, one, one, 1 << 31, null, 40, one, 1 << 40, ctypes.byref(sel), null) == -5
    assert lib.pm_sort_pairs(null, null, null, null, 0, null, 44, null, 0, ctypes.byref(sel), null) == 0 and sel.value == 0
    assert lib.pm_composite_fwd(one, one, one, one, 0, 5, null, null, one, one, 10, null, 4, 1, 32, 32, 16, one, one, one, null, null, null) == -2
    assert lib.pm_tile_sort_capacity() == 8192
    assert lib.pm_tile_sort(one, null, one, 10, null, 1, 2, 2, -1, one, one, null, null) == -1  # overflow flag is mandatory
    assert lib.pm_bin_scatter(one, one, one, one, one, 4, 1, 8, 2, 2, one, 10, null, null, 0, null) == -2  # tile_size 8
    assert lib.pm_bin_scatter(one, one, one, one, one, 4, 1, 16, 2, 2, one, 10, one, null, 0, null) == -1  # conics without opacities
    assert lib.pm_sort_workspace_bytes(1 << 20) >= (1 << 20) // 4096 * 1024 * 4
    assert lib.pm_scan_workspace_bytes(1 << 20) >= (1 << 20) // 2048 * 4


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: handing the product CPU tensors is an error, not a slow path."""
    import torch

    import prometheus_b200 as pb

    N = 8
    with pytest.raises(RuntimeError, match="no CPU"):
        pb.rasterization(torch.zeros(N, 3), torch.ones(N, 4), torch.ones(N, 3), torch.ones(N), torch.ones(N, 3),
                         torch.eye(4)[None], torch.eye(3)[None], 32, 32, packed=False)
    with pytest.raises(NotImplementedError):
        pb.rasterization(torch.zeros(N, 3), torch.ones(N, 4), torch.ones(N, 3), torch.ones(N), torch.ones(N, 3),
                         torch.eye(4)[None], torch.eye(3)[None], 32, 32)  # packed=True default upstream
    with pytest.raises(AssertionError):
        pb.rasterization(torch.zeros(N, 2), torch.ones(N, 4), torch.ones(N, 3), torch.ones(N), torch.ones(N, 3),
                         torch.eye(4)[None], torch.eye(3)[None], 32, 32, packed=False)


def test_widened_entry_points_refuse_cpu_tensors_and_bad_configurations():
    """The rows added around the path keep the rule: CPU tensors are an error, and configurations a CUDA-graph
    capture cannot hold are rejected up front."""
    import torch

    import prometheus_b200 as pb

    with pytest.raises(RuntimeError, match="no CPU"):
        pb.render_loss_terms(torch.zeros(1, 1, 3, 8, 8), torch.ones(1, 1, 1, 8, 8), torch.ones(1, 1, 1, 8, 8),
                             torch.zeros(1, 1, 3, 8, 8))
    conv, rend = pb.GaussianConverter(), pb.GaussianRenderer()
    with pytest.raises(RuntimeError, match="no CPU"):
        conv(torch.zeros(1, 1, 14, 8, 8, dtype=torch.bfloat16), torch.zeros(1, 1, 18))
    with pytest.raises(ValueError, match="random background"):
        pb.GraphedDecodeRender(conv, rend, 64, 64, bg_color="random")
    rend.fused = False
    with pytest.raises(ValueError, match="fused path"):
        pb.GraphedDecodeRender(conv, rend, 64, 64)


def test_param_dtype_codes_match_the_header():
    """ops.PARAM_DTYPES (torch dtype -> enum pm_dtype) against include/prometheus_gs.h."""
    import re

    import torch

    from prometheus_b200 import ops

    text = open(os.path.join(ROOT, "include", "prometheus_gs.h")).read()
    enum = dict((k, int(v)) for k, v in re.findall(r"(PM_DTYPE_\w+)\s*=\s*(\d+)", text))
    assert enum == {"PM_DTYPE_F32": 0, "PM_DTYPE_F16": 1, "PM_DTYPE_BF16": 2}
    assert ops.PARAM_DTYPES == {torch.float32: enum["PM_DTYPE_F32"], torch.float16: enum["PM_DTYPE_F16"],
                                torch.bfloat16: enum["PM_DTYPE_BF16"]}
