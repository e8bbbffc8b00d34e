"""ctypes loader for libprometheus_gs.so — the C-ABI declared in include/prometheus_gs.h.

There is NO fallback: if the library cannot be found (or built with nvcc), every
op raises. PyTorch is only used for device memory and streams; the kernels see
raw device pointers.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
# PM_GS_LIB: developer knob for A/B-timing an alternative build of the same C-ABI (tools/stage_bench.py)
_SO = os.environ.get("PM_GS_LIB") or os.path.join(_PKG, "libprometheus_gs.so")
_CSRC = os.path.join(_PKG, "csrc")
_lib = None

PM_OK = 0
ABI_VERSION = 6

# name -> (restype, argtypes); mirrors include/prometheus_gs.h one to one
_F, _I, _L, _P, _S = c_float, c_int, c_int64, c_void_p, c_size_t
SIGNATURES = {
    "pm_abi_version": (c_int, []),
    "pm_status_string": (c_char_p, [c_int]),
    "pm_last_cuda_error": (c_char_p, []),
    "pm_unproject_fwd": (c_int, [_P, _P, _P, _I, _F, _F, _F, _F, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pm_unproject_fwd_dt": (c_int, [_P, _I, _P, _P, _I, _F, _F, _F, _F, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pm_unproject_fwd_multi": (c_int, [_P, _I, _P, _P, _I, _F, _F, _F, _F, _I, _I, _I, _I, _P, _P]),
    "pm_unproject_bwd": (c_int, [_P, _P, _P, _I, _F, _F, _F, _F, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "pm_unproject_bwd_dt": (c_int, [_P, _I, _P, _P, _I, _F, _F, _F, _F, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "pm_camera_prep": (c_int, [_P, _I, _I, _I, _P, _P, _P]),
    "pm_project_fwd": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P]),
    "pm_project_count_fwd": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "pm_project_scatter_fwd": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _I, _P, _P]),
    "pm_bin_count": (c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "pm_tile_scan": (c_int, [_P, _L, _P, _P, _P, _P]),
    "pm_bin_scatter": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _L, _P, _P, _I, _P]),
    "pm_tile_sort_capacity": (c_int, []),
    "pm_tile_sort": (c_int, [_P, _P, _P, _L, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "pm_project_bwd": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pm_scan_workspace_bytes": (c_size_t, [_L]),
    "pm_scan_tiles": (c_int, [_P, _L, _P, _P, _P, _S, _P]),
    "pm_bin_emit": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _L, _P]),
    "pm_sort_workspace_bytes": (c_size_t, [_L]),
    "pm_sort_pairs": (c_int, [_P, _P, _P, _P, _L, _P, _I, _P, _S, POINTER(c_int), _P]),
    "pm_tile_offsets": (c_int, [_P, _L, _P, _I, _I, _I, _P, _P]),
    "pm_composite_fwd": (c_int, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _L, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pm_splat_stream_build": (c_int, [_P, _P, _P, _P, _I, _I, _P, _P, _L, _P, _I, _I, _I, _P, _P]),
    "pm_tile_render_fwd": (c_int, [_P, _P, _L, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pm_tile_render_packed_fwd": (c_int, [_P, _L, _P, _L, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _L, _P]),
    "pm_hit_words_stride": (_L, [_L, _I, _I, _I, _I]),
    "pm_composite_bwd_packed": (c_int, [_P, _I, _P, _P, _L, _P, _I, _I, _I, _I, _I,
                                        _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _L, _P]),
    "pm_splat_keys": (c_int, [_P, _P, _L, _I, _P, _P, _P]),
    "pm_splat_pack": (c_int, [_P, _P, _P, _P, _P, _P, _L, _I, _P, _P]),
    "pm_loss_terms_fwd": (c_int, [_P, _P, _P, _P, _L, _I, _I, _P, _P, _P, _P]),
    "pm_loss_terms_bwd": (c_int, [_P, _P, _P, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "pm_densify_stats": (c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "pm_sample_cameras": (c_int, [_P, _P, _I, _I, _I, _P, _P]),
    "pm_pack_frames_u8": (c_int, [_P, _L, _I, _I, _P, _P]),
    "pm_composite_bwd": (c_int, [_P, _I, _I, _I, _P, _P, _L, _P, _I, _I, _I, _I, _I,
                                 _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}


def library_path() -> str:
    return _SO


def kernel_source_sha() -> str:
    """sha256 (first 16 hex digits) over csrc/*.cu, *.cuh and the header: stamps profiler captures with the
    kernel sources they describe (profiles/traffic.json), so a stale capture is never quoted for newer code."""
    import hashlib

    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh")))
    for f in files:
        h.update(f.encode())
        h.update(open(os.path.join(_CSRC, f), "rb").read())
    h.update(open(os.path.join(os.path.dirname(_PKG), "include", "prometheus_gs.h"), "rb").read())
    return h.hexdigest()[:16]


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libprometheus_gs.so (in-tree)."""
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        raise RuntimeError("nvcc not found: cannot build libprometheus_gs.so")
    env = dict(os.environ)
    if shutil.which("nvcc") is None:
        env["PATH"] = "/usr/local/cuda/bin:" + env.get("PATH", "")
    subprocess.run(["make", "-C", _CSRC, "-j8", "all"], check=True, env=env,
                   stdout=None if verbose else subprocess.DEVNULL)
    return _SO


def _stale() -> bool:
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    for f in os.listdir(_CSRC):
        if f.endswith((".cu", ".cuh")) and os.path.getmtime(os.path.join(_CSRC, f)) > t:
            return True
    return False


def load():
    """Return the loaded CDLL; build it first when sources are newer and nvcc exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("PM_GS_LIB") and _stale():
        try:
            build()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(_SO):
                raise RuntimeError(
                    "libprometheus_gs.so is missing and could not be built "
                    f"({e}); run `python __graft_entry__.py` or `make -C prometheus_b200/csrc`. "
                    "There is no CPU fallback.") from e
    lib = ctypes.CDLL(_SO)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if lib.pm_abi_version() != ABI_VERSION:
        raise RuntimeError("libprometheus_gs.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == PM_OK:
        return
    lib = load()
    msg = lib.pm_status_string(rc).decode()
    if rc == -4:
        msg += ": " + lib.pm_last_cuda_error().decode()
    raise RuntimeError(f"{what} failed: {msg} (status {rc})")
