#!/usr/bin/env python
"""bench.py — headline benchmark of the GS-decoder rendering hot path.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

Workload (BASELINE.json configs[1], "C2"): 8 input views x 512^2 RGB-D maps ->
2,097,152 pixel-aligned Gaussians (unproject) -> 8 target views at 512^2 (project,
tile-bin, radix sort, alpha composite). One STEP = one such decode+render batch of
synthetic input. At N>1 every rank renders its own scene (scene sharding, no
data-path collective) => weak scaling; value = all ranks' pixels / max-over-ranks time.

metric: rendered Mpix/s, forward. Extra keys: Gaussians/s, fwd+bwd numbers, the
per-stage breakdown with algorithmic bytes, `roofline` for the dominant kernel,
`cpu_baseline` (the C oracle on all host cores, bounded sample) and `e2e`
(host buffers in, host buffers out, copies inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rendered Mpix/s (fwd: unproject+project+bin/sort+composite)"
UNIT = "Mpix/s"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = "/tmp/pm_clocks_%d_%d.csv" % (os.getpid(), gpu_index)

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_reference_step(O, scene, gauss, n_views):
    """One bounded sample of the workload on the CPU oracle: unproject all input views, then
    render `n_views` of the target views (project + bin + sort + composite)."""
    import numpy as np

    from prometheus_b200 import scenes

    t0 = time.perf_counter()
    if gauss is None:
        gauss = O.unproject(scene.params.numpy(), scene.cameras_in.numpy(), mode="mvsplat", z_near=scenes.Z_NEAR,
                            z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX)
    t_unproj = time.perf_counter() - t0
    S = scene.size_out
    t0 = time.perf_counter()
    vm, K = O.camera_prep(scene.cameras_out[0, :n_views].numpy(), S, S)
    O.forward_pipeline(gauss["xyz"][0], gauss["rotations"][0], gauss["scales"][0], gauss["opacity"][0, :, 0],
                       gauss["features"][0, :, 0], vm, K, S, S, backgrounds=np.ones((n_views, 3), np.float32),
                       with_depth=True, want_fragile=False)
    t_render = time.perf_counter() - t0
    return gauss, t_unproj, t_render


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. The reference is pure
    Python whose kernels live in the un-vendored gsplat wheel (not installable here), so this arm
    times the oracle port (oracle/gs_oracle.c, OpenMP, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import gs_oracle as O
    from prometheus_b200 import scenes

    O.build()
    # all the host threads this process may use — torchrun exports OMP_NUM_THREADS=1 to every rank, which would
    # otherwise turn the N > 1 reference runs into single-thread runs
    O.set_num_threads(len(os.sched_getaffinity(0)))
    cores = O.num_threads()
    scene = scenes.make_scene(args.config)
    S, V = scene.size_out, scene.cameras_out.shape[1]
    # One step = the whole workload (unproject + ALL V target views) unless that cannot finish in a few minutes on
    # this box: the first step is timed, and if (warmup + steps) such steps would exceed REF_BUDGET_S the step is
    # bounded to k < V views (+ the unproject pro rata). --ref-views k forces k.
    REF_BUDGET_S = 150.0
    k = args.ref_views or V
    pix_total, t_total, n_timed = 0.0, 0.0, 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, t_u, t_r = cpu_reference_step(O, scene, None, k)
        if it >= args.warmup:
            pix_total += k * S * S
            t_total += t_u * k / V + t_r  # unproject cost attributed pro rata to the rendered views
            n_timed += 1
        if it == 0 and not args.ref_views:
            est = (time.perf_counter() - t0) * (args.warmup + args.steps)
            if est > REF_BUDGET_S:
                k = max(1, int(V * REF_BUDGET_S / est))
    t = t_total / n_timed
    value = pix_total / t_total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, scene), "parallelism": "rank 0 only, %d host threads" % cores,
                   "numa_bound": False, "l2": "n/a (CPU)", "checks": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "per step: unproject of all input views (pro-rated) + %d of %d target views "
                                   "rendered by oracle/gs_oracle.c (OpenMP)" % (k, V)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gaussians_per_s": k * scene.n_gaussians / t,
    }
    print(json.dumps(line), flush=True)


def torch_cpu_baseline(O):
    """Pure-PyTorch CPU restatement (oracle/torch_ref.py) on 1 target view of config C1
    (4 x 256^2 -> 262,144 Gaussians, one 256^2 view), all host cores."""
    import torch

    from oracle import torch_ref
    from prometheus_b200 import scenes

    torch.set_num_threads(os.cpu_count() or 1)
    sc = scenes.make_scene("C1")
    g = O.unproject(sc.params.numpy(), sc.cameras_in.numpy(), mode="mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR,
                    s_min=scenes.S_MIN, s_max=scenes.S_MAX)
    vm, K = O.camera_prep(sc.cameras_out[0, :1].numpy(), 256, 256)
    args = [torch.from_numpy(g[k][0]) for k in ("xyz", "rotations", "scales")]
    args += [torch.from_numpy(g["opacity"][0, :, 0]), torch.from_numpy(g["features"][0, :, 0])]
    t0 = time.perf_counter()
    with torch.no_grad():
        torch_ref.rasterization(*args, torch.from_numpy(vm), torch.from_numpy(K), 256, 256,
                                backgrounds=torch.ones(1, 3), render_mode="RGB+ED")
    dt = time.perf_counter() - t0
    # forward + backward of the same view (torch autograd through the restatement = the oracle of the backward
    # kernels), with the synthetic render loss on all three outputs
    leaves = [a.clone().requires_grad_(True) for a in args]
    t0 = time.perf_counter()
    out = torch_ref.rasterization(*leaves, torch.from_numpy(vm), torch.from_numpy(K), 256, 256,
                                  backgrounds=torch.ones(1, 3), render_mode="RGB+ED")
    render, alpha = out[0], out[1]
    loss = (render[..., :3] ** 2).mean() + (1.0 / render[..., 3:].clamp(0.1, 100.0)).mean() + (1.0 - alpha).mean()
    loss.backward()
    dt_b = time.perf_counter() - t0
    return {"value": 256 * 256 / dt / 1e6, "unit": UNIT, "cores": os.cpu_count(),
            "sample": "oracle/torch_ref.py, 1 target view of C1 (262144 Gaussians -> 256^2) in %.2fs" % dt,
            "fwd_bwd": {"value": 256 * 256 / dt_b / 1e6, "unit": UNIT,
                        "sample": "the same view forward + autograd backward in %.2fs" % dt_b}}


def workload_name(cfg, scene):
    B, Vi = scene.params.shape[:2]
    s = scene.params.shape[-1]
    return "%s: %d scene x %d views x %d^2 RGB-D -> %d Gaussians, %d target views %d^2" % (
        cfg, B, Vi, s, scene.n_gaussians, scene.cameras_out.shape[1], scene.size_out)


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--ref-views", type=int, default=0,
                    help="reference arm: target views rendered per step (0 = all, bounded only if the run would take > 150 s)")
    ap.add_argument("--no-extra", action="store_true", help="skip the C5 / C4 sub-runs and the operator-path figures")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", choices=["eager", "graph"], default="graph",
                    help="e2e leg: DecodeRenderPipeline launching the modules eagerly, or replaying CUDA-graph plans")
    ap.add_argument("--loss", choices=["fused", "eager"], default="fused",
                    help="fwd+bwd leg: per-pixel loss glue through render_loss_terms (one kernel each way) or plain ATen ops")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay side figure")
    ap.add_argument("--no-half-maps", action="store_true", help="skip the bf16-maps e2e side figure")
    ap.add_argument("--no-bwd", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="C5 exchange step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from prometheus_b200 import GaussianConverter, GaussianRenderer, ops, scenes

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU: there is no CPU fallback"
    if world > 1:
        from prometheus_b200 import parallel as _par

        numa_bound = _par.bind_to_gpu_numa(local_rank)  # pinned host buffers of each rank on the NUMA node of its GPU
    else:
        numa_bound = False
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout, in front of the JSON line
        dist.init_process_group("nccl", device_id=dev)

    if args.config == "C5":
        line = run_c5(args, rank, world, dev, args.steps, args.warmup)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # every rank owns one scene of the named config (scene sharding; seeds differ per rank). C4 (text-to-3D batch of
    # 64 scenes) is split over the ranks instead: 64 / N scenes each, a fixed total => strong scaling.
    if args.config == "C4":
        assert 64 % world == 0
        scene = scenes.make_scene("C4", seed=1234 + 1000 * rank, scenes=64 // world)
    else:
        scene = scenes.make_scene(args.config, seed=1234 + 1000 * rank)
    S = scene.size_out
    B, V_out = scene.cameras_out.shape[:2]
    pixels = B * V_out * S * S
    conv = GaussianConverter("mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX).to(dev)
    rend = GaussianRenderer(background=[1, 1, 1]).to(dev)
    params_h = scene.params.pin_memory()
    cams_in_h = scene.cameras_in.pin_memory()
    cams_out_h = scene.cameras_out.pin_memory()
    params_d, cams_in_d, cams_out_d = params_h.to(dev), cams_in_h.to(dev), cams_out_h.to(dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # The renderer sizes its intersection buffers from the previous call and checks the device-side count
    # afterwards. In the timed loops the checks are deferred (rend.defer_checks): the forward-only loop
    # validates every step's render once the loop is queued, the training-style step validates after
    # loss.backward() is queued (before an optimizer would step) — the host never stalls the GPU mid-step.
    invalid = {"n": 0}

    def step_fwd(params, cin, cout):
        with torch.no_grad():
            g = conv(params, cin)
            return rend(cout, g, bg_color=None, h=S, w=S)

    loss_fn = scenes.render_loss_fused if args.loss == "fused" else scenes.render_loss

    def step_fwdbwd(params, cin, cout, target):
        p = params.detach().requires_grad_(True)
        g = conv(p, cin)
        images, depths, masks, _, _ = rend(cout, g, bg_color=None, h=S, w=S)
        loss = loss_fn(images, depths, masks, target)
        loss.backward()
        if not rend.validate():
            invalid["n"] += 1
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        import gc

        for _ in range(warmup):
            fn()
        barrier()
        gc.collect()
        gc.disable()  # as timeit does: a generation-2 collection in the middle of a step starves the GPU for milliseconds
        try:
            return _timed_body(fn, steps, profile)
        finally:
            gc.enable()

    def _timed_body(fn, steps, profile):
        ops.PROFILE = [] if profile else None
        n0 = ops.LAUNCHES
        evs = []
        for _ in range(steps):
            flush_buf.fill_(1)  # L2 flush, outside the timed events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        launches = ops.LAUNCHES - n0
        prof, ops.PROFILE = ops.PROFILE, None
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        if os.environ.get("PM_BENCH_DEBUG"):
            ms = torch.cuda.memory_stats()
            print("[debug] step ms:", [round(a.elapsed_time(b), 2) for a, b in evs], "device_alloc/free/retries:",
                  ms.get("num_device_alloc"), ms.get("num_device_free"), ms.get("num_alloc_retries"),
                  "hints:", dict(rend._cap_hint), file=sys.stderr)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, launches, prof

    sampler = ClockSampler(local_rank)
    sampler.start()
    step_fwd(params_d, cams_in_d, cams_out_d)  # first call of this shape reads the count back and sets the capacity hint
    rend.defer_checks = True
    total_ms, launches, prof = timed(lambda: step_fwd(params_d, cams_in_d, cams_out_d), args.steps, args.warmup, profile=True)
    assert rend.validate(), "a deferred capacity / overflow check failed inside the timed region"
    rend.defer_checks = False
    # records the fast path binned in the last timed step, all scenes (usable tiles: a subset of gsplat's intersections)
    binned = rend.last_stats[0] if rend.last_stats else None
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = world * pixels / (ms_per_step * 1e-3) / 1e6

    # ---- per-stage breakdown (CUDA events on the launching stream, inside the timed region)
    stage_ms = {}
    for name, e0, e1 in prof:
        stage_ms.setdefault(name, []).append(e0.elapsed_time(e1))
    stage_avg = {k: sum(v) / args.steps for k, v in stage_ms.items()}  # per step (a stage may run B times)

    # ---- algorithmic bytes from one instrumented forward
    with torch.no_grad():
        g = conv(params_d, cams_in_d)
        vm, K = ops.camera_prep(cams_out_d[0], S, S)
        from prometheus_b200.rasterization import _rasterize_core

        r, a, meta = _rasterize_core(g[0][0], g[4][0], g[3][0], g[2][0, :, 0], g[1][0, :, 0], vm, K, S, S,
                                     backgrounds=torch.ones(V_out, 3, device=dev), render_mode="RGB+ED", packed=False)
        I = int(meta["isect_ids"].numel())
        G = scene.n_gaussians
        P = V_out * G
        X = V_out * S * S
        off = meta["isect_offsets"].flatten().long()
        last = meta_last_ids(r, meta, ops, g, vm, K, S, V_out, dev)
        tw = meta["tile_width"]
        tiles_last = last.reshape(V_out, S // 16, 16, S // 16, 16).amax(dim=(2, 4)).flatten().long()
        nxt = torch.cat([off[1:], torch.tensor([I], device=dev)])
        nonempty = nxt > off
        I_need = int(((tiles_last - off + 1).clamp(min=0) * nonempty).sum().item())
        # the same two counts for the module fast path's own lists (single-pass binning of the usable tiles)
        I_bin1, I_need_bin = fast_path_counts(ops, g, vm, K, S, V_out, dev, I, int((nxt - off).max().item()), rend.usable_tiles)
    tb, cb = ops.key_bits(S // 16, S // 16, V_out)
    passes = (32 + tb + cb + 7) // 8
    # the module fast path bins only the tiles a splat can be seen in: ITS kernels are charged for the records they
    # really move (I_bin <= I), never for gsplat's larger count
    I_bin = min(I, -(-binned // B)) if binned else I_bin1
    bytes_alg = {
        "unproject_fwd": 112 * G * B,
        # module fast path: projection with the per-tile histogram fused in, no tiles_per_gauss array
        "project_fwd": (40 * G + 28 * P) * B,
        # single-pass binning: the projection also writes the 8-byte intersection records (no bin_scatter)
        "project_scatter_fwd": (40 * G + 28 * P + 8 * I_bin) * B,
        "scan_tiles": 12 * P * B,
        "bin_emit": (20 * P + 12 * I) * B,
        "sort_pairs": (passes * 24 + 8) * I * B,
        "tile_offsets": 8 * I * B,
        "bin_count": 12 * P * B,
        "bin_scatter": (16 * P + 8 * I) * B,
        "tile_sort": 20 * I * B,
        # fused sort-as-you-go: one pass over the tile's records at minimum + composite traffic
        "tile_render_fwd": (8 * I_bin + 4 * I_need_bin + 44 * I_need_bin + 24 * X) * B,
        "composite_fwd": (44 * I_need + 24 * X) * B,
    }
    peak, peak_src = read_peaks()
    stages = {}
    for k, ms in stage_avg.items():
        if k in bytes_alg and ms > 0:
            gbs = bytes_alg[k] / (ms * 1e-3) / 1e9
            stages[k] = {"ms": round(ms, 4), "alg_bytes": bytes_alg[k], "GBps": round(gbs, 1), "frac_hbm": round(gbs / peak, 4)}
        else:
            stages[k] = {"ms": round(ms, 4)}
    # dominant kernel: the single kernel with the largest share of the step. sort_pairs is a
    # sequence of `passes`+2 kernels, so compare composite_fwd with one sort pass.
    comp_ms = stage_avg.get("composite_fwd", 0.0)
    tsort_ms = stage_avg.get("tile_sort", 0.0)
    sort_pass_ms = stage_avg.get("sort_pairs", 0.0) / max(passes, 1)
    fused_ms = stage_avg.get("tile_render_fwd", 0.0)
    if fused_ms >= max(comp_ms, sort_pass_ms, tsort_ms) and fused_ms > 0:
        dom, dom_ms, dom_bytes = "tile_render_fwd_kernel", fused_ms / B, bytes_alg["tile_render_fwd"] / B
    elif comp_ms >= max(sort_pass_ms, tsort_ms):
        dom, dom_ms, dom_bytes = "composite_fwd_kernel", comp_ms / B, bytes_alg["composite_fwd"] / B
    elif tsort_ms >= sort_pass_ms:
        dom, dom_ms, dom_bytes = "tile_sort_kernel", tsort_ms / B, bytes_alg["tile_sort"] / B
    else:
        dom, dom_ms, dom_bytes = "rs_onesweep_kernel", sort_pass_ms / B, 24 * I
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/traffic.json, written
    # by tools/ncu_traffic.py). The capture is stamped with a hash of the kernel sources; a capture of OTHER sources
    # is not quoted (traffic = null, the reason is in traffic_source).
    traffic, traffic_src, tj = None, None, {}
    try:
        from prometheus_b200._lib import kernel_source_sha

        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.config != "C2":
            traffic_src = "captured on C2 only"
        elif tj.get("src_sha") != kernel_source_sha():
            traffic_src = "profiles/traffic.json is stale: captured for kernel sources %s, running %s" % (tj.get("src_sha"), kernel_source_sha())
            tj = {}
        else:
            traffic = tj.get(dom)
            traffic_src = "ncu --set full, profiles/traffic.json (%s), kernel sources %s" % (", ".join(tj.get("reports", [])), tj.get("src_sha"))
    except Exception as e:  # noqa: BLE001
        traffic_src = "unavailable: %s" % str(e)[:80]
    # the kernel's real ceiling is instruction issue: warp instructions of the committed ncu capture /
    # (live launch time x 148 SMs x 4 schedulers x sampled SM clock)
    issue = None
    try:
        wi = tj.get("_warp_instructions", {}).get(dom)
        if wi and args.config == "C2" and clocks.get("sm_mhz"):
            peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
            issue = {"warp_instructions": wi, "peak_warp_inst_per_s": peak_issue,
                     "frac": round(wi / (dom_ms * 1e-3) / peak_issue, 4), "source": "ncu smsp__inst_executed.sum (profiles/)"}
    except Exception:  # noqa: BLE001
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "issue": issue,
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "avg_launch_ms": round(dom_ms, 4),
                "alg_bytes_per_launch": int(dom_bytes),
                "note": "dominant kernel by time; composite and the shared-memory tile sort are issue/latency bound, "
                        "not HBM bound (see DESIGN.md); I_need=%d of I=%d" % (I_need, I)}

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region. Goes through the
    # public streaming API (DecodeRenderPipeline): every step copies its own inputs from pinned host
    # memory and its images back to pinned host memory; the copy of step i+1 / i-1 overlaps the
    # compute of step i on separate streams (double buffering), as a data loader would.
    from prometheus_b200 import DecodeRenderPipeline

    out_h = [torch.empty(B, V_out, c, S, S).pin_memory() for c in (3, 1, 1)]
    pipe = DecodeRenderPipeline(conv, rend, S, S, device=dev, graph=args.pipeline == "graph")
    pipe_default, out_h_default = pipe, out_h

    def e2e_loop(n, maps_h=None, pipe=None, out_h=None):
        pipe = pipe if pipe is not None else pipe_default
        out_h = out_h if out_h is not None else out_h_default
        inputs = (params_h if maps_h is None else maps_h, cams_in_h, cams_out_h)
        dbg = [] if os.environ.get("PM_BENCH_DEBUG") else None
        if B > rend.scene_chunk:  # a big host batch (C4) streams through the pipeline in scene_chunk-sized slices
            for i in range(n):
                t0 = time.perf_counter()
                pipe.run_batch(inputs, out_h)
                if dbg is not None:
                    dbg.append(round((time.perf_counter() - t0) * 1e3, 1))
            pipe.flush()
            if dbg is not None:
                print("[debug] run_batch host ms:", dbg, "captures/failed:", pipe.graphed and (pipe.graphed.n_captures, pipe.graphed.n_failed_checks),
                      "device_alloc:", torch.cuda.memory_stats().get("num_device_alloc"), file=sys.stderr)
            return
        pipe.prefetch(*inputs)
        for i in range(n):
            t0 = time.perf_counter()
            pipe.run(out_h, next_inputs_h=inputs if i + 1 < n else None)
            if dbg is not None:
                dbg.append(round((time.perf_counter() - t0) * 1e3, 2))
        pipe.flush()
        if dbg is not None:
            ms = torch.cuda.memory_stats()
            hs = torch.cuda.host_memory_stats() if hasattr(torch.cuda, "host_memory_stats") else {}
            print("[debug] e2e host ms per run():", dbg, "device_alloc:", ms.get("num_device_alloc"),
                  "host_alloc:", hs.get("num_host_alloc"), file=sys.stderr)

    e2e_loop(max(8, args.warmup))  # long enough for the caching allocator to own every buffer the pipeline rotates through
    barrier()
    flush_buf.fill_(1)
    import gc

    gc.collect()
    gc.disable()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    gc.enable()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_ms /= args.steps
    # the same, strictly serial (copy in -> compute -> copy out, one stream): the un-overlapped cost
    def step_serial():
        p = params_h.to(dev, non_blocking=True)
        ci = cams_in_h.to(dev, non_blocking=True)
        co = cams_out_h.to(dev, non_blocking=True)
        outs = step_fwd(p, ci, co)
        for dst, src in zip(out_h, outs[:3]):
            dst.copy_(src, non_blocking=True)

    ser_ms, _, _ = timed(step_serial, max(3, args.steps // 2), 3)
    ser_ms /= max(3, args.steps // 2)
    h2d = params_h.numel() * 4 + cams_in_h.numel() * 4 + cams_out_h.numel() * 4
    d2h = sum(t.numel() * 4 for t in out_h)
    e2e = {"value": world * pixels / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms, 4), "api": "DecodeRenderPipeline (3 streams, double buffered%s)" % (", CUDA-graph plans" if args.pipeline == "graph" else ""),
           "serial_ms_per_step": round(ser_ms, 4)}

    # ---- fwd+bwd (training-style step with the synthetic render loss)
    extra = {}
    if not args.no_bwd:
        target = (torch.rand(B, V_out, 3, S, S, device=dev) * 2 - 1)
        bsteps = max(3, args.steps // 2)
        rend.defer_checks = True
        bwd_ms, bl, bprof = timed(lambda: step_fwdbwd(params_d, cams_in_d, cams_out_d, target), bsteps, max(3, args.warmup), profile=True)
        rend.defer_checks = False
        assert invalid["n"] == 0, "a deferred capacity / overflow check failed inside the timed region"
        bwd_ms /= bsteps
        bst = {}
        for name, e0, e1 in bprof:
            bst[name] = bst.get(name, 0.0) + e0.elapsed_time(e1) / bsteps
        extra["fwd_bwd"] = {"value": world * pixels / (bwd_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": round(bwd_ms, 4),
                            "gaussians_per_s": world * B * V_out * G / (bwd_ms * 1e-3),
                            "stages_ms": {k: round(v, 4) for k, v in bst.items()},
                            "loss": "mse(images) + mean(1 / clip(depth, 0.1, 100)) + mean(1 - mask); per-pixel glue: %s"
                                    % ("prometheus_b200.render_loss_terms" if args.loss == "fused" else "ATen ops")}

    # side figure: the same forward step replayed from a CUDA graph (GraphedDecodeRender; inputs resident, checks
    # deferred and validated after the loop like the headline loop)
    if not args.no_graph:
        from prometheus_b200 import GraphedDecodeRender

        gr = GraphedDecodeRender(conv, rend, S, S)
        ins = (params_d, cams_in_d, cams_out_d)
        gr.run(*ins)
        g_ms, g_launches, _ = timed(lambda: gr.run(*ins, check=False), args.steps, max(3, args.warmup))
        assert gr.validate(), "the graphed replay overflowed its captured buffers"
        g_ms /= args.steps
        extra["graph_replay"] = {"value": world * pixels / (g_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": round(g_ms, 4),
                                 "gpu_launches": g_launches,
                                 "note": "decode+render captured once, replayed as one CUDA graph (fwd, inputs resident)"}
        del gr

    # side figure (measured last so that its allocations cannot disturb the figures above; not the headline, whose maps are fp32 as BASELINE's config says): the same loop fed with the
    # maps in bf16, the storage type an autocast decoder hands over; the unproject kernel reads them natively
    half = None
    if not args.no_half_maps and params_h.numel() * 4 <= (1 << 30):  # bounded extra pinned memory
        maps_bf16 = params_h.to(torch.bfloat16).pin_memory()
        e2e_loop(3, maps_bf16)
        barrier()
        gc.disable()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        e2e_loop(args.steps, maps_bf16)
        h1.record()
        barrier()
        gc.enable()
        half_ms = h0.elapsed_time(h1)
        if world > 1:
            t = torch.tensor([half_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            half_ms = float(t.item())
        half_ms /= args.steps
        half = {"value": world * pixels / (half_ms * 1e-3) / 1e6, "ms_per_step": round(half_ms, 4),
                "h2d_bytes_per_step": maps_bf16.numel() * 2 + cams_in_h.numel() * 4 + cams_out_h.numel() * 4,
                "note": "decoder maps rounded to bf16 on the host; a different input precision, reported beside the "
                        "fp32 headline only"}
        del maps_bf16
    if half is not None:
        e2e["bf16_maps"] = half
        # side figure: bf16 maps in AND compact outputs (uint8 frames + fp16 depth / mask, DecodeRenderPipeline(
        # compact_outputs=True)): 7 instead of 20 bytes per pixel back over PCIe
        maps_bf16 = params_h.to(torch.bfloat16).pin_memory()
        cpipe = DecodeRenderPipeline(conv, rend, S, S, device=dev, graph=args.pipeline == "graph", compact_outputs=True)
        cout_h = [torch.empty(B, V_out, S, S, 3, dtype=torch.uint8).pin_memory(),
                  torch.empty(B, V_out, 1, S, S, dtype=torch.float16).pin_memory(),
                  torch.empty(B, V_out, 1, S, S, dtype=torch.float16).pin_memory()]
        e2e_loop(3, maps_bf16, cpipe, cout_h)
        barrier()
        gc.disable()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        e2e_loop(args.steps, maps_bf16, cpipe, cout_h)
        h1.record()
        barrier()
        gc.enable()
        c_ms = h0.elapsed_time(h1)
        if world > 1:
            t = torch.tensor([c_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c_ms = float(t.item())
        c_ms /= args.steps
        e2e["compact"] = {"value": world * pixels / (c_ms * 1e-3) / 1e6, "ms_per_step": round(c_ms, 4),
                          "h2d_bytes_per_step": maps_bf16.numel() * 2 + cams_in_h.numel() * 4 + cams_out_h.numel() * 4,
                          "d2h_bytes_per_step": sum(t.numel() * t.element_size() for t in cout_h),
                          "note": "bf16 maps in, uint8 frames + fp16 depth / mask out: a different I/O precision, beside the fp32 headline only"}
        del maps_bf16, cpipe, cout_h
    if pipe.graphed is not None:
        e2e["graph_captures"], e2e["failed_checks"] = pipe.graphed.n_captures, pipe.graphed.n_failed_checks

    # ---- the operator (Level-1 drop-in) beside the module fast path, inputs resident, same scene:
    #   operator_path   = prometheus_b200.rasterization(...) called ONCE for all target views (fused renderer; gsplat's
    #                     sorted meta arrays — isect_ids, flatten_ids, tiles_per_gauss — only when read: with_sorted_meta);
    #   operator_dropin = the reference's own GaussianRenderer.forward control flow (gaussians_renderer.py:475-512:
    #                     per-view .cpu(), numpy inverse, host K, C = 1 rasterization calls, final synchronize),
    #                     restated in tools/dropin_loop.py, over that operator — what install_as_gsplat() gives a
    #                     checkout whose gs_decoder.py:62-63 still builds the reference's renderer.
    if not args.no_extra and args.config in ("C1", "C2", "C3"):
        import prometheus_b200 as pb
        from tools.dropin_loop import reference_style_forward

        with torch.no_grad():
            g_all = conv(params_d, cams_in_d)
        bg1 = torch.ones(3, device=dev)

        def op_batched(read_sorted=False):
            with torch.no_grad():
                for b in range(B):
                    vmb, Kb = ops.camera_prep(cams_out_d[b], S, S)
                    meta = pb.rasterization(g_all[0][b], g_all[4][b], g_all[3][b], g_all[2][b, :, 0], g_all[1][b, :, 0], vmb, Kb, S, S,
                                            backgrounds=bg1.expand(V_out, 3), render_mode="RGB+ED", packed=False)[2]
                    if read_sorted:
                        meta["isect_ids"]  # the sorted arrays are computed when one of them is read

        def op_dropin():
            with torch.no_grad():
                reference_style_forward(pb.rasterization, cams_out_d, g_all, bg1, S, S)

        osteps = max(3, args.steps // 2)
        op_ms, op_l, _ = timed(op_batched, osteps, 3)
        fs_ms, fs_l, _ = timed(lambda: op_batched(True), osteps, 3)
        dr_ms, dr_l, _ = timed(op_dropin, osteps, 3)
        op_ms, fs_ms, dr_ms = op_ms / osteps, fs_ms / osteps, dr_ms / osteps
        render_ms = ms_per_step - stage_avg.get("unproject_fwd", 0.0)
        extra["operator_path"] = {"ms_per_step": round(op_ms, 4), "value": world * pixels / (op_ms * 1e-3) / 1e6, "unit": UNIT,
                                  "gpu_launches": op_l, "vs_module_fast_path": round(op_ms / render_ms, 3),
                                  "note": "rasterization(C = all views), render only (Gaussians resident), compared with the module "
                                          "step minus its unproject: exact-size two-pass binning (one count read-back, as gsplat) + "
                                          "the fused renderer; gsplat's sorted meta arrays are produced when read",
                                  "with_sorted_meta": {"ms_per_step": round(fs_ms, 4), "gpu_launches": fs_l,
                                                       "note": "the same call + meta['isect_ids'] read: projection with per-Gaussian "
                                                               "tile counts, binning and the full per-tile sort on top"}}
        extra["operator_dropin"] = {"ms_per_step": round(dr_ms, 4), "value": world * pixels / (dr_ms * 1e-3) / 1e6, "unit": UNIT,
                                    "gpu_launches": dr_l, "vs_module_fast_path": round(dr_ms / render_ms, 3),
                                    "note": "reference-style host loop (tools/dropin_loop.py): per-view .cpu() + numpy inverse + C = 1 "
                                            "operator call + final synchronize"}
        del g_all

    # ---- parity of THIS run against the oracle (rank 0, N = 1): target view 0 of the benchmark scene through
    # oracle/gs_oracle.c on the Gaussians / camera the GPU path used; the image compared is the timed fast path's
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        parity = parity_check(step_fwd(params_d, cams_in_d, cams_out_d), conv, ops, params_d, cams_in_d, cams_out_d, S, dev,
                              workload_name(args.config, scene))

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on all host cores, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import gs_oracle as O

        O.build()
        k = 1
        gauss, t_u, t_r = cpu_reference_step(O, scene, None, k)       # warm-up + measure view 0
        gauss, _, t_r2 = cpu_reference_step(O, scene, gauss, k)
        t_r = min(t_r, t_r2)
        t = t_u * k / V_out + t_r
        cpu_baseline = {"value": k * S * S / t / 1e6, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                        "sample": "oracle/gs_oracle.c (OpenMP): full unproject %.2fs (pro-rated) + 1 of %d target "
                                  "views rendered in %.2fs" % (t_u, V_out, t_r)}
        # BASELINE.json also names the pure-PyTorch rasterisation path (gsplat's _torch_impl analogue):
        # oracle/torch_ref.py on all host cores, bounded sample = 1 target view of config C1.
        try:
            cpu_baseline["torch_restatement"] = torch_cpu_baseline(O)
        except Exception as e:  # noqa: BLE001
            cpu_baseline["torch_restatement"] = {"error": str(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.config == "C4" else "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.config, scene), "parallelism": "scene-sharded x%d" % world,
                       "numa_bound": numa_bound,
                       "l2": "256 MiB buffer written between timed steps (L2 flush, outside the events)",
                       "checks": "intersection-capacity / sort-overflow checks deferred and validated after the timed "
                                 "loop (fwd) or after loss.backward() is queued (fwd+bwd); all passed"},
            "gaussians_per_s": world * B * V_out * G / (ms_per_step * 1e-3),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity": parity, "stages": stages,
            "n_isects": I, "n_isects_needed": I_need, "n_isects_binned": I_bin, "n_isects_needed_binned": I_need_bin,
        }
        line.update(extra)
    # ---- the other multi-GPU configurations of BASELINE.json ride on the same line: C5 (input views and trajectory
    # frames sharded, ONE exchange of 56 B/Gaussian — the path's only collective) and C4 (64 scenes sharded by scene)
    if not args.no_extra and args.config == "C2":
        torch.cuda.empty_cache()
        xs = max(3, args.steps // 4)
        c5 = run_c5(args, rank, world, dev, xs, 2)
        torch.cuda.empty_cache()
        c4 = run_c4(args, rank, world, dev, max(2, args.steps // 10), 1)
        if rank == 0:
            keep = ("value", "unit", "ms_per_step", "scaling", "e2e", "exchange", "gpu_launches", "config", "steps", "warmup")
            line["c5"] = {k: c5[k] for k in keep if k in c5}
            line["c4"] = {k: c4[k] for k in keep if k in c4}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_c5(args, rank, world, dev, steps, warmup):
    """Config C5 (large scene): 32 input views x 512^2 -> 8.4M Gaussians, 120-frame 1024^2 trajectory.
    Input views AND target frames are sharded over the ranks; the one exchange step is a single NCCL
    all-gather of the packed per-rank Gaussian block (56 B/Gaussian), which each rank's unproject
    kernel writes in place. Strong scaling: total work is fixed."""
    import torch
    import torch.distributed as dist

    from prometheus_b200 import GaussianRenderer, ops, parallel, scenes

    cfg = scenes.CONFIGS["C5"]
    V_in, S_in, V_out, S = cfg["views_in"], cfg["size_in"], cfg["views_out"], cfg["size_out"]
    v0, v1 = parallel.shard_range(V_in, world, rank)
    assert (v1 - v0) * world == V_in, "C5 needs the input views to divide evenly over the ranks"
    params = torch.stack([scenes.make_params(1, S_in, 4321 + v)[0] for v in range(v0, v1)]).to(dev)
    cams_in = scenes.make_cameras(V_in, S_in, 4321)[v0:v1].to(dev)
    cams_out = scenes.make_cameras(V_out, S, 4322, first_identity=False).to(dev)
    rend = GaussianRenderer(background=[1, 1, 1]).to(dev)
    muls = torch.ones(14, device=dev)
    n_local = (v1 - v0) * S_in * S_in
    f0, f1 = parallel.shard_range(V_out, world, rank)

    # exchange step: fused unproject + all-gather over NVLink peer memory (parallel.PeerGather); the NCCL
    # all-gather path stays selectable (--exchange nccl) and is used if symmetric memory cannot be set up
    peer, exchange = None, "none (1 GPU)"
    if world > 1:
        exchange = "nccl all_gather_into_tensor"
        if args.exchange == "p2p":
            try:
                peer = parallel.PeerGather(n_local, dev)
                exchange = "fused unproject + peer-memory stores (symmetric memory, 2 device barriers)"
            except Exception as e:  # noqa: BLE001
                exchange += " (peer memory unavailable: %s)" % str(e)[:80]
    ucfg = ("mvsplat", scenes.Z_NEAR, scenes.Z_FAR, scenes.S_MIN, scenes.S_MAX)

    def step():
        with torch.no_grad():
            if peer is not None:
                full = peer.gather(lambda n, ptrs: ops.unproject_fwd_multi(params, cams_in, muls, *ucfg, n, ptrs))
            else:
                full = parallel.gather_gaussians(
                    lambda slots: ops.unproject_fwd(params, cams_in, muls, *ucfg, out=slots), n_local, dev)
            xyz, rgb, opac, scl, quat = full
            g = (xyz[None], rgb[None, :, None, :], opac[None, :, None], scl[None], quat[None])
            return rend(cams_out[None, f0:f1], g, bg_color=None, h=S, w=S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def device_ms(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / n)

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    n0 = ops.LAUNCHES
    ms = device_ms(step, steps)
    n_launch = ops.LAUNCHES - n0
    clocks = sampler.stop()

    # ---- the exchange step isolated: fused unproject + peer stores (and the NCCL all-gather form) against the plain
    # unproject of the local views, and a bit-for-bit self-check of the two exchange forms (the pytest version needs
    # two GPUs and is skipped on the 1-GPU test box)
    xch = None
    with torch.no_grad():
        local_ms = device_ms(lambda: ops.unproject_fwd(params, cams_in, muls, *ucfg), 10)
        if world > 1:
            nccl_fn = lambda: parallel.gather_gaussians(lambda slots: ops.unproject_fwd(params, cams_in, muls, *ucfg, out=slots), n_local, dev)
            nccl_fn()
            nccl_ms = device_ms(nccl_fn, 10)
            sent = n_local * 56 * (world - 1)
            xch = {"local_unproject_ms": round(local_ms, 4), "nccl_ms": round(nccl_ms, 4), "nvlink_bytes_per_rank": sent,
                   "nvlink_floor_ms": round(sent / 900e9 * 1e3, 4), "floor": "bytes each rank sends (= receives) / 900 GB/s per direction"}
            if peer is not None:
                p2p_fn = lambda: peer.gather(lambda n, ptrs: ops.unproject_fwd_multi(params, cams_in, muls, *ucfg, n, ptrs))
                p2p_fn()
                xch["p2p_ms"] = round(device_ms(p2p_fn, 10), 4)
                xch["p2p_multicast"] = bool(peer.multicast)
                a, b = p2p_fn(), nccl_fn()
                same = all(torch.equal(x, y) for x, y in zip(a, b))
                flag = torch.tensor([0 if same else 1], device=dev)
                dist.all_reduce(flag)
                xch["self_check"] = "p2p gather == nccl all-gather, bit for bit, on every rank" if int(flag.item()) == 0 else "MISMATCH"
                assert int(flag.item()) == 0, "fused peer-memory exchange differs from the NCCL all-gather"
        else:
            xch = {"local_unproject_ms": round(local_ms, 4), "note": "1 GPU: no exchange"}
    # e2e of this config = the trajectory video API (prometheus_b200.TrajectoryRenderer) behind the decode: every rank
    # unprojects its input views into the gathered scene (the exchange step), samples the poses of the video schedule
    # on the device along a path through the target cameras, renders ITS share of the frames in chunks, packs them to
    # uint8 and copies them (+ fp32 depth) to pinned host memory on a side stream. Wall clock around the whole
    # call between barriers, max over ranks.
    from prometheus_b200 import TrajectoryRenderer

    tr = TrajectoryRenderer(rend)
    path = cams_out[::15].contiguous()  # 8 key poses of the arc
    frames_h = torch.empty(f1 - f0, S, S, 3, dtype=torch.uint8).pin_memory()
    depths_h = torch.empty(f1 - f0, S, S, 1).pin_memory()

    def e2e_step():
        with torch.no_grad():
            if peer is not None:
                full = peer.gather(lambda n, ptrs: ops.unproject_fwd_multi(params, cams_in, muls, *ucfg, n, ptrs))
            else:
                full = parallel.gather_gaussians(
                    lambda slots: ops.unproject_fwd(params, cams_in, muls, *ucfg, out=slots), n_local, dev)
            xyz, rgb, opac, scl, quat = full
            g = (xyz[None], rgb[None, :, None, :], opac[None, :, None], scl[None], quat[None])
            tr.render(g, path, num_frames=V_out, render_size=S, frames_out=frames_h, depths_out=depths_h,
                      frame_range=(f0, f1))

    e2e_step()
    e2e_step()
    barrier()
    n_e2e = max(2, steps)
    t0 = time.perf_counter()
    laps = []
    for _ in range(n_e2e):
        e2e_step()
        if os.environ.get("PM_BENCH_DEBUG"):
            torch.cuda.synchronize()
            laps.append(round((time.perf_counter() - t0) * 1e3, 1))
    barrier()
    dt = (time.perf_counter() - t0) / n_e2e
    if laps:
        print("[debug] rank %d e2e laps (cumulative ms): %s" % (rank, laps), file=sys.stderr)
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": V_out * S * S / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(path.numel() * 4),
           "d2h_bytes_per_step": int(frames_h.numel() + depths_h.numel() * 4), "ms_per_step": round(dt * 1e3, 3),
           "api": "unproject + exchange + TrajectoryRenderer.render(frame_range = this rank's frames): device-side pose "
                  "sampling, uint8 frames + fp32 depth to pinned host; d2h bytes are per rank"}
    return {
        "e2e": e2e,
        "metric": METRIC, "value": V_out * S * S / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5: 32 views x 512^2 RGB-D -> 8388608 Gaussians, 120-frame 1024^2 trajectory",
                   "parallelism": "input views and target frames sharded x%d; exchange of %d B/rank: %s" % (
                       world, n_local * 56, exchange),
                   "l2": "working set >> L2 (8.4M Gaussians = 470 MB, 1.3 GB of sort records per 8-frame chunk)"},
        "exchange": xch,
        "gaussians_per_s": V_out * V_in * S_in * S_in / (ms * 1e-3), "clocks": clocks,
        "gpu_launches": n_launch}


def run_c4(args, rank, world, dev, steps, warmup):
    """Config C4 (text-to-3D batch): 64 scenes x (8 x 512^2 -> 2.1M Gaussians -> 8 views 512^2), sharded by scene over
    the ranks (64 / N each, no data-path collective): strong scaling. Each rank generates min(8, 64 / N) distinct
    scenes and tiles them to its share (generating 64 distinct scenes on one rank costs minutes of host time and
    changes nothing the GPU sees: every scene has its own buffers)."""
    import torch
    import torch.distributed as dist

    from prometheus_b200 import DecodeRenderPipeline, GaussianConverter, GaussianRenderer, ops, scenes

    assert 64 % world == 0
    n_local = 64 // world
    n_gen = min(8, n_local)
    base = scenes.make_scene("C4", seed=1234 + 1000 * rank, scenes=n_gen)
    rep = n_local // n_gen
    S = base.size_out
    V_out = base.cameras_out.shape[1]
    params_h = base.params.repeat(rep, 1, 1, 1, 1).pin_memory()
    cin_h = base.cameras_in.repeat(rep, 1, 1).pin_memory()
    cout_h = base.cameras_out.repeat(rep, 1, 1).pin_memory()
    conv = GaussianConverter("mvsplat", z_near=scenes.Z_NEAR, z_far=scenes.Z_FAR, s_min=scenes.S_MIN, s_max=scenes.S_MAX).to(dev)
    rend = GaussianRenderer(background=[1, 1, 1]).to(dev)
    pixels = 64 * V_out * S * S

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    params_d, cin_d, cout_d = params_h.to(dev), cin_h.to(dev), cout_h.to(dev)

    def step():
        with torch.no_grad():
            return rend(cout_d, conv(params_d, cin_d), bg_color=None, h=S, w=S)

    step()
    rend.defer_checks = True
    for _ in range(warmup):
        step()
    barrier()
    n0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    n_launch = ops.LAUNCHES - n0
    assert rend.validate(), "a deferred capacity / overflow check failed inside the timed region (C4)"
    rend.defer_checks = False
    ms = max_over_ranks(e0.elapsed_time(e1) / steps)
    del params_d, cin_d, cout_d
    # e2e: the host batch streams through DecodeRenderPipeline.run_batch in scene_chunk-sized slices
    out_h = [torch.empty(n_local, V_out, c, S, S).pin_memory() for c in (3, 1, 1)]
    pipe = DecodeRenderPipeline(conv, rend, S, S, device=dev, graph=args.pipeline == "graph")
    ins = (params_h, cin_h, cout_h)
    pipe.run_batch(ins, out_h)
    pipe.flush()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        pipe.run_batch(ins, out_h)
    pipe.flush()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / steps * 1e3)
    h2d = sum(t.numel() * 4 for t in ins)
    d2h = sum(t.numel() * 4 for t in out_h)
    G = base.n_gaussians
    return {
        "metric": METRIC, "value": pixels / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "scaling": "strong", "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4: 64 scenes x 8 views x 512^2 RGB-D -> 64 x %d Gaussians, 8 target views 512^2 each" % G,
                   "parallelism": "scene-sharded x%d (%d scenes per rank, %d distinct, no data-path collective)" % (world, n_local, n_gen),
                   "l2": "working set >> L2 (%.1f GB of maps per rank)" % (params_h.numel() * 4 / 1e9)},
        "gaussians_per_s": 64 * V_out * G / (ms * 1e-3), "gpu_launches": n_launch,
        "e2e": {"value": pixels / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": round(e2e_ms, 3),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "DecodeRenderPipeline.run_batch (pinned host batch in %d-scene slices, 3 streams); bytes are per rank; "
                       "wall clock between barriers, max over ranks" % max(1, min(rend.scene_chunk, -(-n_local // 4)))}}


def parity_check(outs, conv, ops, params_d, cams_in_d, cams_out_d, S, dev, workload):
    """`parity` key of the bench line: view 0 of scene 0 of the benchmark workload, GPU vs oracle/gs_oracle.c on the same
    Gaussians and camera. Integers (radii, tiles_per_gauss, isect_ids, flatten_ids, isect_offsets) must be equal; the
    fast path's image / alpha / expected depth are compared on the pixels the oracle does not flag fragile."""
    import numpy as np
    import torch

    from oracle import gs_oracle as O
    from prometheus_b200.rasterization import _rasterize_core

    O.build()
    images, depths, masks = outs[:3]
    with torch.no_grad():
        g = conv(params_d[:1], cams_in_d[:1])
        vm, K = ops.camera_prep(cams_out_d[0, :1], S, S)
        ga = [t.contiguous() for t in (g[0][0], g[4][0], g[3][0], g[2][0, :, 0], g[1][0, :, 0])]
        r, a, meta = _rasterize_core(*ga, vm, K, S, S, backgrounds=torch.ones(1, 3, device=dev), render_mode="RGB+D", packed=False)
    st = O.forward_pipeline(*[t.cpu().numpy() for t in ga], vm.cpu().numpy(), K.cpu().numpy(), S, S,
                            backgrounds=np.ones((1, 3), np.float32), with_depth=True)
    ints = {k: bool(np.array_equal(meta[k].cpu().numpy(), st[k])) for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets")}
    keep = st["fragile"][0] == 0
    rgb = ((images[0, 0].permute(1, 2, 0) + 1) / 2).cpu().numpy().astype(np.float64)
    alpha = masks[0, 0, 0].cpu().numpy().astype(np.float64)
    ed = depths[0, 0, 0].cpu().numpy().astype(np.float64)
    ref_rgb, ref_a = st["render"][0, ..., :3].astype(np.float64), st["alphas"][0].astype(np.float64)
    ref_ed = st["render"][0, ..., 3].astype(np.float64) / np.maximum(ref_a, 1e-10)
    max_abs, max_rel, worst = 0.0, 0.0, 0.0
    for got, ref in ((rgb, ref_rgb), (alpha, ref_a), (ed, ref_ed)):
        d = np.abs(got - ref)[keep]
        b = np.abs(ref)[keep]
        max_abs = max(max_abs, float(d.max()))
        max_rel = max(max_rel, float((d / np.maximum(b, 1e-3)).max()))
        worst = max(worst, float((d / (1e-5 + 1e-4 * b)).max()))
    return {"config": workload + " — scene 0, target view 0 (%d intersections)" % len(st["isect_ids"]),
            "oracle": "oracle/gs_oracle.c (restatement of gsplat 1.4.0; parity unpinned for stages 2-4)",
            "ints_equal": all(ints.values()), "ints": ints, "max_abs": max_abs, "max_rel": max_rel,
            "worst_ratio_to_1e-5_abs_plus_1e-4_rel": worst, "within_tolerance": worst <= 1.0,
            "fragile_frac": float(1.0 - keep.mean()),
            "compared": "rgb, alpha, expected depth of the timed fast path on non-fragile pixels; operator-path integers"}


def meta_last_ids(r, meta, ops, g, vm, K, S, V_out, dev):
    """last_ids of the forward (needed for I_need): re-run the composite stage alone."""
    import torch

    _, _, last, _ = ops.composite_fwd(meta["means2d"].contiguous(), meta["conics"].contiguous(), g[2][0, :, 0].contiguous(),
                                   g[1][0, :, 0].contiguous(), meta["depths"].contiguous(),
                                   torch.ones(V_out, 3, device=dev), meta["isect_offsets"], meta["flatten_ids"], S, S)
    return last


def fast_path_counts(ops, g, vm, K, S, V_out, dev, I, max_count, usable):
    """{records binned, records at or before the last one a pixel of their tile needs} of the module fast path on scene 0:
    single-pass binning (of the usable tiles when `usable`) + the fused renderer, by hand."""
    import torch

    opac, col = g[2][0, :, 0].contiguous(), g[1][0, :, 0].contiguous()
    out = ops.project_fwd(g[0][0], g[4][0], g[3][0], vm, K, S, S, count_tiles=False, tile_histogram=True,
                          tile_capacity=max_count, opacities=opac if usable else None)
    means2d, depths, conics, counts, records = out[1], out[2], out[3], out[5], out[6]
    offs, stats, order = ops.tile_scan(counts, want_order=True)
    last = ops.tile_render_packed_fwd(records, offs, means2d, conics, opac, col, depths, torch.ones(V_out, 3, device=dev), S, S,
                                      d_n=stats, tile_order=order, record_stride=max_count, capacity=I)[3]
    n = int(stats[0].item())
    off = offs.flatten().long()
    cnt = counts.flatten().long()
    tiles_last = last.reshape(V_out, S // 16, 16, S // 16, 16).amax(dim=(2, 4)).flatten().long()
    need = int(((tiles_last - off + 1).clamp(min=0) * (cnt > 0)).sum().item())
    return n, need


if __name__ == "__main__":
    main()
