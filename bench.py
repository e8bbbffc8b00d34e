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
    k = args.ref_views  # bounded sample: k of the V target views per step (+ the full unproject)
    times = []
    for it in range(args.warmup + args.steps):
        _, t_u, t_r = cpu_reference_step(O, scene, None, k)
        if it >= args.warmup:
            times.append(t_u * k / V + t_r)  # unproject cost attributed pro rata to the rendered views
    t = sum(times) / len(times)
    value = k * S * S / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, scene), "l2": "n/a (CPU)"},
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
    ap.add_argument("--ref-views", type=int, default=1)
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
        run_c5(args, rank, world, dev)
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
    tb, cb = ops.key_bits(S // 16, S // 16, V_out)
    passes = (32 + tb + cb + 7) // 8
    bytes_alg = {
        "unproject_fwd": 112 * G * B,
        # module fast path: projection with the per-tile histogram fused in, no tiles_per_gauss array
        "project_fwd": (40 * G + 28 * P) * B,
        "scan_tiles": 12 * P * B,
        "bin_emit": (20 * P + 12 * I) * B,
        "sort_pairs": (passes * 24 + 8) * I * B,
        "tile_offsets": 8 * I * B,
        "bin_count": 12 * P * B,
        "bin_scatter": (16 * P + 8 * I) * B,
        "tile_sort": 20 * I * B,
        # fused sort-as-you-go: one pass over the tile's records at minimum + composite traffic
        "tile_render_fwd": (8 * I + 4 * I_need + 44 * I_need + 24 * X) * B,
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
    traffic = None  # DRAM bytes per launch of that kernel from the committed ncu capture (profiles/traffic.json)
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom) if args.config == "C2" else None
    except Exception:  # noqa: BLE001
        pass
    # the kernel's real ceiling is instruction issue: warp instructions of the committed ncu capture /
    # (live launch time x 148 SMs x 4 schedulers x sampled SM clock)
    issue = None
    try:
        wi = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("_warp_instructions", {}).get(dom)
        if wi and args.config == "C2" and clocks.get("sm_mhz"):
            peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
            issue = {"warp_instructions": wi, "peak_warp_inst_per_s": peak_issue,
                     "frac": round(wi / (dom_ms * 1e-3) / peak_issue, 4), "source": "ncu smsp__inst_executed.sum (profiles/)"}
    except Exception:  # noqa: BLE001
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "issue": issue,
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
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

    def e2e_loop(n, maps_h=None):
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
    if world == 1 and not args.no_half_maps and params_h.numel() * 4 <= (1 << 30):  # bounded extra pinned memory
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
        half_ms = h0.elapsed_time(h1) / args.steps
        half = {"value": pixels / (half_ms * 1e-3) / 1e6, "ms_per_step": round(half_ms, 4),
                "h2d_bytes_per_step": maps_bf16.numel() * 2 + cams_in_h.numel() * 4 + cams_out_h.numel() * 4,
                "note": "decoder maps rounded to bf16 on the host; a different input precision, reported beside the "
                        "fp32 headline only"}
        del maps_bf16
    if half is not None:
        e2e["bf16_maps"] = half
    if pipe.graphed is not None:
        e2e["graph_captures"], e2e["failed_checks"] = pipe.graphed.n_captures, pipe.graphed.n_failed_checks

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
            "cpu_baseline": cpu_baseline, "stages": stages,
            "n_isects": I, "n_isects_needed": I_need,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_c5(args, rank, world, dev):
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

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    n0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms = total_ms / args.steps
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
    n_e2e = max(2, args.steps)
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
    if rank == 0:
        print(json.dumps({
            "e2e": e2e,
            "metric": METRIC, "value": V_out * S * S / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5: 32 views x 512^2 RGB-D -> 8388608 Gaussians, 120-frame 1024^2 trajectory",
                       "parallelism": "input views and target frames sharded x%d; exchange of %d B/rank: %s" % (
                           world, n_local * 56, exchange),
                       "l2": "working set >> L2 (8.4M Gaussians = 470 MB, 1.3 GB of sort records per 8-frame chunk)"},
            "gaussians_per_s": V_out * V_in * S_in * S_in / (ms * 1e-3), "clocks": clocks,
            "gpu_launches": ops.LAUNCHES - n0}), flush=True)


def meta_last_ids(r, meta, ops, g, vm, K, S, V_out, dev):
    """last_ids of the forward (needed for I_need): re-run the composite stage alone."""
    import torch

    _, _, last = ops.composite_fwd(meta["means2d"].contiguous(), meta["conics"].contiguous(), g[2][0, :, 0].contiguous(),
                                   g[1][0, :, 0].contiguous(), meta["depths"].contiguous(),
                                   torch.ones(V_out, 3, device=dev), meta["isect_offsets"], meta["flatten_ids"], S, S)
    return last


if __name__ == "__main__":
    main()
