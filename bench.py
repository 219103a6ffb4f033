#!/usr/bin/env python
"""bench.py -- fwd+bwd Mpixel/s (and train iters/s) of the Gaussian-splat render path on B200.

    python bench.py --gpus N --steps K --warmup W            # ours (sm_100a kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU back end (oracle/_ref)

Workload (BASELINE.json configs[1]): 1M synthetic Gaussians, 1920x1080, SH degree 3, fp32, one camera
view per GPU (data-parallel over views; weak scaling).  A "step" is one pass of the hot path over one
view: SH fwd -> project fwd -> scan/emit/sort/bins -> blend fwd -> MSE -> blend bwd -> project bwd ->
SH bwd (+ one NCCL all-reduce of the flat per-Gaussian gradient buffer when N > 1).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, W, H, scale, opacity range)
    "c2_1M_1080p_sh3": (1_000_000, 1920, 1080, 0.02, (0.05, 0.95)),
    "c3_3M_4k_sh3": (3_000_000, 3840, 2160, 0.01, (0.05, 0.95)),        # same pixel footprints as c2 (focal doubles)
    "c5_5M_1440p_dense": (5_000_000, 2560, 1440, 0.023, (0.05, 0.95)),   # ~200 blended-candidate splats / pixel
    "c1_1k_256": (1_000, 256, 256, 0.5, (0.05, 0.95)),
}
# kernels of OURS launched per fwd+bwd step (counted from the C-ABI implementations, bucket fast path):
#   project_fwd 1, cumsum 3, tile_count 1, tile_scan 1, sh_fwd(+clamp) 1, bucket_emit 1, tile_sort_pack 1,
#   blend_fwd 1, mse 1, blend_bwd 1, row_reduce 1, project_bwd 1, sh_bwd 1 (N>1 fused exchange: mask 1 + multiview 1)
def launches_per_step(world=1, fused=True, train=False):
    n = 1 + 3 + 1 + 1 + 1 + 1 + 1 + 1 + 1 + 1 + 1 + 1
    n += 2 if (world > 1 and fused) else 1
    return n + (1 if train else 0)


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(workload, steps=1, warmup=1):
    """Times the reference's own CPU back end (oracle/_ref: unmodified gsplat_cpu.cpp + operator .cpp
    files) on a bounded sample of the workload: a 1/16-area window (W/4 x H/4) holding N/16 Gaussians with
    the same per-pixel splat density and the same pixel footprint distribution (scale x4 because the focal
    length scales with W).  Returns (Mpixel/s, dict)."""
    import numpy as np
    import torch
    from oracle import ref
    from opensplat_b200.scene import make_scene
    n, W, H, scale, opac = WORKLOADS[workload]
    div = 2 if n >= 16_000 else 1
    ns, Ws, Hs = n // (div * div), W // div, H // div
    sc = make_scene(ns, Ws, Hs, scale=scale * div, sh_degree=3, opacity=opac, seed=0)
    o = ref.ops()
    # all the host threads the reference can use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    torch.set_num_threads(max(torch.get_num_threads(), min(os.cpu_count() or 1, 64)))
    t = lambda a, g=False: torch.from_numpy(np.ascontiguousarray(a)).requires_grad_(g)
    target = torch.zeros(Hs, Ws, 3)
    times = []
    for it in range(warmup + steps):
        means, scales, quats = t(sc["means"], True), t(sc["scales"], True), t(sc["quats"], True)
        coeffs, opacity = t(sc["coeffs"], True), t(sc["opacities"], True)
        t0 = time.perf_counter()
        rgbs = torch.clamp_min(o.sh_cpu(3, t(sc["viewdirs"]), coeffs) + 0.5, 0.0)
        p = o.project_cpu(means, scales, 1.0, quats, t(sc["viewmat"]), t(sc["projmat"]), sc["fx"], sc["fy"],
                          sc["cx"], sc["cy"], Hs, Ws, 0.01)
        img = o.rasterize_cpu(p[0], p[1], p[2], rgbs, opacity, p[3], p[4].contiguous(), Hs, Ws, torch.zeros(3))
        loss = torch.nn.functional.mse_loss(img, target)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    info = {"cores": int(o.num_threads()), "kind": "reference",
            "sample": f"{ns} Gaussians at {Ws}x{Hs} (1/{div*div}-area window of {workload}, same splats/pixel and "
                      f"pixel footprints), fwd+bwd, {steps} timed step(s); raster loops of the reference are "
                      f"single-threaded, ATen ops use {int(o.num_threads())} threads; host has {os.cpu_count()} cpus",
            "seconds_per_step": sec}
    return Ws * Hs / sec / 1e6, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    v, info = cpu_reference_sample(args.workload, steps=steps, warmup=warmup)
    n, W, H, scale, opac = WORKLOADS[args.workload]
    out = {"impl": "reference", "metric": "fwd_bwd_mpixel_per_s", "value": v, "unit": "Mpixel/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": info["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "gaussians": n, "width": W, "height": H, "sh_degree": 3,
                      "note": f"bounded sample, {steps} timed step(s) (requested {args.steps})"},
           "cpu_baseline": {"value": v, "unit": "Mpixel/s", "cores": info["cores"], "kind": "reference",
                            "sample": info["sample"]},
           "e2e": {"value": v, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from opensplat_b200 import capi, ops
    from opensplat_b200.pipeline import SplatPipeline
    from opensplat_b200.scene import make_scene, rotated_camera

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi.lib()  # fail loudly if the CUDA library is missing

    n, W, H, scale, opac = WORKLOADS[args.workload]
    sc = make_scene(n, W, H, scale=scale, sh_degree=3, opacity=opac, seed=0)  # same Gaussians on every rank
    pipe = SplatPipeline(n, W, H, sh_degree=3, device=dev, stage_timing=True)
    pipe.load_scene(sc)
    if world > 1:  # one camera view per GPU (config C4): orbit the scene
        cam = rotated_camera(W, H, rank, n_views=max(world, 8))
        pipe.set_camera(cam)
        vd = sc["means"] - cam["cam_pos"]
        vd = (vd / np.linalg.norm(vd, axis=-1, keepdims=True)).astype(np.float32)
        pipe.viewdirs.copy_(torch.from_numpy(vd).to(dev))
    rng = np.random.default_rng(1 + rank)
    target_host = torch.from_numpy(rng.uniform(0, 1, (H, W, 3)).astype(np.float32)).pin_memory()
    pipe.target.copy_(target_host, non_blocking=True)

    if world > 1 and args.exchange == "fused":
        from opensplat_b200.multigpu import ViewParallelExchange
        pipe.exchange = ViewParallelExchange(pipe, cam["cam_pos"])

    def step_fwd_bwd():
        pipe.forward()
        pipe.backward()       # with pipe.exchange: fused multi-view SH backward + NVLink exchange inside
        if world > 1 and pipe.exchange is None:
            dist.all_reduce(pipe.grad_flat, op=dist.ReduceOp.SUM)
        pipe._collect()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    for _ in range(max(args.warmup, 3)):
        step_fwd_bwd()
    pipe.stage_ms.clear(); pipe._steps_ev = []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_fwd_bwd, args.steps)
    stage_ms = pipe.resolve_stage_times()
    m_timed = pipe.m
    alg, passes = pipe.algorithmic_bytes()
    ms_step = ms_total / args.steps
    value = world * W * H / (ms_step * 1e-3) / 1e6

    # ---- train iters/s: + fused Adam (and the all-reduce average) ----
    pipe.stage_timing = False
    for _ in range(2):
        pipe.train_step(world_size=world)
    ms_train = timed(lambda: pipe.train_step(world_size=world), args.steps) / args.steps

    # ---- e2e: through the autograd operators (ops.py), host buffers in the timed region ----
    P = {k: pipe.p[k].detach().clone().requires_grad_() for k in ("means", "scales", "quats", "coeffs", "opacities")}
    view_host = pipe.viewmat.cpu().pin_memory()
    proj_host = pipe.projmat.cpu().pin_memory()
    fx, fy, cx, cy = pipe.intr
    bg = torch.zeros(3, device=dev)
    loss_host = torch.zeros(1).pin_memory()

    from opensplat_b200 import cpp_ops
    use_cpp = cpp_ops.available()
    cops = cpp_ops.ops() if use_cpp else None

    # H2D of every step's inputs (target image + camera) from pinned memory runs on a copy stream, one step
    # ahead of the compute (double buffer), so PCIe overlaps the kernels; it is still inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [dict(tgt=torch.empty_like(pipe.target), vm=torch.empty_like(pipe.viewmat),
                  pm=torch.empty_like(pipe.projmat), ev=torch.cuda.Event(), used=torch.cuda.Event()) for _ in range(2)]
    state = {"i": 0}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(slot["used"])      # the compute that last read this slot has finished
            slot["tgt"].copy_(target_host, non_blocking=True)
            slot["vm"].copy_(view_host, non_blocking=True)
            slot["pm"].copy_(proj_host, non_blocking=True)
            slot["ev"].record(copy_stream)

    for sl in slots:
        sl["used"].record(torch.cuda.current_stream())
    prefetch(slots[0])

    def step_e2e():
        cur = slots[state["i"] % 2]
        prefetch(slots[(state["i"] + 1) % 2])                    # next step's inputs, overlapped
        state["i"] += 1
        torch.cuda.current_stream().wait_event(cur["ev"])       # this step's H2D has landed
        tgt, vm, pm = cur["tgt"], cur["vm"], cur["pm"]
        for t in P.values():
            t.grad = None
        if use_cpp:   # the libtorch autograd operators a C++ caller of the reference API uses
            rgbs = torch.clamp_min(cops.spherical_harmonics(3, pipe.viewdirs, P["coeffs"]) + 0.5, 0.0)
            xys, depths, radii, conics, nth, _ = cops.project_gaussians(
                P["means"], P["scales"], 1.0, P["quats"], vm, pm, fx, fy, cx, cy, H, W, 0.01)
            img = cops.rasterize_gaussians(xys, depths, radii, conics, nth, rgbs, P["opacities"], H, W, bg)
        else:
            rgbs = torch.clamp_min(ops.SphericalHarmonics.apply(3, pipe.viewdirs, P["coeffs"]) + 0.5, 0.0)
            xys, depths, radii, conics, nth, _ = ops.ProjectGaussians.apply(
                P["means"], P["scales"], 1.0, P["quats"], vm, pm, fx, fy, cx, cy, H, W, pipe.tb)
            img = ops.RasterizeGaussians.apply(xys, depths, radii, conics, nth, rgbs, P["opacities"], H, W, bg)
        loss = torch.nn.functional.mse_loss(img, tgt)
        loss.backward()
        if world > 1:
            for t in P.values():
                dist.all_reduce(t.grad, op=dist.ReduceOp.SUM)
        cur["used"].record(torch.cuda.current_stream())
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)  # D2H: the step's result
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * W * H / (ms_e2e * 1e-3) / 1e6

    # ---- roofline of the dominant stage ----
    peak, peak_src = peaks()
    # the fused two-level binning stage stands for SURVEY's emit + sort + bins rows
    alg_stage = dict(alg)
    alg_stage["bucket_sort_pack"] = alg["emit"] + alg["sort"] + alg["bins"]
    dom = max(stage_ms, key=lambda k: stage_ms[k]) if stage_ms else "raster_bwd"
    dom_key = dom if dom in alg_stage else "raster_bwd"
    alg = alg_stage if dom_key == "bucket_sort_pack" else alg
    ach = alg[dom_key] / (stage_ms.get(dom_key, ms_step) * 1e-3) / 1e9
    path_bytes = sum(v for k, v in alg.items() if k != "bucket_sort_pack")
    path_ach = path_bytes / (ms_step * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get(dom_key)
        except Exception:
            traffic = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, info = cpu_reference_sample(args.workload, steps=2, warmup=1)
            cpu = {"value": v, "unit": "Mpixel/s", "cores": info["cores"], "kind": info["kind"],
                   "sample": info["sample"]}
        except Exception as ex:  # the checker is optional for the number; say why it is missing
            cpu = {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}

    out = {
        "metric": "fwd_bwd_mpixel_per_s", "value": value, "unit": "Mpixel/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "gaussians": n, "width": W, "height": H, "sh_degree": 3,
                   "intersections_M": m_timed, "views_per_gpu": 1, "parallelism": f"dp{world}-views",
                   "step": "sh+project+scan/emit/sort/bins+blend fwd, mse, blend+project+sh bwd"
                           + ((", fused multi-view SH bwd over NVLink peer loads + nccl allreduce(geometry grads)"
                               if args.exchange == "fused" else ", nccl allreduce(flat grads)") if world > 1 else ""),
                   "l2_policy": "inputs larger than L2 (>300 MB of parameters/records per step vs 126 MB L2)"},
        "train_iters_per_s": 1e3 / ms_train, "train_ms_per_iter": ms_train,
        "e2e": {"value": e2e_value, "unit": "Mpixel/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(target_host.numel() * 4 + 128), "d2h_bytes_per_step": 4 + 4,
                "api": ("C++ libtorch autograd operators ProjectGaussians/RasterizeGaussians/SphericalHarmonics "
                        "(libopensplat_b200_ops.so via torch.ops)") if use_cpp else
                       "opensplat_b200.ops python autograd operators"},
        "gpu_launches": launches_per_step(world, args.exchange == "fused") * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom_key, "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg[dom_key], "ms_per_launch": stage_ms.get(dom_key)},
        "roofline_path": {"achieved": path_ach, "peak": peak, "unit": "GB/s", "frac": path_ach / peak,
                          "algorithmic_bytes_per_step": path_bytes, "sort_passes": passes},
        "stages_ms": {k: round(v, 4) for k, v in stage_ms.items()},
        "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _watchdog(seconds):
    """A bench run that stalls (hung kernel / collective) must not block the driver: exit hard after `seconds`."""
    def fire():
        sys.stderr.write(f"bench.py watchdog: no result after {seconds}s, aborting\n")
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


if __name__ == "__main__":
    _watchdog(1500)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_1M_1080p_sh3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1: fused multi-view SH backward over peer memory (default) or plain NCCL all-reduce")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
