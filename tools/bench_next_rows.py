"""Times the "next" rows of SURVEY.md 8f that have device kernels -- topology edits (densify) and the scene
writers -- on one GPU with CUDA events, next to their algorithmic HBM bytes, and the reference's CPU
implementation of the same work (oracle/scene_edit.py = the reference's ATen op sequence) on the host.

    python tools/bench_next_rows.py [--n 1000000] [--reps 20] [--no-cpu]"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.add_(1.0)                      # > L2: evict the inputs between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    from opensplat_b200 import densify, export
    from util import scene_edit_inputs
    dev = "cuda:0"
    n, k = a.n, 16
    p, m, v, draws = scene_edit_inputs(n, k, 3)
    P = {x: torch.from_numpy(y).to(dev) for x, y in p.items()}
    M = {x: torch.from_numpy(y).to(dev) for x, y in m.items()}
    V = {x: torch.from_numpy(y).to(dev) for x, y in v.items()}
    flush = torch.zeros(64 * 1024 * 1024, device=dev)       # 256 MB
    cfg = densify.RefineConfig()
    dn = densify.Densifier(cfg)
    H, W = 480, 640
    for v_xy, radii in draws:
        dn.accumulate(torch.from_numpy(v_xy).to(dev), torch.from_numpy(radii).to(dev), H, W)
    out = {"n": n, "sh_bases": k}
    v_xy, radii = torch.from_numpy(draws[0][0]).to(dev), torch.from_numpy(draws[0][1]).to(dev)
    t = timed(lambda: dn.accumulate(v_xy, radii, H, W), a.reps, flush)
    out["densify_stats_update"] = {"ms": t, "alg_bytes": 36 * n, "GBps": 36 * n / t / 1e6}
    res = {}

    def do_classify():
        res["c"] = densify.classify(P["scales"], P["opacities"], dn.xys_grad_norm, dn.vis_counts, dn.max_2d_size,
                                    max(H, W), cfg, True, True, True)
    t = timed(do_classify, a.reps, flush)
    src_map, split_rank, counts = res["c"]
    cnt = counts.cpu().tolist()
    n_splits, new_n = cnt[0], cnt[4]
    by = 28 * n + 4 * new_n + 4 * n                # scales 12, opacity 4, 3 stats 12 ; src_map + split_rank written
    out["densify_classify"] = {"ms": t, "alg_bytes": by, "GBps": by / t / 1e6, "n_splits": n_splits, "new_n": new_n,
                               "n_dups": cnt[5]}
    samples = torch.randn(2 * n_splits, 3, device=dev)

    def do_apply():
        densify.means_scales(src_map, split_rank, new_n, n_splits, samples, P["means"], P["scales"], P["quats"],
                             cfg.size_fac)
        for x in ("quats", "featuresDc", "featuresRest", "opacities"):
            densify.gather_rows(src_map, new_n, P[x])
        for S in (M, V):
            for x in S:
                densify.gather_rows(src_map, new_n, S[x], zero_children=True)
    t = timed(do_apply, a.reps, flush)
    row = 59 * 4                                    # floats per Gaussian over the six tensors
    by = 3 * (2 * row * new_n) + 12 * 4 * new_n     # read + write of params, m, v ; the row map re-read per tensor
    out["densify_apply_params_and_adam"] = {"ms": t, "alg_bytes": by, "GBps": by / t / 1e6}
    t = timed(lambda: export.pack_ply_rows(P), a.reps, flush)
    by = 2 * 59 * 4 * n + 12 * n
    out["pack_ply_rows"] = {"ms": t, "alg_bytes": by, "GBps": by / t / 1e6}
    t = timed(lambda: export.pack_splat_rows(P), a.reps, flush)
    by = (16 + 44 + 32 + 24) * n                    # key pass, row inputs, rows, ~sort traffic (4 passes x (8+4) B ...)
    out["pack_splat_rows_sorted"] = {"ms": t, "alg_bytes_min": by, "GBps_min": by / t / 1e6}
    if not a.no_cpu:
        from oracle import scene_edit as se
        ncpu = min(n, 200_000)
        pc = {x: y[:ncpu] for x, y in p.items()}
        mc = {x: y[:ncpu] for x, y in m.items()}
        vc = {x: y[:ncpu] for x, y in v.items()}
        stats = None
        for vxy, rad in draws:
            stats = se.densify_stats(stats, vxy[:ncpu], rad[:ncpu], H, W)
        ocfg = types.SimpleNamespace(**{f: getattr(cfg, f) for f in (
            "densify_grad_thresh", "densify_size_thresh", "split_screen_size", "cull_alpha_thresh",
            "cull_scale_thresh", "cull_screen_size", "size_fac")})
        t0 = time.perf_counter()
        se.refine(pc, mc, vc, stats, max(H, W), ocfg, True, True, lambda ns: torch.randn(2 * ns, 3))
        t_ref = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        se.ply_body(pc["means"], pc["featuresDc"], pc["featuresRest"], pc["opacities"], pc["scales"], pc["quats"])
        t_ply = (time.perf_counter() - t0) * 1e3
        out["cpu_reference_ops"] = {"sample_n": ncpu, "threads": torch.get_num_threads(),
                                    "refine_ms_scaled_to_n": t_ref * n / ncpu, "ply_body_ms_scaled_to_n": t_ply * n / ncpu,
                                    "kind": "port (the reference's ATen op sequence, oracle/scene_edit.py)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
