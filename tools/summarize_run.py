"""Turns what a GPU run left in gpurun_out/ into the tracked evidence under profiles/ (run here, no GPU):
  * bench JSON lines -> profiles/r02_bench_*.json (verbatim) + a short markdown table on stdout
  * ncu launch list -> profiles/r02_launches_*.txt (tools/ncu_launches.py)
  * .ncu-rep captures -> profiles/r02_ncu_*.txt (tools/ncu_summary.py) and profiles/dram_traffic.json
    (dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by workload and stage name)
usage: python tools/summarize_run.py"""
import csv, glob, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

STAGE_OF = {"rasterize_backward_kernel": "raster_bwd", "rasterize_forward_kernel": "raster_fwd",
            "tile_dsort_pack_kernel": "tile_dsort_pack", "tile_sort_pack_kernel": "tile_sort_pack",
            "bin_count_kernel": "bin_count", "count_scan_kernel": "count_scan", "bucket_emit_kernel": "bucket_emit",
            "tile_scan_kernel": "tile_scan", "reduce_grad_rows_kernel": "reduce_grad_rows"}


def last_json(path):
    try:
        for ln in reversed(open(path).read().strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
    except Exception:
        pass
    return None


def bench_lines():
    rows = []
    for f in sorted(glob.glob(os.path.join(G, "r2_bench*.log"))):
        d = last_json(f)
        if not d:
            continue
        name = os.path.basename(f).replace("r2_", "r02_").replace(".log", ".json")
        json.dump(d, open(os.path.join(P, name), "w"), indent=1)
        rows.append((name, d))
    for name, d in rows:
        e = d.get("e2e", {})
        print(f"| `{name}` | N={d['n_gpus']} | {d['value']:.0f} | {d['ms_per_step']:.3f} | {e.get('value', 0):.0f} | "
              f"{d.get('train_iters_per_s', 0):.0f} | {d['config'].get('intersections_M')} |")
        print("   stages:", d.get("stages_ms"))
        r = d.get("roofline", {})
        print("   roofline:", r.get("kernel"), f"{r.get('achieved', 0):.0f} GB/s frac {r.get('frac', 0):.3f}",
              "pairs:", {k: (f"{v:.3g}" if isinstance(v, float) else v) for k, v in (r.get("pairs") or {}).items()})
        for wl, s in (d.get("other_configs") or {}).items():
            if "error" in s:
                print("   ", wl, "ERROR", s["error"])
            else:
                print(f"    {wl}: {s['value']:.0f} Mpixel/s {s['ms_per_step']:.3f} ms M={s['intersections_binned']} "
                      f"(ref {s['intersections_reference']}) longest {s['longest_tile_list']} stages {s['stages_ms']}")
        if d.get("per_rank") and d["n_gpus"] > 1:
            print("   per_rank:", [(p["rank"], p["intersections_binned"], p["compute_ms_without_exchange"],
                                    p["exchange_stage_ms"]) for p in d["per_rank"]])
            print("   exchange_check:", d.get("exchange_check"))
        print("   cpu_baseline:", d.get("cpu_baseline"))


def launches():
    for f in sorted(glob.glob(os.path.join(G, "r2_launches*.csv"))):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_launches.py"), f], capture_output=True,
                             text=True).stdout
        name = os.path.basename(f).replace("r2_", "r02_").replace(".csv", ".txt")
        open(os.path.join(P, name), "w").write(out)
        print(name); print(out[:1800])


def ncu_reps():
    traffic, counters = {}, {}
    tfile = os.path.join(P, "dram_traffic.json")
    for f, wl in ((os.path.join(G, "r2_prof_c2.ncu-rep"), "c2_1M_1080p_sh3"),
                  (os.path.join(G, "r2_prof_c5.ncu-rep"), "c5_5M_1440p_dense")):
        if not os.path.exists(f):
            continue
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), f], capture_output=True,
                             text=True).stdout
        name = "r02_ncu_" + wl.split("_")[0] + ".txt"
        open(os.path.join(P, name), "w").write(out)
        print(name, len(out))
        raw = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        idx = {h: i for i, h in enumerate(rows[0])}
        units = rows[1]
        def val(r, key):
            v = float(r[idx[key]].replace(",", ""))
            u = units[idx[key]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        for r in rows[2:]:
            kn = r[idx["Kernel Name"]]
            st = next((s for k, s in STAGE_OF.items() if k in kn), None)
            if st and "dram__bytes_read.sum" in idx:
                traffic.setdefault(wl, {})[st] = int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
            if st and "smsp__inst_executed.sum" in idx:
                counters.setdefault(wl, {})[st] = {
                    "warp_instructions": int(float(r[idx["smsp__inst_executed.sum"]].replace(",", ""))),
                    "issue_active_pct": float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                    "sm_throughput_pct": float(r[idx["sm__throughput.avg.pct_of_peak_sustained_elapsed"]])}
    if counters:
        json.dump(counters, open(os.path.join(P, "ncu_counters.json"), "w"), indent=1)
    if traffic:
        json.dump(traffic, open(tfile, "w"), indent=1)
        print("dram_traffic.json", traffic)


if __name__ == "__main__":
    bench_lines()
    launches()
    ncu_reps()
