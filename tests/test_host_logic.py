"""CPU-only host logic: workspace size queries of the C ABI (no kernel launches), the synthetic scene generator,
the flat parameter layout."""
import numpy as np

from opensplat_b200 import capi, parallel
from opensplat_b200.scene import make_scene, rotated_camera


def test_workspace_queries_are_monotone_and_aligned():
    L = capi.lib()
    for fn in (L.gsb_sort_workspace_bytes, L.gsb_raster_records_bytes, L.gsb_raster_grad_rows_bytes,
               (lambda m: L.gsb_bucket_workspace_bytes(1000, m, 8160)), L.gsb_cumsum_workspace_bytes):
        prev = 0
        for m in (0, 1, 100, 2048, 2049, 1_000_000, 50_000_000):
            b = fn(m)
            assert b >= prev and b % 256 == 0
            prev = b
    assert L.gsb_raster_records_bytes(1000) >= 1000 * 48 + 256        # 48-B records + the scratch words
    assert L.gsb_raster_grad_rows_bytes(1000) >= 1000 * 48
    # bucket workspace: header + scan states + one 128-B counter line per tile + 48-B attribute records + 12 B / slot
    assert L.gsb_bucket_workspace_bytes(1000, 5000, 8160) >= 256 + 8160 * 128 + 1000 * 48 + 5000 * 12
    assert L.gsb_bucket_max_tile_len() == 16384
    assert L.gsb_ssim_workspace_bytes(1080, 1920) >= 3 * 1080 * 1920 * 3 * 4


def test_scene_generator_conventions():
    sc = make_scene(5000, 320, 200, scale=0.1, sh_degree=3, seed=3)
    tz = sc["means"][:, 2] + np.float32(8.0)
    assert len(np.unique(tz)) == 5000                       # strictly distinct view depths
    assert tz.min() >= 7.0 and tz.max() <= 9.0
    assert np.allclose(np.linalg.norm(sc["quats"], axis=-1), 1, atol=1e-6)
    assert np.allclose(np.linalg.norm(sc["viewdirs"], axis=-1), 1, atol=1e-6)
    assert sc["coeffs"].shape == (5000, 16, 3) and sc["opacities"].shape == (5000, 1)
    assert np.array_equal(sc["viewmat"], sc["projmat"]) and sc["viewmat"][2, 3] == 8.0   # w == 1 convention
    assert sc["cx"] == 160 and sc["cy"] == 100 and abs(sc["fx"] - 160.0) < 1e-9
    # same seed -> same scene
    assert np.array_equal(make_scene(5000, 320, 200, scale=0.1, sh_degree=3, seed=3)["means"], sc["means"])
    cam = rotated_camera(320, 200, 2, n_views=8)
    R = cam["viewmat"][:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)


def test_flat_layout_matches_pipeline_order():
    offs, total = parallel.flat_layout(1000, 16)
    assert total == 1000 * 59
    names = sorted(offs, key=lambda k: offs[k][0])
    assert names == ["means", "scales", "quats", "opacities", "coeffs"]   # geometry prefix, SH coefficients last
    assert offs["coeffs"][0] == 1000 * 11
    # odd Gaussian counts (any refinement can leave one): every slice still starts on a 16-byte boundary
    for n in (1, 3, 257, 1001, 99_999):
        offs, total = parallel.flat_layout(n, 16)
        assert all(o % 4 == 0 for o, _, _ in offs.values()) and total % 4 == 0
        ends = sorted((o, o + c) for o, c, _ in offs.values())
        assert all(a[1] <= b[0] for a, b in zip(ends[:-1], ends[1:])) and ends[-1][1] <= total


def test_refine_schedule_matches_reference_defaults():
    """Model::afterTrain's step arithmetic (model.cpp:339-341,349,441,472) with the CLI defaults (opensplat.cpp:30-43):
    refine every 100 steps after 500, reset interval 3000, densify unless within numCameras+100 steps after a reset
    boundary, stop splitting at 15000, screen-size rules until 4000, huge cull after 3000."""
    from opensplat_b200.densify import Densifier, RefineConfig
    dn = Densifier(RefineConfig(num_cameras=50))
    sch = dn.schedule
    assert sch(499)[0] is False and sch(500)[0] is False and sch(550)[0] is False
    assert sch(600)[:3] == (True, True, False)                   # 600 % 3000 = 600 > 150
    assert sch(3000)[:3] == (True, False, False)                 # 0 > 150 is false: no densification on the boundary
    assert sch(3100)[:3] == (True, False, True)                  # alpha reset at boundary + refineEvery
    assert sch(3200)[:3] == (True, True, False)
    assert sch(3900)[3] is True and sch(4000)[3] is False        # screen-size rules
    assert sch(3000)[4] is False and sch(3100)[4] is True        # huge cull only after refineEvery * resetAlphaEvery
    assert sch(15000)[:3] == (True, False, False) and sch(14900)[1] is True
    assert RefineConfig(max_steps=30001).stop_split_at == 15000


def test_committed_bench_line_has_the_contract_keys():
    """The last bench line measured on the B200 (profiles/r02_bench_final.json) carries every key of the bench
    contract; guards bench.py's output format against accidental regressions between GPU runs."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = [json.loads(l) for l in open(os.path.join(root, "profiles", "r02_bench_final.json")) if l.startswith("{")]
    assert len(lines) == 1
    d = lines[0]
    # round 2: pixel-pair throughput beside the HBM fraction, C3 / C5 sub-records, per-rank work, tile occupancy
    p = d["roofline"]["pairs"]
    assert p["pairs_blended"] > 0 and p["bwd_blended_pairs_per_s"] > 0 and p["fwd_evaluated_pairs_per_s"] > 0
    assert set(d["other_configs"]) == {"c3_3M_4k_sh3", "c5_5M_1440p_dense"}
    assert all(v["value"] > 0 and v["roofline"]["frac"] > 0 for v in d["other_configs"].values())
    assert d["per_rank"][0]["intersections_binned"] < d["per_rank"][0]["intersections_reference"]
    assert d["ms_per_step_instrumented"] >= d["ms_per_step"] and d["tile_occupancy"]
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "Mpixel/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f32"
    assert "workload" in d["config"] and d["warmup"] >= 3 and d["n_gpus"] == 1
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert abs(d["value"] - 1920 * 1080 / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]
