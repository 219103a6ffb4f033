"""oracle/scene_edit.py (the CPU restatement of Model::afterTrain and the scene writers) against golden vectors
produced by the UNMODIFIED reference model.cpp (tests/golden/make_golden_scene_edit.py).  CPU only."""
import os
import sys
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import scene_edit as se  # noqa: E402
from util import PARAM_NAMES, load_golden, scene_edit_inputs  # noqa: E402

CASES = ["scene_edit_densify_screen", "scene_edit_densify_huge", "scene_edit_densify_all", "scene_edit_alpha_reset"]


def cfg_of(g):
    c = g["cfg"]
    return types.SimpleNamespace(
        num_cameras=int(c[0]), refine_every=int(c[1]), warmup_length=int(c[2]), reset_alpha_every=int(c[3]),
        densify_grad_thresh=float(np.float32(c[4])), densify_size_thresh=float(np.float32(c[5])),
        stop_screen_size_at=int(c[6]), split_screen_size=float(np.float32(c[7])), max_steps=int(c[8]),
        cull_alpha_thresh=float(np.float32(0.1)), cull_scale_thresh=0.5, cull_screen_size=float(np.float32(0.15)),
        size_fac=float(np.float32(1.6)), stop_split_at=int(c[8]) // 2)


def schedule(cfg, step):
    """model.cpp:339-341,349,441,472 (same function as opensplat_b200.densify.Densifier.schedule)."""
    refine = step % cfg.refine_every == 0 and step > cfg.warmup_length
    interval = cfg.reset_alpha_every * cfg.refine_every
    densify = refine and step < cfg.stop_split_at and step % interval > cfg.num_cameras + cfg.refine_every
    reset = refine and step < cfg.stop_split_at and step % interval == cfg.refine_every
    return refine, densify, reset, step < cfg.stop_screen_size_at, step > cfg.refine_every * cfg.reset_alpha_every


def run_oracle(g):
    """Replays a golden case through the restatement; returns (params, m, v, per-step stats)."""
    n, k, seed = int(g["n"]), int(g["k"]), int(g["seed"])
    H, W = (int(x) for x in g["hw"])
    cfg = cfg_of(g)
    p, m, v, draws = scene_edit_inputs(n, k, seed, max(H, W))
    p = {a: torch.from_numpy(b) for a, b in p.items()}
    m = {a: torch.from_numpy(b) for a, b in m.items()}
    v = {a: torch.from_numpy(b) for a, b in v.items()}
    stats, per_step = None, []
    for si, step in enumerate(int(s) for s in g["steps"]):
        v_xy, radii = draws[si]
        if step < cfg.stop_split_at:
            stats = se.densify_stats(stats, v_xy, radii, H, W)
        refine, densify, reset, chk_screen, chk_huge = schedule(cfg, step)
        if refine:
            if densify:
                def draw(n_splits):
                    torch.manual_seed(int(g["seed_randn"]))
                    return torch.randn(2 * n_splits, 3)
                p, m, v, _ = se.refine(p, m, v, stats, max(H, W), cfg, chk_screen, chk_huge, draw)
            if reset:
                p["opacities"] = se.reset_opacity(p["opacities"], cfg.cull_alpha_thresh)
            stats = None
        per_step.append(stats)
    return p, m, v, per_step


@pytest.mark.parametrize("name", CASES)
def test_after_train_restatement_matches_reference(name):
    g = load_golden(name)
    p, m, v, per_step = run_oracle(g)
    for si, st in enumerate(per_step):
        for i, x in enumerate(("xysGradNorm", "visCounts", "max2DSize")):
            ref = g[f"s{si}_{x}"]
            if st is None:
                assert ref.size == 0
            else:
                np.testing.assert_array_equal(st[i].numpy(), ref)
    for x in PARAM_NAMES:
        assert p[x].shape == g["p_" + x].shape, x
        np.testing.assert_array_equal(p[x].numpy(), g["p_" + x], err_msg=x)
        if name != "scene_edit_alpha_reset":     # D14: the reference drops the zeroed opacity state; moments untouched
            np.testing.assert_array_equal(m[x].numpy(), g["m_" + x], err_msg="m_" + x)
            np.testing.assert_array_equal(v[x].numpy(), g["v_" + x], err_msg="v_" + x)
    assert p["means"].shape[0] != int(g["n"]) or name == "scene_edit_alpha_reset"


def test_alpha_reset_reference_keeps_adam_state():
    """Pins divergence D14: Model::afterTrain builds a zeroed AdamParamState for the opacities and then drops it
    (model.cpp:477-486), so the reference's moments are unchanged by an alpha reset."""
    g = load_golden("scene_edit_alpha_reset")
    _, m, v, _ = scene_edit_inputs(int(g["n"]), int(g["k"]), int(g["seed"]))
    np.testing.assert_array_equal(g["m_opacities"], m["opacities"])
    np.testing.assert_array_equal(g["v_opacities"], v["opacities"])
    assert float(g["p_opacities"].max()) <= float(torch.logit(torch.tensor(0.2))) + 1e-7


@pytest.mark.parametrize("name", ["scene_edit_save", "scene_edit_save_crs"])
def test_scene_writers_match_reference_bytes(name):
    g = load_golden(name)
    n, k = int(g["n"]), int(g["k"])
    p, _, _, _ = scene_edit_inputs(n, k, int(g["seed"]))
    keep, scale, tr = bool(g["keep_crs"]), float(g["scale"]), tuple(float(x) for x in g["translation"])
    ply = se.ply_header(n, 3 * (k - 1), int(g["step"])) + se.ply_body(
        p["means"], p["featuresDc"], p["featuresRest"], p["opacities"], p["scales"], p["quats"], keep, scale, tr)
    assert ply == g["ply"].tobytes()
    body, order = se.splat_body(p["means"], p["featuresDc"], p["opacities"], p["scales"], p["quats"], keep, scale, tr)
    ref_rows = g["splat"].reshape(n, 32)
    rows = np.frombuffer(body, np.uint8).reshape(n, 32)
    # std::sort is unstable: equal keys may come out in another order -> compare as sorted row sets + key order
    assert sorted(map(bytes, rows)) == sorted(map(bytes, ref_rows))
    _, key = se.splat_rows(p["means"], p["featuresDc"], p["opacities"], p["scales"], p["quats"], keep, scale, tr)
    lookup = {bytes(r): i for i, r in enumerate(se.splat_rows(p["means"], p["featuresDc"], p["opacities"], p["scales"],
                                                              p["quats"], keep, scale, tr)[0])}
    ref_keys = np.array([key[lookup[bytes(r)]] for r in ref_rows])
    assert np.all(np.diff(ref_keys) <= 0)
    if len(set(key.tolist())) == n:
        assert body == g["splat"].tobytes()


@pytest.mark.parametrize("name", ["scene_edit_save", "scene_edit_save_crs"])
def test_ply_loader_restatement_matches_reference_loadply(name):
    """oracle ply_load == Model::loadPly run on the reference's own file (resume path, model.cpp:614-778)."""
    g = load_golden(name)
    keep, scale, tr = bool(g["keep_crs"]), float(g["scale"]), tuple(float(x) for x in g["translation"])
    p, step = se.ply_load(g["ply"].tobytes(), keep, scale, tr)
    assert step == int(g["ld_step"]) == int(g["step"])
    for x in PARAM_NAMES:
        np.testing.assert_array_equal(p[x].numpy(), g["ld_" + x], err_msg=x)


def test_ply_header_checks_host_logic():
    """opensplat_b200.export.parse_ply_header accepts what Model::savePly writes and rejects what loadPly rejects."""
    from opensplat_b200.export import parse_ply_header, ply_header
    h = ply_header(5, 16, 1234)
    assert h == se.ply_header(5, 45, 1234)
    assert parse_ply_header(h + b"\0" * 10) == (1234, 5, 3, 45, len(h))
    assert parse_ply_header(ply_header(7, 1, 0))[:4] == (0, 7, 3, 0)
    for bad in (h.replace(b"ply\n", b"plx\n", 1), h.replace(b"binary_little_endian", b"ascii"),
                h.replace(b" at iteration 1234", b""), h.replace(b"property float nx\n", b""),
                h.replace(b"property float rot_3\n", b""), h.replace(b"end_header\n", b"")):
        with pytest.raises(ValueError):
            parse_ply_header(bad)
