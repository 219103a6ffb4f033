"""GPU parity at the FULL sizes of BASELINE.json configs[2] (C3: 3M Gaussians, 3840x2160) and configs[4] (C5: 5M
Gaussians, 2560x1440, dense overlap) -- the scenes bench.py times.
 * integer artefacts at full size, bit-exact vs the plain-C oracle: radii, num_tiles_hit, M, the sorted key order,
   gaussian_ids_sorted, tile_bins (generic path) and, bit-exact vs the generic path, the fast path's (cull = 0) bins,
   order and record stream;
 * image and gradients through the operator the bench drives (culled fast path): the oracle is evaluated on two tile
   WINDOWS (image centre, image corner) of the same full-size problem -- the window's tile lists index the full sorted
   list, pixel coordinates are shifted -- and the backward pass gets a v_output that is zero outside the window
   (the backward map is linear in v_output), so the per-Gaussian gradients are comparable exactly as at C2;
 * size-independent properties: bins partition [0, M), linearity of the backward pass, determinism.
Tolerances as at C2 (tests/test_gpu_parity.py): image 2e-5 on >= 99.8 % of pixels / flips <= 4.5e-3, gradients 1e-3 rel-L2
(2e-3 in the dense config where alpha saturates at 0.99)."""
import numpy as np
import pytest
import torch

from bench import WORKLOADS
from oracle import oracle as orc
from opensplat_b200 import ops
from opensplat_b200.scene import make_scene
from util import rel_l2, image_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


def _window(bins, xys, tb, tx0, ty0, tw, th):
    """Sub-problem of the tiles [tx0, tx0+tw) x [ty0, ty0+th): its tile_bins (same indices into the full sorted
    list) and pixel-shifted centres."""
    b = bins.reshape(tb[1], tb[0], 2)[ty0:ty0 + th, tx0:tx0 + tw].reshape(-1, 2).copy()
    x = xys.copy()
    x[:, 0] -= 16.0 * tx0
    x[:, 1] -= 16.0 * ty0
    return b, x


@pytest.mark.parametrize("workload,gtol", [("c3_3M_4k_sh3", 1e-3), ("c5_5M_1440p_dense", 2e-3)])
def test_full_size_config_parity(workload, gtol):
    n, W, H, scale, opac = WORKLOADS[workload]
    sc = make_scene(n, W, H, scale=scale, sh_degree=3, opacity=opac, seed=0)
    tb = ops.tile_bounds(W, H)
    T = tb[0] * tb[1]
    # ---- projection + binning, full size, bit-exact vs the oracle ----
    cov3d, xys, depths, radii, conics, nth = ops.project_gaussians_forward(
        cu(sc["means"]), cu(sc["scales"]), 1.0, cu(sc["quats"]), cu(sc["viewmat"]), cu(sc["projmat"]), sc["fx"],
        sc["fy"], sc["cx"], sc["cy"], H, W, tb)
    o = orc.project_forward(sc["means"], sc["scales"], 1.0, sc["quats"], sc["viewmat"], sc["projmat"], sc["fx"],
                            sc["fy"], sc["cx"], sc["cy"], H, W)
    assert np.array_equal(npy(radii), o["radii"]) and np.array_equal(npy(nth), o["num_tiles_hit"])
    assert np.array_equal(npy(xys), o["xys"]) and np.array_equal(npy(conics), o["conics"])
    cum = ops.cumsum_tiles_hit(nth)
    ocum, m = orc.cumsum(o["num_tiles_hit"])
    assert int(cum[-1]) == m and np.array_equal(npy(cum), ocum)
    isect, gids, ks, gs, bins, idx = ops.binAndSortGaussians(n, m, xys, depths, radii, cum, tb, return_index=True)
    ob = orc.bin_and_sort(o["xys"], o["depths"], o["radii"], ocum, H, W)
    binsn, gsn = npy(bins), npy(gs)
    assert np.array_equal(npy(ks), ob["isect_ids_sorted"]) and np.array_equal(gsn, ob["gaussian_ids_sorted"])
    assert np.array_equal(binsn, ob["tile_bins"])
    nz = binsn[:, 1] > binsn[:, 0]
    assert (binsn[nz, 1] - binsn[nz, 0]).sum() == m                     # bins partition [0, M)
    del isect, gids, ks, ob
    # ---- colours as the pipeline computes them ----
    rgbs = torch.clamp_min(ops.compute_sh_forward(3, 3, cu(sc["viewdirs"]), cu(sc["coeffs"])) + 0.5, 0.0)
    opc = cu(sc["opacities"])
    bg = cu(np.zeros(3, np.float32))
    # ---- fast path without culling == generic path, bit for bit, at this size ----
    _, _, st0, _ = ops.bucket_tile_ranges(xys, radii, conics, rgbs, opc, tb, 0, 0, cull=False)
    m0, len0 = (int(v) for v in st0.tolist()[:2])
    assert m0 == m and len0 == int((binsn[:, 1] - binsn[:, 0]).max())
    bins_b, cum_b, st, ws = ops.bucket_tile_ranges(xys, radii, conics, rgbs, opc, tb, m, len0, cull=False)
    rec_b, idx_b, gs_b = ops.bucket_sort_pack(n, m, len0, depths, radii, cum_b, tb, bins_b, st, ws, cull=False,
                                              want_index=True)
    assert torch.equal(bins_b, bins) and torch.equal(cum_b, cum)
    assert torch.equal(idx_b, idx) and torch.equal(gs_b, gs)
    out_g, fT_g, fI_g, rec_g = ops.rasterize_forward(tb, (W, H, 1), gs, idx, bins, xys, conics, rgbs, opc, bg)
    assert torch.equal(rec_b[: m * 48], rec_g[: m * 48])
    del rec_b, idx_b, gs_b, ws, idx
    # ---- the operator (culled fast path): image identical to the generic path, finite, deterministic ----
    colt, opt = rgbs.clone().requires_grad_(), opc.clone().requires_grad_()
    xyt, cont = xys.clone().requires_grad_(), conics.clone().requires_grad_()
    img = ops.RasterizeGaussians.apply(xyt, depths, radii, cont, nth, colt, opt, H, W, bg)
    assert torch.equal(img, out_g) and bool(torch.isfinite(img).all())
    # ---- oracle on two tile windows of the full problem ----
    tw, th = 24, 16
    windows = [((tb[0] - tw) // 2, (tb[1] - th) // 2), (0, 0)]
    rng = np.random.default_rng(0)
    xysn, conn, rgbn = npy(xys), npy(conics), npy(rgbs)
    v_out = np.zeros((H, W, 3), np.float32)
    win_res = []
    for tx0, ty0 in windows:
        wb, wx = _window(binsn, xysn, tb, tx0, ty0, tw, th)
        f = orc.rasterize_forward(16 * th, 16 * tw, gsn, wb, wx, conn, rgbn, sc["opacities"], [0, 0, 0], exp_mode=1)
        sub = npy(img)[16 * ty0:16 * (ty0 + th), 16 * tx0:16 * (tx0 + tw)]
        ok, stats = image_close(sub, f["out_img"], tol=2e-5, frac=2e-3)
        assert ok, (workload, tx0, ty0, stats)
        wv = rng.uniform(-1, 1, (16 * th, 16 * tw, 3)).astype(np.float32)
        v_out[16 * ty0:16 * (ty0 + th), 16 * tx0:16 * (tx0 + tw)] = wv
        win_res.append((wb, wx, f, wv))
    (img * cu(v_out)).sum().backward()
    ref = {k: np.zeros_like(npy(t)) for k, t in (("v_xy", xyt), ("v_conic", cont), ("v_colors", colt))}
    ref["v_opacity"] = np.zeros((n,), np.float32)
    for wb, wx, f, wv in win_res:     # the windows are disjoint: their gradient contributions add
        r = orc.rasterize_backward(16 * th, 16 * tw, gsn, wb, wx, conn, rgbn, sc["opacities"], [0, 0, 0], f["final_Ts"],
                                   f["final_idx"], wv, exp_mode=1)
        for k in ref:
            ref[k] += r[k].reshape(ref[k].shape)
    assert rel_l2(npy(xyt.grad), ref["v_xy"]) <= gtol
    assert rel_l2(npy(cont.grad), ref["v_conic"]) <= gtol
    assert rel_l2(npy(colt.grad), ref["v_colors"]) <= gtol
    assert rel_l2(npy(opt.grad).reshape(-1), ref["v_opacity"]) <= gtol
    # Gaussians that touch neither window get exactly zero gradient
    touched = np.zeros(n, bool)
    for wb, _, _, _ in win_res:
        for a, b in wb:
            touched[gsn[a:b]] = True
    assert float(colt.grad[cu(~touched)].abs().sum()) == 0.0
    # ---- backward is linear in v_output and bit-reproducible (generic entry point, full image) ----
    w_full = cu(rng.uniform(-1, 1, (H, W, 3)).astype(np.float32))
    v1 = ops.rasterize_backward(H, W, n, m, bins, conics, opc, rec_g, cum, bg, fT_g, fI_g, w_full)
    v2 = ops.rasterize_backward(H, W, n, m, bins, conics, opc, rec_g, cum, bg, fT_g, fI_g, w_full * 2)
    v3 = ops.rasterize_backward(H, W, n, m, bins, conics, opc, rec_g, cum, bg, fT_g, fI_g, w_full)
    for a, b, c in zip(v1, v2, v3):
        assert torch.equal(a * 2, b) and torch.equal(a, c) and bool(torch.isfinite(a).all())
