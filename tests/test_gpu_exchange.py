"""Single-GPU parity of the kernels the default N>1 path runs (`gsb_exchange_gradients`, `gsb_sh_backward_multiview`,
`gsb_mask_rgb_grad`): they take plain device-pointer arrays, so the peers' buffers are emulated by local ones.
 * multi-view SH VJP (1, 3, 8, 9 views, invisible views, degrees_to_use < degree) against the sum of single-view
   gsb_sh_backward calls and against the plain-C oracle;
 * the two-shot all-reduce role: G emulated ranks run their slice one after the other over G local buffers -- every
   buffer must end up holding scale * sum exactly (fixed summation order);
 * the clamp-gradient mask.
The reference has no counterpart (single-GPU only, README.md:268); the oracle is sum_r of its SH VJP (sh.cuh:126-216)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from opensplat_b200 import capi

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _views(n, num_views, seed):
    rng = np.random.default_rng(seed)
    means = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    cams = (rng.standard_normal((num_views, 3)) * 6 + np.array([0, 0, -8])).astype(np.float32)
    v = rng.standard_normal((num_views, n, 3)).astype(np.float32)
    v[rng.uniform(size=(num_views, n)) < 0.3] = 0.0     # Gaussians not visible in a view contribute nothing
    return means, cams, v


@pytest.mark.parametrize("num_views,n,deg,use", [(1, 1000, 3, 3), (3, 5003, 3, 3), (8, 4097, 3, 2), (9, 777, 3, 3),
                                                 (2, 129, 4, 4), (5, 300, 1, 1), (8, 1, 0, 0)])
def test_multiview_sh_backward_matches_sum_of_single_view_vjps(num_views, n, deg, use):
    L = capi.lib()
    K = (deg + 1) ** 2
    means, cams, v = _views(n, num_views, 17 * num_views + n)
    bufs = [cu(v[r]) for r in range(num_views)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    out = torch.full((n, K, 3), 9.0, device=DEV)
    scale = 1.0 / num_views
    means_d, cams_d = cu(means), cu(cams)      # keep the device copies alive across the (asynchronous) launch
    capi.check(L.gsb_sh_backward_multiview(n, deg, use, capi.ptr(means_d), num_views, capi.ptr(cams_d),
                                           ptrs.data_ptr(), scale, capi.ptr(out), capi.stream()))
    # (a) sum of the single-view kernel (the path a single GPU runs), (b) the oracle in float64
    acc = torch.zeros((n, K, 3), device=DEV, dtype=torch.float64)
    ref64 = np.zeros((n, K, 3), np.float64)
    one = torch.empty((n, K, 3), device=DEV)
    for r in range(num_views):
        vd = (means - cams[r]).astype(np.float32)
        vd_d = cu(vd)
        capi.check(L.gsb_sh_backward(n, deg, use, capi.ptr(vd_d), capi.ptr(bufs[r]), capi.ptr(one), capi.stream()))
        acc += one.double()
        ref64 += orc.sh_backward(use, K, vd, v[r]).astype(np.float64)
    got = out.cpu().numpy().astype(np.float64)
    tol = 2e-6 * max(1.0, np.abs(ref64).max())
    assert np.abs(got - scale * acc.cpu().numpy()).max() <= tol
    assert np.abs(got - scale * ref64).max() <= 5e-6 * max(1.0, np.abs(ref64).max())
    assert np.all(got[:, (use + 1) ** 2:, :] == 0)                  # unused bases are written as zeros


@pytest.mark.parametrize("world,floats", [(1, 4 * 1000), (2, 4 * 12345), (3, 4 * 7), (8, 4 * 100_003)])
def test_two_shot_allreduce_role_on_emulated_ranks(world, floats):
    """Rank r sums slice r of all buffers and writes it to all buffers (peer-pointer flavour of the role)."""
    L = capi.lib()
    rng = np.random.default_rng(world)
    host = [rng.standard_normal(floats).astype(np.float32) for _ in range(world)]
    bufs = [cu(h) for h in host]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    scale = 1.0 / world
    dummy = torch.zeros(3, device=DEV)
    for r in range(world):      # n = 0: only the all-reduce role runs
        capi.check(L.gsb_exchange_gradients(0, 3, 3, None, 1, capi.ptr(dummy), None, scale, None, r, world, floats,
                                            ptrs.data_ptr(), None, capi.stream()))
    acc = torch.zeros(floats, device=DEV)
    for h in host:              # same order as the kernel: ((0 + b0) + b1) + ...
        acc = acc + cu(h)
    expect = acc * scale
    for b in bufs:
        assert torch.equal(b, expect)


def test_exchange_launch_does_both_roles_at_once():
    """One launch: multi-view SH VJP into v_coeffs AND the all-reduce of the geometry prefix (world = 1 here, so the
    reduction is the identity times scale)."""
    L = capi.lib()
    n, deg, views = 3001, 3, 4
    means, cams, v = _views(n, views, 5)
    bufs = [cu(v[r]) for r in range(views)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    out_a, out_b = torch.empty((n, 16, 3), device=DEV), torch.empty((n, 16, 3), device=DEV)
    geom = torch.randn(4 * 5000, device=DEV)
    g0 = geom.clone()
    gp = torch.tensor([geom.data_ptr()], dtype=torch.int64, device=DEV)
    means_d, cams_d = cu(means), cu(cams)
    capi.check(L.gsb_sh_backward_multiview(n, deg, deg, capi.ptr(means_d), views, capi.ptr(cams_d), ptrs.data_ptr(),
                                           0.5, capi.ptr(out_a), capi.stream()))
    capi.check(L.gsb_exchange_gradients(n, deg, deg, capi.ptr(means_d), views, capi.ptr(cams_d), ptrs.data_ptr(),
                                        0.5, capi.ptr(out_b), 0, 1, geom.numel(), gp.data_ptr(), None, capi.stream()))
    assert torch.equal(out_a, out_b) and torch.equal(geom, g0 * 0.5)


def test_mask_rgb_grad_is_the_clamp_gradient():
    n = 10_001
    rgbs = torch.clamp_min(torch.randn(n, 3, device=DEV), 0.0)
    v = torch.randn(n, 3, device=DEV)
    expect = v * (rgbs > 0)
    capi.check(capi.lib().gsb_mask_rgb_grad(n, capi.ptr(rgbs), capi.ptr(v), capi.stream()))
    assert torch.equal(v, expect) and float((expect == 0).float().mean()) > 0.3
