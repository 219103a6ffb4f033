"""CPU, world_size 2, gloo: host logic of the data-parallel path (flat gradient bucket, view sharding,
one all-reduce per step, replicas stay identical after the same update)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opensplat_b200 import parallel


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, K = 257, 16
    offs, total = parallel.flat_layout(n, K)
    assert total == sum((c + 3) // 4 * 4 for _, c, _ in offs.values()) and total >= n * 59   # 16-byte aligned slices
    torch.manual_seed(0)                      # same replica everywhere
    param = torch.randn(total)
    grad = torch.zeros(total)
    g = parallel.flat_views(grad, offs)
    p = parallel.flat_views(param, offs)
    assert p["coeffs"].data_ptr() == param[offs["coeffs"][0]:].data_ptr()   # views alias the flat buffer
    # every rank renders different views -> different local gradients
    my_views = parallel.views_for_rank(8, rank, world)
    for name in g:
        g[name].fill_(0)
        for v in my_views:
            g[name].add_(float(v + 1))
    parallel.allreduce_gradients(grad, world, average=False)
    expect = float(sum(v + 1 for v in range(8)))
    ok = all(bool(torch.all(g[name] == expect)) for name in g)   # (the alignment padding stays zero)
    param.add_(grad, alpha=-0.01)             # identical update on every rank
    ok = ok and parallel.replicas_in_sync(param, world)
    # a rank that diverges is detected
    if rank == 1:
        param[3] += 1.0
    diverged = not parallel.replicas_in_sync(param, world)
    ret[rank] = (ok, diverged, my_views)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()   # never fork a multi-threaded pytest process
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] and ret[1][1]
    assert ret[0][2] == [0, 2, 4, 6] and ret[1][2] == [1, 3, 5, 7]


def test_view_sharding_covers_all_views_once():
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in parallel.views_for_rank(8, r, world))
        assert seen == list(range(8))


def _stats_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opensplat_b200.densify import Densifier
    dn = Densifier()
    n = 1000
    g = torch.Generator().manual_seed(100 + rank)     # every rank saw different views
    dn.xys_grad_norm = torch.rand(n, generator=g)
    dn.vis_counts = torch.randint(1, 5, (n,), generator=g).float()
    dn.max_2d_size = torch.rand(n, generator=g)
    mine = (dn.xys_grad_norm.clone(), dn.vis_counts.clone(), dn.max_2d_size.clone())
    dn.sync_stats()
    ret[rank] = tuple(t.numpy() for t in mine) + tuple(t.numpy() for t in (dn.xys_grad_norm, dn.vis_counts, dn.max_2d_size))
    dist.destroy_process_group()


def test_densify_stats_sync_world2():
    """Replicas must classify identically: statistics are summed / maxed over the ranks before a refinement."""
    import numpy as np
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_stats_worker, args=(world, port, ret), nprocs=world, join=True)
    a, b = ret[0], ret[1]
    for r in (a, b):
        assert np.array_equal(r[3], a[0] + b[0]) and np.array_equal(r[4], a[1] + b[1])
        assert np.array_equal(r[5], np.maximum(a[2], b[2]))
    assert all(np.array_equal(a[i], b[i]) for i in (3, 4, 5))       # identical on every rank


def _grads_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    ts = [torch.randn(7, 3, requires_grad=True), torch.randn(7, 15, 3, requires_grad=True), torch.randn(7, 1, requires_grad=True)]
    ts[0].grad = torch.full((7, 3), float(rank + 1))
    ts[1].grad = torch.full((7, 15, 3), 10.0 * (rank + 1))
    if rank == 0:
        ts[2].grad = torch.full((7, 1), 5.0)          # rank 1 saw nothing: undefined gradient there
    parallel.allreduce_tensor_grads(ts, average=True)
    ret[rank] = [t.grad.clone().numpy() for t in ts]
    dist.destroy_process_group()


def test_separate_tensor_gradients_one_bucket_world2():
    import numpy as np
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_grads_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in (0, 1):
        assert np.all(ret[r][0] == 1.5) and np.all(ret[r][1] == 15.0) and np.all(ret[r][2] == 2.5)


def _lockstep_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opensplat_b200.densify import Densifier, RefineConfig

    class HostDensifier(Densifier):
        """The control flow of Densifier.after_train with the two device calls replaced by host arithmetic."""
        def accumulate(self, v_xy, radii, img_h, img_w):
            n = radii.shape[0]
            if self.xys_grad_norm is None:
                self.xys_grad_norm, self.vis_counts, self.max_2d_size = torch.zeros(n), torch.zeros(n), torch.zeros(n)
            vis = radii > 0
            self.xys_grad_norm += v_xy.norm(dim=-1) * vis
            self.vis_counts += vis.float()

        def refine(self, params, adam_m, adam_v, max_dim, check_split_screen, check_huge):
            # the decision every replica must agree on: Gaussians whose reduced gradient statistic is large
            keep = (self.xys_grad_norm / self.vis_counts.clamp_min(1)) < 0.5
            newp = {k: v[keep] for k, v in params.items()}
            return newp, adam_m, adam_v, {"n": int(keep.sum()), "added": 0, "culled": int((~keep).sum())}

        def reset_opacity(self, opacities, exp_avg=None, exp_avg_sq=None):
            opacities.clamp_(max=-1.0)

    cfg = RefineConfig(refine_every=2, warmup_length=1, num_cameras=0, reset_alpha_every=1000)
    dn = HostDensifier(cfg)
    n = 64
    params = {"means": torch.arange(n * 3, dtype=torch.float32).view(n, 3), "opacities": torch.zeros(n, 1)}
    g = torch.Generator().manual_seed(5)
    counts = []
    for step in range(1, 7):
        # rank 1's view never hits a Gaussian: its xys.grad is undefined (v_xy None) on every step
        v_xy = torch.rand(params["means"].shape[0], 2, generator=g) if rank == 0 else None
        radii = torch.ones(params["means"].shape[0], dtype=torch.int32)
        params, _, _, info = dn.after_train(step, params, None, None, v_xy, radii, 100, 100)
        counts.append(params["means"].shape[0])
    ret[rank] = (counts, params["means"].numpy().copy())
    dist.destroy_process_group()


def test_refinement_stays_in_lockstep_when_one_rank_sees_nothing_world2():
    """a rank whose view hits no Gaussian must still take part in the refine step's collectives and end up
    with the same Gaussian set as the others (no hang, no divergence)."""
    import numpy as np
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_lockstep_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] == ret[1][0] and ret[0][0][-1] < 64          # both replicas culled the same Gaussians
    assert np.array_equal(ret[0][1], ret[1][1])
