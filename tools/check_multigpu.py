"""torchrun --nproc-per-node N tools/check_multigpu.py [--no-resize] : data-parallel correctness on N GPUs.
Checks that the fused exchange launch (multi-view SH backward over peer memory + two-shot all-reduce of the geometry
gradients, NVSwitch multimem when available; GSB_EXCHANGE_MULTICAST=0 forces the peer-pointer flavour) produces the
same averaged gradients as sh_backward + one NCCL all-reduce of the whole flat buffer, that replicas stay
bit-identical after Adam steps, and that both still hold after a lock-step change of the Gaussian count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from opensplat_b200 import parallel
from opensplat_b200.multigpu import ViewParallelExchange
from opensplat_b200.pipeline import SplatPipeline
from opensplat_b200.scene import make_scene, rotated_camera

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, W, H = 200_000, 640, 360
sc = make_scene(n, W, H, scale=0.06, sh_degree=3, opacity=(0.05, 0.95), seed=0)
cam = rotated_camera(W, H, rank, n_views=8)


def make_pipe():
    p = SplatPipeline(n, W, H, device=dev)
    p.load_scene(sc)
    p.set_camera(cam)
    vd = sc["means"] - cam["cam_pos"]
    p.viewdirs.copy_(torch.from_numpy((vd / np.linalg.norm(vd, axis=-1, keepdims=True)).astype(np.float32)).to(dev))
    p.target.copy_(torch.from_numpy(np.random.default_rng(rank).uniform(0, 1, (H, W, 3)).astype(np.float32)).to(dev))
    return p


# baseline: full flat all-reduce
a = make_pipe()
a.forward(); a.backward()
parallel.allreduce_gradients(a.grad_flat, world, average=True)
# fused exchange
b = make_pipe()
b.exchange = ViewParallelExchange(b, cam["cam_pos"])
for _ in range(3):   # several steps: the symmetric buffers are reused every step
    b.forward(); b.backward()
torch.cuda.synchronize()


def compare(ref, got, geo):
    """geometry prefix: same sums in a possibly different order (exact at world 2, where the order cannot matter);
    SH block: expanded from the exchanged colour gradients instead of reduced -> fp32 re-association only."""
    d = (ref[:geo] - got[:geo])
    e_geo = float(d.abs().max())
    r_geo = float(d.norm() / ref[:geo].norm())
    r_sh = float((ref[geo:] - got[geo:]).norm() / ref[geo:].norm())
    good = (e_geo == 0.0 if world <= 2 else r_geo < 1e-6) and r_sh < 1e-5
    return good, e_geo, r_sh


ok, err_geo, rel_sh = compare(a.grad_flat, b.grad_flat, b.geom_numel)
mc = bool(b.exchange.multicast_ptr)
# replicas stay identical through training steps
for _ in range(3):
    b.train_step(world_size=world)
sync = parallel.replicas_in_sync(b.param_flat, world)
# a refinement changes the Gaussian count on every replica in lock-step (densify.Densifier.sync_stats makes the
# decisions identical): new flat layout + new symmetric buffers, and the fused exchange keeps matching plain NCCL
do_resize = "--no-resize" not in sys.argv
ok2 = True
if do_resize:
    keep = torch.arange(0, n, 2, device=dev)
    idx = torch.cat([keep, keep[:1001]])     # odd count: the aligned flat layout must cope
    views_m = {name: b.adam_m[o:o + c].view(shp) for name, (o, c, shp) in b.offs.items()}
    views_v = {name: b.adam_v[o:o + c].view(shp) for name, (o, c, shp) in b.offs.items()}
    newp = {k: v[idx].clone() for k, v in b.p.items()}
    newm = {k: v[idx].clone() for k, v in views_m.items()}
    newv = {k: v[idx].clone() for k, v in views_v.items()}
    b.resize_gaussians(newp, newm, newv)
    # the fused kernel derives every view's direction from the CURRENT means; give both paths the same directions
    cp = torch.from_numpy(np.asarray(cam["cam_pos"], np.float32)).to(dev)
    vd_new = torch.nn.functional.normalize(b.p["means"] - cp, dim=-1)
    b.viewdirs.copy_(vd_new)
    for _ in range(2):
        b.forward(); b.backward()
    c = SplatPipeline(b.n, W, H, device=dev)
    c.set_camera(cam)
    c.param_flat.copy_(b.param_flat); c.viewdirs.copy_(b.viewdirs); c.target.copy_(b.target)
    c.forward(); c.backward()
    parallel.allreduce_gradients(c.grad_flat, world, average=True)
    torch.cuda.synchronize()
    ok2, e_geo2, r_sh2 = compare(c.grad_flat, b.grad_flat, b.geom_numel)
    if rank == 0:
        print(f"after resize to n={b.n}: geometry max|d|={e_geo2:.3g} sh rel-L2={r_sh2:.3g}")
    b.train_step(world_size=world)
    sync = sync and parallel.replicas_in_sync(b.param_flat, world)
ok = ok and ok2
res = torch.tensor([int(ok), int(sync)], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"multigpu check world={world} multicast={mc}: geometry max|d|={err_geo:.3g} sh rel-L2={rel_sh:.3g} after_resize_ok={ok2} "
          f"fused_ok={bool(res[0])} replicas_in_sync={bool(res[1])}")
dist.destroy_process_group()
sys.exit(0 if bool(res[0]) and bool(res[1]) else 1)
