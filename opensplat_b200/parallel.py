"""Host-side logic of the data-parallel (over camera views) path, SURVEY.md section 8e.

One process per GPU; every rank holds a full replica of the Gaussians and renders its own view(s);
gradients of all per-Gaussian parameters live in ONE flat fp32 buffer so a step needs a single all-reduce.
Device-agnostic (runs under gloo on CPU in the tests, NCCL on the GPUs)."""
import torch
import torch.distributed as dist


def flat_layout(n, sh_bases, align=4):
    """Per-tensor (offset, count, shape) of the flat parameter / gradient buffer and its total length (floats).
    Every slice starts on a multiple of `align` floats (16 bytes), whatever n is: the projection kernels use
    128-bit accesses on the quaternion slices, and a refinement may leave an odd Gaussian count."""
    sizes = [("means", (n, 3)), ("scales", (n, 3)), ("quats", (n, 4)), ("opacities", (n, 1)),
             ("coeffs", (n, sh_bases, 3))]
    offs, o = {}, 0
    for name, shp in sizes:
        c = 1
        for d in shp:
            c *= d
        offs[name] = (o, c, shp)
        o += (c + align - 1) // align * align
    return offs, o


def flat_views(flat, offs):
    return {name: flat[o:o + c].view(shp) for name, (o, c, shp) in offs.items()}


def views_for_rank(num_views, rank, world):
    """Round-robin assignment of camera views to ranks (view v -> rank v % world)."""
    return [v for v in range(num_views) if v % world == rank]


def allreduce_gradients(grad_flat, world, average=True):
    """The path's single exchange step: SUM all-reduce of the flat per-Gaussian gradient buffer."""
    if world > 1:
        dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM)
        if average:
            grad_flat.mul_(1.0 / world)
    return grad_flat


def replicas_in_sync(param_flat, world):
    """True iff every rank holds bit-identical parameters (checksum exchange; debugging aid)."""
    if world <= 1:
        return True
    cs = torch.stack([param_flat.double().sum(), param_flat.double().abs().sum()]).to(param_flat.device)
    lo, hi = cs.clone(), cs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def allreduce_tensor_grads(tensors, world=None, average=True, group=None):
    """Data-parallel exchange for a model whose parameters are separate tensors (model.GaussianModel: the reference's
    six tensors): pack every defined .grad into ONE flat bucket, one all-reduce, unpack in place.  A tensor whose
    gradient is undefined on this rank (Model::forward returned only the background) contributes zeros, and gets
    the reduced gradient if any rank had one -- so every replica takes the same Adam step."""
    if world is None:
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world <= 1:
        return
    tensors = list(tensors)
    if not tensors:
        return
    total = sum(t.numel() for t in tensors)
    bucket = torch.zeros(total, dtype=torch.float32, device=tensors[0].device)
    o = 0
    for t in tensors:
        if t.grad is not None:
            bucket[o:o + t.numel()].copy_(t.grad.reshape(-1))
        o += t.numel()
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    if average:
        bucket.mul_(1.0 / world)
    o = 0
    for t in tensors:
        g = bucket[o:o + t.numel()].view_as(t)
        if t.grad is None:
            t.grad = g.clone()
        else:
            t.grad.copy_(g)
        o += t.numel()
