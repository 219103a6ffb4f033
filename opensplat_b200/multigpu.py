"""Data-parallel (over camera views) gradient exchange on NVLink, fused with the SH backward pass.

Baseline (parallel.allreduce_gradients): every rank runs sh_backward and then ONE NCCL all-reduce over the
flat 59-float/Gaussian gradient buffer (236 MB at 1M Gaussians, SH degree 3).

Fused path (this module): the SH VJP is rank-1 in (view basis) x (colour gradient), so instead of reducing the
48 coefficient gradients per Gaussian, every rank exposes only its view's colour gradient v_rgb [N,3] in
symmetric (peer-mapped) memory, and `gsb_sh_backward_multiview` forms sum_r Y_r (x) v_rgb_r itself, loading the
peers' v_rgb over NVLink (P2P) while it computes.  The remaining 11 floats/Gaussian (means, scales, quats,
opacity) go through one small NCCL all-reduce on a side stream, overlapped with the fused kernel.
NVLink bytes per rank per step at G ranks: (G-1) x 12 B + 2(G-1)/G x 44 B per Gaussian instead of
2(G-1)/G x 236 B (G = 8: 161 MB instead of 413 MB at 1M Gaussians), and one kernel fewer.
"""
import torch
import torch.distributed as dist

from . import capi


class ViewParallelExchange:
    def __init__(self, pipe, cam_pos, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.pipe = pipe
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        dev = pipe.dev
        self._alloc_symmetric(pipe.n)
        # every rank's camera centre (tiny, exchanged once)
        cp = torch.as_tensor(cam_pos, dtype=torch.float32, device=dev).reshape(1, 3)
        allcp = [torch.zeros_like(cp) for _ in range(self.world)]
        dist.all_gather(allcp, cp, group=self.group)
        self.cam_positions = torch.cat(allcp, 0).contiguous()
        self.side = torch.cuda.Stream(device=dev)
        assert pipe.sizes[-1][0] == "coeffs"

    def _alloc_symmetric(self, n):
        import torch.distributed._symmetric_memory as symm_mem
        dev = self.pipe.dev
        # two symmetric buffers (double-buffered so ONE barrier per step is enough, see exchange())
        self.bufs, self.hdls, self.ptrs = [], [], []
        for _ in range(2):
            t = symm_mem.empty((n, 3), dtype=torch.float32, device=dev)
            t.zero_()
            h = symm_mem.rendezvous(t, self.group.group_name)
            self.bufs.append(t)
            self.hdls.append(h)
            self.ptrs.append(h.buffer_ptrs_dev)  # device array of world_size pointers (peer-mapped)
        self.step = 0
        self.geom_numel = n * 11  # means 3 + scales 3 + quats 4 + opacity 1: the prefix of the flat buffer

    def resize(self, pipe):
        """After a refinement changed the Gaussian count (collective: every rank refines in lock-step, see
        densify.Densifier.sync_stats): new symmetric buffers + rendezvous."""
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)
        self._alloc_symmetric(pipe.n)

    def v_rgbs_buffer(self):
        """Where this step's rasterize-backward must write its colour gradient."""
        return self.bufs[self.step % 2]

    def exchange(self, average=True):
        """Call after project_backward: masks v_rgbs, synchronises the ranks, runs the fused multi-view SH
        backward (peer loads over NVLink) and, concurrently, the NCCL all-reduce of the geometry gradients."""
        p = self.pipe
        L = capi.lib()
        i = self.step % 2
        buf, hdl = self.bufs[i], self.hdls[i]
        cur = torch.cuda.current_stream()
        scale = 1.0 / self.world if average else 1.0
        # geometry gradients: small NCCL all-reduce on a side stream, overlapped with the fused kernel
        geom = p.grad_flat[: self.geom_numel]
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            dist.all_reduce(geom, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                geom.mul_(scale)
        capi.check(L.gsb_mask_rgb_grad(p.n, capi.ptr(p.rgbs), capi.ptr(buf), capi.stream()))
        # all ranks have finished writing this step's v_rgbs.  (Double buffering: the buffer written at step
        # t is last READ by peers in step t's fused kernel, which every rank has passed in stream order
        # before it reaches the barrier of step t+1, i.e. before anyone writes that buffer again at t+2.)
        hdl.barrier(channel=0)
        capi.check(L.gsb_sh_backward_multiview(p.n, p.deg, p.deg, capi.ptr(p.p["means"]), self.world,
                                               capi.ptr(self.cam_positions), self.ptrs[i], scale,
                                               capi.ptr(p.g["coeffs"]), capi.stream()))
        cur.wait_stream(self.side)
        self.step += 1
