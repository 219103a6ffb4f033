"""Data-parallel (over camera views) gradient exchange on NVLink / NVSwitch, fused with the SH backward pass.

Baseline (parallel.allreduce_gradients): every rank runs sh_backward and then ONE NCCL all-reduce over the
flat 59-float/Gaussian gradient buffer (236 MB at 1M Gaussians, SH degree 3).

Fused path (this module): the same kernel (`sh_backward_multiview_kernel`, two CTA roles) either as ONE launch per step
(`exchange()`: `gsb_exchange_gradients` between two cross-rank barriers) or, in the pipeline, as its two roles on two
streams (`start_colour()` right after rasterize-backward, `finish()` after project-backward) so that the colour pulls
overlap project_backward and the geometry all-reduce:
 * the SH VJP is rank-1 in (view basis) x (colour gradient), so instead of reducing the 48 coefficient gradients
   per Gaussian every rank exposes only its view's colour gradient v_rgb [N,3] in symmetric (peer-mapped) memory
   and the kernel forms sum_r Y_r (x) v_rgb_r itself, pulling the peers' v_rgb over NVLink with coalesced loads
   while it computes;
 * the remaining 11 floats/Gaussian (means, scales, quats, opacity -- the prefix of the flat gradient buffer, which
   lives in the same symmetric allocation so the backward kernels write it in place) are all-reduced by the first
   CTAs of the same launch: two-shot, rank r owns slice r, with NVSwitch multicast one `multimem.ld_reduce` (sum
   formed inside the switch) and one `multimem.st` (broadcast) per 16 bytes; without multicast through the peers'
   mapped pointers.
NVLink bytes per rank per step at G ranks (N Gaussians): received (G-1) x 12 N (colour gradients) + 44 N
(reduced slice + broadcasts), sent the same -- against 2(G-1)/G x 236 N each way for the flat all-reduce
(G = 8: 128 MB instead of 413 MB at 1M Gaussians), and no separate sh_backward / NCCL kernels.
"""
import os

import torch
import torch.distributed as dist

from . import capi


class ViewParallelExchange:
    def __init__(self, pipe, cam_pos, group=None):
        self.pipe = pipe
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        dev = pipe.dev
        # every rank's camera centre (tiny, exchanged once)
        cp = torch.as_tensor(cam_pos, dtype=torch.float32, device=dev).reshape(1, 3)
        allcp = [torch.zeros_like(cp) for _ in range(self.world)]
        dist.all_gather(allcp, cp, group=self.group)
        self.cam_positions = torch.cat(allcp, 0).contiguous()
        self.use_multicast = os.environ.get("GSB_EXCHANGE_MULTICAST", "1") != "0"
        self.overlap = os.environ.get("GSB_EXCHANGE_OVERLAP", "1") != "0"
        self.side = torch.cuda.Stream(device=dev)
        self._rgb_ready, self._colour_done = torch.cuda.Event(), torch.cuda.Event()
        self._alloc_symmetric(pipe)

    def _alloc_symmetric(self, pipe):
        import torch.distributed._symmetric_memory as symm_mem
        dev, n = pipe.dev, pipe.n
        # ONE symmetric allocation: [flat gradient buffer | this view's colour gradient v_rgb [n,3]]
        rgb_off = (pipe.numel + 3) // 4 * 4
        total = rgb_off + (3 * n + 3) // 4 * 4
        t = symm_mem.empty(total, dtype=torch.float32, device=dev)
        t.zero_()
        hdl = symm_mem.rendezvous(t, self.group.group_name)
        self.buf, self.hdl = t, hdl
        pipe.rebind_grad_flat(t[:pipe.numel])
        self.v_rgb = t[rgb_off:rgb_off + 3 * n].view(n, 3)
        base_off = int(getattr(hdl, "offset", 0) or 0)      # the tensor's offset inside its symmetric allocation block
        ptrs = [int(p) + base_off for p in hdl.buffer_ptrs]
        self.geom_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)              # peers' flat buffers
        self.rgb_ptrs = torch.tensor([p + 4 * rgb_off for p in ptrs], dtype=torch.int64, device=dev)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if self.use_multicast else 0
        mc = mc + base_off if mc else 0
        assert ptrs[self.rank] == t.data_ptr(), "symmetric-memory handle does not describe this tensor"
        self.multicast_ptr = mc                                                           # 0: no NVSwitch multicast
        self.geom_numel = pipe.geom_numel   # means, scales, quats, opacities (16-byte aligned slices)
        assert self.geom_numel % 4 == 0

    def resize(self, pipe):
        """After a refinement changed the Gaussian count (collective: every rank refines in lock-step, see
        densify.Densifier.sync_stats): new symmetric buffers + rendezvous."""
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)
        self._alloc_symmetric(pipe)

    def v_rgbs_buffer(self):
        """Where this step's rasterize-backward must write its colour gradient."""
        return self.v_rgb

    def start_colour(self, average=True):
        """Call right after rasterize-backward (v_rgb is final, the geometry gradients are not yet): masks v_rgb with
        the clamp's gradient and starts the multi-view SH backward -- the part of the exchange that moves most bytes,
        (G-1) x 12 B per Gaussian -- on a side stream, so that it overlaps project_backward and, afterwards, the
        all-reduce of the geometry gradients."""
        p = self.pipe
        L = capi.lib()
        self._scale = 1.0 / self.world if average else 1.0
        cur = torch.cuda.current_stream()
        capi.check(L.gsb_mask_rgb_grad(p.n, capi.ptr(p.rgbs), capi.ptr(self.v_rgb), capi.stream()))
        self._rgb_ready.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self._rgb_ready)
            self.hdl.barrier(channel=0)          # every rank's v_rgb is complete
            capi.check(L.gsb_sh_backward_multiview(
                p.n, p.deg, p.deg, capi.ptr(p.p["means"]), self.world, capi.ptr(self.cam_positions),
                self.rgb_ptrs.data_ptr(), self._scale, capi.ptr(p.g["coeffs"]), self.side.cuda_stream))
            self._colour_done.record(self.side)

    def finish(self):
        """Call after project_backward: all-reduces the geometry gradients (two-shot; NVSwitch multimem when mapped),
        joins the colour half and closes the step with the barrier that lets every rank overwrite its buffers."""
        p = self.pipe
        L = capi.lib()
        self.hdl.barrier(channel=1)              # every rank's geometry gradients are complete
        capi.check(L.gsb_exchange_gradients(
            0, p.deg, p.deg, None, 1, capi.ptr(self.cam_positions), None, self._scale, None, self.rank, self.world,
            self.geom_numel, self.geom_ptrs.data_ptr(), self.multicast_ptr if self.multicast_ptr else None,
            capi.stream()))
        torch.cuda.current_stream().wait_event(self._colour_done)
        # every rank's slice of the reduced geometry gradients has landed everywhere, and nobody still reads the v_rgb /
        # geometry buffers of this step (so the next backward pass may overwrite them)
        self.hdl.barrier(channel=2)

    def exchange(self, average=True):
        """The whole exchange after project_backward, as ONE fused launch (gsb_exchange_gradients: both CTA roles in one
        grid) between two barriers -- no overlap with the backward kernels; what callers use that cannot split the
        step (bench.py's operator-level e2e arm)."""
        p = self.pipe
        L = capi.lib()
        scale = 1.0 / self.world if average else 1.0
        capi.check(L.gsb_mask_rgb_grad(p.n, capi.ptr(p.rgbs), capi.ptr(self.v_rgb), capi.stream()))
        # every rank has finished writing this step's v_rgb and geometry gradients
        self.hdl.barrier(channel=0)
        capi.check(L.gsb_exchange_gradients(
            p.n, p.deg, p.deg, capi.ptr(p.p["means"]), self.world, capi.ptr(self.cam_positions),
            self.rgb_ptrs.data_ptr(), scale, capi.ptr(p.g["coeffs"]), self.rank, self.world, self.geom_numel,
            self.geom_ptrs.data_ptr(), self.multicast_ptr if self.multicast_ptr else None, capi.stream()))
        self.hdl.barrier(channel=2)
