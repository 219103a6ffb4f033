"""Runs a few fwd+bwd steps of the bench workload (no timing, no CPU baseline) -- the command wrapped
by ncu under gpurun (see profiles/README.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS
from opensplat_b200.pipeline import SplatPipeline
from opensplat_b200.scene import make_scene

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_1M_1080p_sh3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, W, H, scale, opac = WORKLOADS[wl]
sc = make_scene(n, W, H, scale=scale, sh_degree=3, opacity=opac, seed=0)
pipe = SplatPipeline(n, W, H, device="cuda:0")
pipe.load_scene(sc)
pipe.target.copy_(torch.from_numpy(np.random.default_rng(1).uniform(0, 1, (H, W, 3)).astype(np.float32)))
for _ in range(steps):
    pipe.forward_backward()
torch.cuda.synchronize()
print("M", pipe.m, "loss", float(pipe.loss))
