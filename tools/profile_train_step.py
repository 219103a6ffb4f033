"""A few full train steps (fwd + bwd + fused Adam) for ncu captures of the streaming kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import WORKLOADS
from opensplat_b200.pipeline import SplatPipeline
from opensplat_b200.scene import make_scene
n, W, H, scale, opac = WORKLOADS["c2_1M_1080p_sh3"]
sc = make_scene(n, W, H, scale=scale, sh_degree=3, opacity=opac, seed=0)
pipe = SplatPipeline(n, W, H, device="cuda:0")
pipe.load_scene(sc)
pipe.target.copy_(torch.from_numpy(np.random.default_rng(1).uniform(0, 1, (H, W, 3)).astype(np.float32)))
for _ in range(2):
    pipe.train_step(lr=1e-4)
torch.cuda.synchronize()
print("M", pipe.m)
