"""Generates tests/golden/scene_edit_*.npz by running the REFERENCE ITSELF: the unmodified model.cpp
(Model::afterTrain, Model::save) compiled from /root/reference into oracle/_ref/libopensplat_ref_model.so and
driven through tests/native/model_driver.cpp.  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden_scene_edit.py

Inputs are regenerated from a seed (tests/util.py scene_edit_inputs), so the fixtures hold only the reference's
outputs.  They pin oracle/scene_edit.py (tests/test_oracle_vs_golden.py) and are compared directly against the
CUDA path (tests/test_gpu_scene_edit.py)."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import PARAM_NAMES, scene_edit_inputs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.ops.load_library(os.path.join(ROOT, "oracle", "_ref", "libopensplat_ref_model.so"))
REF = torch.ops.opensplat_ref_model

# name -> (n, sh_bases, seed, H, W, cfg, steps): each step runs Model::afterTrain once, statistics carried over.
# cfg = numCameras, refineEvery, warmupLength, resetAlphaEvery, densifyGradThresh, densifySizeThresh,
#       stopScreenSizeAt, splitScreenSize, maxSteps
CASES = {
    # two accumulate steps then densify with the screen-size rules active and no "huge" cull (step <= 30)
    "scene_edit_densify_screen": (700, 16, 11, 480, 640, (1, 10, 5, 3, 0.0002, 0.01, 40, 0.05, 200), (18, 19, 20)),
    # densify + huge cull, screen-size rules off (step >= stopScreenSizeAt)
    "scene_edit_densify_huge": (900, 4, 12, 640, 400, (1, 10, 5, 3, 0.0002, 0.01, 40, 0.05, 200), (48, 49, 50)),
    # densify + huge cull + screen-size rules (both on)
    "scene_edit_densify_all": (900, 4, 13, 480, 640, (1, 10, 5, 3, 0.0002, 0.01, 100, 0.05, 200), (49, 50)),
    # refine step that only resets alpha (step % resetInterval == refineEvery)
    "scene_edit_alpha_reset": (500, 4, 14, 480, 640, (1, 10, 5, 3, 0.0002, 0.01, 100, 0.05, 200), (39, 40)),
}
SEED_RANDN = 1234


def run_case(name, n, k, seed, H, W, cfg, steps):
    p, m, v, draws = scene_edit_inputs(n, k, seed, max(H, W))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    params = [t(p[x]) for x in PARAM_NAMES]
    ms = [t(m[x]) for x in PARAM_NAMES]
    vs = [t(v[x]) for x in PARAM_NAMES]
    stats = []
    out = {}
    for si, step in enumerate(steps):
        v_xy, radii = draws[si]
        if params[0].shape[0] != n:
            raise RuntimeError("only the last step of a case may change the Gaussian count")
        r = REF.after_train(params, ms, vs, stats, t(v_xy), t(radii), H, W, step, SEED_RANDN, *cfg)
        params, ms, vs, st = r[0:6], r[6:12], r[12:18], r[18:21]
        stats = [] if st[0].numel() == 0 else list(st)
        for i, x in enumerate(("xysGradNorm", "visCounts", "max2DSize")):
            out[f"s{si}_{x}"] = st[i].numpy()
    for i, x in enumerate(PARAM_NAMES):
        out["p_" + x], out["m_" + x], out["v_" + x] = params[i].numpy(), ms[i].numpy(), vs[i].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n=n, k=k, seed=seed, hw=np.array([H, W]),
                        cfg=np.array(cfg, np.float64), steps=np.array(steps), seed_randn=SEED_RANDN, **out)
    print(name, "n", n, "->", params[0].shape[0])


def save_case(name, n, k, seed, keep_crs, scale, translation):
    p, _, _, _ = scene_edit_inputs(n, k, seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    params = [t(p[x]) for x in PARAM_NAMES]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for ext in ("ply", "splat"):
            fn = os.path.join(d, "scene." + ext)
            REF.save(params, fn, 7000, keep_crs, scale, t(np.asarray(translation, np.float32)))
            out[ext] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
        # ... and what the reference's own Model::loadPly makes of its file (resume path)
        ld = REF.load(os.path.join(d, "scene.ply"), keep_crs, scale, t(np.asarray(translation, np.float32)), torch.zeros(1))
        out["ld_step"] = np.array(int(ld[0]))
        for i, x in enumerate(PARAM_NAMES):
            out["ld_" + x] = ld[1 + i].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n=n, k=k, seed=seed, keep_crs=keep_crs, scale=scale,
                        translation=np.asarray(translation, np.float32), step=7000, **out)
    print(name, {k_: v_.size for k_, v_ in out.items()})


if __name__ == "__main__":
    for name, args in CASES.items():
        run_case(name, *args)
    save_case("scene_edit_save", 400, 16, 21, False, 1.0, (0.0, 0.0, 0.0))
    save_case("scene_edit_save_crs", 300, 4, 22, True, 0.37, (12.5, -3.25, 100.0))
