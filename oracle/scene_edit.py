"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's topology edits and scene writers.

Restates, with the same ATen (torch CPU) operations in the same order as the reference issues them through
libtorch (the reference's own arithmetic here IS a sequence of ATen calls; torch 2.11.0 is the pinned dependency):
  * densification statistics            Model::afterTrain   model.cpp:317-337
  * split / duplicate / cull + Adam     Model::afterTrain   model.cpp:339-470, addToOptimizer :253-279,
    state surgery                                           removeFromOptimizer :281-308
  * opacity reset                       model.cpp:472-487
  * PLY / .splat bodies                 Model::savePly :505-558, Model::saveSplat :560-594
Pinned against the unmodified reference model.cpp compiled into oracle/_ref (tests/golden/make_golden_scene_edit.py
-> tests/golden/scene_edit_*.npz, checked by tests/test_oracle_vs_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""

import numpy as np
import torch

SH_C0 = 0.28209479177387814


def _t(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x)).to(dtype) if not isinstance(x, torch.Tensor) else x.detach().cpu().to(dtype)


def quat_to_rotmat(quat):
    """tensor_math.cpp:5-28."""
    u = torch.unbind(torch.nn.functional.normalize(quat, dim=-1), -1)
    w, x, y, z = u
    return torch.stack([
        torch.stack([1.0 - 2.0 * (y.pow(2) + z.pow(2)), 2.0 * (x * y - w * z), 2.0 * (x * z + w * y)], -1),
        torch.stack([2.0 * (x * y + w * z), 1.0 - 2.0 * (x.pow(2) + z.pow(2)), 2.0 * (y * z - w * x)], -1),
        torch.stack([2.0 * (x * z - w * y), 2.0 * (y * z + w * x), 1.0 - 2.0 * (x.pow(2) + y.pow(2))], -1),
    ], -2)


def densify_stats(state, v_xy, radii, img_h, img_w):
    """model.cpp:317-337.  state = None (first step after a clear) or (xysGradNorm, visCounts, max2DSize)."""
    v_xy, radii = _t(v_xy), _t(radii, torch.int32)
    visible = (radii > 0).flatten()
    grads = torch.linalg.vector_norm(v_xy, 2, dim=-1)
    if state is None:
        gn, vc = grads.clone(), torch.ones_like(grads)
        ms = torch.zeros_like(radii, dtype=torch.float32)
    else:
        gn, vc, ms = (_t(s).clone() for s in state)
        vc[visible] = vc[visible] + 1
        gn[visible] = grads[visible] + gn[visible]
    new_radii = radii[visible]
    ms[visible] = torch.maximum(ms[visible], new_radii / float(max(img_h, img_w)))
    return gn, vc, ms


def refine(params, adam_m, adam_v, stats, max_dim, cfg, check_screen, check_huge, samples):
    """model.cpp:339-470 on dicts of CPU tensors.  `samples` = the torch::randn({2*nSplits,3}) draw, or a callable
    n_splits -> samples.  Returns (new_params, new_m, new_v, info) with info["src_map"] in the C ABI's encoding and
    info["margin"] = per-parent distance of the closest threshold comparison (tests skip knife-edge parents)."""
    p = {k: _t(v).clone() for k, v in params.items()}
    m = {k: _t(v).clone() for k, v in (adam_m or {}).items()}
    v = {k: _t(v_).clone() for k, v_ in (adam_v or {}).items()}
    gn, vc, ms = (_t(s) for s in stats)
    n = p["means"].shape[0]
    f32 = torch.float32
    avg = (gn / vc) * 0.5 * float(max_dim)
    high = (avg > cfg.densify_grad_thresh).squeeze()
    mx = p["scales"].exp().max(-1)[0]
    splits = (mx > cfg.densify_size_thresh).squeeze()
    if check_screen:
        splits = splits | (ms > cfg.split_screen_size).squeeze()
    splits = splits & high
    n_splits = int(splits.sum())
    if callable(samples):
        samples = samples(n_splits)
    samples = _t(samples).reshape(2 * n_splits, 3)
    scaled = torch.exp(p["scales"][splits].repeat(2, 1)) * samples
    qs = p["quats"][splits] / torch.linalg.vector_norm(p["quats"][splits], 2, dim=-1, keepdim=True)
    rots = quat_to_rotmat(qs.repeat(2, 1))
    rotated = torch.bmm(rots, scaled[..., None]).squeeze(-1)
    split_means = rotated + p["means"][splits].repeat(2, 1)
    split_scales = torch.log(torch.exp(p["scales"][splits]) / cfg.size_fac).repeat(2, 1)
    dups = (mx <= cfg.densify_size_thresh).squeeze() & high
    idx = torch.arange(n)
    parent = torch.cat([idx, idx[splits].repeat(2), idx[dups]])
    kind = torch.cat([torch.zeros(n, dtype=torch.int64), torch.ones(n_splits, dtype=torch.int64),
                      torch.full((n_splits,), 2, dtype=torch.int64), torch.full((int(dups.sum()),), 3, dtype=torch.int64)])
    cat = {}
    for k, t in p.items():
        if k == "means":
            cat[k] = torch.cat([t, split_means, t[dups]], 0)
        elif k == "scales":
            cat[k] = torch.cat([t, split_scales, t[dups]], 0)
        else:
            reps = (2,) + (1,) * (t.dim() - 1)
            cat[k] = torch.cat([t, t[splits].repeat(*reps), t[dups]], 0)
    n_add = 2 * n_splits + int(dups.sum())

    def grow_state(s):
        return {k: torch.cat([t, torch.zeros((n_add,) + tuple(t.shape[1:]), dtype=f32)], 0) for k, t in s.items()}
    m, v = grow_state(m), grow_state(v)
    ms_cat = torch.cat([ms, torch.zeros(n_add)])
    splits_mask = torch.cat([splits, torch.zeros(n_add, dtype=torch.bool)])
    # cull (model.cpp:437-468)
    sig = torch.sigmoid(cat["opacities"]).squeeze(-1)
    culls = (sig < cfg.cull_alpha_thresh) | splits_mask
    mx_cat = torch.exp(cat["scales"]).max(-1)[0]
    if check_huge:
        huge = mx_cat > cfg.cull_scale_thresh
        if check_screen:
            huge = huge | (ms_cat > cfg.cull_screen_size)
        culls = culls | huge
    keep = ~culls
    new_p = {k: t[keep] for k, t in cat.items()}
    new_m = {k: t[keep] for k, t in m.items()}
    new_v = {k: t[keep] for k, t in v.items()}
    src_map = (parent[keep] | (kind[keep] << 30)).to(torch.int32)
    # knife-edge margins per parent (relative distance of every compared quantity to its threshold)
    def rel(a, thr):
        return (a - thr).abs() / max(abs(thr), 1e-30)
    margin = torch.minimum(rel(avg.squeeze(), cfg.densify_grad_thresh), rel(mx, cfg.densify_size_thresh))
    margin = torch.minimum(margin, rel(sig[:n], cfg.cull_alpha_thresh))
    if check_screen:
        margin = torch.minimum(margin, rel(ms, cfg.split_screen_size))
    if check_huge:
        margin = torch.minimum(margin, rel(mx, cfg.cull_scale_thresh))
        margin = torch.minimum(margin, rel(mx / cfg.size_fac, cfg.cull_scale_thresh))
        if check_screen:
            margin = torch.minimum(margin, rel(ms, cfg.cull_screen_size))
    info = {"n_splits": n_splits, "n_dups": int(dups.sum()), "new_n": int(keep.sum()), "src_map": src_map,
            "splits": splits, "dups": dups, "culls": culls, "margin": margin, "samples": samples}
    return new_p, new_m, new_v, info


def reset_opacity(opacities, cull_alpha_thresh=0.1):
    """model.cpp:472-475."""
    reset_value = cull_alpha_thresh * 2.0
    return torch.clamp_max(_t(opacities), float(torch.logit(torch.tensor(reset_value, dtype=torch.float32))))


# ---- scene writers ---------------------------------------------------------------------------------------------
def ply_header(n, num_rest, step):
    """model.cpp:509-545 (std::endl = '\\n')."""
    lines = ["ply", "format binary_little_endian 1.0", f"comment Generated by opensplat at iteration {step}",
             f"element vertex {n}", "property float x", "property float y", "property float z", "property float nx",
             "property float ny", "property float nz"]
    lines += [f"property float f_dc_{i}" for i in range(3)]
    lines += [f"property float f_rest_{i}" for i in range(num_rest)]
    lines += ["property float opacity", "property float scale_0", "property float scale_1", "property float scale_2",
              "property float rot_0", "property float rot_1", "property float rot_2", "property float rot_3",
              "end_header"]
    return ("\n".join(lines) + "\n").encode()


def ply_body(means, features_dc, features_rest, opacities, scales, quats, keep_crs=False, scale=1.0,
             translation=(0.0, 0.0, 0.0)):
    """model.cpp:525,547-557 -> bytes of the vertex rows."""
    means, dc, rest, op, sc, q = (_t(a) for a in (means, features_dc, features_rest, opacities, scales, quats))
    n = means.shape[0]
    rest_t = rest.transpose(1, 2).reshape(n, -1)
    means_c = (means / scale) + _t(np.asarray(translation, dtype=np.float32)) if keep_crs else means
    scales_c = torch.log(torch.exp(sc) / scale) if keep_crs else sc
    rows = torch.cat([means_c, torch.zeros(n, 3), dc, rest_t, op.reshape(n, 1), scales_c, q], 1)
    return rows.contiguous().numpy().astype("<f4").tobytes()


def ply_load(blob, keep_crs=False, scale=1.0, translation=(0.0, 0.0, 0.0)):
    """Model::loadPly's data path (model.cpp:724-746) on a file produced by ply_header + ply_body:
    returns (dict of the six tensors, step)."""
    end = blob.index(b"end_header\n") + len(b"end_header\n")
    lines = blob[:end].decode().split("\n")
    step = int(lines[2].rsplit(" ", 1)[1])
    n = int(lines[3].rsplit(" ", 1)[1])
    n_rest = sum(1 for l in lines if l.startswith("property float f_rest_"))
    rows = torch.from_numpy(np.frombuffer(blob[end:], "<f4").reshape(n, 17 + n_rest).copy())
    means, dc = rows[:, 0:3].clone(), rows[:, 6:9].clone()
    rest = rows[:, 9:9 + n_rest].clone()
    opac = rows[:, 9 + n_rest:10 + n_rest].clone()
    sc = rows[:, 10 + n_rest:13 + n_rest].clone()
    q = rows[:, 13 + n_rest:17 + n_rest].clone()
    if keep_crs:
        means = (means - _t(np.asarray(translation, dtype=np.float32))) * scale
        sc = torch.log(scale * torch.exp(sc))
    rest = rest.reshape(n, 3, n_rest // 3).transpose(2, 1).contiguous()
    return {"means": means, "scales": sc, "quats": q, "featuresDc": dc, "featuresRest": rest, "opacities": opac}, step


def splat_rows(means, features_dc, opacities, scales, quats, keep_crs=False, scale=1.0,
               translation=(0.0, 0.0, 0.0)):
    """model.cpp:560-583 -> (unordered rows as a [n,32] uint8 array, sort key [n] f32)."""
    means, dc, op, sc, q = (_t(a) for a in (means, features_dc, opacities, scales, quats))
    n = means.shape[0]
    means_c = (means / scale) + _t(np.asarray(translation, dtype=np.float32)) if keep_crs else means
    scales_c = (torch.exp(sc) / scale) if keep_crs else torch.exp(sc)
    rgbs = (torch.clamp((dc * SH_C0) + 0.5, 0.0, 1.0) * 255.0).to(torch.uint8)   # sh2rgb, spherical_harmonics.cpp:25-28
    opac = 1.0 + torch.exp(-op)
    alpha = torch.clamp((1.0 / opac) * 255.0, 0.0, 255.0).to(torch.uint8)
    quats8 = torch.clamp(q * 128.0 + 128.0, 0.0, 255.0).to(torch.uint8)
    order_key = (scales_c[..., 0] + scales_c[..., 1] + scales_c[..., 2]) / opac[..., 0]
    rows = np.zeros((n, 32), dtype=np.uint8)
    rows[:, 0:12] = means_c.contiguous().numpy().astype("<f4").view(np.uint8).reshape(n, 12)
    rows[:, 12:24] = scales_c.contiguous().numpy().astype("<f4").view(np.uint8).reshape(n, 12)
    rows[:, 24:27] = rgbs.numpy()
    rows[:, 27:28] = alpha.numpy().reshape(n, 1)
    rows[:, 28:32] = quats8.numpy()
    return rows, order_key.numpy()


def splat_body(*args, **kw):
    """Rows in the reference's order (descending key; ties by ascending index, where std::sort is unspecified)."""
    rows, key = splat_rows(*args, **kw)
    order = np.lexsort((np.arange(len(key)), -key.astype(np.float64)))
    return rows[order].tobytes(), order
