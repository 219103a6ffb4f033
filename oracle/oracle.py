"""ctypes/numpy front end of oracle/gsplat_oracle.c (plain-C restatement, TEST INFRASTRUCTURE ONLY).

Every function mirrors one reference entry point; see gsplat_oracle.c for file:line citations.
Pinned against the reference itself through tests/golden (tests/test_oracle_vs_golden.py).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "gsplat_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_num_sh_bases.restype = C.c_int
        L.orc_cumsum_i32.restype = C.c_int64
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _fp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def num_sh_bases(degree):
    return lib().orc_num_sh_bases(int(degree))


def sh_forward(degrees_to_use, viewdirs, coeffs):
    viewdirs, coeffs = _f(viewdirs), _f(coeffs)
    n, K = coeffs.shape[0], coeffs.shape[1]
    degree = {1: 0, 4: 1, 9: 2, 16: 3}.get(K, 4)
    out = np.zeros((n, 3), np.float32)
    lib().orc_sh_forward(n, degree, int(degrees_to_use), _fp(viewdirs), _fp(coeffs), _fp(out))
    return out


def sh_backward(degrees_to_use, K, viewdirs, v_colors):
    viewdirs, v_colors = _f(viewdirs), _f(v_colors)
    n = viewdirs.shape[0]
    degree = {1: 0, 4: 1, 9: 2, 16: 3}.get(K, 4)
    out = np.zeros((n, K, 3), np.float32)
    lib().orc_sh_backward(n, degree, int(degrees_to_use), _fp(viewdirs), _fp(v_colors), _fp(out))
    return out


def project_forward(means, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_h, img_w,
                    clip_thresh=0.01):
    means, scales, quats = _f(means), _f(scales), _f(quats)
    viewmat, projmat = _f(viewmat).reshape(16), _f(projmat).reshape(16)
    n = means.shape[0]
    tx, ty = (img_w + 15) // 16, (img_h + 15) // 16
    cov3d = np.zeros((n, 6), np.float32)
    xys = np.zeros((n, 2), np.float32)
    depths = np.zeros((n,), np.float32)
    radii = np.zeros((n,), np.int32)
    conics = np.zeros((n, 3), np.float32)
    nth = np.zeros((n,), np.int32)
    lib().orc_project_forward(
        n, _fp(means), _fp(scales), C.c_float(glob_scale), _fp(quats), _fp(viewmat), _fp(projmat),
        C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), int(img_h), int(img_w), tx, ty,
        C.c_float(clip_thresh), _fp(cov3d), _fp(xys), _fp(depths), _fp(radii), _fp(conics), _fp(nth))
    return dict(cov3d=cov3d, xys=xys, depths=depths, radii=radii, conics=conics, num_tiles_hit=nth)


def project_backward(means, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_h, img_w,
                     radii, conics, v_xy, v_depth, v_conic):
    means, scales, quats = _f(means), _f(scales), _f(quats)
    viewmat, projmat = _f(viewmat).reshape(16), _f(projmat).reshape(16)
    radii, conics, v_xy, v_conic = _i(radii), _f(conics), _f(v_xy), _f(v_conic)
    v_depth = _f(v_depth) if v_depth is not None else None
    n = means.shape[0]
    vm = np.zeros((n, 3), np.float32)
    vs = np.zeros((n, 3), np.float32)
    vq = np.zeros((n, 4), np.float32)
    lib().orc_project_backward(
        n, _fp(means), _fp(scales), C.c_float(glob_scale), _fp(quats), _fp(viewmat), _fp(projmat),
        C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), int(img_h), int(img_w),
        _fp(radii), _fp(conics), _fp(v_xy), _fp(v_depth), _fp(v_conic), _fp(vm), _fp(vs), _fp(vq))
    return dict(v_mean3d=vm, v_scale=vs, v_quat=vq)


def cumsum(num_tiles_hit):
    a = _i(num_tiles_hit)
    out = np.zeros_like(a)
    total = lib().orc_cumsum_i32(a.shape[0], _fp(a), _fp(out))
    return out, int(total)


def bin_and_sort(xys, depths, radii, cum_tiles_hit, img_h, img_w):
    """binAndSortGaussians (rasterize_gaussians.cpp:6-37) -> dict of the five reference outputs."""
    xys, depths, radii, cum = _f(xys), _f(depths), _i(radii), _i(cum_tiles_hit)
    n = xys.shape[0]
    m = int(cum[-1]) if n else 0
    tx, ty = (img_w + 15) // 16, (img_h + 15) // 16
    isect = np.zeros((m,), np.int64)
    gids = np.zeros((m,), np.int32)
    lib().orc_map_gaussian_to_intersects(n, _fp(xys), _fp(depths), _fp(radii), _fp(cum), tx, ty,
                                         _fp(isect), _fp(gids))
    ks = np.zeros((m,), np.int64)
    idx = np.zeros((m,), np.int32)
    lib().orc_sort_isects(C.c_int64(m), _fp(isect), _fp(ks), _fp(idx))
    gsorted = gids[idx] if m else gids
    bins = np.zeros((tx * ty, 2), np.int32)
    lib().orc_tile_bin_edges(C.c_int64(m), tx * ty, _fp(ks), _fp(bins))
    return dict(isect_ids=isect, gaussian_ids=gids, isect_ids_sorted=ks, sorted_index=idx,
                gaussian_ids_sorted=np.ascontiguousarray(gsorted), tile_bins=bins)


def rasterize_forward(img_h, img_w, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities,
                      background, exp_mode=0):
    gs, tb = _i(gaussian_ids_sorted), _i(tile_bins)
    xys, conics, colors = _f(xys), _f(conics), _f(colors)
    opac, bg = _f(opacities).reshape(-1), _f(background)
    tx, ty = (img_w + 15) // 16, (img_h + 15) // 16
    out = np.zeros((img_h, img_w, 3), np.float32)
    fT = np.zeros((img_h, img_w), np.float32)
    fI = np.zeros((img_h, img_w), np.int32)
    lib().orc_rasterize_forward(int(img_h), int(img_w), tx, ty, _fp(gs), _fp(tb), _fp(xys), _fp(conics),
                                _fp(colors), _fp(opac), _fp(bg), int(exp_mode), _fp(out), _fp(fT), _fp(fI))
    return dict(out_img=out, final_Ts=fT, final_idx=fI)


def rasterize_backward(img_h, img_w, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities,
                       background, final_Ts, final_idx, v_output, v_output_alpha=None, exp_mode=0):
    gs, tb = _i(gaussian_ids_sorted), _i(tile_bins)
    xys, conics, colors = _f(xys), _f(conics), _f(colors)
    opac, bg = _f(opacities).reshape(-1), _f(background)
    fT, fI, vo = _f(final_Ts), _i(final_idx), _f(v_output)
    voa = _f(v_output_alpha) if v_output_alpha is not None else None
    n = xys.shape[0]
    tx, ty = (img_w + 15) // 16, (img_h + 15) // 16
    v_xy = np.zeros((n, 2), np.float32)
    v_conic = np.zeros((n, 3), np.float32)
    v_col = np.zeros((n, 3), np.float32)
    v_op = np.zeros((n, 1), np.float32)
    lib().orc_rasterize_backward(int(img_h), int(img_w), tx, ty, _fp(gs), _fp(tb), _fp(xys), _fp(conics),
                                 _fp(colors), _fp(opac), _fp(bg), _fp(fT), _fp(fI), _fp(vo), _fp(voa),
                                 int(exp_mode), n, _fp(v_xy), _fp(v_conic), _fp(v_col), _fp(v_op))
    return dict(v_xy=v_xy, v_conic=v_conic, v_colors=v_col, v_opacity=v_op)


# ------------------------------------------------------------------------------------------------
# Training loss: Model::mainLoss model.cpp:780-784 = (1-w) * l1 + w * (1 - SSIM), with
# SSIM::eval ssim.cpp:8-32, window ssim.cpp:34-47 (gaussian(1.5) at floor((i-11)/2), normalised), zero padding 5,
# l1 model.cpp:54-56.  numpy/scipy restatement in float64 with the analytic gradient w.r.t. `rendered`.
# ------------------------------------------------------------------------------------------------
def ssim_window():
    i = np.arange(11, dtype=np.float32)
    d = np.floor((i - 11.0) / 2.0).astype(np.float32)
    g = np.exp(-(d ** 2) / np.float32(2.0 * 1.5 * 1.5)).astype(np.float32)
    return (g / g.sum()).astype(np.float64)


def main_loss(rendered, gt, ssim_weight):
    """rendered, gt: [H,W,3].  Returns dict(loss, l1, ssim, v_rendered)."""
    from scipy.ndimage import correlate1d
    w = ssim_window()
    y = np.asarray(rendered, np.float64)
    x = np.asarray(gt, np.float64)

    def filt(a, k):  # conv2d(padding=5) of the reference == separable cross-correlation, zero padded
        a = correlate1d(a, k, axis=0, mode="constant", cval=0.0)
        return correlate1d(a, k, axis=1, mode="constant", cval=0.0)

    mx, my = filt(x, w), filt(y, w)
    sxx, syy, sxy = filt(x * x, w) - mx * mx, filt(y * y, w) - my * my, filt(x * y, w) - mx * my
    C1, C2 = np.float32(0.01 * 0.01), np.float32(0.03 * 0.03)
    A1, A2 = 2 * mx * my + C1, 2 * sxy + C2
    B1, B2 = mx * mx + my * my + C1, sxx + syy + C2
    S = A1 * A2 / (B1 * B2)
    count = y.size
    ssim = S.mean()
    l1 = np.abs(x - y).mean()
    loss = (1 - ssim_weight) * l1 + ssim_weight * (1 - ssim)
    d_e12 = 2 * A1 / (B1 * B2)
    d_e22 = -S / B2
    d_mu = 2 * mx * (A2 - A1) / (B1 * B2) - 2 * my * S / B1 + 2 * my * S / B2
    wt = w[::-1].copy()  # transposed (adjoint) filter
    dssim = filt(d_mu, wt) + 2 * y * filt(d_e22, wt) + x * filt(d_e12, wt)
    v = -ssim_weight / count * dssim + (1 - ssim_weight) / count * np.sign(y - x)
    return dict(loss=float(loss), l1=float(l1), ssim=float(ssim), v_rendered=v.astype(np.float32))
