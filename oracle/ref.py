"""Loader for oracle/_ref/libopensplat_ref_cpu.so -- the UNMODIFIED reference CPU back end
(rasterizer/gsplat-cpu + the three operator .cpp files) compiled by oracle/Makefile from
/root/reference, exposed through our shim oracle/ref_driver.cpp as torch.ops.opensplat_ref.*.
TEST INFRASTRUCTURE ONLY; also bench.py's `--impl reference` arm.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libopensplat_ref_cpu.so")
_loaded = False


def available():
    return os.path.exists(_SO)


def build():
    """(Re)build from /root/reference when it is present (this container only)."""
    if os.path.exists("/root/reference/rasterizer/gsplat-cpu/gsplat_cpu.cpp"):
        subprocess.check_call(["make", "-C", _HERE, "-j5", "ref"], stdout=subprocess.DEVNULL)
    return available()


def ops():
    global _loaded
    import torch
    if not _loaded:
        if not available():
            raise RuntimeError("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
        torch.ops.load_library(_SO)
        _loaded = True
    return torch.ops.opensplat_ref


def render(means, scales, quats, viewmat, projmat, fx, fy, cx, cy, H, W, colors, opacity, background,
           dense_depths=True):
    """ProjectGaussiansCPU -> RasterizeGaussiansCPU exactly as simple_trainer.cpp:152-170, except
    that camDepths is made dense (as shipped the stride-3 view is read through a raw
    pointer and scrambles the compositing order)."""
    o = ops()
    p = o.project_cpu(means, scales, 1.0, quats, viewmat, projmat, fx, fy, cx, cy, H, W, 0.01)
    xys, radii, conics, cov2d, cam_depths = p
    if dense_depths:
        cam_depths = cam_depths.contiguous()
    img = o.rasterize_cpu(xys, radii, conics, colors, opacity, cov2d, cam_depths, H, W, background)
    return img, dict(xys=xys, radii=radii, conics=conics, cov2d=cov2d, cam_depths=cam_depths)
