"""Loader for the C++/libtorch operator layer (opensplat_b200/lib/libopensplat_b200_ops.so):
torch.ops.opensplat_b200.{project_gaussians, rasterize_gaussians, spherical_harmonics,
bin_and_sort_gaussians} call the SAME autograd classes a C++ caller of the reference API uses
(opensplat_b200/csrc/ops/*.hpp)."""
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libopensplat_b200_ops.so")
_loaded = False


def available():
    return os.path.exists(_SO)


def ops():
    global _loaded
    if not _loaded:
        if not available():
            raise RuntimeError(f"{_SO} not built -- run `python -m opensplat_b200.build_ops`")
        torch.ops.load_library(_SO)
        _loaded = True
    return torch.ops.opensplat_b200
