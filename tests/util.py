import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def image_close(a, b, tol=1e-5, flip=4.5e-3, frac=1e-3):
    """SURVEY 8c tolerance statement: >= (1-frac) of pixels within tol, the rest (isolated
    alpha-threshold flips) bounded by 1/255 (+ slack)."""
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max(axis=-1)
    bad = d > tol
    return bad.mean() <= frac and d.max() <= flip, (float(bad.mean()), float(d.max()))
