"""opensplat_b200 -- Blackwell (sm_100a) differentiable Gaussian-splat render path.

Drop-in for the hot path behind OpenSplat's autograd operators (ProjectGaussians,
RasterizeGaussians, SphericalHarmonics).  See DESIGN.md / INTEGRATION.md.
"""
__version__ = "0.1.0"
