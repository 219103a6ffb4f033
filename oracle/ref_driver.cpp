// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// Thin torch-library shim (our code) that exposes the UNMODIFIED reference CPU
// back end (compiled from /root/reference where it lies, see oracle/Makefile)
// to Python as `torch.ops.opensplat_ref.*`.  It calls exactly the entry points
// the reference's own simple_trainer.cpp:152-170 / model.cpp:123-205 call:
//   ProjectGaussiansCPU::apply       (project_gaussians.cpp:94)
//   RasterizeGaussiansCPU::apply     (rasterize_gaussians.cpp:144, bwd :182)
//   SphericalHarmonicsCPU::apply     (spherical_harmonics.cpp:66)
// Used (a) to pin oracle/gsplat_oracle.c, (b) to generate tests/golden/*.npz,
// (c) as bench.py's `--impl reference` / cpu_baseline (kind "reference").
#include <torch/torch.h>
#include <torch/library.h>
#include "project_gaussians.hpp"
#include "rasterize_gaussians.hpp"
#include "spherical_harmonics.hpp"
#include "ssim.hpp"

static std::vector<torch::Tensor> ref_project_cpu(
    torch::Tensor means, torch::Tensor scales, double globScale, torch::Tensor quats,
    torch::Tensor viewMat, torch::Tensor projMat, double fx, double fy, double cx, double cy,
    int64_t imgHeight, int64_t imgWidth, double clipThresh) {
    // returns {xys, radii, conics, cov2d, camDepths}; differentiable through torch autograd
    return ProjectGaussiansCPU::apply(means, scales, (float)globScale, quats, viewMat, projMat,
                                      (float)fx, (float)fy, (float)cx, (float)cy,
                                      (int)imgHeight, (int)imgWidth, (float)clipThresh);
}

static torch::Tensor ref_rasterize_cpu(
    torch::Tensor xys, torch::Tensor radii, torch::Tensor conics, torch::Tensor colors,
    torch::Tensor opacity, torch::Tensor cov2d, torch::Tensor camDepths,
    int64_t imgHeight, int64_t imgWidth, torch::Tensor background) {
    return RasterizeGaussiansCPU::apply(xys, radii, conics, colors, opacity, cov2d, camDepths,
                                        (int)imgHeight, (int)imgWidth, background);
}

static torch::Tensor ref_sh_cpu(int64_t degreesToUse, torch::Tensor viewDirs, torch::Tensor coeffs) {
    return SphericalHarmonicsCPU::apply((int)degreesToUse, viewDirs, coeffs);
}

// Model::mainLoss (model.cpp:780-784) with the reference's own SSIM class (ssim.cpp, compiled unmodified) and
// its l1 helper (model.cpp:54-56: (rendered - gt).abs().mean()); model.cpp itself needs OpenCV/nanoflann headers
// that are unavailable offline, so these three lines are restated here around the reference's SSIM::eval.
static torch::Tensor ref_main_loss_cpu(torch::Tensor rgb, torch::Tensor gt, double ssimWeight) {
    static SSIM ssim(11, 3);  // model.hpp:32
    torch::Tensor ssimLoss = 1.0f - ssim.eval(rgb, gt);
    torch::Tensor l1Loss = (rgb - gt).abs().mean();
    return (1.0f - (float)ssimWeight) * l1Loss + (float)ssimWeight * ssimLoss;
}

static int64_t ref_num_threads() { return (int64_t)at::get_num_threads(); }

TORCH_LIBRARY(opensplat_ref, m) {
    m.def("project_cpu", &ref_project_cpu);
    m.def("rasterize_cpu", &ref_rasterize_cpu);
    m.def("sh_cpu", &ref_sh_cpu);
    m.def("main_loss_cpu", &ref_main_loss_cpu);
    m.def("num_threads", &ref_num_threads);
}
