// project_gaussians.hpp -- autograd operator ProjectGaussians on the B200 back end.
// Same class name, argument list and return slots as the reference operator
// (/root/reference/project_gaussians.hpp:12-30, project_gaussians.cpp:5-90) so that model.cpp:147-169
// and simple_trainer.cpp:173-180 compile unchanged.
#pragma once
#include <torch/torch.h>
#include "tile_bounds.hpp"
#include "gsplat.hpp"

using namespace torch::autograd;

class ProjectGaussians : public Function<ProjectGaussians> {
public:
    // returns { xys [N,2], depths [N], radii [N] i32, conics [N,3], numTilesHit [N] i32, cov3d [N,6] }
    static variable_list forward(AutogradContext *ctx, torch::Tensor means, torch::Tensor scales,
                                 float globScale, torch::Tensor quats, torch::Tensor viewMat,
                                 torch::Tensor projMat, float fx, float fy, float cx, float cy,
                                 int imgHeight, int imgWidth, TileBounds tileBounds,
                                 float clipThresh = 0.01);
    // 14 slots; gradients for means (0), scales (1), quats (3)
    static tensor_list backward(AutogradContext *ctx, tensor_list grad_outputs);
};

// The CPU flavour belongs to the reference's rasterizer/gsplat-cpu back end.  It is declared so that
// callers which branch on the device (model.cpp:123-141, simple_trainer.cpp:151-170) still compile;
// this back end's definition fails loudly -- there is no CPU path here.
class ProjectGaussiansCPU {
public:
    static variable_list apply(torch::Tensor means, torch::Tensor scales, float globScale,
                               torch::Tensor quats, torch::Tensor viewMat, torch::Tensor projMat, float fx,
                               float fy, float cx, float cy, int imgHeight, int imgWidth,
                               float clipThresh = 0.01);
};
