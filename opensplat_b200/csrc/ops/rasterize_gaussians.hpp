// rasterize_gaussians.hpp -- autograd operator RasterizeGaussians (+ binAndSortGaussians) on the B200
// back end.  API as /root/reference/rasterize_gaussians.hpp:11-37, rasterize_gaussians.cpp:6-140.
#pragma once
#include <torch/torch.h>
#include "tile_bounds.hpp"

using namespace torch::autograd;

// -> { isectIds [M] i64, gaussianIds [M] i32, isectIdsSorted [M] i64, gaussianIdsSorted [M] i32,
//      tileBins [tiles,2] i32 }
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
binAndSortGaussians(int numPoints, int numIntersects, torch::Tensor xys, torch::Tensor depths,
                    torch::Tensor radii, torch::Tensor cumTilesHit, TileBounds tileBounds);

class RasterizeGaussians : public Function<RasterizeGaussians> {
public:
    // returns the rendered image [H,W,3]
    static torch::Tensor forward(AutogradContext *ctx, torch::Tensor xys, torch::Tensor depths,
                                 torch::Tensor radii, torch::Tensor conics, torch::Tensor numTilesHit,
                                 torch::Tensor colors, torch::Tensor opacity, int imgHeight, int imgWidth,
                                 torch::Tensor background);
    // 10 slots; gradients for xys (0), conics (3), colors (5), opacity (6)
    static tensor_list backward(AutogradContext *ctx, tensor_list grad_outputs);
};

// Declared for callers that branch on the device (model.cpp:195-205); defined to fail loudly.
class RasterizeGaussiansCPU : public Function<RasterizeGaussiansCPU> {
public:
    static torch::Tensor forward(AutogradContext *ctx, torch::Tensor xys, torch::Tensor radii,
                                 torch::Tensor conics, torch::Tensor colors, torch::Tensor opacity,
                                 torch::Tensor cov2d, torch::Tensor camDepths, int imgHeight, int imgWidth,
                                 torch::Tensor background);
    static tensor_list backward(AutogradContext *ctx, tensor_list grad_outputs);
};
