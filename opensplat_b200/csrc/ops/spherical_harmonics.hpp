// spherical_harmonics.hpp -- autograd operator SphericalHarmonics and the SH helpers on the B200 back
// end.  API as /root/reference/spherical_harmonics.hpp:9-31, spherical_harmonics.cpp:3-63.
#pragma once
#include <torch/torch.h>
#include "gsplat.hpp"

using namespace torch::autograd;

int degFromSh(int numBases);                    // 1,4,9,16 -> 0..3, otherwise 4
torch::Tensor rgb2sh(const torch::Tensor &rgb);  // (rgb - 0.5) / C0
torch::Tensor sh2rgb(const torch::Tensor &sh);   // clamp(sh * C0 + 0.5, 0, 1)

class SphericalHarmonics : public Function<SphericalHarmonics> {
public:
    static torch::Tensor forward(AutogradContext *ctx, int degreesToUse, torch::Tensor viewDirs,
                                 torch::Tensor coeffs);
    static tensor_list backward(AutogradContext *ctx, tensor_list grad_outputs);  // {none, none, v_coeffs}
};

// Declared for callers that branch on the device (model.cpp:181-182); defined to fail loudly.
class SphericalHarmonicsCPU {
public:
    static torch::Tensor apply(int degreesToUse, torch::Tensor viewDirs, torch::Tensor coeffs);
};
