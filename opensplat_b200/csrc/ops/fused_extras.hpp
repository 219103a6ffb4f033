// fused_extras.hpp -- OPTIONAL one-line replacements for the ATen glue around the three operators in the reference's
// Model.  Nothing in model.cpp needs them (it compiles and runs unchanged on this back end);
// a maintainer who wants the fused kernels swaps, e.g.,
//     torch::Tensor ssimLoss = 1.0f - ssim.eval(rgb, gt); ... return (1-w)*l1Loss + w*ssimLoss;   (model.cpp:780-784)
// for
//     return gsb::MainLoss::apply(rgb, gt, ssimWeight);
// and the six `xxxOpt->step()` calls (model.cpp:236-243) for gsb::adamStep on each parameter.
#pragma once
#include <torch/torch.h>

namespace gsb {

// Model::mainLoss = (1-w) * mean|rgb - gt| + w * (1 - SSIM(rgb, gt)) with the reference's SSIM (ssim.cpp:8-47),
// forward + gradient w.r.t. rgb in two fused tile kernels.  rgb, gt: [H,W,3] CUDA tensors.  Returns a scalar.
class MainLoss : public torch::autograd::Function<MainLoss> {
public:
    static torch::Tensor forward(torch::autograd::AutogradContext *ctx, torch::Tensor rgb, torch::Tensor gt,
                                 double ssimWeight);
    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext *ctx,
                                                 torch::autograd::tensor_list grad_outputs);
};

// One torch::optim::Adam step (no weight decay / amsgrad) on `param` in place with caller-held moments;
// `step` is the 1-based step count of this parameter (AdamParamState::step after the increment).
void adamStep(torch::Tensor param, const torch::Tensor &grad, torch::Tensor expAvg, torch::Tensor expAvgSq, double lr,
              int64_t step, double beta1 = 0.9, double beta2 = 0.999, double eps = 1e-8);

// Model::afterTrain statistics (model.cpp:317-337) in one pass; `first` = the three tensors are being (re)created.
void densifyStats(const torch::Tensor &xysGrad, const torch::Tensor &radii, int imgHeight, int imgWidth, bool first,
                  torch::Tensor xysGradNorm, torch::Tensor visCounts, torch::Tensor max2DSize);

}  // namespace gsb
