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

// The parameter activations of Model::forward in one pass each way (model.cpp:148-150,176-177,200):
//   {scales = exp(logScales), quats = rawQuats / |rawQuats|, opacities [N,1] = sigmoid(opacityLogits),
//    viewDirs = normalize(means - camPos)  (detached: no gradient)}.       camPos: 3-element tensor (any device).
//   auto a = gsb::ActivateGaussians::apply(means, scales, quats, opacities, camPos);   // a[0..3]
class ActivateGaussians : public torch::autograd::Function<ActivateGaussians> {
public:
    static torch::autograd::tensor_list forward(torch::autograd::AutogradContext *ctx, torch::Tensor means,
                                                torch::Tensor logScales, torch::Tensor rawQuats,
                                                torch::Tensor opacityLogits, torch::Tensor camPos);
    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext *ctx,
                                                 torch::autograd::tensor_list grad_outputs);
};

// The colour pass of Model::forward without its ATen glue (model.cpp:176-177,186-192):
//   rgbs = clamp_min(SphericalHarmonics(degreesToUse, means - camPos, cat(featuresDc[:,None,:], featuresRest)) + 0.5, 0)
// reading featuresDc [N,3] / featuresRest [N,K-1,3] where they lie (no 12K-byte/Gaussian cat, no split in autograd)
// and writing their two gradients directly.
//   torch::Tensor rgbs = gsb::SphericalHarmonicsRgb::apply(degreesToUse, means, camPos, featuresDc, featuresRest);
class SphericalHarmonicsRgb : public torch::autograd::Function<SphericalHarmonicsRgb> {
public:
    static torch::Tensor forward(torch::autograd::AutogradContext *ctx, int64_t degreesToUse, torch::Tensor means,
                                 torch::Tensor camPos, torch::Tensor featuresDc, torch::Tensor featuresRest);
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
