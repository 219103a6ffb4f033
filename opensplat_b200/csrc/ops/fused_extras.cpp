// fused_extras.cpp -- see fused_extras.hpp.  Thin libtorch wrappers over the C ABI (gsb_ssim_l1_loss, gsb_adam_step,
// gsb_densify_stats_*); Python has the same through opensplat_b200/ops.py and model.py.
#include "fused_extras.hpp"

#include <cmath>

#include "gsb_torch.hpp"

namespace gsb {

torch::Tensor MainLoss::forward(torch::autograd::AutogradContext *ctx, torch::Tensor rgb, torch::Tensor gt,
                                double ssimWeight) {
    TORCH_CHECK(rgb.dim() == 3 && rgb.size(2) == 3 && rgb.sizes() == gt.sizes(), "rgb and gt must be [H,W,3]");
    c10::cuda::CUDAGuard guard(rgb.device());
    const int H = (int)rgb.size(0), W = (int)rgb.size(1);
    torch::Tensor r = f32(rgb), g = f32(gt);
    torch::Tensor v = torch::empty_like(r);
    torch::Tensor out = torch::empty({3}, like(r, torch::kFloat32));
    const size_t wsBytes = gsb_ssim_workspace_bytes(H, W);
    torch::Tensor ws = torch::empty({(int64_t)wsBytes + 256}, like(r, torch::kUInt8));
    uint8_t *wp = ws.data_ptr<uint8_t>();
    const size_t off = (256 - ((uintptr_t)wp % 256)) % 256;
    check(gsb_ssim_l1_loss(H, W, fp(r), fp(g), (float)ssimWeight, fpw(v), fpw(out), wp + off, wsBytes, stream()),
          "gsb_ssim_l1_loss");
    ctx->save_for_backward({v});
    return out[0].clone();
}

torch::autograd::tensor_list MainLoss::backward(torch::autograd::AutogradContext *ctx,
                                                torch::autograd::tensor_list grad_outputs) {
    torch::Tensor v = ctx->get_saved_variables()[0];
    return {v * grad_outputs[0], torch::Tensor(), torch::Tensor()};
}

void adamStep(torch::Tensor param, const torch::Tensor &grad, torch::Tensor expAvg, torch::Tensor expAvgSq, double lr,
              int64_t step, double beta1, double beta2, double eps) {
    TORCH_CHECK(param.is_cuda() && param.is_contiguous() && param.scalar_type() == torch::kFloat32,
                "adamStep: param must be a contiguous fp32 CUDA tensor");
    TORCH_CHECK(expAvg.is_contiguous() && expAvgSq.is_contiguous() && expAvg.numel() == param.numel() &&
                    expAvgSq.numel() == param.numel() && grad.numel() == param.numel(),
                "adamStep: moment / gradient size mismatch");
    TORCH_CHECK(step >= 1, "adamStep: step is 1-based");
    c10::cuda::CUDAGuard guard(param.device());
    torch::NoGradGuard noGrad;
    torch::Tensor g = f32(grad);
    check(gsb_adam_step(param.numel(), param.data_ptr<float>(), fp(g), expAvg.data_ptr<float>(),
                        expAvgSq.data_ptr<float>(), (float)lr, (float)beta1, (float)beta2, (float)eps,
                        (float)(1.0 - std::pow(beta1, (double)step)), (float)(1.0 - std::pow(beta2, (double)step)),
                        stream()),
          "gsb_adam_step");
}

void densifyStats(const torch::Tensor &xysGrad, const torch::Tensor &radii, int imgHeight, int imgWidth, bool first,
                  torch::Tensor xysGradNorm, torch::Tensor visCounts, torch::Tensor max2DSize) {
    const int n = (int)radii.numel();
    TORCH_CHECK(xysGradNorm.numel() == n && visCounts.numel() == n && max2DSize.numel() == n,
                "densifyStats: statistics tensors must have one entry per Gaussian");
    c10::cuda::CUDAGuard guard(radii.device());
    torch::Tensor g = f32(xysGrad), r = i32(radii);
    auto fn = first ? gsb_densify_stats_init : gsb_densify_stats_update;
    check(fn(n, fp(g), r.data_ptr<int32_t>(), imgHeight, imgWidth, xysGradNorm.data_ptr<float>(),
             visCounts.data_ptr<float>(), max2DSize.data_ptr<float>(), stream()),
          "gsb_densify_stats");
}

}  // namespace gsb
