// spherical_harmonics.cpp -- SphericalHarmonics over the C ABI (gsb_sh_forward / gsb_sh_backward).
// Replaces the reference's spherical_harmonics.cpp:3-63 + bindings.cu:68-124.
#include "spherical_harmonics.hpp"
#include "gsb_torch.hpp"

int degFromSh(int numBases) {
    if (numBases == 1) return 0;
    if (numBases == 4) return 1;
    if (numBases == 9) return 2;
    if (numBases == 16) return 3;
    return 4;
}

// public helper of the reference's CPU bindings that model.hpp:44 calls (gsplat_cpu.cpp:409-423)
int numShBases(int degree) {
    if (degree < 0 || degree > 4) return 25;
    return (degree + 1) * (degree + 1);
}

static const double kShC0 = 0.28209479177387814;

torch::Tensor rgb2sh(const torch::Tensor &rgb) { return (rgb - 0.5) / kShC0; }

torch::Tensor sh2rgb(const torch::Tensor &sh) { return torch::clamp(sh * kShC0 + 0.5, 0.0f, 1.0f); }

torch::Tensor SphericalHarmonics::forward(AutogradContext *ctx, int degreesToUse, torch::Tensor viewDirs,
                                          torch::Tensor coeffs) {
    const int n = (int)coeffs.size(0);
    const int degree = degFromSh((int)coeffs.size(-2));
    TORCH_CHECK(coeffs.dim() == 3 && coeffs.size(2) == 3, "coeffs must have dimensions (N, D, 3)");
    TORCH_CHECK(degreesToUse >= 0 && degreesToUse <= degree, "degreesToUse out of range");
    c10::cuda::CUDAGuard guard(coeffs.device());
    torch::Tensor vd = gsb::f32(viewDirs), co = gsb::f32(coeffs);
    torch::Tensor colors = torch::empty({n, 3}, gsb::like(co, torch::kFloat32));
    gsb::check(gsb_sh_forward(n, degree, degreesToUse, gsb::fp(vd), gsb::fp(co), gsb::fpw(colors), gsb::stream()),
               "gsb_sh_forward");
    ctx->saved_data["degreesToUse"] = degreesToUse;
    ctx->saved_data["degree"] = degree;
    ctx->save_for_backward({vd});
    return colors;
}

tensor_list SphericalHarmonics::backward(AutogradContext *ctx, tensor_list grad_outputs) {
    torch::Tensor vd = ctx->get_saved_variables()[0];
    const int degree = (int)ctx->saved_data["degree"].toInt();
    const int degreesToUse = (int)ctx->saved_data["degreesToUse"].toInt();
    torch::Tensor v_colors = gsb::f32(grad_outputs[0]);
    const int n = (int)v_colors.size(0);
    TORCH_CHECK(v_colors.dim() == 2 && v_colors.size(1) == 3, "v_colors must have dimensions (N, 3)");
    c10::cuda::CUDAGuard guard(v_colors.device());
    const int K = (degree + 1) * (degree + 1);
    torch::Tensor v_coeffs = torch::empty({n, K, 3}, gsb::like(v_colors, torch::kFloat32));
    gsb::check(gsb_sh_backward(n, degree, degreesToUse, gsb::fp(vd), gsb::fp(v_colors), gsb::fpw(v_coeffs),
                               gsb::stream()),
               "gsb_sh_backward");
    torch::Tensor none;
    return {none, none, v_coeffs};
}

torch::Tensor SphericalHarmonicsCPU::apply(int, torch::Tensor, torch::Tensor) {
    TORCH_CHECK(false, "SphericalHarmonicsCPU: the gsplat_b200 back end has no CPU path; link the reference's "
                       "rasterizer/gsplat-cpu for CPU execution");
    return {};
}
