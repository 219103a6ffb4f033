// project_gaussians.cpp -- ProjectGaussians over the C ABI (gsb_project_forward / gsb_project_backward).
// Replaces the reference's project_gaussians.cpp:5-90 + bindings.cu:133-277.
#include "project_gaussians.hpp"
#include "gsb_torch.hpp"

variable_list ProjectGaussians::forward(AutogradContext *ctx, torch::Tensor means, torch::Tensor scales,
                                        float globScale, torch::Tensor quats, torch::Tensor viewMat,
                                        torch::Tensor projMat, float fx, float fy, float cx, float cy,
                                        int imgHeight, int imgWidth, TileBounds tileBounds, float clipThresh) {
    const int n = (int)means.size(0);
    c10::cuda::CUDAGuard guard(means.device());
    torch::Tensor m = gsb::f32(means), s = gsb::f32(scales), q = gsb::f32(quats);
    torch::Tensor V = gsb::f32(viewMat), P = gsb::f32(projMat);
    // every output element is written by the kernel (culled Gaussians get zeros) -> empty, not zeros
    torch::Tensor cov3d = torch::empty({n, 6}, gsb::like(m, torch::kFloat32));
    torch::Tensor xys = torch::empty({n, 2}, gsb::like(m, torch::kFloat32));
    torch::Tensor depths = torch::empty({n}, gsb::like(m, torch::kFloat32));
    torch::Tensor radii = torch::empty({n}, gsb::like(m, torch::kInt32));
    torch::Tensor conics = torch::empty({n, 3}, gsb::like(m, torch::kFloat32));
    torch::Tensor numTilesHit = torch::empty({n}, gsb::like(m, torch::kInt32));
    gsb::check(gsb_project_forward(n, gsb::fp(m), gsb::fp(s), globScale, gsb::fp(q), gsb::fp(V), gsb::fp(P), fx,
                                   fy, cx, cy, imgHeight, imgWidth, std::get<0>(tileBounds),
                                   std::get<1>(tileBounds), clipThresh, gsb::fpw(cov3d), gsb::fpw(xys),
                                   gsb::fpw(depths), radii.data_ptr<int32_t>(), gsb::fpw(conics),
                                   numTilesHit.data_ptr<int32_t>(), gsb::stream()),
               "gsb_project_forward");

    ctx->saved_data["imgHeight"] = imgHeight;
    ctx->saved_data["imgWidth"] = imgWidth;
    ctx->saved_data["globScale"] = (double)globScale;
    ctx->saved_data["fx"] = (double)fx;
    ctx->saved_data["fy"] = (double)fy;
    ctx->saved_data["cx"] = (double)cx;
    ctx->saved_data["cy"] = (double)cy;
    ctx->save_for_backward({m, s, q, V, P, radii, conics});
    ctx->mark_non_differentiable({radii, numTilesHit});
    return {xys, depths, radii, conics, numTilesHit, cov3d};
}

tensor_list ProjectGaussians::backward(AutogradContext *ctx, tensor_list grad_outputs) {
    variable_list saved = ctx->get_saved_variables();
    torch::Tensor m = saved[0], s = saved[1], q = saved[2], V = saved[3], P = saved[4];
    torch::Tensor radii = saved[5], conics = saved[6];
    const int n = (int)m.size(0);
    c10::cuda::CUDAGuard guard(m.device());
    // cotangents of xys (0), depths (1), conics (3); undefined == zeros
    torch::Tensor v_xy = grad_outputs[0].defined() ? gsb::f32(grad_outputs[0])
                                                   : torch::zeros({n, 2}, gsb::like(m, torch::kFloat32));
    torch::Tensor v_depth = grad_outputs[1].defined() ? gsb::f32(grad_outputs[1]) : torch::Tensor();
    torch::Tensor v_conic = grad_outputs[3].defined() ? gsb::f32(grad_outputs[3])
                                                      : torch::zeros({n, 3}, gsb::like(m, torch::kFloat32));
    torch::Tensor v_mean = torch::empty({n, 3}, gsb::like(m, torch::kFloat32));
    torch::Tensor v_scale = torch::empty({n, 3}, gsb::like(m, torch::kFloat32));
    torch::Tensor v_quat = torch::empty({n, 4}, gsb::like(m, torch::kFloat32));
    gsb::check(gsb_project_backward(
                   n, gsb::fp(m), gsb::fp(s), (float)ctx->saved_data["globScale"].toDouble(), gsb::fp(q),
                   gsb::fp(V), gsb::fp(P), (float)ctx->saved_data["fx"].toDouble(),
                   (float)ctx->saved_data["fy"].toDouble(), (float)ctx->saved_data["cx"].toDouble(),
                   (float)ctx->saved_data["cy"].toDouble(), (int)ctx->saved_data["imgHeight"].toInt(),
                   (int)ctx->saved_data["imgWidth"].toInt(), nullptr, radii.data_ptr<int32_t>(), gsb::fp(conics),
                   gsb::fp(v_xy), v_depth.defined() ? gsb::fp(v_depth) : nullptr, gsb::fp(v_conic),
                   gsb::fpw(v_mean), gsb::fpw(v_scale), gsb::fpw(v_quat), gsb::stream()),
               "gsb_project_backward");
    torch::Tensor none;
    return {v_mean, v_scale, none, v_quat, none, none, none, none, none, none, none, none, none, none};
}

variable_list ProjectGaussiansCPU::apply(torch::Tensor, torch::Tensor, float, torch::Tensor, torch::Tensor,
                                         torch::Tensor, float, float, float, float, int, int, float) {
    TORCH_CHECK(false, "ProjectGaussiansCPU: the gsplat_b200 back end has no CPU path; link the reference's "
                       "rasterizer/gsplat-cpu for CPU execution");
    return {};
}
