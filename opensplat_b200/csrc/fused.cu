// fused.cu -- small streaming kernels around the hot path ("next" rows of SURVEY.md section 8f):
//   gsb_mse_loss_grad : loss = mean((img - target)^2) and v_img = 2 (img - target) / count in one pass
//                       (simple_trainer.cpp:199-201 does this with torch::nn::MSELoss + autograd)
//   gsb_adam_step     : one fused Adam update over a flat parameter buffer
//                       (simple_trainer.cpp:146,202 torch::optim::Adam; model.cpp:236-243 runs six of them)
// Both are HBM-bound elementwise passes: 128-bit accesses, grid = multiple of the SM count.
#include "gsb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
mse_loss_grad_kernel(long long n4, long long n, const float *__restrict__ img,
                     const float *__restrict__ target, float *__restrict__ v_img,
                     float *__restrict__ loss_out, float inv_count) {
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 a = reinterpret_cast<const float4 *>(img)[i];
        const float4 b = reinterpret_cast<const float4 *>(target)[i];
        const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        const float s = 2.f * inv_count;
        reinterpret_cast<float4 *>(v_img)[i] = make_float4(s * d.x, s * d.y, s * d.z, s * d.w);
    }
    // tail (n not a multiple of 4)
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = img[i] - target[i];
        acc += d * d;
        v_img[i] = 2.f * inv_count * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = sm[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
        if (threadIdx.x == 0) atomicAdd(loss_out, v * inv_count);
    }
}

__global__ void __launch_bounds__(256)
adam_kernel(long long n4, long long n, float *__restrict__ p, const float *__restrict__ g,
            float *__restrict__ m, float *__restrict__ v, float lr, float b1, float b2, float eps,
            float inv_bc1, float inv_sqrt_bc2) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    auto upd = [&](float &pp, float gg, float &mm, float &vv) {
        mm = b1 * mm + (1.f - b1) * gg;
        vv = b2 * vv + (1.f - b2) * gg * gg;
        // torch.optim.Adam: p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
        pp -= lr * inv_bc1 * __fdividef(mm, sqrtf(vv) * inv_sqrt_bc2 + eps);
    };
    // two float4 per thread per trip: 8 independent 128-bit loads in flight before the first use
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) {
        const long long j = i + stride;
        float4 P0 = reinterpret_cast<float4 *>(p)[i], P1 = reinterpret_cast<float4 *>(p)[j];
        const float4 G0 = ldg_stream4(reinterpret_cast<const float4 *>(g) + i);
        const float4 G1 = ldg_stream4(reinterpret_cast<const float4 *>(g) + j);
        float4 M0 = reinterpret_cast<float4 *>(m)[i], M1 = reinterpret_cast<float4 *>(m)[j];
        float4 V0 = reinterpret_cast<float4 *>(v)[i], V1 = reinterpret_cast<float4 *>(v)[j];
        upd(P0.x, G0.x, M0.x, V0.x); upd(P0.y, G0.y, M0.y, V0.y); upd(P0.z, G0.z, M0.z, V0.z); upd(P0.w, G0.w, M0.w, V0.w);
        upd(P1.x, G1.x, M1.x, V1.x); upd(P1.y, G1.y, M1.y, V1.y); upd(P1.z, G1.z, M1.z, V1.z); upd(P1.w, G1.w, M1.w, V1.w);
        reinterpret_cast<float4 *>(p)[i] = P0; reinterpret_cast<float4 *>(p)[j] = P1;
        reinterpret_cast<float4 *>(m)[i] = M0; reinterpret_cast<float4 *>(m)[j] = M1;
        reinterpret_cast<float4 *>(v)[i] = V0; reinterpret_cast<float4 *>(v)[j] = V1;
    }
    for (; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4 *>(p)[i];
        const float4 G = ldg_stream4(reinterpret_cast<const float4 *>(g) + i);
        float4 M = reinterpret_cast<float4 *>(m)[i];
        float4 V = reinterpret_cast<float4 *>(v)[i];
        upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
        reinterpret_cast<float4 *>(p)[i] = P;
        reinterpret_cast<float4 *>(m)[i] = M;
        reinterpret_cast<float4 *>(v)[i] = V;
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        upd(p[i], g[i], m[i], v[i]);
}

// ---- parameter activations of Model::forward (model.cpp:114,148-150,176-177,200), one pass each way ----
__global__ void __launch_bounds__(256)
activate_forward_kernel(int n, const float *__restrict__ means, const float *__restrict__ log_scales,
                        const float *__restrict__ raw_quats, const float *__restrict__ opacity_logits,
                        const float *__restrict__ cam_pos, float *__restrict__ scales,
                        float4 *__restrict__ quats, float *__restrict__ opacities, float *__restrict__ viewdirs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    scales[3 * i] = expf(log_scales[3 * i]);
    scales[3 * i + 1] = expf(log_scales[3 * i + 1]);
    scales[3 * i + 2] = expf(log_scales[3 * i + 2]);
    const float4 q = reinterpret_cast<const float4 *>(raw_quats)[i];
    const float inv = 1.f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    quats[i] = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
    opacities[i] = 1.f / (1.f + expf(-opacity_logits[i]));
    const float dx = means[3 * i] - __ldg(cam_pos), dy = means[3 * i + 1] - __ldg(cam_pos + 1),
                dz = means[3 * i + 2] - __ldg(cam_pos + 2);
    const float dn = 1.f / sqrtf(dx * dx + dy * dy + dz * dz);
    viewdirs[3 * i] = dx * dn; viewdirs[3 * i + 1] = dy * dn; viewdirs[3 * i + 2] = dz * dn;
}

__global__ void __launch_bounds__(256)
activate_backward_kernel(int n, const float *__restrict__ scales, const float *__restrict__ raw_quats,
                         const float *__restrict__ opacities, const float *__restrict__ v_scales,
                         const float4 *__restrict__ v_quats, const float *__restrict__ v_opacities,
                         float *__restrict__ v_log_scales, float4 *__restrict__ v_raw_quats,
                         float *__restrict__ v_opacity_logits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // d exp(s) = exp(s)
    v_log_scales[3 * i] = v_scales[3 * i] * scales[3 * i];
    v_log_scales[3 * i + 1] = v_scales[3 * i + 1] * scales[3 * i + 1];
    v_log_scales[3 * i + 2] = v_scales[3 * i + 2] * scales[3 * i + 2];
    // d (q / |q|) : (I - q^ q^T) / |q|
    const float4 q = reinterpret_cast<const float4 *>(raw_quats)[i];
    const float inv = 1.f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    const float4 h = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv), g = v_quats[i];
    const float dot = h.x * g.x + h.y * g.y + h.z * g.z + h.w * g.w;
    v_raw_quats[i] = make_float4((g.x - h.x * dot) * inv, (g.y - h.y * dot) * inv, (g.z - h.z * dot) * inv,
                                 (g.w - h.w * dot) * inv);
    // d sigmoid = o (1 - o)
    const float o = opacities[i];
    v_opacity_logits[i] = v_opacities[i] * o * (1.f - o);
}

// ---- densification statistics of Model::afterTrain (model.cpp:317-337), one pass, no boolean-mask indexing ----
__global__ void __launch_bounds__(256)
densify_stats_kernel(int n, const float2 *__restrict__ v_xy, const int *__restrict__ radii, float max_hw,
                     float *__restrict__ xys_grad_norm, float *__restrict__ vis_counts,
                     float *__restrict__ max_2d_size) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = radii[i];
    if (r <= 0) return;  // visibleMask = radii > 0
    const float2 g = v_xy[i];
    xys_grad_norm[i] += sqrtf(g.x * g.x + g.y * g.y);
    vis_counts[i] += 1.f;
    max_2d_size[i] = fmaxf(max_2d_size[i], (float)r / max_hw);
}

// first step after a refinement (model.cpp:321-323,328-330): xysGradNorm = |v_xy| and visCounts = 1 for EVERY
// Gaussian (visible or not), max2DSize = 0 then the visible update
__global__ void __launch_bounds__(256)
densify_stats_init_kernel(int n, const float2 *__restrict__ v_xy, const int *__restrict__ radii, float max_hw,
                          float *__restrict__ xys_grad_norm, float *__restrict__ vis_counts,
                          float *__restrict__ max_2d_size) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = radii[i];
    const float2 g = v_xy[i];
    xys_grad_norm[i] = sqrtf(g.x * g.x + g.y * g.y);
    vis_counts[i] = 1.f;
    max_2d_size[i] = r > 0 ? fmaxf(0.f, (float)r / max_hw) : 0.f;
}

int sm_count() {
    static thread_local int cached = 0;
    if (!cached) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        if (cached <= 0) cached = 148;
    }
    return cached;
}

}  // namespace

// loss_out (device float) is zeroed here (stream-ordered) and then accumulated into by the kernel.
extern "C" int gsb_mse_loss_grad(long long n, const float *img, const float *target, float *v_img,
                                 float *loss_out, float inv_count, gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0);
    if (n == 0) return 0;
    GSB_CHECK_ARG(img && target && v_img && loss_out);
    const bool vec = (((uintptr_t)img | (uintptr_t)target | (uintptr_t)v_img) % 16) == 0;
    const long long n4 = vec ? n / 4 : 0;
    GSB_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), (cudaStream_t)stream));
    mse_loss_grad_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(n4, n, img, target, v_img, loss_out,
                                                                        inv_count);
    GSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gsb_adam_step(long long n, float *param, const float *grad, float *exp_avg,
                             float *exp_avg_sq, float lr, float beta1, float beta2, float eps,
                             float bias_correction1, float bias_correction2, gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0 && bias_correction1 > 0.f && bias_correction2 > 0.f);
    if (n == 0) return 0;
    GSB_CHECK_ARG(param && grad && exp_avg && exp_avg_sq);
    const bool vec = (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16) == 0;
    const long long n4 = vec ? n / 4 : 0;
    adam_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(n4, n, param, grad, exp_avg, exp_avg_sq, lr, beta1,
                                                               beta2, eps, 1.f / bias_correction1,
                                                               1.f / sqrtf(bias_correction2));
    GSB_LAUNCH_CHECK();
    return 0;
}

// Parameter activations of Model::forward fused into one pass: scales = exp(log_scales)
// (model.cpp:148), quats = raw / |raw| (:150), opacities = sigmoid(logits) (:200), viewdirs =
// normalize(means - cam_pos) (:176-177; detached, no gradient).  cam_pos is a device float[3].
extern "C" int gsb_activate_forward(int n, const float *means, const float *log_scales, const float *raw_quats,
                                    const float *opacity_logits, const float *cam_pos, float *scales, float *quats,
                                    float *opacities, float *viewdirs, gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0);
    if (n == 0) return 0;
    GSB_CHECK_ARG(means && log_scales && raw_quats && opacity_logits && cam_pos && scales && quats && opacities &&
                  viewdirs);
    GSB_CHECK_ARG(((uintptr_t)raw_quats % 16) == 0 && ((uintptr_t)quats % 16) == 0);
    activate_forward_kernel<<<gsb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, means, log_scales, raw_quats, opacity_logits, cam_pos, scales, reinterpret_cast<float4 *>(quats), opacities,
        viewdirs);
    GSB_LAUNCH_CHECK();
    return 0;
}

// VJP of gsb_activate_forward: takes the forward OUTPUTS scales / opacities and the raw quaternions.
extern "C" int gsb_activate_backward(int n, const float *scales, const float *raw_quats, const float *opacities,
                                     const float *v_scales, const float *v_quats, const float *v_opacities,
                                     float *v_log_scales, float *v_raw_quats, float *v_opacity_logits,
                                     gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0);
    if (n == 0) return 0;
    GSB_CHECK_ARG(scales && raw_quats && opacities && v_scales && v_quats && v_opacities && v_log_scales &&
                  v_raw_quats && v_opacity_logits);
    GSB_CHECK_ARG(((uintptr_t)raw_quats % 16) == 0 && ((uintptr_t)v_quats % 16) == 0 &&
                  ((uintptr_t)v_raw_quats % 16) == 0);
    activate_backward_kernel<<<gsb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, scales, raw_quats, opacities, v_scales, reinterpret_cast<const float4 *>(v_quats), v_opacities, v_log_scales,
        reinterpret_cast<float4 *>(v_raw_quats), v_opacity_logits);
    GSB_LAUNCH_CHECK();
    return 0;
}

// Densification statistics (Model::afterTrain model.cpp:317-337): for visible Gaussians
// (radii > 0): xys_grad_norm += |v_xy|, vis_counts += 1, max_2d_size = max(max_2d_size, radii / max(H, W)).
// The reference does this with boolean-mask index / index_put (each a host-synchronising nonzero()).
extern "C" int gsb_densify_stats_update(int n, const float *v_xy, const int32_t *radii, int img_h, int img_w,
                                        float *xys_grad_norm, float *vis_counts, float *max_2d_size,
                                        gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0 && img_h > 0 && img_w > 0);
    if (n == 0) return 0;
    GSB_CHECK_ARG(v_xy && radii && xys_grad_norm && vis_counts && max_2d_size && ((uintptr_t)v_xy % 8) == 0);
    const float max_hw = (float)(img_h > img_w ? img_h : img_w);
    densify_stats_kernel<<<gsb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, reinterpret_cast<const float2 *>(v_xy), radii, max_hw, xys_grad_norm, vis_counts, max_2d_size);
    GSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gsb_densify_stats_init(int n, const float *v_xy, const int32_t *radii, int img_h, int img_w,
                                      float *xys_grad_norm, float *vis_counts, float *max_2d_size,
                                      gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0 && img_h > 0 && img_w > 0);
    if (n == 0) return 0;
    GSB_CHECK_ARG(v_xy && radii && xys_grad_norm && vis_counts && max_2d_size && ((uintptr_t)v_xy % 8) == 0);
    const float max_hw = (float)(img_h > img_w ? img_h : img_w);
    densify_stats_init_kernel<<<gsb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, reinterpret_cast<const float2 *>(v_xy), radii, max_hw, xys_grad_norm, vis_counts, max_2d_size);
    GSB_LAUNCH_CHECK();
    return 0;
}
