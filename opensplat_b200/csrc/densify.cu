// densify.cu -- topology edits of Model::afterTrain (reference model.cpp:339-470) as stream compaction.
//
// The reference refines the Gaussian set with ~60 boolean-mask index / cat / repeat ops (each mask index is a
// host-synchronising nonzero()), once per tensor and again per Adam state tensor.  Here one classification pass
// decides, per parent Gaussian, what survives -- itself, its two split children, its duplicate -- and three
// kernels (flag + per-block counts, single-block scan of the block counts, scatter) turn that into a row map
//      src_map[j] = parent | kind << 30        j in [0, new_n),  kind: 0 survivor, 1/2 split sample 0/1, 3 dup
// in exactly the reference's output order:  cat(originals, split sample 0, split sample 1, dups)[~culls].
// Every per-Gaussian tensor (parameters, Adam moments) is then rebuilt by one gather pass over that map.
// Only the counts travel to the host (one read-back, where the reference has ~10 .item() calls), because the
// caller has to size the new tensors and draw the 2*n_splits normal samples (torch::randn, model.cpp:359).
//
// Children inherit the parent's opacity, so the alpha cull of a child is the parent's; the "huge" cull of a split
// child is evaluated on its shrunk scales log(exp(s)/1.6) and a child's max2DSize is 0 (model.cpp:399-403).
#include "gsb_common.cuh"

namespace {

constexpr int DF_SPLIT = 1, DF_DUP = 2, DF_KEEP_SELF = 4, DF_KEEP_SPLIT = 8, DF_KEEP_DUP = 16;
constexpr int DB = 1024;       // Gaussians per block (one per thread)
constexpr int KIND_SHIFT = 30;

struct DensifyCfg {
    float max_dim, grad_thresh, size_thresh, split_screen, cull_alpha, cull_scale, cull_screen, size_fac;
    int check_split_screen, check_huge, check_cull_screen;
};

__device__ __forceinline__ int classify_one(int i, const float *__restrict__ scales,
                                            const float *__restrict__ opac, const float *__restrict__ gn,
                                            const float *__restrict__ vc, const float *__restrict__ m2d,
                                            const DensifyCfg &c) {
    const float s0 = scales[3 * i], s1 = scales[3 * i + 1], s2 = scales[3 * i + 2];
    const float mx = fmaxf(fmaxf(expf(s0), expf(s1)), expf(s2));
    const float avg = ((gn[i] / vc[i]) * 0.5f) * c.max_dim;  // model.cpp:343
    const bool high = avg > c.grad_thresh;
    const float m2 = m2d ? m2d[i] : 0.f;
    bool split = mx > c.size_thresh;
    if (c.check_split_screen) split = split || (m2 > c.split_screen);
    split = split && high;
    const bool dup = (mx <= c.size_thresh) && high;
    const float sig = 1.f / (1.f + expf(-opac[i]));
    const bool lowa = sig < c.cull_alpha;
    const bool huge_self = c.check_huge && (mx > c.cull_scale || (c.check_cull_screen && m2 > c.cull_screen));
    int f = 0;
    if (split) {
        f |= DF_SPLIT;
        const float c0 = expf(logf(expf(s0) / c.size_fac)), c1 = expf(logf(expf(s1) / c.size_fac)),
                    c2 = expf(logf(expf(s2) / c.size_fac));
        const float mxc = fmaxf(fmaxf(c0, c1), c2);
        const bool huge_child = c.check_huge && (mxc > c.cull_scale || (c.check_cull_screen && 0.f > c.cull_screen));
        if (!lowa && !huge_child) f |= DF_KEEP_SPLIT;
    }
    if (dup) {
        f |= DF_DUP;
        const bool huge_child = c.check_huge && (mx > c.cull_scale || (c.check_cull_screen && 0.f > c.cull_screen));
        if (!lowa && !huge_child) f |= DF_KEEP_DUP;
    }
    if (!lowa && !split && !huge_self) f |= DF_KEEP_SELF;
    return f;
}

// four 16-bit counters packed in one 64-bit word: split | keep_self | keep_split | keep_dup
__device__ __forceinline__ unsigned long long pack_counts(int f) {
    return (unsigned long long)((f & DF_SPLIT) != 0) | ((unsigned long long)((f & DF_KEEP_SELF) != 0) << 16) |
           ((unsigned long long)((f & DF_KEEP_SPLIT) != 0) << 32) | ((unsigned long long)((f & DF_KEEP_DUP) != 0) << 48);
}

// exclusive block scan of the packed counters (1024 threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ unsigned long long block_scan_packed(unsigned long long v, unsigned long long *total,
                                                                unsigned long long *warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_sums[lane];
        unsigned long long winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += o;
        }
        warp_sums[lane] = winc - w;              // exclusive
        if (lane == 31) warp_sums[32] = winc;    // block total
    }
    __syncthreads();
    *total = warp_sums[32];
    return warp_sums[warp] + inc - v;
}

__global__ void __launch_bounds__(DB)
densify_flag_kernel(int n, const float *__restrict__ scales, const float *__restrict__ opac,
                    const float *__restrict__ gn, const float *__restrict__ vc, const float *__restrict__ m2d,
                    DensifyCfg cfg, uint8_t *__restrict__ flags, unsigned long long *__restrict__ block_counts) {
    __shared__ unsigned long long warp_sums[33];
    const int i = blockIdx.x * DB + threadIdx.x;
    int f = 0;
    if (i < n) {
        f = classify_one(i, scales, opac, gn, vc, m2d, cfg);
        flags[i] = (uint8_t)f;
    }
    unsigned long long total;
    block_scan_packed(pack_counts(f), &total, warp_sums);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// one CTA: exclusive scan of the per-block packed counts, 4 x 32-bit running sums; writes counts[8]
__global__ void __launch_bounds__(DB)
densify_scan_kernel(int nblocks, const unsigned long long *__restrict__ block_counts,
                    int4 *__restrict__ block_offsets, int32_t *__restrict__ counts) {
    __shared__ int4 warp_tot[33];
    int4 carry = make_int4(0, 0, 0, 0);
    for (int base = 0; base < nblocks; base += DB) {
        const int b = base + threadIdx.x;
        int4 v = make_int4(0, 0, 0, 0);
        if (b < nblocks) {
            const unsigned long long c = block_counts[b];
            v = make_int4((int)(c & 0xffff), (int)((c >> 16) & 0xffff), (int)((c >> 32) & 0xffff), (int)(c >> 48));
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int4 inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int ox = __shfl_up_sync(0xffffffffu, inc.x, d), oy = __shfl_up_sync(0xffffffffu, inc.y, d),
                      oz = __shfl_up_sync(0xffffffffu, inc.z, d), ow = __shfl_up_sync(0xffffffffu, inc.w, d);
            if (lane >= d) { inc.x += ox; inc.y += oy; inc.z += oz; inc.w += ow; }
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int4 w = warp_tot[lane];
            int4 wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int ox = __shfl_up_sync(0xffffffffu, wi.x, d), oy = __shfl_up_sync(0xffffffffu, wi.y, d),
                          oz = __shfl_up_sync(0xffffffffu, wi.z, d), ow = __shfl_up_sync(0xffffffffu, wi.w, d);
                if (lane >= d) { wi.x += ox; wi.y += oy; wi.z += oz; wi.w += ow; }
            }
            warp_tot[lane] = make_int4(wi.x - w.x, wi.y - w.y, wi.z - w.z, wi.w - w.w);
            if (lane == 31) warp_tot[32] = wi;
        }
        __syncthreads();
        const int4 wo = warp_tot[warp];
        if (b < nblocks)
            block_offsets[b] = make_int4(carry.x + wo.x + inc.x - v.x, carry.y + wo.y + inc.y - v.y,
                                         carry.z + wo.z + inc.z - v.z, carry.w + wo.w + inc.w - v.w);
        const int4 t = warp_tot[32];
        carry.x += t.x; carry.y += t.y; carry.z += t.z; carry.w += t.w;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // carry = { n_splits, kept originals, kept split parents, kept dups }
        counts[0] = carry.x;                              // n_splits  (rows of samples = 2 * n_splits)
        counts[1] = carry.y;                              // survivors among the originals
        counts[2] = carry.z;                              // split parents whose children survive
        counts[3] = carry.w;                              // surviving duplicates
        counts[4] = carry.y + 2 * carry.z + carry.w;      // new_n
        counts[5] = 0; counts[6] = 0; counts[7] = 0;      // [5] = n_dups, filled by the scatter kernel's atomics
    }
}

__global__ void __launch_bounds__(DB)
densify_scatter_kernel(int n, const uint8_t *__restrict__ flags, const int4 *__restrict__ block_offsets,
                       int32_t *__restrict__ counts, int32_t *__restrict__ src_map,
                       int32_t *__restrict__ split_rank) {
    __shared__ unsigned long long warp_sums[33];
    const int i = blockIdx.x * DB + threadIdx.x;
    const int f = i < n ? flags[i] : 0;
    unsigned long long total;
    const unsigned long long ex = block_scan_packed(pack_counts(f), &total, warp_sums);
    const int4 bo = block_offsets[blockIdx.x];
    const int kself = counts[1], ksplit = counts[2];
    if (i < n) {
        const int r_split = bo.x + (int)(ex & 0xffff), r_self = bo.y + (int)((ex >> 16) & 0xffff),
                  r_ks = bo.z + (int)((ex >> 32) & 0xffff), r_kd = bo.w + (int)(ex >> 48);
        split_rank[i] = (f & DF_SPLIT) ? r_split : -1;
        if (f & DF_KEEP_SELF) src_map[r_self] = i;
        if (f & DF_KEEP_SPLIT) {
            src_map[kself + r_ks] = i | (1 << KIND_SHIFT);
            src_map[kself + ksplit + r_ks] = i | (2 << KIND_SHIFT);
        }
        if (f & DF_KEEP_DUP) src_map[kself + 2 * ksplit + r_kd] = i | (3 << KIND_SHIFT);
    }
    // n_dups (reported only; "Added N gaussians" of model.cpp:432 = 2 * n_splits + n_dups)
    const unsigned nd = __syncthreads_count((f & DF_DUP) != 0);
    if (threadIdx.x == 0 && nd) atomicAdd(&counts[5], (int)nd);
}

// means / scales of the new set (model.cpp:359-373): split children get mean + R(q/|q|) (exp(s) * sample) and
// log(exp(s)/size_fac); survivors and duplicates are copies.
__global__ void __launch_bounds__(256)
densify_means_scales_kernel(int new_n, int n_splits, const int32_t *__restrict__ src_map,
                            const int32_t *__restrict__ split_rank, const float *__restrict__ samples,
                            const float *__restrict__ means, const float *__restrict__ scales,
                            const float *__restrict__ quats, float size_fac, float *__restrict__ new_means,
                            float *__restrict__ new_scales) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= new_n) return;
    const int e = src_map[j];
    const int p = e & ((1 << KIND_SHIFT) - 1), kind = (unsigned)e >> KIND_SHIFT;
    float m0 = means[3 * p], m1 = means[3 * p + 1], m2 = means[3 * p + 2];
    float s0 = scales[3 * p], s1 = scales[3 * p + 1], s2 = scales[3 * p + 2];
    if (kind == 1 || kind == 2) {
        const int row = (kind - 1) * n_splits + split_rank[p];
        const float e0 = expf(s0), e1 = expf(s1), e2 = expf(s2);
        const float v0 = e0 * samples[3 * row], v1 = e1 * samples[3 * row + 1], v2 = e2 * samples[3 * row + 2];
        float w = quats[4 * p], x = quats[4 * p + 1], y = quats[4 * p + 2], z = quats[4 * p + 3];
        const float nrm = sqrtf(w * w + x * x + y * y + z * z);          // linalg_vector_norm, model.cpp:361
        w /= nrm; x /= nrm; y /= nrm; z /= nrm;
        const float n2 = fmaxf(sqrtf(w * w + x * x + y * y + z * z), 1e-12f);  // F.normalize in quatToRotMat
        w /= n2; x /= n2; y /= n2; z /= n2;
        const float r00 = 1.f - 2.f * (y * y + z * z), r01 = 2.f * (x * y - w * z), r02 = 2.f * (x * z + w * y);
        const float r10 = 2.f * (x * y + w * z), r11 = 1.f - 2.f * (x * x + z * z), r12 = 2.f * (y * z - w * x);
        const float r20 = 2.f * (x * z - w * y), r21 = 2.f * (y * z + w * x), r22 = 1.f - 2.f * (x * x + y * y);
        m0 += r00 * v0 + r01 * v1 + r02 * v2;
        m1 += r10 * v0 + r11 * v1 + r12 * v2;
        m2 += r20 * v0 + r21 * v1 + r22 * v2;
        s0 = logf(e0 / size_fac); s1 = logf(e1 / size_fac); s2 = logf(e2 / size_fac);
    }
    new_means[3 * j] = m0; new_means[3 * j + 1] = m1; new_means[3 * j + 2] = m2;
    new_scales[3 * j] = s0; new_scales[3 * j + 1] = s1; new_scales[3 * j + 2] = s2;
}

// generic row gather: dst[j, :] = src[parent(j), :], or 0 for children when zero_children (Adam moments of
// new Gaussians, model.cpp:253-279).  One thread per element; RF > 0 fixes the row length at compile time.
template <int RF>
__global__ void __launch_bounds__(256)
gather_rows_kernel(long long total, int row_floats, const int32_t *__restrict__ src_map,
                   const float *__restrict__ src, float *__restrict__ dst, int zero_children) {
    const int rf = RF > 0 ? RF : row_floats;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long j = idx / rf;
        const int c = (int)(idx - j * rf);
        const int e = src_map[j];
        const int p = e & ((1 << KIND_SHIFT) - 1);
        const bool child = ((unsigned)e >> KIND_SHIFT) != 0;
        dst[idx] = (child && zero_children) ? 0.f : src[(long long)p * rf + c];
    }
}

__global__ void __launch_bounds__(256)
reset_opacity_kernel(int n, float max_logit, float *__restrict__ opac, float *__restrict__ m, float *__restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    opac[i] = fminf(opac[i], max_logit);
    if (m) m[i] = 0.f;
    if (v) v[i] = 0.f;
}

}  // namespace

extern "C" size_t gsb_densify_workspace_bytes(int n) {
    if (n < 0) return 0;
    const size_t nb = (size_t)gsb_div_up(n > 0 ? n : 1, DB);
    return gsb_align_up((size_t)(n > 0 ? n : 1), 256) + gsb_align_up(nb * sizeof(unsigned long long), 256) +
           gsb_align_up(nb * sizeof(int4), 256);
}

extern "C" int gsb_densify_classify(int n, const float *scales, const float *opacities, const float *xys_grad_norm,
                                    const float *vis_counts, const float *max_2d_size, float max_dim,
                                    float densify_grad_thresh, float densify_size_thresh, int check_split_screen,
                                    float split_screen_size, float cull_alpha_thresh, int check_huge,
                                    float cull_scale_thresh, int check_cull_screen, float cull_screen_size,
                                    float size_fac, void *workspace, size_t workspace_bytes, int32_t *src_map,
                                    int32_t *split_rank, int32_t *counts, gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0 && n < (1 << (KIND_SHIFT - 1)));  // parent in 30 bits, new_n <= 3n in an int32
    GSB_CHECK_ARG(counts != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        GSB_CUDA(cudaMemsetAsync(counts, 0, 8 * sizeof(int32_t), st));
        return 0;
    }
    GSB_CHECK_ARG(scales && opacities && xys_grad_norm && vis_counts && workspace && src_map && split_rank);
    GSB_CHECK_ARG(workspace_bytes >= gsb_densify_workspace_bytes(n) && ((uintptr_t)workspace % 16) == 0);
    GSB_CHECK_ARG(size_fac > 0.f);
    const int nb = gsb_div_up(n, DB);
    uint8_t *flags = static_cast<uint8_t *>(workspace);
    unsigned long long *block_counts =
        reinterpret_cast<unsigned long long *>(flags + gsb_align_up((size_t)n, 256));
    int4 *block_offsets = reinterpret_cast<int4 *>(reinterpret_cast<uint8_t *>(block_counts) +
                                                   gsb_align_up((size_t)nb * sizeof(unsigned long long), 256));
    DensifyCfg cfg;
    cfg.max_dim = max_dim; cfg.grad_thresh = densify_grad_thresh; cfg.size_thresh = densify_size_thresh;
    cfg.split_screen = split_screen_size; cfg.cull_alpha = cull_alpha_thresh; cfg.cull_scale = cull_scale_thresh;
    cfg.cull_screen = cull_screen_size; cfg.size_fac = size_fac; cfg.check_split_screen = check_split_screen;
    cfg.check_huge = check_huge; cfg.check_cull_screen = check_cull_screen;
    densify_flag_kernel<<<nb, DB, 0, st>>>(n, scales, opacities, xys_grad_norm, vis_counts, max_2d_size, cfg, flags,
                                           block_counts);
    GSB_LAUNCH_CHECK();
    densify_scan_kernel<<<1, DB, 0, st>>>(nb, block_counts, block_offsets, counts);
    GSB_LAUNCH_CHECK();
    densify_scatter_kernel<<<nb, DB, 0, st>>>(n, flags, block_offsets, counts, src_map, split_rank);
    GSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gsb_densify_means_scales(int new_n, int n_splits, const int32_t *src_map, const int32_t *split_rank,
                                        const float *samples, const float *means, const float *scales,
                                        const float *quats, float size_fac, float *new_means, float *new_scales,
                                        gsb_stream_t stream) {
    GSB_CHECK_ARG(new_n >= 0 && n_splits >= 0 && size_fac > 0.f);
    if (new_n == 0) return 0;
    GSB_CHECK_ARG(src_map && split_rank && means && scales && quats && new_means && new_scales);
    GSB_CHECK_ARG(n_splits == 0 || samples != nullptr);
    densify_means_scales_kernel<<<gsb_div_up(new_n, 256), 256, 0, (cudaStream_t)stream>>>(
        new_n, n_splits, src_map, split_rank, samples, means, scales, quats, size_fac, new_means, new_scales);
    GSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gsb_densify_gather_rows(int new_n, int row_floats, const int32_t *src_map, const float *src,
                                       float *dst, int zero_children, gsb_stream_t stream) {
    GSB_CHECK_ARG(new_n >= 0 && row_floats > 0);
    if (new_n == 0) return 0;
    GSB_CHECK_ARG(src_map && src && dst);
    const long long total = (long long)new_n * row_floats;
    long long blocks = (total + 255) / 256;
    if (blocks > (1 << 20)) blocks = 1 << 20;
    cudaStream_t st = (cudaStream_t)stream;
#define GSB_GATHER(RF) gather_rows_kernel<RF><<<(int)blocks, 256, 0, st>>>(total, row_floats, src_map, src, dst, zero_children)
    switch (row_floats) {
        case 1: GSB_GATHER(1); break;
        case 3: GSB_GATHER(3); break;
        case 4: GSB_GATHER(4); break;
        case 45: GSB_GATHER(45); break;
        case 48: GSB_GATHER(48); break;
        default: GSB_GATHER(0); break;
    }
#undef GSB_GATHER
    GSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gsb_reset_opacity(int n, float max_logit, float *opacities, float *exp_avg, float *exp_avg_sq,
                                 gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0);
    if (n == 0) return 0;
    GSB_CHECK_ARG(opacities != nullptr);
    reset_opacity_kernel<<<gsb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, max_logit, opacities, exp_avg,
                                                                              exp_avg_sq);
    GSB_LAUNCH_CHECK();
    return 0;
}
