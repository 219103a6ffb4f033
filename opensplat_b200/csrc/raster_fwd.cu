// raster_fwd.cu -- record packing + front-to-back alpha compositing (R1 of SURVEY.md section 8a).
//
// Replaces rasterize_forward (reference rasterizer/gsplat/forward.cu:256-378, host bindings.cu:338-410).
// Per-pixel semantics are the reference's: integer pixel coordinates, sigma = .5(a dx^2 + c dy^2) + b dx dy,
// alpha = min(0.999, opac * exp(-sigma)), skip if sigma < 0 or alpha < 1/255, stop BEFORE blending
// once T(1-alpha) <= 1e-4, out = colour + T * background, final_idx = sorted index of the last
// blended Gaussian (0 if none).
//
// Blackwell design (differs from the reference's 256-thread CTA with __syncthreads batching and
// per-pixel global colour gathers):
//  * a pack kernel turns the sorted id list into a contiguous 48-B record stream (raster_common.cuh),
//    so a tile's list is one contiguous byte range;
//  * one warp owns one tile and streams its range into a private 4-stage shared-memory ring with
//    1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on mbarriers -- no __syncthreads,
//    no per-thread gather instructions in the blend loop;
//  * each lane carries 8 pixels in registers (8 independent dependency chains), and every record
//    is read once per warp with three broadcast 128-bit shared loads.
#include "raster_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
pack_records_kernel(int m, const int *__restrict__ gaussian_ids_sorted,
                    const int *__restrict__ sorted_index, const float2 *__restrict__ xys,
                    const float *__restrict__ conics, const float *__restrict__ colors,
                    const float *__restrict__ opacities, GsbRecord *__restrict__ records) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int g = gaussian_ids_sorted[i];
    const int k = sorted_index ? sorted_index[i] : i;
    const float2 xy = __ldg(xys + g);
    GsbRecord r;
    const float opac = __ldg(opacities + g);
    // alpha >= 1/255  <=>  sigma <= ln(255 * opac); +1e-3 keeps the pre-test conservative w.r.t. the
    // rounding of sigma and of ex2.approx (the exact alpha test still runs inside the branch)
    const float smax = fmaxf(0.f, __logf(255.f * fmaxf(opac, 1e-30f)) + 1e-3f);  // >= +0: compared as unsigned bits
    r.q0 = make_float4(xy.x, xy.y, opac, __int_as_float(k));
    r.q1 = make_float4(0.5f * __ldg(conics + 3 * g), __ldg(conics + 3 * g + 1), 0.5f * __ldg(conics + 3 * g + 2),
                       smax);
    r.q2 = make_float4(__ldg(colors + 3 * g), __ldg(colors + 3 * g + 1), __ldg(colors + 3 * g + 2),
                       __int_as_float(g));
    float4 *dst = reinterpret_cast<float4 *>(records + i);
    stg_stream4(dst, r.q0);
    stg_stream4(dst + 1, r.q1);
    stg_stream4(dst + 2, r.q2);
}

struct __align__(128) WarpRing {
    GsbRecord rec[RK_STAGES][RK_CHUNK];
    uint64_t full[RK_STAGES];
};

__global__ void __launch_bounds__(RK_THREADS)
rasterize_forward_kernel(int img_h, int img_w, int tiles_x, int num_tiles,
                         const int2 *__restrict__ tile_bins, const GsbRecord *__restrict__ records,
                         const float *__restrict__ background, float *__restrict__ out_img,
                         float *__restrict__ final_Ts, int *__restrict__ final_idx) {
    __shared__ WarpRing rings[RK_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * RK_WARPS + warp;
    if (tile >= num_tiles) return;
    WarpRing &ring = rings[warp];
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < RK_STAGES; ++s) mbar_init(&ring.full[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int X = tx * GSB_TILE + (lane & 15);
    const int Y0 = ty * GSB_TILE + (lane >> 4);
    const float px = (float)X;
    const int2 range = tile_bins[tile];
    const int L = range.y - range.x;
    const int nchunks = (L + RK_CHUNK - 1) / RK_CHUNK;

    // T[j] > 0: transmittance of a live pixel.  A finished pixel keeps its final transmittance NEGATED
    // (sign bit == "done"), so the hot loop needs no separate done test: T*(1-alpha) <= 1e-4 sends it
    // to the (rare) terminate branch, which ignores pixels that are already negative.
    float T[RK_PIX], cr[RK_PIX], cg[RK_PIX], cb[RK_PIX], py[RK_PIX];
    int last[RK_PIX];
    unsigned done = 0;  // bit j: pixel j finished (or outside the image)
#pragma unroll
    for (int j = 0; j < RK_PIX; ++j) {
        T[j] = 1.f; cr[j] = cg[j] = cb[j] = 0.f; last[j] = 0;
        py[j] = (float)(Y0 + 2 * j);
        if (X >= img_w || Y0 + 2 * j >= img_h) { done |= 1u << j; T[j] = -1.f; }
    }

    auto issue = [&](int c) {
        if (lane == 0) {
            const int s = c % RK_STAGES;
            const int cnt = min(RK_CHUNK, L - c * RK_CHUNK);
            const uint32_t bytes = (uint32_t)cnt * (uint32_t)sizeof(GsbRecord);
            mbar_arrive_expect_tx(&ring.full[s], bytes);
            tma_load_1d(&ring.rec[s][0], records + range.x + c * RK_CHUNK, bytes, &ring.full[s]);
        }
    };

    const int pro = min(RK_STAGES, nchunks);
    for (int c = 0; c < pro; ++c) issue(c);
    int issued = pro;

    int c = 0;
    for (; c < nchunks; ++c) {
        const int s = c % RK_STAGES;
        mbar_wait(&ring.full[s], (uint32_t)(c / RK_STAGES) & 1u);
        const int cnt = min(RK_CHUNK, L - c * RK_CHUNK);
        const int idx0 = range.x + c * RK_CHUNK;
        for (int t = 0; t < cnt; ++t) {
            const float4 q0 = ring.rec[s][t].q0;
            const float4 q1 = ring.rec[s][t].q1;
            const float4 q2 = ring.rec[s][t].q2;
            const float dx = q0.x - px;
            const float adx2 = q1.x * dx * dx;   // (a/2) dx^2
            const float bdx = q1.y * dx;
#pragma unroll
            for (int j = 0; j < RK_PIX; ++j) {
                const float dy = q0.y - py[j];
                // sigma = (a/2)dx^2 + (c/2)dy^2 + b dx dy   (forward.cu:340-342)
                const float sigma = fmaf(dy, fmaf(q1.z, dy, bdx), adx2);
                if (__float_as_uint(sigma) <= __float_as_uint(q1.w)) {     // 0 <= sigma <= smax in ONE compare, no exp
                    const float alpha = fminf(0.999f, q0.z * ex2_approx(sigma * -1.4426950408889634f));
                    if (alpha >= (1.f / 255.f)) {
                        const float next_T = T[j] * (1.f - alpha);
                        if (next_T <= 1e-4f) {                          // terminate BEFORE blending
                            if (T[j] > 0.f) { T[j] = -T[j]; done |= 1u << j; }
                        } else {
                            const float vis = alpha * T[j];
                            cr[j] = fmaf(q2.x, vis, cr[j]);
                            cg[j] = fmaf(q2.y, vis, cg[j]);
                            cb[j] = fmaf(q2.z, vis, cb[j]);
                            T[j] = next_T;
                            last[j] = idx0 + t;
                        }
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, done == 0xffu)) { ++c; break; }  // whole tile saturated
        __syncwarp();
        if (issued < nchunks) { issue(issued); ++issued; }
    }
    // drain bulk copies that are still in flight before the ring's shared memory is released
    for (; c < issued; ++c) mbar_wait(&ring.full[c % RK_STAGES], (uint32_t)(c / RK_STAGES) & 1u);

    const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
#pragma unroll
    for (int j = 0; j < RK_PIX; ++j) {
        const int Y = Y0 + 2 * j;
        if (X < img_w && Y < img_h) {
            const size_t p = (size_t)Y * img_w + X;
            const float Tf = fabsf(T[j]);
            final_Ts[p] = Tf;
            final_idx[p] = last[j];
            out_img[3 * p] = cr[j] + Tf * bg0;
            out_img[3 * p + 1] = cg[j] + Tf * bg1;
            out_img[3 * p + 2] = cb[j] + Tf * bg2;
        }
    }
}

}  // namespace

extern "C" size_t gsb_raster_records_bytes(int m) {
    return gsb_align_up((size_t)(m > 0 ? m : 0) * sizeof(GsbRecord) + 256, 256);
}

extern "C" int gsb_rasterize_forward(int img_h, int img_w, int tiles_x, int tiles_y, int m,
                                     const int32_t *gaussian_ids_sorted, const int32_t *sorted_index,
                                     const int32_t *tile_bins, const float *xys, const float *conics,
                                     const float *colors, const float *opacities,
                                     const float *background, void *records, float *out_img,
                                     float *final_Ts, int32_t *final_idx, gsb_stream_t stream) {
    GSB_CHECK_ARG(img_h > 0 && img_w > 0 && m >= 0);
    GSB_CHECK_ARG(tiles_x == gsb_div_up(img_w, GSB_TILE) && tiles_y == gsb_div_up(img_h, GSB_TILE));
    GSB_CHECK_ARG(tile_bins && background && out_img && final_Ts && final_idx);
    cudaStream_t s = (cudaStream_t)stream;
    if (m > 0) {
        GSB_CHECK_ARG(gaussian_ids_sorted && xys && conics && colors && opacities && records);
        GSB_CHECK_ARG(((uintptr_t)records % 16) == 0 && ((uintptr_t)xys % 8) == 0);
        pack_records_kernel<<<gsb_div_up(m, 256), 256, 0, s>>>(
            m, gaussian_ids_sorted, sorted_index, reinterpret_cast<const float2 *>(xys), conics, colors,
            opacities, reinterpret_cast<GsbRecord *>(records));
    }
    const int num_tiles = tiles_x * tiles_y;
    rasterize_forward_kernel<<<gsb_div_up(num_tiles, RK_WARPS), RK_THREADS, 0, s>>>(
        img_h, img_w, tiles_x, num_tiles, reinterpret_cast<const int2 *>(tile_bins),
        reinterpret_cast<const GsbRecord *>(records), background, out_img, final_Ts, final_idx);
    GSB_LAUNCH_CHECK();
    return 0;
}
