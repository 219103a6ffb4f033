// export.cu -- on-device packing of the reference's two scene formats (Model::savePly model.cpp:505-558,
// Model::saveSplat model.cpp:560-594), so a checkpoint is one packed buffer copied D2H on a side stream instead of
// six .cpu() copies + a per-Gaussian ofstream::write loop on the training thread.
//
//   PLY body row (binary_little_endian float32):  x y z | nx ny nz (=0) | f_dc_0..2 | f_rest_* channel-major
//       ([K-1,3] -> [3,K-1], "Match Inria's version" model.cpp:525) | opacity | scale_0..2 | rot_0..3
//       = 14 + 3K floats (62 at degree 3).
//   .splat row (32 B):  mean 3 f32 | exp(scale) 3 f32 | rgb 3 u8 | alpha u8 | quat 4 u8, rows ordered by
//       descending (sum exp(scale)) / (1 + exp(-opacity))  (model.cpp:571-583).
//
// featuresDc / featuresRest are addressed with a row stride so that both the reference's two tensors
// ([n,3] and [n,K-1,3]) and this repo's merged coefficient block ([n,K,3]; dc = block, rest = block + 3) work.
#include "gsb_common.cuh"

namespace {

constexpr float SH_C0 = 0.28209479177387814f;  // spherical_harmonics.cpp sh2rgb

struct Crs {
    int keep;          // keepCrs (model.cpp:548-551,564-565)
    float scale, tx, ty, tz;
};

__global__ void __launch_bounds__(256)
pack_ply_kernel(long long total, int rf, int k, const float *__restrict__ means, const float *__restrict__ dc,
                int dc_stride, const float *__restrict__ rest, int rest_stride, const float *__restrict__ opac,
                const float *__restrict__ scales, const float *__restrict__ quats, Crs crs,
                float *__restrict__ out) {
    const int nrest = 3 * (k - 1);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / rf;
        int c = (int)(idx - i * rf);
        float v;
        if (c < 3) {
            v = means[3 * i + c];
            if (crs.keep) v = v / crs.scale + (c == 0 ? crs.tx : c == 1 ? crs.ty : crs.tz);
        } else if (c < 6) {
            v = 0.f;
        } else if (c < 9) {
            v = dc[i * dc_stride + (c - 6)];
        } else if (c < 9 + nrest) {
            const int r = c - 9, ch = r / (k - 1), b = r - ch * (k - 1);   // out[ch][b] = rest[b][ch]
            v = rest[i * rest_stride + 3 * b + ch];
        } else {
            c -= 9 + nrest;
            if (c == 0) v = opac[i];
            else if (c < 4) {
                v = scales[3 * i + (c - 1)];
                if (crs.keep) v = logf(expf(v) / crs.scale);
            } else v = quats[4 * i + (c - 4)];
        }
        out[idx] = v;
    }
}

// inverse of pack_ply_kernel (Model::loadPly, model.cpp:724-746): rows -> the six parameter tensors; keepCrs applies
// means = (means - translation) * scale, scales = log(scale * exp(scales)).
__global__ void __launch_bounds__(256)
unpack_ply_kernel(long long total, int rf, int k, const float *__restrict__ rows, Crs crs, float *__restrict__ means,
                  float *__restrict__ dc, int dc_stride, float *__restrict__ rest, int rest_stride,
                  float *__restrict__ opac, float *__restrict__ scales, float *__restrict__ quats) {
    const int nrest = 3 * (k - 1);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / rf;
        int c = (int)(idx - i * rf);
        float v = rows[idx];
        if (c < 3) {
            if (crs.keep) v = (v - (c == 0 ? crs.tx : c == 1 ? crs.ty : crs.tz)) * crs.scale;
            means[3 * i + c] = v;
        } else if (c < 6) {
            // normals: ignored
        } else if (c < 9) {
            dc[i * dc_stride + (c - 6)] = v;
        } else if (c < 9 + nrest) {
            const int r = c - 9, ch = r / (k - 1), b = r - ch * (k - 1);
            rest[i * rest_stride + 3 * b + ch] = v;
        } else {
            c -= 9 + nrest;
            if (c == 0) opac[i] = v;
            else if (c < 4) scales[3 * i + (c - 1)] = crs.keep ? logf(crs.scale * expf(v)) : v;
            else quats[4 * i + (c - 4)] = v;
        }
    }
}

__device__ __forceinline__ float splat_alpha_den(float o) { return 1.f + expf(-o); }

__global__ void __launch_bounds__(256)
splat_keys_kernel(int n, const float *__restrict__ scales, const float *__restrict__ opac, Crs crs,
                  int64_t *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float e0 = expf(scales[3 * i]), e1 = expf(scales[3 * i + 1]), e2 = expf(scales[3 * i + 2]);
    if (crs.keep) { e0 /= crs.scale; e1 /= crs.scale; e2 /= crs.scale; }
    const float key = ((e0 + e1) + e2) / splat_alpha_den(opac[i]);
    uint32_t b = __float_as_uint(key);
    b ^= (b >> 31) ? 0xffffffffu : 0x80000000u;   // monotone in the float order
    keys[i] = (int64_t)(uint32_t)~b;              // ascending integer order = descending float order
}

// torch's float -> uint8 conversion: truncate toward zero, keep the low byte
__device__ __forceinline__ uint32_t to_u8(float x) {
    if (!(fabsf(x) < 2147483648.f)) return 0u;
    return (uint32_t)((int)x) & 0xffu;
}

__global__ void __launch_bounds__(256)
pack_splat_kernel(int n, const int32_t *__restrict__ order, const float *__restrict__ means,
                  const float *__restrict__ scales, const float *__restrict__ dc, int dc_stride,
                  const float *__restrict__ opac, const float *__restrict__ quats, Crs crs,
                  uint4 *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const long long i = order ? order[j] : j;
    float m0 = means[3 * i], m1 = means[3 * i + 1], m2 = means[3 * i + 2];
    float e0 = expf(scales[3 * i]), e1 = expf(scales[3 * i + 1]), e2 = expf(scales[3 * i + 2]);
    if (crs.keep) {
        m0 = m0 / crs.scale + crs.tx; m1 = m1 / crs.scale + crs.ty; m2 = m2 / crs.scale + crs.tz;
        e0 /= crs.scale; e1 /= crs.scale; e2 /= crs.scale;
    }
    uint32_t rgba = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {   // sh2rgb clamps to [0,1] (spherical_harmonics.cpp:25-28)
        const float rgb = fminf(fmaxf(dc[i * dc_stride + c] * SH_C0 + 0.5f, 0.f), 1.f);
        rgba |= to_u8(rgb * 255.0f) << (8 * c);
    }
    const float a = fminf(fmaxf((1.0f / splat_alpha_den(opac[i])) * 255.0f, 0.f), 255.f);
    rgba |= to_u8(a) << 24;
    uint32_t q = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        q |= to_u8(fminf(fmaxf(quats[4 * i + c] * 128.0f + 128.0f, 0.f), 255.f)) << (8 * c);
    out[2 * (long long)j] = make_uint4(__float_as_uint(m0), __float_as_uint(m1), __float_as_uint(m2), __float_as_uint(e0));
    out[2 * (long long)j + 1] = make_uint4(__float_as_uint(e1), __float_as_uint(e2), rgba, q);
}

Crs make_crs(int keep, float scale, const float *t) {
    Crs c;
    c.keep = keep; c.scale = scale;
    c.tx = t ? t[0] : 0.f; c.ty = t ? t[1] : 0.f; c.tz = t ? t[2] : 0.f;
    return c;
}

}  // namespace

extern "C" int gsb_ply_row_floats(int sh_bases) { return sh_bases > 0 ? 14 + 3 * sh_bases : 0; }

extern "C" int gsb_pack_ply_rows(int n, int sh_bases, const float *means, const float *features_dc, int dc_stride,
                                 const float *features_rest, int rest_stride, const float *opacities,
                                 const float *scales, const float *quats, int keep_crs, float crs_scale,
                                 const float *crs_translation, float *out_rows, gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0 && sh_bases >= 1 && dc_stride >= 3);
    if (n == 0) return 0;
    GSB_CHECK_ARG(means && features_dc && opacities && scales && quats && out_rows);
    GSB_CHECK_ARG(sh_bases == 1 || (features_rest != nullptr && rest_stride >= 3 * (sh_bases - 1)));
    GSB_CHECK_ARG(!keep_crs || crs_scale != 0.f);
    const int rf = gsb_ply_row_floats(sh_bases);
    const long long total = (long long)n * rf;
    long long blocks = (total + 255) / 256;
    if (blocks > (1 << 20)) blocks = 1 << 20;
    pack_ply_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        total, rf, sh_bases, means, features_dc, dc_stride, features_rest, rest_stride, opacities, scales, quats,
        make_crs(keep_crs, crs_scale, crs_translation), out_rows);
    GSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gsb_unpack_ply_rows(int n, int sh_bases, const float *rows, int keep_crs, float crs_scale,
                                   const float *crs_translation, float *means, float *features_dc, int dc_stride,
                                   float *features_rest, int rest_stride, float *opacities, float *scales,
                                   float *quats, gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0 && sh_bases >= 1 && dc_stride >= 3);
    if (n == 0) return 0;
    GSB_CHECK_ARG(rows && means && features_dc && opacities && scales && quats);
    GSB_CHECK_ARG(sh_bases == 1 || (features_rest != nullptr && rest_stride >= 3 * (sh_bases - 1)));
    const int rf = gsb_ply_row_floats(sh_bases);
    const long long total = (long long)n * rf;
    long long blocks = (total + 255) / 256;
    if (blocks > (1 << 20)) blocks = 1 << 20;
    unpack_ply_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        total, rf, sh_bases, rows, make_crs(keep_crs, crs_scale, crs_translation), means, features_dc, dc_stride,
        features_rest, rest_stride, opacities, scales, quats);
    GSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gsb_splat_order_keys(int n, const float *scales, const float *opacities, int keep_crs,
                                    float crs_scale, int64_t *keys, gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0);
    if (n == 0) return 0;
    GSB_CHECK_ARG(scales && opacities && keys);
    GSB_CHECK_ARG(!keep_crs || crs_scale != 0.f);
    splat_keys_kernel<<<gsb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, scales, opacities, make_crs(keep_crs, crs_scale, nullptr), keys);
    GSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int gsb_pack_splat_rows(int n, const int32_t *order, const float *means, const float *scales,
                                   const float *features_dc, int dc_stride, const float *opacities,
                                   const float *quats, int keep_crs, float crs_scale, const float *crs_translation,
                                   void *out_rows, gsb_stream_t stream) {
    GSB_CHECK_ARG(n >= 0 && dc_stride >= 3);
    if (n == 0) return 0;
    GSB_CHECK_ARG(means && scales && features_dc && opacities && quats && out_rows);
    GSB_CHECK_ARG(((uintptr_t)out_rows % 16) == 0);
    GSB_CHECK_ARG(!keep_crs || crs_scale != 0.f);
    pack_splat_kernel<<<gsb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, order, means, scales, features_dc, dc_stride, opacities, quats,
        make_crs(keep_crs, crs_scale, crs_translation), static_cast<uint4 *>(out_rows));
    GSB_LAUNCH_CHECK();
    return 0;
}
