// raster_common.cuh -- record stream layout and tile/pixel mapping shared by the blend kernels.
#pragma once
#include "gsb_common.cuh"

// One depth-sorted intersection = 48 bytes = three 16-byte quads, contiguous per tile, so that a
// tile's whole list is ONE contiguous range that the blend kernels pull with 1-D TMA bulk copies.
//   q0 = { x, y, opacity, bits(k) }   k = slot of this intersection in the UNSORTED (Gaussian-major)
//                                      order, i.e. cum_tiles_hit[g-1] + position inside g's tile bbox
//   q1 = { a/2, b, c/2, smax }        conic with the 1/2 of sigma folded in; smax = ln(255*opacity) + 1e-3
//                                      is a CONSERVATIVE bound: sigma > smax  =>  alpha < 1/255, so the
//                                      blend loops reject most (pixel, Gaussian) pairs before the exp
//   q2 = { r, g, b, bits(g) }         g = Gaussian id (debug / tooling only)
struct __align__(16) GsbRecord {
    float4 q0, q1, q2;
};
static_assert(sizeof(GsbRecord) == 48, "record must be 48 bytes");

// Blend-kernel geometry: one WARP owns one 16x16 tile; lane l owns column (l & 15) and the 8 rows
// 2*j + (l >> 4), j = 0..7.  No block-level synchronisation anywhere in the blend loops.
constexpr int RK_WARPS = 4;            // tiles per CTA
constexpr int RK_THREADS = RK_WARPS * 32;
constexpr int RK_PIX = 8;              // pixels per lane
constexpr int RK_CHUNK = 32;           // records per TMA bulk copy (1536 B)
constexpr int RK_STAGES = 4;           // ring depth per warp

// per-intersection gradient row written by the backward blend kernel (48 B, indexed by k):
//   { v_x, v_y, v_conic_a, v_conic_b, v_conic_c, v_r, v_g, v_b, v_opacity, 0, 0, 0 }
constexpr int GSB_GRAD_ROW_FLOATS = 12;
