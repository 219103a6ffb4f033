// rasterize_gaussians.cpp -- RasterizeGaussians / binAndSortGaussians over the C ABI.
// Replaces the reference's rasterize_gaussians.cpp:6-140 + bindings.cu:279-632 and the ATen calls in
// between (torch::cumsum :62, .item<int>() :63, torch::sort :25, torch::gather :32).
#include "rasterize_gaussians.hpp"
#include "gsplat.hpp"
#include "gsb_torch.hpp"

namespace {

struct Binned {
    torch::Tensor isectIds, gaussianIds, isectIdsSorted, sortedIndex, gaussianIdsSorted, tileBins;
};

Binned bin_and_sort(int numPoints, int numIntersects, const torch::Tensor &xys, const torch::Tensor &depths,
                    const torch::Tensor &radii, const torch::Tensor &cumTilesHit, TileBounds tileBounds) {
    const int tilesX = std::get<0>(tileBounds), tilesY = std::get<1>(tileBounds);
    const int numTiles = tilesX * tilesY, m = numIntersects;
    torch::Tensor x = gsb::f32(xys), d = gsb::f32(depths), r = gsb::i32(radii), cum = gsb::i32(cumTilesHit);
    Binned b;
    b.isectIds = torch::empty({m}, gsb::like(x, torch::kInt64));
    b.gaussianIds = torch::empty({m}, gsb::like(x, torch::kInt32));
    b.isectIdsSorted = torch::empty({m}, gsb::like(x, torch::kInt64));
    b.sortedIndex = torch::empty({m}, gsb::like(x, torch::kInt32));
    b.gaussianIdsSorted = torch::empty({m}, gsb::like(x, torch::kInt32));
    b.tileBins = torch::empty({numTiles, 2}, gsb::like(x, torch::kInt32));
    gsb::check(gsb_map_gaussian_to_intersects(numPoints, m, gsb::fp(x), gsb::fp(d), r.data_ptr<int32_t>(),
                                              cum.data_ptr<int32_t>(), tilesX, tilesY,
                                              b.isectIds.data_ptr<int64_t>(), b.gaussianIds.data_ptr<int32_t>(),
                                              gsb::stream()),
               "gsb_map_gaussian_to_intersects");
    const size_t wsBytes = gsb_sort_workspace_bytes(m);
    torch::Tensor ws = torch::empty({(int64_t)wsBytes + 256}, gsb::like(x, torch::kUInt8));
    char *wsPtr = (char *)ws.data_ptr();
    wsPtr += (256 - ((uintptr_t)wsPtr % 256)) % 256;
    gsb::check(gsb_sort_intersects(m, numTiles, b.isectIds.data_ptr<int64_t>(),
                                   b.isectIdsSorted.data_ptr<int64_t>(), b.sortedIndex.data_ptr<int32_t>(), wsPtr,
                                   wsBytes, gsb::stream()),
               "gsb_sort_intersects");
    gsb::check(gsb_gather_bin_edges(m, numTiles, b.isectIdsSorted.data_ptr<int64_t>(),
                                    b.sortedIndex.data_ptr<int32_t>(), b.gaussianIds.data_ptr<int32_t>(),
                                    b.gaussianIdsSorted.data_ptr<int32_t>(), b.tileBins.data_ptr<int32_t>(),
                                    gsb::stream()),
               "gsb_gather_bin_edges");
    return b;
}

}  // namespace

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
binAndSortGaussians(int numPoints, int numIntersects, torch::Tensor xys, torch::Tensor depths,
                    torch::Tensor radii, torch::Tensor cumTilesHit, TileBounds tileBounds) {
    c10::cuda::CUDAGuard guard(xys.device());
    Binned b = bin_and_sort(numPoints, numIntersects, xys, depths, radii, cumTilesHit, tileBounds);
    return std::make_tuple(b.isectIds, b.gaussianIds, b.isectIdsSorted, b.gaussianIdsSorted, b.tileBins);
}

torch::Tensor RasterizeGaussians::forward(AutogradContext *ctx, torch::Tensor xys, torch::Tensor depths,
                                          torch::Tensor radii, torch::Tensor conics, torch::Tensor numTilesHit,
                                          torch::Tensor colors, torch::Tensor opacity, int imgHeight,
                                          int imgWidth, torch::Tensor background) {
    const int n = (int)xys.size(0);
    TORCH_CHECK(colors.size(-1) == 3, "RasterizeGaussians: only 3 colour channels are supported");
    c10::cuda::CUDAGuard guard(xys.device());
    const TileBounds tileBounds =
        std::make_tuple((imgWidth + BLOCK_X - 1) / BLOCK_X, (imgHeight + BLOCK_Y - 1) / BLOCK_Y, 1);
    const int tilesX = std::get<0>(tileBounds), tilesY = std::get<1>(tileBounds);
    torch::Tensor x = gsb::f32(xys), con = gsb::f32(conics), col = gsb::f32(colors), op = gsb::f32(opacity);
    torch::Tensor bg = gsb::f32(background), nth = gsb::i32(numTilesHit);

    // inclusive scan (slot offsets of the gradient rows) ...
    torch::Tensor cum = torch::empty({n}, gsb::like(x, torch::kInt32));
    torch::Tensor d = gsb::f32(depths), r = gsb::i32(radii);
    const int numTiles = tilesX * tilesY;
    torch::Tensor tileBins = torch::empty({numTiles, 2}, gsb::like(x, torch::kInt32));
    torch::Tensor stats = torch::zeros({2}, gsb::like(x, torch::kInt32));
    torch::Tensor tileCursor = torch::empty({(int64_t)gsb_bucket_cursor_bytes(numTiles) / 4}, gsb::like(x, torch::kInt32));
    if (n > 0) {
        const size_t sb = gsb_cumsum_workspace_bytes(n);
        torch::Tensor sws = torch::empty({(int64_t)sb}, gsb::like(x, torch::kUInt8));
        gsb::check(gsb_cumsum_tiles_hit(n, nth.data_ptr<int32_t>(), cum.data_ptr<int32_t>(), sws.data_ptr(), sb,
                                        nullptr, gsb::stream()),
                   "gsb_cumsum_tiles_hit");
    }
    // ... tile sizes -> tile_bins, and the path's single device->host read-back (rasterize_gaussians.cpp:63):
    // M together with the longest tile list
    gsb::check(gsb_bucket_tile_ranges(n, gsb::fp(x), r.data_ptr<int32_t>(), tilesX, tilesY,
                                      tileBins.data_ptr<int32_t>(), tileCursor.data_ptr<int32_t>(),
                                      stats.data_ptr<int32_t>(), gsb::stream()),
               "gsb_bucket_tile_ranges");
    torch::Tensor statsHost = stats.cpu();
    const int m = statsHost[0].item<int>(), maxLen = statsHost[1].item<int>();

    torch::Tensor records = torch::empty({(int64_t)gsb_raster_records_bytes(m)}, gsb::like(x, torch::kUInt8));
    torch::Tensor outImg = torch::empty({imgHeight, imgWidth, 3}, gsb::like(x, torch::kFloat32));
    torch::Tensor finalTs = torch::empty({imgHeight, imgWidth}, gsb::like(x, torch::kFloat32));
    torch::Tensor finalIdx = torch::empty({imgHeight, imgWidth}, gsb::like(x, torch::kInt32));
    if (maxLen <= gsb_bucket_max_tile_len()) {
        // fast path: two-level bucket sort fused with the record packing
        const size_t wsBytes = gsb_bucket_workspace_bytes(n, m);
        torch::Tensor ws = torch::empty({(int64_t)wsBytes + 256}, gsb::like(x, torch::kUInt8));
        char *wp = (char *)ws.data_ptr();
        wp += (256 - ((uintptr_t)wp % 256)) % 256;
        gsb::check(gsb_bucket_sort_pack(n, m, maxLen, gsb::fp(x), gsb::fp(d), r.data_ptr<int32_t>(),
                                        cum.data_ptr<int32_t>(), tilesX, tilesY, tileBins.data_ptr<int32_t>(),
                                        tileCursor.data_ptr<int32_t>(), gsb::fp(con), gsb::fp(col), gsb::fp(op), wp, wsBytes, records.data_ptr(),
                                        nullptr, nullptr, gsb::stream()),
                   "gsb_bucket_sort_pack");
        gsb::check(gsb_rasterize_forward_packed(imgHeight, imgWidth, tilesX, tilesY, m,
                                                tileBins.data_ptr<int32_t>(), gsb::fp(bg), records.data_ptr(),
                                                gsb::fpw(outImg), gsb::fpw(finalTs), finalIdx.data_ptr<int32_t>(),
                                                gsb::stream()),
                   "gsb_rasterize_forward_packed");
    } else {
        // pathological tile lists: generic global radix sort
        Binned b = bin_and_sort(n, m, x, d, r, cum, tileBounds);
        tileBins = b.tileBins;
        gsb::check(gsb_rasterize_forward(imgHeight, imgWidth, tilesX, tilesY, m,
                                         b.gaussianIdsSorted.data_ptr<int32_t>(),
                                         b.sortedIndex.data_ptr<int32_t>(), b.tileBins.data_ptr<int32_t>(),
                                         gsb::fp(x), gsb::fp(con), gsb::fp(col), gsb::fp(op), gsb::fp(bg),
                                         records.data_ptr(), gsb::fpw(outImg), gsb::fpw(finalTs),
                                         finalIdx.data_ptr<int32_t>(), gsb::stream()),
                   "gsb_rasterize_forward");
    }

    ctx->saved_data["imgWidth"] = imgWidth;
    ctx->saved_data["imgHeight"] = imgHeight;
    ctx->saved_data["numIntersects"] = m;
    ctx->save_for_backward({tileBins, con, op, records, cum, bg, finalTs, finalIdx});
    return outImg;
}

tensor_list RasterizeGaussians::backward(AutogradContext *ctx, tensor_list grad_outputs) {
    const int imgHeight = (int)ctx->saved_data["imgHeight"].toInt();
    const int imgWidth = (int)ctx->saved_data["imgWidth"].toInt();
    const int m = (int)ctx->saved_data["numIntersects"].toInt();
    variable_list saved = ctx->get_saved_variables();
    torch::Tensor tileBins = saved[0], con = saved[1], op = saved[2], records = saved[3], cum = saved[4];
    torch::Tensor bg = saved[5], finalTs = saved[6], finalIdx = saved[7];
    const int n = (int)con.size(0);
    c10::cuda::CUDAGuard guard(con.device());
    torch::Tensor v_out = gsb::f32(grad_outputs[0]);  // may arrive as an expanded (stride-0) tensor
    torch::Tensor rows = torch::empty({(int64_t)gsb_raster_grad_rows_bytes(m)}, gsb::like(con, torch::kUInt8));
    torch::Tensor v_xy = torch::empty({n, 2}, gsb::like(con, torch::kFloat32));
    torch::Tensor v_conic = torch::empty({n, 3}, gsb::like(con, torch::kFloat32));
    torch::Tensor v_colors = torch::empty({n, 3}, gsb::like(con, torch::kFloat32));
    torch::Tensor v_opacity = torch::empty({n, 1}, gsb::like(con, torch::kFloat32));
    // v_output_alpha is identically zero in the reference (rasterize_gaussians.cpp:108) -> NULL
    gsb::check(gsb_rasterize_backward(imgHeight, imgWidth, (imgWidth + BLOCK_X - 1) / BLOCK_X,
                                      (imgHeight + BLOCK_Y - 1) / BLOCK_Y, n, m, tileBins.data_ptr<int32_t>(),
                                      gsb::fp(con), gsb::fp(op), records.data_ptr(), cum.data_ptr<int32_t>(),
                                      gsb::fp(bg), gsb::fp(finalTs), finalIdx.data_ptr<int32_t>(), gsb::fp(v_out),
                                      nullptr, rows.data_ptr(), gsb::fpw(v_xy), gsb::fpw(v_conic),
                                      gsb::fpw(v_colors), gsb::fpw(v_opacity), gsb::stream()),
               "gsb_rasterize_backward");
    torch::Tensor none;
    return {v_xy, none, none, v_conic, none, v_colors, v_opacity, none, none, none};
}

torch::Tensor RasterizeGaussiansCPU::forward(AutogradContext *, torch::Tensor, torch::Tensor, torch::Tensor,
                                             torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, int, int,
                                             torch::Tensor) {
    TORCH_CHECK(false, "RasterizeGaussiansCPU: the gsplat_b200 back end has no CPU path; link the reference's "
                       "rasterizer/gsplat-cpu for CPU execution");
    return {};
}

tensor_list RasterizeGaussiansCPU::backward(AutogradContext *, tensor_list) {
    TORCH_CHECK(false, "RasterizeGaussiansCPU: no CPU path in the gsplat_b200 back end");
    return {};
}
