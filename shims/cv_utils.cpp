#include "cv_utils.hpp"

cv::Mat tensorToImage(const torch::Tensor &t) {
    torch::Tensor u8 = (t.detach().cpu().clamp(0.0, 1.0) * 255.0).to(torch::kUInt8).contiguous();
    cv::Mat m((int)u8.size(0), (int)u8.size(1), CV_8UC3);
    std::memcpy(m.data, u8.data_ptr<uint8_t>(), (size_t)u8.numel());
    return m;
}
