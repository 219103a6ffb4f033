#pragma once
#include "core/core.hpp"
namespace cv {
inline void cvtColor(const Mat &src, Mat &dst, int /*code: RGB<->BGR swap*/) {
    Mat out(src.rows, src.cols, CV_8UC3);
    for (size_t i = 0; i < (size_t)src.rows * src.cols; ++i) {
        out.data[3 * i] = src.data[3 * i + 2];
        out.data[3 * i + 1] = src.data[3 * i + 1];
        out.data[3 * i + 2] = src.data[3 * i];
    }
    dst = out;
    dst.reset_data();
}
}  // namespace cv
