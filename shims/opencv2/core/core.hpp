// shims/opencv2 -- minimal stand-in for the OpenCV types simple_trainer.cpp:12-14,206-210 touches
// (cv::Mat, cv::cvtColor, cv::imwrite).  OpenCV C++ headers are not available offline; image output
// is outside the hot path.  imwrite writes a binary PPM next to the requested name.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#define CV_8UC3 16
namespace cv {
enum { COLOR_RGB2BGR = 4, COLOR_BGR2RGB = 4 };
class Mat {
public:
    Mat() = default;
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_((size_t)r * c * 3) { data = buf_.data(); }
    int rows = 0, cols = 0;
    uint8_t *data = nullptr;
    int type() const { return type_; }
    bool empty() const { return rows == 0 || cols == 0; }
    void reset_data() { data = buf_.data(); }

private:
    int type_ = CV_8UC3;
    std::vector<uint8_t> buf_;
};
}  // namespace cv
