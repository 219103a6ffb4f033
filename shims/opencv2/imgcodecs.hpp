#pragma once
#include <cstdio>
#include "core/core.hpp"
namespace cv {
// writes a PPM (P6) with the given path + ".ppm"; channel order as stored
inline bool imwrite(const std::string &path, const Mat &m) {
    FILE *f = std::fopen((path + ".ppm").c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "P6\n%d %d\n255\n", m.cols, m.rows);
    std::fwrite(m.data, 1, (size_t)m.rows * m.cols * 3, f);
    std::fclose(f);
    return true;
}
}  // namespace cv
