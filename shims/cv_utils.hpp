// shims/cv_utils.hpp -- stand-in for the reference's cv_utils.hpp (image IO helpers, out of the hot
// path's scope) declaring the one function simple_trainer.cpp:207 uses.
#pragma once
#include <torch/torch.h>
#include <opencv2/core/core.hpp>

cv::Mat tensorToImage(const torch::Tensor &t);  // [H,W,3] float in [0,1] -> 8-bit RGB
