// shims/model_deps/opencv2/calib3d.hpp -- BUILD SHIM (like shims/cxxopts.hpp).  input_data.hpp:8 includes it but declares nothing
// with OpenCV types; model.cpp needs none of it.
#pragma once
