// oracle/ref_shims/opencv2/calib3d.hpp -- TEST INFRASTRUCTURE.  input_data.hpp:8 includes it but declares nothing
// with OpenCV types; model.cpp needs none of it.
#pragma once
