// shims/model_deps/nlohmann/json_fwd.hpp -- BUILD SHIM (like shims/cxxopts.hpp).  Forward declaration only: the reference's
// nerfstudio.hpp:11-37 mentions `json` in function signatures that model.cpp never calls.
#pragma once
namespace nlohmann { class json; }
