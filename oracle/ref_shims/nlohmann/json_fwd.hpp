// oracle/ref_shims/nlohmann/json_fwd.hpp -- TEST INFRASTRUCTURE.  Forward declaration only: the reference's
// nerfstudio.hpp:11-37 mentions `json` in function signatures that model.cpp never calls.
#pragma once
namespace nlohmann { class json; }
