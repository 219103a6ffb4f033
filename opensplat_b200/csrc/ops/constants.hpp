// constants.hpp -- constants the reference's callers pull from this header
// (/root/reference/constants.hpp; simple_trainer.cpp:17 needs PI and APP_VERSION).
#pragma once

#define PI 3.14159265358979323846
#define FLOAT_EPS 1e-9f
#ifndef APP_VERSION
#define APP_VERSION "gsplat_b200"
#endif
#ifndef APP_REVISION
#define APP_REVISION "gsplat_b200"
#endif
