// gsplat.hpp -- back-end include switch of the reference (/root/reference/gsplat.hpp).  With the
// B200 back end there is exactly one GPU flavour: the C ABI of libgsplat_b200.so.
#pragma once
#include <gsplat/config.h>
#include "../../../include/gsplat_b200.h"
