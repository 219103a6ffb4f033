// gsplat.hpp -- back-end include switch of the reference (/root/reference/gsplat.hpp).  With the
// B200 back end there is exactly one GPU flavour: the C ABI of libgsplat_b200.so.
#pragma once
#include <gsplat/config.h>
#include "../../../include/gsplat_b200.h"

// The one helper of the reference's always-compiled CPU bindings (rasterizer/gsplat-cpu/bindings.h:75,
// gsplat_cpu.cpp:409-423) that callers use with every back end: model.hpp:44 sizes the SH tensor with it.
int numShBases(int degree);   // 0..4 -> 1, 4, 9, 16, 25 (anything else 25)
