// gsplat/config.h -- tile geometry macros that leak into the callers (model.cpp:144,
// simple_trainer.cpp:91 use BLOCK_X / BLOCK_Y).  Values fixed by the C ABI (GSB_TILE == 16).
#pragma once
#define BLOCK_X 16
#define BLOCK_Y 16
#define BLOCK_SIZE (BLOCK_X * BLOCK_Y)
#define N_THREADS 256
