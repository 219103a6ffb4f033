// tile_bounds.hpp -- (tiles_x, tiles_y, 1); same typedef the reference's callers use
// (/root/reference/tile_bounds.hpp:6, model.cpp:144, simple_trainer.cpp:91).
#pragma once
#include <tuple>

using TileBounds = std::tuple<int, int, int>;
