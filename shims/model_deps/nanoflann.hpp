// shims/model_deps/nanoflann.hpp -- BUILD SHIM (like shims/cxxopts.hpp).  Just enough of nanoflann's names for the reference's
// kdtree_tensor.hpp:47-50 type alias to parse when model.cpp is compiled for oracle/_ref (nanoflann is a
// FetchContent dependency of the reference, absent offline).  No k-d tree is ever built: the ref driver supplies
// its own PointsTensor::scales() (used only by Model's constructor for the initial scales, which the driver
// overwrites).
#pragma once
#include <cstddef>
#include <initializer_list>
namespace nanoflann {
template <class T, class DataSource> struct L2_Simple_Adaptor {};
template <class Distance, class DatasetAdaptor, int DIM, class IndexType> struct KDTreeSingleIndexAdaptor {
    KDTreeSingleIndexAdaptor(int, const DatasetAdaptor &, std::initializer_list<int>) {}
};
}  // namespace nanoflann
