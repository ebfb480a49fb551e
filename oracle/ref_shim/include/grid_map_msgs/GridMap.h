#pragma once
// Stand-in for the wire message: the test driver hands the already-decoded map through it.
#include <grid_map_core/GridMap.hpp>
namespace grid_map_msgs {
struct GridMap {
  const grid_map::GridMap* decoded;
  GridMap() : decoded(0) {}
};
}  // namespace grid_map_msgs
