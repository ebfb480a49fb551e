#pragma once
#include <cstddef>
#include <vector>
#include <grid_map_core/GridMap.hpp>
namespace grid_map {
class Polygon {
 public:
  void addVertex(const Position& vertex) { vertices_.push_back(vertex); }
  const Position& getVertex(const size_t index) const { return vertices_.at(index); }
  size_t nVertices() const { return vertices_.size(); }
 private:
  std::vector<Position> vertices_;
};
}  // namespace grid_map
