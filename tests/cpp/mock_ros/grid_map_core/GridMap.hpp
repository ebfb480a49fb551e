#pragma once
// grid_map::Position is Eigen::Vector2d upstream: x(), y(), operator()(i)
namespace grid_map {
class Position {
 public:
  Position() { v_[0] = v_[1] = 0; }
  Position(double x, double y) { v_[0] = x; v_[1] = y; }
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double& x() { return v_[0]; }
  double& y() { return v_[1]; }
  double operator()(int i) const { return v_[i]; }
 private:
  double v_[2];
};
class GridMap {};
}  // namespace grid_map
