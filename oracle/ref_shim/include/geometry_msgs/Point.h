#pragma once
#include <iostream>
#include <vector>
#include <std_msgs/Header.h>
namespace geometry_msgs {
struct Point {
  double x, y, z;
  Point() : x(0), y(0), z(0) {}
};
inline std::ostream& operator<<(std::ostream& o, const Point& p) { return o << "x: " << p.x << "\ny: " << p.y << "\nz: " << p.z << "\n"; }
struct Vector3 {
  double x, y, z;
  Vector3() : x(0), y(0), z(0) {}
};
struct Quaternion {
  double x, y, z, w;
  Quaternion() : x(0), y(0), z(0), w(0) {}
};
struct Point32 {
  float x, y, z;
  Point32() : x(0), y(0), z(0) {}
};
struct Pose {
  Point position;
  Quaternion orientation;
};
struct PointStamped {
  std_msgs::Header header;
  Point point;
};
inline std::ostream& operator<<(std::ostream& o, const PointStamped& p) { return o << p.point; }
struct PoseStamped {
  std_msgs::Header header;
  Pose pose;
};
struct Polygon {
  std::vector<Point32> points;
};
struct PolygonStamped {
  std_msgs::Header header;
  Polygon polygon;
};
}  // namespace geometry_msgs
