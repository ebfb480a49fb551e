#pragma once
namespace geometry_msgs {
struct Point { double x, y, z; Point() : x(0), y(0), z(0) {} };
struct Quaternion { double x, y, z, w; Quaternion() : x(0), y(0), z(0), w(1) {} };
struct Pose { Point position; Quaternion orientation; };
}  // namespace geometry_msgs
