#pragma once
#include <geometry_msgs/Point.h>
#include <grid_map_core/GridMap.hpp>
#include <grid_map_msgs/GridMap.h>
namespace grid_map {
struct GridMapRosConverter {
  static bool fromMessage(const grid_map_msgs::GridMap& msg, grid_map::GridMap& map) {
    if (!msg.decoded) return false;
    map = *msg.decoded;
    return true;
  }
  static void toMessage(const grid_map::GridMap& map, grid_map_msgs::GridMap& msg) { msg.decoded = &map; }
};
struct PolygonRosConverter {
  static void toMessage(const grid_map::Polygon& polygon, geometry_msgs::PolygonStamped& message) {
    message.header.frame_id = polygon.getFrameId();
    message.polygon.points.clear();
    for (size_t i = 0; i < polygon.vertices().size(); ++i) {
      geometry_msgs::Point32 p;
      p.x = static_cast<float>(polygon.vertices()[i].x);
      p.y = static_cast<float>(polygon.vertices()[i].y);
      message.polygon.points.push_back(p);
    }
  }
};
}  // namespace grid_map
