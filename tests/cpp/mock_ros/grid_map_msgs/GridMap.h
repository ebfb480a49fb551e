#pragma once
#include <geometry_msgs/Point.h>
#include <std_msgs/Float32MultiArray.h>
#include <std_msgs/Header.h>
namespace grid_map_msgs {
struct GridMapInfo {
  std_msgs::Header header;
  double resolution, length_x, length_y;
  geometry_msgs::Pose pose;
  GridMapInfo() : resolution(0), length_x(0), length_y(0) {}
};
struct GridMap {
  GridMapInfo info;
  std::vector<std::string> layers, basic_layers;
  std::vector<std_msgs::Float32MultiArray> data;
  uint16_t outer_start_index, inner_start_index;
  GridMap() : outer_start_index(0), inner_start_index(0) {}
};
}  // namespace grid_map_msgs
