#pragma once
#include <string>
#include <vector>
#include <geometry_msgs/Point.h>
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, POINTS = 8 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
  std_msgs::Header header;
  std::string ns;
  int id;
  int type;
  int action;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  ros::Duration lifetime;
  bool frame_locked;
  std::vector<geometry_msgs::Point> points;
  std::vector<std_msgs::ColorRGBA> colors;
  std::string text;
  Marker() : id(0), type(0), action(0), frame_locked(false) {}
};
struct MarkerArray {
  std::vector<Marker> markers;
};
}  // namespace visualization_msgs
