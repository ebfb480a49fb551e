#pragma once
#include <string>
#include <ros/ros.h>
namespace std_msgs {
struct Header {
  unsigned seq;
  ros::Time stamp;
  std::string frame_id;
  Header() : seq(0) {}
};
struct ColorRGBA {
  float r, g, b, a;
  ColorRGBA() : r(0), g(0), b(0), a(0) {}
};
}  // namespace std_msgs
