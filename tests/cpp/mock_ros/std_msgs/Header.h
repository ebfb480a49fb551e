#pragma once
#include <ros/ros.h>
namespace std_msgs {
struct Header { uint32_t seq; ros::Time stamp; std::string frame_id; Header() : seq(0) {} };
}
