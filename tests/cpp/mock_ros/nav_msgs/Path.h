#pragma once
#include <std_msgs/Header.h>
namespace nav_msgs {
struct Path { std_msgs::Header header; };
}
