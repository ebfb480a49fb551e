#pragma once
#include <geometry_msgs/Point.h>
#include <stdint.h>
namespace foothold_planner_msgs {
struct Foothold { geometry_msgs::Point point; uint8_t foot_id, gait_cycle_id; Foothold() : foot_id(0), gait_cycle_id(0) {} };
}
