#pragma once
// generated-message stand-in: foothold_planner_msgs/msg/Foothold.msg
#include <cstdint>
#include <geometry_msgs/Point.h>
namespace foothold_planner_msgs {
struct Foothold {
  geometry_msgs::Point point;
  uint8_t foot_id;
  uint8_t gait_cycle_id;
  Foothold() : foot_id(0), gait_cycle_id(0) {}
};
}  // namespace foothold_planner_msgs
