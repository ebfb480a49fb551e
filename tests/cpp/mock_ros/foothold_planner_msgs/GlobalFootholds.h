#pragma once
#include <vector>
#include <foothold_planner_msgs/Foothold.h>
#include <std_msgs/Header.h>
namespace foothold_planner_msgs {
struct GlobalFootholds {
  std_msgs::Header header;
  uint8_t success, gait_cycles, gait_cycles_succeed;  // bool / uint8 / uint8 (roscpp generates uint8_t for bool)
  std::vector<Foothold> footholds;
  GlobalFootholds() : success(0), gait_cycles(0), gait_cycles_succeed(0) {}
};
}  // namespace foothold_planner_msgs
