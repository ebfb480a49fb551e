#pragma once
// generated-message stand-in: foothold_planner_msgs/msg/GlobalFootholds.msg
#include <vector>
#include <foothold_planner_msgs/Foothold.h>
namespace foothold_planner_msgs {
struct GlobalFootholds {
  std_msgs::Header header;
  uint8_t success;  // bool on the wire
  uint8_t gait_cycles;
  uint8_t gait_cycles_succeed;
  std::vector<Foothold> footholds;
  GlobalFootholds() : success(0), gait_cycles(0), gait_cycles_succeed(0) {}
};
}  // namespace foothold_planner_msgs
