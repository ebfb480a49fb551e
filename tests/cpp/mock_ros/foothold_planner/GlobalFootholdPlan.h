#pragma once
#include <foothold_planner_msgs/GlobalFootholds.h>
namespace foothold_planner {
struct GlobalFootholdPlan {
  struct Request {
    geometry_msgs::Point FR_current_foothold, RR_current_foothold, RL_current_foothold, FL_current_foothold;
    uint8_t gait_cycles;
    Request() : gait_cycles(0) {}
  };
  struct Response { foothold_planner_msgs::GlobalFootholds footholds; };
};
}  // namespace foothold_planner
