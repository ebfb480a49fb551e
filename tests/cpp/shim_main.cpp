// Drives the C++ host shim (FootholdPlannerGpu.hpp) exactly like a ROS node would: map update, then the
// plan_global_footholds service body.  Reads a raw map file written by the test, prints the response rows.
//   shim_main <map.bin> <rows> <cols> <res> <start_row> <start_col> <x> <y> <z> <gait_cycles> [msg]
// With the trailing "msg" the map arrives as a grid_map_msgs/GridMap (the reference's callback argument) instead of
// an already converted grid_map::GridMap view.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "FootholdPlannerGpu.hpp"

int main(int argc, char** argv) {
  if (argc != 11 && argc != 12) return 2;
  const int rows = atoi(argv[2]), cols = atoi(argv[3]);
  const double res = atof(argv[4]);
  std::vector<float> trav(static_cast<size_t>(rows) * cols), elev(trav.size());
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  if (fread(&trav[0], 4, trav.size(), f) != trav.size() || fread(&elev[0], 4, elev.size(), f) != elev.size()) return 4;
  fclose(f);
  foothold_planner::FootholdPlannerGpu planner(0);
  foothold_planner::GridMapView m;
  m.rows = rows;
  m.cols = cols;
  m.resolution = res;
  m.length_x = rows * res;
  m.length_y = cols * res;
  m.pos_x = 0;
  m.pos_y = 0;
  m.start_row = atoi(argv[5]);
  m.start_col = atoi(argv[6]);
  m.layer_names.push_back("traversability");
  m.layer_names.push_back("elevation");
  m.layer_data.push_back(&trav[0]);
  m.layer_data.push_back(&elev[0]);
  if (argc == 12) {
    grid_map_msgs::GridMap msg;
    msg.info.header.frame_id = "odom";
    msg.info.resolution = res;
    msg.info.length_x = rows * res;
    msg.info.length_y = cols * res;
    msg.layers = m.layer_names;
    msg.outer_start_index = static_cast<uint16_t>(m.start_row);
    msg.inner_start_index = static_cast<uint16_t>(m.start_col);
    for (int l = 0; l < 2; ++l) {  // GridMapRosConverter::toMessage layout of a column-major matrix
      std_msgs::Float32MultiArray a;
      a.layout.dim.resize(2);
      a.layout.dim[0].label = "column_index";
      a.layout.dim[0].size = cols;
      a.layout.dim[0].stride = rows * cols;
      a.layout.dim[1].label = "row_index";
      a.layout.dim[1].size = rows;
      a.layout.dim[1].stride = rows;
      a.data = l == 0 ? trav : elev;
      msg.data.push_back(a);
    }
    planner.gridmapCallback(msg);
    if (planner.gridMapInfo_.row != rows || planner.gridMapInfo_.col != cols || planner.gridMapInfo_.frameID != "odom") return 6;
  } else {
    planner.gridmapCallback(m);
  }
  planner.initialPose_[0] = atof(argv[7]);
  planner.initialPose_[1] = atof(argv[8]);
  planner.initialPose_[2] = atof(argv[9]);
  foothold_planner::GlobalFootholdPlanRequest req;
  req.gait_cycles = static_cast<uint8_t>(atoi(argv[10]));
  foothold_planner::GlobalFootholdPlanResponse resp;
  if (!planner.globalFootholdPlan(req, resp)) return 5;
  printf("%d %d %d %zu\n", resp.footholds.success ? 1 : 0, resp.footholds.gait_cycles, resp.footholds.gait_cycles_succeed,
         resp.footholds.footholds.size());
  for (size_t i = 0; i < resp.footholds.footholds.size(); ++i) {
    const foothold_planner::Foothold& h = resp.footholds.footholds[i];
    printf("%.17g %.17g %.17g %d %d\n", h.point.x, h.point.y, h.point.z, h.foot_id, h.gait_cycle_id);
  }
  // the fine-grained public members, once each (values checked by the test against the oracle)
  foothold_planner::Point init[4], cur[4], next[4], poly[4][4], centre, cand;
  if (!planner.setFirstGait(init, cur)) return 7;
  planner.getFootholdSearchGridMap(cur, -0.007, next, poly, centre);
  const bool candOk = planner.checkCandidateFoothold(next[0].x, next[0].y, next[0].x + 0.01, next[0].y - 0.02, cand);
  const bool cpOk = planner.checkCirclePolygonFoothold(next[1].x, next[1].y, next[1].x, next[1].y);
  printf("members %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d %.17g %.17g %d\n", cur[0].x, cur[3].y, centre.x, centre.y,
         next[2].x, next[2].y, poly[3][0].x, poly[3][2].y, candOk ? 1 : 0, candOk ? cand.x : 0.0, candOk ? cand.y : 0.0,
         cpOk ? 1 : 0);
  return 0;
}
