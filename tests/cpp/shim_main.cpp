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
  if (!planner.globalFootholdPlan(req, resp)) {  // the reference's opt-chain abort at cycle 0 (FP.cpp:931-934)
    printf("aborted %d %zu\n", planner.optSubmapWouldFail_ ? 1 : 0, planner.nominalGlobalFootholdsMsg_.footholds.size());
    return planner.optSubmapWouldFail_ ? 0 : 5;
  }
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
  // marker payloads (FP.cpp:2731-2790): one point list per foot for each chain
  for (int chain = 0; chain < 3; ++chain) {
    const std::vector<foothold_planner::Point>* mk =
        chain == 0 ? planner.nominalMarkerPoints : (chain == 1 ? planner.centroidMarkerPoints : planner.defaultMarkerPoints);
    printf("markers %d", chain);
    for (int ft = 0; ft < 4; ++ft) {
      printf(" %zu", mk[ft].size());
      for (size_t i = 0; i < mk[ft].size(); ++i) printf(" %.17g %.17g %.17g", mk[ft][i].x, mk[ft][i].y, mk[ft][i].z);
    }
    printf("\n");
  }
  // the same members through the reference's own signatures (FootholdPlanner.hpp:51-448): identical answers
  {
    using foothold_planner::PointStamped;
    grid_map::GridMap gm;  // ignored: the map is device-resident
    const float r = planner.params.foot_radius, R = planner.params.search_radius;
    PointStamped ri[4], rc[4], nx[4];
    for (int ft = 0; ft < 4; ++ft) ri[ft].point = init[ft];
    bool same = planner.setFirstGait(ri[0], ri[1], ri[2], ri[3], rc[0], rc[1], rc[2], rc[3]);
    for (int ft = 0; ft < 4; ++ft) same = same && rc[ft].point.x == cur[ft].x && rc[ft].point.y == cur[ft].y && rc[ft].point.z == cur[ft].z;
    nav_msgs::Path path;
    grid_map::Polygon pg[4];
    ros::Publisher pub;
    planner.ajustedPoseY_ = -0.007;
    same = same && planner.getFootholdSearchGridMap(rc[0], rc[1], rc[2], rc[3], path, nx[0], nx[1], nx[2], nx[3], pg[0], pg[1],
                                                     pg[2], pg[3], pub, pub, pub, pub, pub, pub, pub, pub);
    for (int ft = 0; ft < 4; ++ft) {
      same = same && nx[ft].point.x == next[ft].x && nx[ft].point.y == next[ft].y && pg[ft].nVertices() == 4;
      for (int v = 0; v < 4; ++v) same = same && pg[ft].getVertex(v).x() == poly[ft][v].x && pg[ft].getVertex(v).y() == poly[ft][v].y;
    }
    const foothold_planner::Point c2 = planner.getPolygonCenter(rc[0], rc[1], rc[2], rc[3]);
    same = same && c2.x == centre.x && c2.y == centre.y;
    for (int ft = 0; ft < 4; ++ft) {
      const grid_map::Position at(next[ft].x, next[ft].y);
      PointStamped a;
      foothold_planner::Point b;
      bool okA = false, okB = false;
      planner.checkFoothold(gm, at, r, R, pg[ft], a, okA);
      planner.checkFoothold(next[ft].x, next[ft].y, next[ft].x, next[ft].y, b, okB);
      same = same && okA == okB && a.point.x == b.x && a.point.y == b.y && (!okA || a.point.z == b.z);
      same = same && planner.checkDefaultFoothold(gm, at, r) == planner.checkDefaultFoothold(next[ft].x, next[ft].y);
      same = same && planner.getFootholdMeanHeight(gm, at, r, planner.params.h) == planner.getFootholdMeanHeight(next[ft].x, next[ft].y);
      same = same && planner.checkCirclePolygonFoothold(gm, at, r, pg[ft]) ==
                         planner.checkCirclePolygonFoothold(next[ft].x, next[ft].y, next[ft].x, next[ft].y);
      PointStamped shifted;
      shifted.point.x = next[ft].x + 0.01;
      shifted.point.y = next[ft].y - 0.02;
      PointStamped ca;
      foothold_planner::Point cb;
      const bool k1 = planner.checkCandidateFoothold(gm, at, r, R, planner.getSearchPolygon(shifted, R), ca);
      const bool k2 = planner.checkCandidateFoothold(next[ft].x, next[ft].y, next[ft].x + 0.01, next[ft].y - 0.02, cb);
      same = same && k1 == k2 && (!k1 || (ca.point.x == cb.x && ca.point.y == cb.y));
      PointStamped pa, din;
      foothold_planner::Point pb;
      int br = -1, er = -1;
      din.point = next[ft];
      const bool m1 = planner.checkFootholdUseCentroidMethod(gm, din, pa, br, er);
      const bool m2 = planner.checkFootholdUseCentroidMethod(next[ft].x, next[ft].y, pb);
      same = same && m1 == m2 && pa.point.x == pb.x && pa.point.y == pb.y && pa.point.z == pb.z;
    }
    bool threw = false;
    try {  // a polygon that getSearchPolygon cannot have built is refused, not guessed
      grid_map::Polygon bad = pg[0];
      bad.addVertex(grid_map::Position(0, 0));
      PointStamped a;
      bool okA;
      planner.checkFoothold(gm, grid_map::Position(next[0].x, next[0].y), r, R, bad, a, okA);
    } catch (const std::invalid_argument&) {
      threw = true;
    }
    printf("refsig %d %d\n", same ? 1 : 0, threw ? 1 : 0);
  }
  return 0;
}
