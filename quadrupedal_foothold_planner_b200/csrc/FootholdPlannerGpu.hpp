// FootholdPlannerGpu.hpp — the host side a ROS node links: a C++11, header-only class with the member names,
// argument meaning and error behaviour of foothold_planner::FootholdPlanner for the hot path
// (foothold_planner/include/foothold_planner/FootholdPlanner.hpp:38-757), implemented on the libfhp C-ABI
// (include/fhp.h).  ROS and grid_map stay on the host: with FHP_WITH_ROS defined the ROS/grid_map types are used
// directly; without it (this image has no ROS) plain structs of the same shape stand in, so the file compiles and
// is exercised by tests/test_cpp_shim.py.
//
//   reference member                               here
//   ---------------------------------------------  -------------------------------------------------------------
//   gridmapCallback(grid_map_msgs::GridMap)         gridmapCallback(const GridMapView&)      FP.cpp:504-536
//   globalFootholdPlan(Request&, Response&)         globalFootholdPlan(Request&, Response&)  FP.cpp:539-1602
//   checkFoothold(map, center, r, R, poly, out, ok) checkFoothold(center, polyCenter, out, ok)     FP.cpp:2001-2036
//   checkDefaultFoothold(map, center, r)            checkDefaultFoothold(center)             FP.cpp:2039-2082
//   checkFootholdUseCentroidMethod(map, p, out,..)  checkFootholdUseCentroidMethod(p, out)   FP.cpp:1605-1997
//   getFootholdMeanHeight(map, center, r, h)        getFootholdMeanHeight(center)            FP.cpp:2520-2554
// The map arguments disappear because the map is device-resident after gridmapCallback (the reference passes
// copies of the same member map); radii/thresholds come from the rosparams as in the reference.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "fhp.h"

namespace foothold_planner {

#ifndef FHP_WITH_ROS
// stand-ins with the field names of geometry_msgs/Point, foothold_planner_msgs/Foothold and GlobalFootholds
struct Point {
  double x, y, z;
  Point() : x(0), y(0), z(0) {}
};
struct Foothold {
  Point point;
  uint8_t foot_id;
  uint8_t gait_cycle_id;
  Foothold() : foot_id(0), gait_cycle_id(0) {}
};
struct GlobalFootholds {
  bool success;
  uint8_t gait_cycles;
  uint8_t gait_cycles_succeed;
  std::vector<Foothold> footholds;
  GlobalFootholds() : success(false), gait_cycles(0), gait_cycles_succeed(0) {}
};
struct GlobalFootholdPlanRequest {  // srv/GlobalFootholdPlan.srv:1-5
  Point FR_current_foothold, RR_current_foothold, RL_current_foothold, FL_current_foothold;  // ignored (FP.cpp:546-549)
  uint8_t gait_cycles;
  GlobalFootholdPlanRequest() : gait_cycles(0) {}
};
struct GlobalFootholdPlanResponse {  // srv/GlobalFootholdPlan.srv:7
  GlobalFootholds footholds;
};
#endif

// What gridmapCallback needs from a grid_map::GridMap after GridMapRosConverter::fromMessage (FP.cpp:506):
// geometry + one column-major float pointer per layer (gridmap_[name].data()).
struct GridMapView {
  int rows, cols;
  double resolution, length_x, length_y, pos_x, pos_y;
  int start_row, start_col;
  std::vector<std::string> layer_names;
  std::vector<const float*> layer_data;
};

class FootholdPlannerGpu {
 public:
  // rosparams the path reads (FP.cpp:249-311); defaults = config/foothold_planner.yaml
  fhp_params params;
  double initialPose_[3];
  GlobalFootholds nominalGlobalFootholdsMsg_;
  GlobalFootholds centroidGlobalFootholdsMsg_;
  GlobalFootholds defaultGlobalFootholds_;  // globalFootholdsResult_.defaultFootholds (FP.cpp:666-671, :1344-1348)
  // footholdsKPI_ (FootholdPlanner.hpp:732-745): two values per successful cycle
  std::vector<double> feetDistance_nominal, cogSpeed_nominal, feetDistance_centroid, cogSpeed_centroid;
  // marker payloads of publishGlobalFootholdSphereMarkers / ...CubeMarkers / ...PointMarkers (FP.cpp:2731-2790):
  // one SPHERE_LIST / CUBE_LIST / POINTS point array per foot, in foot order RF,RH,LH,LF
  std::vector<Point> nominalMarkerPoints[4], centroidMarkerPoints[4], defaultMarkerPoints[4];

  explicit FootholdPlannerGpu(int device = 0) : ctx_(NULL) {
    fhp_default_params(&params);
    initialPose_[0] = initialPose_[1] = initialPose_[2] = 0.0;
    if (fhp_create(device, &ctx_) != FHP_OK) throw std::runtime_error(std::string("fhp_create: ") + fhp_last_error(NULL));
  }
  ~FootholdPlannerGpu() { fhp_destroy(ctx_); }

  // FP.cpp:504-536 — upload once per map update
  void gridmapCallback(const GridMapView& m) {
    std::vector<const char*> names;
    for (size_t i = 0; i < m.layer_names.size(); ++i) names.push_back(m.layer_names[i].c_str());
    fhp_map_desc d;
    d.rows = m.rows;
    d.cols = m.cols;
    d.resolution = m.resolution;
    d.length_x = m.length_x;
    d.length_y = m.length_y;
    d.pos_x = m.pos_x;
    d.pos_y = m.pos_y;
    d.start_row = m.start_row;
    d.start_col = m.start_col;
    d.n_layers = static_cast<int32_t>(names.size());
    d.layer_names = names.empty() ? NULL : &names[0];
    check(fhp_set_map(ctx_, &d, m.layer_data.empty() ? NULL : &m.layer_data[0]), "fhp_set_map");
  }

  // FP.cpp:539-1602 — the plan_global_footholds service body.  Returns the service's bool; planning failures are
  // reported in-band through response.footholds.success like the reference.
  bool globalFootholdPlan(GlobalFootholdPlanRequest& request, GlobalFootholdPlanResponse& response) {
    const int cycles = request.gait_cycles;
    fhp_query q;
    q.x = initialPose_[0];
    q.y = initialPose_[1];
    q.z = initialPose_[2];
    q.param_id = 0;
    q.reserved = 0;
    std::vector<fhp_foothold> nom(FHP_ROWS_PER_QUERY(cycles)), cen(FHP_ROWS_PER_QUERY(cycles)), dfl(FHP_ROWS_PER_QUERY(cycles));
    fhp_query_status st;
    fhp_plan_outputs out;
    out.nominal = &nom[0];
    out.centroid = &cen[0];
    out.default_chain = &dfl[0];
    out.status = &st;
    if (fhp_plan_ex(ctx_, &params, 1, &q, 1, cycles, &out) != FHP_OK) return false;
    fill(nominalGlobalFootholdsMsg_, nom, st, nominalMarkerPoints);
    fill(centroidGlobalFootholdsMsg_, cen, st, centroidMarkerPoints);
    fill(defaultGlobalFootholds_, dfl, st, defaultMarkerPoints);
    const int succ = st.n_rows / 4 - 1;
    feetDistance_nominal.assign(2 * succ, 0.0);
    cogSpeed_nominal.assign(2 * succ, 0.0);
    feetDistance_centroid.assign(2 * succ, 0.0);
    cogSpeed_centroid.assign(2 * succ, 0.0);
    if (succ > 0) {  // getHipDistance / getCogSpeed (FP.cpp:1357-1368, :1422-1433)
      fhp_cycle_kpi(&params, &q, &nom[0], st.n_rows, &feetDistance_nominal[0], &cogSpeed_nominal[0]);
      fhp_cycle_kpi(&params, &q, &cen[0], st.n_rows, &feetDistance_centroid[0], &cogSpeed_centroid[0]);
    }
    response.footholds = nominalGlobalFootholdsMsg_;  // FP.cpp:1588
    return true;
  }

  // FP.cpp:2001-2036
  void checkFoothold(double cx, double cy, double poly_cx, double poly_cy, Point& footholdResult, bool& footholdValidation) {
    const double c[2] = {cx, cy}, rc[2] = {poly_cx, poly_cy};
    fhp_foothold out;
    uint8_t ok = 0;
    check(fhp_check_foothold(ctx_, &params, c, rc, 1, &out, &ok), "fhp_check_foothold");
    footholdResult.x = out.x;
    footholdResult.y = out.y;
    if (ok) footholdResult.z = out.z;  // z untouched on failure (FP.cpp:2028-2033)
    footholdValidation = ok != 0;
  }
  // FP.cpp:2039-2082
  bool checkDefaultFoothold(double cx, double cy) {
    const double c[2] = {cx, cy};
    uint8_t ok = 0;
    check(fhp_check_default(ctx_, &params, c, 1, &ok), "fhp_check_default");
    return ok != 0;
  }
  // FP.cpp:1605-1997 (policy: returns true on the paths where the reference falls off the end, false when the
  // submap fails)
  bool checkFootholdUseCentroidMethod(double px, double py, Point& resultFoothold) {
    const double c[2] = {px, py};
    fhp_foothold out;
    int8_t cs = 0;
    check(fhp_check_centroid(ctx_, &params, c, 1, &out, &cs), "fhp_check_centroid");
    if (cs < 0) return false;
    if (cs <= 3) {
      resultFoothold.x = out.x;
      resultFoothold.y = out.y;
      resultFoothold.z = out.z;
    }
    return true;
  }
  // FP.cpp:2520-2554
  float getFootholdMeanHeight(double cx, double cy) {
    const double c[2] = {cx, cy};
    float z = 0;
    check(fhp_mean_height(ctx_, &params, c, 1, &z), "fhp_mean_height");
    return z;
  }
  fhp_ctx* ctx() { return ctx_; }

 private:
  fhp_ctx* ctx_;
  FootholdPlannerGpu(const FootholdPlannerGpu&);
  FootholdPlannerGpu& operator=(const FootholdPlannerGpu&);
  void check(int rc, const char* what) {
    if (rc == FHP_ERR_MISSING_LAYER) throw std::out_of_range(fhp_last_error(ctx_));  // GridMap::at throws this
    if (rc != FHP_OK) throw std::runtime_error(std::string(what) + ": " + fhp_last_error(ctx_));
  }
  // response assembly (FP.cpp:685-699, :1378-1396) and marker point lists (:2731-2790) in one pass over the rows
  static void fill(GlobalFootholds& msg, const std::vector<fhp_foothold>& rows, const fhp_query_status& st,
                   std::vector<Point> markers[4]) {
    for (int f = 0; f < 4; ++f) markers[f].clear();
    msg.success = st.success != 0;
    msg.gait_cycles = st.gait_cycles;
    msg.gait_cycles_succeed = st.gait_cycles_succeed;
    msg.footholds.clear();
    for (int i = 0; i < st.n_rows; ++i) {  // FP.cpp:685-699, :1378-1396
      Foothold f;
      f.point.x = rows[i].x;
      f.point.y = rows[i].y;
      f.point.z = rows[i].z;
      f.foot_id = rows[i].foot_id;
      f.gait_cycle_id = rows[i].gait_cycle_id;
      msg.footholds.push_back(f);
      if (i >= 4) markers[rows[i].foot_id & 3].push_back(f.point);  // the stance has its own marker (FP.cpp:757)
    }
  }
};

}  // namespace foothold_planner
