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
//   checkCandidateFoothold(map, c, r, R, poly, out) checkCandidateFoothold(c, polyCenter, out)  FP.cpp:2085-2114
//   checkCirclePolygonFoothold(map, c, r, poly)     checkCirclePolygonFoothold(c, polyCenter)   FP.cpp:2117-2163
//   setFirstGait / getPolygonCenter / getDefaultFootholdNext / getSearchPolygon / getFootholdSearchGridMap
//                                                   same names, host-side (fhp_initial_stance, fhp_search_geometry)
// The map arguments disappear because the map is device-resident after gridmapCallback (the reference passes
// copies of the same member map); radii/thresholds come from the rosparams as in the reference.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "fhp.h"

#ifndef FHP_WITH_ROS
// stand-ins with the field names of grid_map_msgs/GridMap and the std_msgs / geometry_msgs types it embeds
namespace std_msgs {
struct Header {
  uint32_t seq;
  struct { uint32_t sec, nsec; } stamp;
  std::string frame_id;
  Header() : seq(0) { stamp.sec = stamp.nsec = 0; }
};
struct MultiArrayDimension {
  std::string label;
  uint32_t size, stride;
  MultiArrayDimension() : size(0), stride(0) {}
};
struct MultiArrayLayout {
  std::vector<MultiArrayDimension> dim;
  uint32_t data_offset;
  MultiArrayLayout() : data_offset(0) {}
};
struct Float32MultiArray {
  MultiArrayLayout layout;
  std::vector<float> data;
};
}  // namespace std_msgs
namespace grid_map_msgs {
struct GridMapInfo {
  std_msgs::Header header;
  double resolution, length_x, length_y;
  struct { struct { double x, y, z; } position; struct { double x, y, z, w; } orientation; } pose;
  GridMapInfo() : resolution(0), length_x(0), length_y(0) {
    pose.position.x = pose.position.y = pose.position.z = 0;
    pose.orientation.x = pose.orientation.y = pose.orientation.z = 0;
    pose.orientation.w = 1;
  }
};
struct GridMap {
  GridMapInfo info;
  std::vector<std::string> layers, basic_layers;
  std::vector<std_msgs::Float32MultiArray> data;
  uint16_t outer_start_index, inner_start_index;
  GridMap() : outer_start_index(0), inner_start_index(0) {}
};
}  // namespace grid_map_msgs
#endif

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

  // gridMapInfo_ of the reference (FootholdPlanner.hpp; filled at FP.cpp:508-515)
  struct GridMapInfo {
    double x, y, length, width, resolution;
    int row, col;
    std::string frameID;
    GridMapInfo() : x(0), y(0), length(0), width(0), resolution(0), row(0), col(0) {}
  } gridMapInfo_;

  // FP.cpp:504-536 with the message itself: GridMapRosConverter::fromMessage and the device upload in one step —
  // the layer arrays are read in place from msg.data[i].data (no Eigen copy, no by-value message copy).
  void gridmapCallback(const grid_map_msgs::GridMap& msg) {
    std::vector<fhp_str> names(msg.layers.size());
    for (size_t i = 0; i < msg.layers.size(); ++i) {
      names[i].p = msg.layers[i].data();
      names[i].n = static_cast<uint32_t>(msg.layers[i].size());
    }
    std::vector<fhp_multiarray> arrs(msg.data.size());
    for (size_t i = 0; i < msg.data.size(); ++i) {
      const std_msgs::Float32MultiArray& a = msg.data[i];
      fhp_multiarray& v = arrs[i];
      v.n_dims = static_cast<uint32_t>(a.layout.dim.size());
      v.reserved = 0;
      v.dim0_label.p = v.dim1_label.p = NULL;
      v.dim0_label.n = v.dim1_label.n = 0;
      v.dim0_size = v.dim1_size = 0;
      if (v.n_dims > 0) {
        v.dim0_label.p = a.layout.dim[0].label.data();
        v.dim0_label.n = static_cast<uint32_t>(a.layout.dim[0].label.size());
        v.dim0_size = a.layout.dim[0].size;
      }
      if (v.n_dims > 1) {
        v.dim1_label.p = a.layout.dim[1].label.data();
        v.dim1_label.n = static_cast<uint32_t>(a.layout.dim[1].label.size());
        v.dim1_size = a.layout.dim[1].size;
      }
      v.data = a.data.empty() ? NULL : &a.data[0];
      v.n = a.data.size();
    }
    fhp_gridmap_msg m;
    m.resolution = msg.info.resolution;
    m.length_x = msg.info.length_x;
    m.length_y = msg.info.length_y;
    m.pos_x = msg.info.pose.position.x;
    m.pos_y = msg.info.pose.position.y;
    m.n_layers = static_cast<int32_t>(names.size());
    m.n_data = static_cast<int32_t>(arrs.size());
    m.layers = names.empty() ? NULL : &names[0];
    m.data = arrs.empty() ? NULL : &arrs[0];
    m.outer_start_index = msg.outer_start_index;
    m.inner_start_index = msg.inner_start_index;
    m.reserved = 0;
    check(fhp_set_map_msg(ctx_, &m), "fhp_set_map_msg");
    fhp_map_desc d;
    check(fhp_map_info(ctx_, &d), "fhp_map_info");
    gridMapInfo_.x = d.pos_x;  // FP.cpp:508-515
    gridMapInfo_.y = d.pos_y;
    gridMapInfo_.length = d.length_x;
    gridMapInfo_.width = d.length_y;
    gridMapInfo_.row = d.rows;
    gridMapInfo_.col = d.cols;
    gridMapInfo_.resolution = d.resolution;
    gridMapInfo_.frameID = msg.info.header.frame_id;
  }

  // the same from an already converted grid_map::GridMap (geometry + gridmap_[layer].data() pointers)
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
  // FP.cpp:2085-2114 — the spiral search alone; resultFoothold.xy is set on a hit only
  bool checkCandidateFoothold(double cx, double cy, double poly_cx, double poly_cy, Point& footholdResult) {
    const double c[2] = {cx, cy}, rc[2] = {poly_cx, poly_cy};
    fhp_foothold out;
    uint8_t ok = 0;
    check(fhp_check_candidate(ctx_, &params, c, rc, 1, &out, &ok), "fhp_check_candidate");
    if (ok) {
      footholdResult.x = out.x;
      footholdResult.y = out.y;
    }
    return ok != 0;
  }
  // FP.cpp:2117-2163
  bool checkCirclePolygonFoothold(double cx, double cy, double poly_cx, double poly_cy) {
    const double c[2] = {cx, cy}, rc[2] = {poly_cx, poly_cy};
    uint8_t ok = 0;
    check(fhp_check_circle_polygon(ctx_, &params, c, rc, 1, &ok), "fhp_check_circle_polygon");
    return ok != 0;
  }
  // FP.cpp:2679-2699 on the stance of initialize() (FP.cpp:350-378): currentFootholds[f] = initial[f] - (stepLength/2, 0, 0)
  bool setFirstGait(Point initialFootholds[4], Point currentFootholds[4]) {
    fhp_query q;
    q.x = initialPose_[0];
    q.y = initialPose_[1];
    q.z = initialPose_[2];
    q.param_id = 0;
    q.reserved = 0;
    double a[12], b[12];
    if (fhp_initial_stance(&params, &q, a, b) != FHP_OK) return false;
    for (int f = 0; f < 4; ++f) {
      initialFootholds[f].x = a[3 * f];
      initialFootholds[f].y = a[3 * f + 1];
      initialFootholds[f].z = a[3 * f + 2];
      currentFootholds[f].x = b[3 * f];
      currentFootholds[f].y = b[3 * f + 1];
      currentFootholds[f].z = b[3 * f + 2];
    }
    return true;
  }
  // FP.cpp:2166-2250: next default footholds and search rectangles (4 vertices each, reference order) of the four
  // current feet RF,RH,LH,LF; ajustedPoseY = ajustedPose_[1] of the cycle.  polygonCenter = getPolygonCenter (:2421-2463).
  void getFootholdSearchGridMap(const Point currentFootholds[4], double ajustedPoseY, Point nextFootholds[4],
                                Point searchPolygons[4][4], Point& polygonCenter) {
    fhp_query q;
    q.x = initialPose_[0];
    q.y = initialPose_[1];
    q.z = initialPose_[2];
    q.param_id = 0;
    q.reserved = 0;
    double cur[12], c[3], nxt[8], poly[32];
    for (int f = 0; f < 4; ++f) {
      cur[3 * f] = currentFootholds[f].x;
      cur[3 * f + 1] = currentFootholds[f].y;
      cur[3 * f + 2] = currentFootholds[f].z;
    }
    check(fhp_search_geometry(&params, &q, ajustedPoseY, cur, c, nxt, poly), "fhp_search_geometry");
    polygonCenter.x = c[0];
    polygonCenter.y = c[1];
    polygonCenter.z = c[2];
    for (int f = 0; f < 4; ++f) {
      nextFootholds[f].x = nxt[2 * f];
      nextFootholds[f].y = nxt[2 * f + 1];
      nextFootholds[f].z = 0.0;  // getDefaultFootholdNext leaves z at 0 (FP.cpp:2411-2418)
      for (int v = 0; v < 4; ++v) {
        searchPolygons[f][v].x = poly[8 * f + 2 * v];
        searchPolygons[f][v].y = poly[8 * f + 2 * v + 1];
        searchPolygons[f][v].z = 0.0;
      }
    }
  }
  Point getPolygonCenter(const Point& RF, const Point& RH, const Point& LH, const Point& LF) {
    const Point cur[4] = {RF, RH, LH, LF};
    Point next[4], poly[4][4], c;
    getFootholdSearchGridMap(cur, 0.0, next, poly, c);
    return c;
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
