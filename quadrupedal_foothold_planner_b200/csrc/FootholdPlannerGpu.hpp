// FootholdPlannerGpu.hpp — the host side a ROS node links: a C++11, header-only class with the member names,
// argument meaning and error behaviour of foothold_planner::FootholdPlanner for the hot path
// (foothold_planner/include/foothold_planner/FootholdPlanner.hpp:38-757), implemented on the libfhp C-ABI
// (include/fhp.h).  ROS and grid_map stay on the host.
//
// Two build modes:
//   -DFHP_WITH_ROS   the real ROS / grid_map headers are included and the class works on geometry_msgs::Point(Stamped),
//                    foothold_planner_msgs::Foothold / GlobalFootholds, foothold_planner::GlobalFootholdPlan::Request /
//                    Response, grid_map_msgs::GridMap, grid_map::Position / Polygon / GridMap, nav_msgs::Path and
//                    ros::Publisher themselves.  (This image has no ROS: tests/test_cpp_shim.py compiles and runs this
//                    mode against tests/cpp/mock_ros/, headers that declare those types with their real field names.)
//   default          plain structs of the same shape stand in, so the file compiles anywhere.
//
// Every member exists twice:
//   * with the reference's OWN signature (FootholdPlanner.hpp:51-448) — by-value maps become const references and are
//     ignored, because the map is device-resident after gridmapCallback (the reference passes copies of the very same
//     member map); radii arguments are honoured per call; a polygon argument must be a rectangle as getSearchPolygon
//     builds it (its centre is recovered exactly, anything else throws std::invalid_argument);
//   * in a short form on plain doubles (used by batch callers and the tests).
//
//   reference member (FootholdPlanner.hpp / FP.cpp)                                   C-ABI entry
//   gridmapCallback(grid_map_msgs::GridMap)                         :491 / :504-536     fhp_set_map_msg / fhp_set_map
//   globalFootholdPlan(Request&, Response&)                         :51  / :539-1602    fhp_plan_ex
//   checkFootholdUseCentroidMethod(map, p, out, beginRow, endRow)   :71  / :1605-1997   fhp_check_centroid
//   checkFoothold(map, center, r, R, polygon, out, ok)              :94  / :2001-2036   fhp_check_foothold
//   checkDefaultFoothold(map, center, r)                            :110 / :2039-2082   fhp_check_default
//   checkCandidateFoothold(map, c, r, R, polygon, out)              :124 / :2085-2114   fhp_check_candidate
//   checkCirclePolygonFoothold(map, c, r, polygon)                  :140 / :2117-2163   fhp_check_circle_polygon
//   getFootholdSearchGridMap(4 feet, path, 4 next, 4 polygons, 8 publishers) :169 / :2166-2250   fhp_search_geometry
//   getDefaultFootholdNext / getPolygonCenter / getSearchPolygon    :218-240 / :2411-2517  fhp_search_geometry
//   getFootholdMeanHeight(map, center, r, h)                        :262 / :2520-2554   fhp_mean_height
//   setFirstGait(4 initial, 4 current)                              :345 / :2679-2699   fhp_initial_stance
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "fhp.h"

#ifdef FHP_WITH_ROS
#include <foothold_planner/GlobalFootholdPlan.h>
#include <foothold_planner_msgs/Foothold.h>
#include <foothold_planner_msgs/GlobalFootholds.h>
#include <geometry_msgs/Point.h>
#include <geometry_msgs/PointStamped.h>
#include <grid_map_core/GridMap.hpp>
#include <grid_map_core/Polygon.hpp>
#include <grid_map_msgs/GridMap.h>
#include <nav_msgs/Path.h>
#include <ros/ros.h>
#include <std_msgs/Float32MultiArray.h>
namespace foothold_planner {
typedef geometry_msgs::Point Point;
typedef geometry_msgs::PointStamped PointStamped;
typedef foothold_planner_msgs::Foothold Foothold;
typedef foothold_planner_msgs::GlobalFootholds GlobalFootholds;
typedef GlobalFootholdPlan::Request GlobalFootholdPlanRequest;
typedef GlobalFootholdPlan::Response GlobalFootholdPlanResponse;
}  // namespace foothold_planner
#else
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
namespace grid_map {  // the three grid_map_core types that appear in the reference's member signatures
struct Position {    // Eigen::Vector2d
  double v[2];
  Position() { v[0] = v[1] = 0; }
  Position(double x, double y) { v[0] = x; v[1] = y; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double& x() { return v[0]; }
  double& y() { return v[1]; }
};
class Polygon {
 public:
  void addVertex(const Position& p) { vertices_.push_back(p); }
  const Position& getVertex(size_t i) const { return vertices_.at(i); }
  size_t nVertices() const { return vertices_.size(); }
 private:
  std::vector<Position> vertices_;
};
class GridMap {};  // the map argument of the reference members: ignored (the map is device-resident)
}  // namespace grid_map
namespace nav_msgs {
struct Path {};
}
namespace ros {
class Publisher {};
}
namespace foothold_planner {
// stand-ins with the field names of geometry_msgs/Point(Stamped), foothold_planner_msgs/Foothold and GlobalFootholds
struct Point {
  double x, y, z;
  Point() : x(0), y(0), z(0) {}
};
struct PointStamped {
  std_msgs::Header header;
  Point point;
};
struct Foothold {
  Point point;
  uint8_t foot_id;
  uint8_t gait_cycle_id;
  Foothold() : foot_id(0), gait_cycle_id(0) {}
};
struct GlobalFootholds {
  uint8_t success;  // bool on the wire
  uint8_t gait_cycles;
  uint8_t gait_cycles_succeed;
  std::vector<Foothold> footholds;
  GlobalFootholds() : success(0), gait_cycles(0), gait_cycles_succeed(0) {}
};
struct GlobalFootholdPlanRequest {  // srv/GlobalFootholdPlan.srv:1-5
  Point FR_current_foothold, RR_current_foothold, RL_current_foothold, FL_current_foothold;  // ignored (FP.cpp:546-549)
  uint8_t gait_cycles;
  GlobalFootholdPlanRequest() : gait_cycles(0) {}
};
struct GlobalFootholdPlanResponse {  // srv/GlobalFootholdPlan.srv:7
  GlobalFootholds footholds;
};
}  // namespace foothold_planner
#endif

namespace foothold_planner {

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

  // true (default): globalFootholdPlan returns false where the reference's service call would (see there)
  bool abortLikeReference;
  bool optSubmapWouldFail_;

  explicit FootholdPlannerGpu(int device = 0)
      : abortLikeReference(true), optSubmapWouldFail_(false), ajustedPoseY_(0.0), ctx_(NULL) {
    fhp_default_params(&params);
    // This class stands in for the reference class: bug-compatible grid_map::SpiralIterator (it ignores the map's
    // circular-buffer start index) so that a moving robot-centric map gives the reference's footholds.
    // Batch callers that want the repaired iterator set params.spiral_mode = FHP_SPIRAL_WRAP_CORRECT.
    params.spiral_mode = FHP_SPIRAL_FAITHFUL;
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
    // The reference aborts the service call when the gait-cycle submap of its opt (NLopt) chain does not fit the map
    // (`return false`, FP.cpp:931-934); libfhp reports that predicate for cycle 0 (FHP_STATUS_OPT_SUBMAP_FAIL).  The
    // reference has only the four stance rows in nominalGlobalFootholdsMsg_ at that point and leaves the response alone.
    optSubmapWouldFail_ = (st.reserved & FHP_STATUS_OPT_SUBMAP_FAIL) != 0;
    if (optSubmapWouldFail_ && abortLikeReference) {
      st.n_rows = 4;
      st.success = 0;
      st.gait_cycles_succeed = 0;
      fill(nominalGlobalFootholdsMsg_, nom, st, nominalMarkerPoints);
      return false;
    }
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

  // ---- the reference's own member signatures (FootholdPlanner.hpp:51-448) ----------------------------------------------
  // FP.hpp:71-75.  beginRow / endRow (the clear band's rows, consumed only by the opt chain's constraints) are set to 0.
  bool checkFootholdUseCentroidMethod(const grid_map::GridMap&, PointStamped defaultFoothold, PointStamped& resultFoothold,
                                      int& beginRow, int& endRow) {
    beginRow = endRow = 0;
    return checkFootholdUseCentroidMethod(defaultFoothold.point.x, defaultFoothold.point.y, resultFoothold.point);
  }
  // FP.hpp:94-100 (returns a default-constructed point, like the reference: FP.cpp:2008, :2035)
  PointStamped checkFoothold(const grid_map::GridMap&, grid_map::Position center, float footRadius, float searchRadius,
                             const grid_map::Polygon& polygon, PointStamped& footholdResult, bool& footholdValidation) {
    const fhp_params p = with_radii(footRadius, searchRadius);
    double rc[2];
    rectangle_centre(polygon, p, rc);
    const double c[2] = {center.x(), center.y()};
    fhp_foothold out;
    uint8_t ok = 0;
    check(fhp_check_foothold(ctx_, &p, c, rc, 1, &out, &ok), "fhp_check_foothold");
    footholdResult.point.x = out.x;
    footholdResult.point.y = out.y;
    if (ok) footholdResult.point.z = out.z;
    footholdValidation = ok != 0;
    return PointStamped();
  }
  // FP.hpp:110-112
  bool checkDefaultFoothold(const grid_map::GridMap&, grid_map::Position center, float footRadius) {
    const fhp_params p = with_radii(footRadius, params.search_radius);
    const double c[2] = {center.x(), center.y()};
    uint8_t ok = 0;
    check(fhp_check_default(ctx_, &p, c, 1, &ok), "fhp_check_default");
    return ok != 0;
  }
  // FP.hpp:124-129
  bool checkCandidateFoothold(const grid_map::GridMap&, grid_map::Position spiralCenter, float footRadius, float searchRadius,
                              const grid_map::Polygon& polygon, PointStamped& footholdResult) {
    const fhp_params p = with_radii(footRadius, searchRadius);
    double rc[2];
    rectangle_centre(polygon, p, rc);
    const double c[2] = {spiralCenter.x(), spiralCenter.y()};
    fhp_foothold out;
    uint8_t ok = 0;
    check(fhp_check_candidate(ctx_, &p, c, rc, 1, &out, &ok), "fhp_check_candidate");
    if (ok) {
      footholdResult.point.x = out.x;
      footholdResult.point.y = out.y;
    }
    return ok != 0;
  }
  // FP.hpp:140-143 (the third argument is the FOOT radius: it is what the reference's CircleIterator is given, FP.cpp:2126)
  bool checkCirclePolygonFoothold(const grid_map::GridMap&, grid_map::Position center, float footRadius,
                                  const grid_map::Polygon& polygon) {
    const fhp_params p = with_radii(footRadius, params.search_radius);
    double rc[2];
    rectangle_centre(polygon, p, rc);
    const double c[2] = {center.x(), center.y()};
    uint8_t ok = 0;
    check(fhp_check_circle_polygon(ctx_, &p, c, rc, 1, &ok), "fhp_check_circle_polygon");
    return ok != 0;
  }
  // FP.hpp:262-265
  float getFootholdMeanHeight(const grid_map::GridMap&, grid_map::Position center, float radius, double h) {
    fhp_params p = with_radii(radius, params.search_radius);
    p.h = h;
    const double c[2] = {center.x(), center.y()};
    float z = 0;
    check(fhp_mean_height(ctx_, &p, c, 1, &z), "fhp_mean_height");
    return z;
  }
  // FP.hpp:345-352.  The reference copies its arguments; the stance itself is initialize()'s (FP.cpp:350-378), which is
  // what the four *_initialPosition arguments hold there.
  bool setFirstGait(PointStamped RF_initialPosition, PointStamped RH_initialPosition, PointStamped LH_initialPosition,
                    PointStamped LF_initialPosition, PointStamped& RF_currentPosition, PointStamped& RH_currentPosition,
                    PointStamped& LH_currentPosition, PointStamped& LF_currentPosition) {
    const PointStamped* in[4] = {&RF_initialPosition, &RH_initialPosition, &LH_initialPosition, &LF_initialPosition};
    PointStamped* out[4] = {&RF_currentPosition, &RH_currentPosition, &LH_currentPosition, &LF_currentPosition};
    const double hs = static_cast<double>(params.step_length / 2);  // float division (FP.cpp:2693)
    const double hx = hs * std::cos(params.heading), hy = hs * std::sin(params.heading);
    for (int f = 0; f < 4; ++f) {
      out[f]->point = in[f]->point;
      if (params.heading != 0.0) {
        out[f]->point.x = in[f]->point.x - hx;
        out[f]->point.y = in[f]->point.y - hy;
      } else {
        out[f]->point.x = in[f]->point.x - hs;
      }
    }
    return true;
  }
  // FP.hpp:227-230
  Point getPolygonCenter(PointStamped RF_point, PointStamped RH_point, PointStamped LH_point, PointStamped LF_point) {
    return getPolygonCenter(RF_point.point, RH_point.point, LH_point.point, LF_point.point);
  }
  // FP.hpp:218-219 (z is left at 0, FP.cpp:2411-2418)
  PointStamped getDefaultFootholdNext(Point feetCenterNext, PointStamped defaultBias) {
    PointStamped f;
    f.point.x = feetCenterNext.x + defaultBias.point.x;
    f.point.y = feetCenterNext.y + defaultBias.point.y;
    return f;
  }
  // FP.hpp:239-240 (vertex order leftup, rightup, rightdown, leftdown, FP.cpp:2496-2517; rotated by params.heading)
  grid_map::Polygon getSearchPolygon(PointStamped center, float radius) {
    const fhp_params p = with_radii(params.foot_radius, radius);
    double v[8];
    rectangle_vertices(p, center.point.x, center.point.y, v);
    grid_map::Polygon poly;
    for (int k = 0; k < 4; ++k) poly.addVertex(grid_map::Position(v[2 * k], v[2 * k + 1]));
    return poly;
  }
  // FP.hpp:169-189.  The path and the eight publishers are host-side visualisation and stay untouched.
  bool getFootholdSearchGridMap(PointStamped RF_position, PointStamped RH_position, PointStamped LH_position,
                                PointStamped LF_position, nav_msgs::Path&, PointStamped& RF_nextPosition,
                                PointStamped& RH_nextPosition, PointStamped& LH_nextPosition, PointStamped& LF_nextPosition,
                                grid_map::Polygon& RF_polygon, grid_map::Polygon& RH_polygon, grid_map::Polygon& LH_polygon,
                                grid_map::Polygon& LF_polygon, ros::Publisher, ros::Publisher, ros::Publisher, ros::Publisher,
                                ros::Publisher, ros::Publisher, ros::Publisher, ros::Publisher) {
    const Point cur[4] = {RF_position.point, RH_position.point, LH_position.point, LF_position.point};
    Point next[4], poly[4][4], c;
    getFootholdSearchGridMap(cur, ajustedPoseY_, next, poly, c);
    PointStamped* nx[4] = {&RF_nextPosition, &RH_nextPosition, &LH_nextPosition, &LF_nextPosition};
    grid_map::Polygon* pg[4] = {&RF_polygon, &RH_polygon, &LH_polygon, &LF_polygon};
    for (int f = 0; f < 4; ++f) {
      nx[f]->point = next[f];
      *pg[f] = grid_map::Polygon();
      for (int v = 0; v < 4; ++v) pg[f]->addVertex(grid_map::Position(poly[f][v].x, poly[f][v].y));
    }
    return true;
  }
  double ajustedPoseY_;  // ajustedPose_[1] (FP.cpp:381, :759, :1578) for the reference-signature getFootholdSearchGridMap

  fhp_ctx* ctx() { return ctx_; }

 private:
  fhp_ctx* ctx_;
  FootholdPlannerGpu(const FootholdPlannerGpu&);
  FootholdPlannerGpu& operator=(const FootholdPlannerGpu&);
  fhp_params with_radii(float footRadius, float searchRadius) const {
    fhp_params p = params;
    p.foot_radius = footRadius;
    p.search_radius = searchRadius;
    return p;
  }
  // getSearchPolygon's vertices for a rectangle centre (FP.cpp:2496-2517; heading extension: rotated offsets, the same
  // expressions as libfhp's derive_params)
  static void rectangle_vertices(const fhp_params& p, double cx, double cy, double v[8]) {
    const double R = p.search_radius, hR = 0.5 * p.search_radius;
    const double ox[4] = {R, R, -R, -R}, oy[4] = {hR, -hR, -hR, hR};
    const double c = std::cos(p.heading), s = std::sin(p.heading);
    for (int k = 0; k < 4; ++k) {
      const double rx = p.heading != 0.0 ? ox[k] * c - oy[k] * s : ox[k];
      const double ry = p.heading != 0.0 ? ox[k] * s + oy[k] * c : oy[k];
      v[2 * k] = cx + rx;
      v[2 * k + 1] = cy + ry;
    }
  }
  // The centre of a rectangle that getSearchPolygon built: libfhp takes search rectangles by their centre.  Any centre
  // whose rectangle reproduces the four vertices bit for bit IS the rectangle (the planner only ever looks at the
  // vertices), so the estimate vertex0 - offset0 and its floating-point neighbours are tried.
  static void rectangle_centre(const grid_map::Polygon& polygon, const fhp_params& p, double rc[2]) {
    if (polygon.nVertices() != 4) throw std::invalid_argument("search polygon: getSearchPolygon builds 4 vertices");
    double off[8];
    rectangle_vertices(p, 0.0, 0.0, off);
    for (int axis = 0; axis < 2; ++axis) {
      double want[4];
      for (int k = 0; k < 4; ++k) want[k] = axis == 0 ? polygon.getVertex(k).x() : polygon.getVertex(k).y();
      const double c0 = want[0] - off[axis];
      const double up1 = std::nextafter(c0, 1e300), dn1 = std::nextafter(c0, -1e300);
      const double cands[5] = {c0, up1, dn1, std::nextafter(up1, 1e300), std::nextafter(dn1, -1e300)};
      bool found = false;
      for (int t = 0; t < 5 && !found; ++t) {
        bool ok = true;
        for (int k = 0; k < 4; ++k) ok = ok && (cands[t] + off[2 * k + axis] == want[k]);
        if (ok) {
          rc[axis] = cands[t];
          found = true;
        }
      }
      if (!found) throw std::invalid_argument("search polygon is not a rectangle built by getSearchPolygon with these radii");
    }
  }
  void check(int rc, const char* what) {
    if (rc == FHP_ERR_MISSING_LAYER) throw std::out_of_range(fhp_last_error(ctx_));  // GridMap::at throws this
    if (rc != FHP_OK) throw std::runtime_error(std::string(what) + ": " + fhp_last_error(ctx_));
  }
  // response assembly (FP.cpp:685-699, :1378-1396) and marker point lists (:2731-2790) in one pass over the rows
  static void fill(GlobalFootholds& msg, const std::vector<fhp_foothold>& rows, const fhp_query_status& st,
                   std::vector<Point> markers[4]) {
    for (int f = 0; f < 4; ++f) markers[f].clear();
    msg.success = st.success != 0 ? 1 : 0;
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
