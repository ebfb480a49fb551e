// ORACLE / TEST INFRASTRUCTURE ONLY.
// Drives the reference's OWN FootholdPlanner.cpp (compiled in place from /root/reference, see Makefile) through its
// real class interface: construct with rosparams -> gridmapCallback(map) -> globalFootholdPlan(request, response).
// Used by tests/test_reference_build.py to check the CPU restatement (oracle/fp_oracle.hpp) against the reference
// code itself.  What is NOT the reference here: ROS / message plumbing (no-op shims), NLopt (stub: the opt chain is
// outside the response path) and grid_map_core (restated in oracle/gm_core.hpp — third party, not vendored).
#include <fcntl.h>
#include <unistd.h>

#include <cstring>

// everything FootholdPlanner.hpp includes, pulled in BEFORE the access hack so only the reference class is affected
#include <pwd.h>
#include <time.h>

#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include <geometry_msgs/PointStamped.h>
#include <geometry_msgs/PolygonStamped.h>
#include <geometry_msgs/Vector3.h>
#include <grid_map_core/GridMap.hpp>
#include <grid_map_ros/GridMapRosConverter.hpp>
#include <grid_map_ros/grid_map_ros.hpp>
#include <nav_msgs/Path.h>
#include <ros/ros.h>
#include <std_srvs/Empty.h>
#include <visualization_msgs/Marker.h>
#include <visualization_msgs/MarkerArray.h>

#include "foothold_planner/GlobalFootholdPlan.h"
#include "foothold_planner_msgs/GlobalFootholds.h"

#define private public  // reach gridmapCallback / members of the reference class from the test driver
#define protected public
#include "foothold_planner/FootholdPlanner.hpp"
#undef private
#undef protected

#include "../../include/fhp.h"

namespace {
struct Quiet {  // the reference prints hundreds of lines per cycle
  int saved;
  Quiet() {
    fflush(stdout);
    saved = dup(1);
    int devnull = open("/dev/null", O_WRONLY);
    dup2(devnull, 1);
    close(devnull);
  }
  ~Quiet() {
    fflush(stdout);
    std::cout.flush();
    dup2(saved, 1);
    close(saved);
  }
};
}  // namespace

extern "C" {

// One plan_global_footholds call per query through the unmodified reference class.
//   wrap_correct_spiral: 0 = upstream SpiralIterator (ignores the start index), 1 = repaired variant.
//   out_nominal / out_centroid: FHP_ROWS_PER_QUERY(gait_cycles) rows per query (x,y,z,foot_id,gait_cycle_id filled;
//   row/col/kind are not part of the reference's response and are left 0).
//   status[q].reserved = 1 when the service itself returned false (opt-chain submap failure, FP.cpp:931-934).
int ref_plan(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params, const fhp_query* queries,
             int64_t n_queries, int32_t gait_cycles, int32_t wrap_correct_spiral, fhp_foothold* out_nominal,
             fhp_foothold* out_centroid, fhp_query_status* status, fhp_foothold* out_default, double* out_kpi) {
  if (!desc || !layers || !params || !queries || !out_nominal || !status) return FHP_ERR_INVALID_ARG;
  Quiet quiet;
  grid_map::SpiralIterator::wrapCorrect() = wrap_correct_spiral != 0;
  grid_map::GridMap map;
  map.setRaw(desc->rows, desc->cols, desc->resolution, desc->length_x, desc->length_y, desc->pos_x, desc->pos_y,
             desc->start_row, desc->start_col);
  map.setFrameId("odom");
  for (int i = 0; i < desc->n_layers; ++i) map.add(desc->layer_names[i], layers[i]);
  const int64_t R = FHP_ROWS_PER_QUERY(gait_cycles);
  for (int64_t q = 0; q < n_queries; ++q) {
    const fhp_params& p = params[queries[q].param_id];
    std::map<std::string, double>& num = ros::ParamTable::get().num;
    num.clear();
    ros::ParamTable::get().str.clear();
    num["debug"] = 0;
    num["debug2"] = 0;
    num["debug3"] = 0;
    num["checkDefaultFoothold_Debug"] = 0;
    num["footRadius"] = p.foot_radius;
    num["defaultFootholdThreshold"] = p.thr_default;
    num["candidateFootholdThreshold"] = p.thr_candidate;
    num["searchRadius"] = p.search_radius;
    num["stepLength"] = p.step_length;
    num["RF_FIRST"] = p.rf_first;
    num["laikago_kinematics/length"] = p.kin_length;
    num["laikago_kinematics/width"] = p.kin_width;
    num["laikago_kinematics/l1"] = p.kin_l1;
    num["laikago_kinematics/skewLength"] = p.skew;
    num["initial_position/x"] = queries[q].x;
    num["initial_position/y"] = queries[q].y;
    num["initial_position/z"] = queries[q].z;
    num["nlopt/manulStepSize"] = 0;
    num["nlopt/useInequalityConstraits"] = 0;
    ros::NodeHandle nh("~");
    foothold_planner::FootholdPlanner planner(nh);
    grid_map_msgs::GridMap msg;
    msg.decoded = &map;
    planner.gridmapCallback(msg);
    foothold_planner::GlobalFootholdPlan::Request req;
    foothold_planner::GlobalFootholdPlan::Response res;
    req.gait_cycles = static_cast<uint8_t>(gait_cycles);
    const bool ok = planner.globalFootholdPlan(req, res);
    fhp_query_status& st = status[q];
    std::memset(&st, 0, sizeof(st));
    const foothold_planner_msgs::GlobalFootholds& nom = ok ? res.footholds : planner.nominalGlobalFootholdsMsg_;
    st.success = nom.success ? 1 : 0;
    st.gait_cycles = nom.gait_cycles;
    st.gait_cycles_succeed = nom.gait_cycles_succeed;
    st.reserved = ok ? 0 : 1;
    st.n_rows = static_cast<int32_t>(nom.footholds.size());
    for (int64_t i = 0; i < R; ++i) std::memset(out_nominal + q * R + i, 0, sizeof(fhp_foothold));
    for (size_t i = 0; i < nom.footholds.size() && static_cast<int64_t>(i) < R; ++i) {
      fhp_foothold& o = out_nominal[q * R + i];
      o.x = nom.footholds[i].point.x;
      o.y = nom.footholds[i].point.y;
      o.z = nom.footholds[i].point.z;
      o.foot_id = nom.footholds[i].foot_id;
      o.gait_cycle_id = nom.footholds[i].gait_cycle_id;
    }
    if (out_default) {  // globalFootholdsResult_.defaultFootholds: one VectorXd(12) = RF,RH,LH,LF xyz per recorded cycle
      for (int64_t i = 0; i < R; ++i) std::memset(out_default + q * R + i, 0, sizeof(fhp_foothold));
      const std::vector<Eigen::VectorXd>& d = planner.globalFootholdsResult_.defaultFootholds;
      for (size_t c = 0; c < d.size() && static_cast<int64_t>(c * 4 + 3) < R; ++c)
        for (int f = 0; f < 4; ++f) {
          fhp_foothold& o = out_default[q * R + c * 4 + f];
          o.x = d[c][3 * f];
          o.y = d[c][3 * f + 1];
          o.z = d[c][3 * f + 2];
          o.foot_id = static_cast<uint8_t>(f);
        }
    }
    if (out_kpi) {  // footholdsKPI_: two values per successful cycle each
      const auto& k = planner.footholdsKPI_;
      double* o = out_kpi + q * static_cast<int64_t>(gait_cycles) * 8;
      for (size_t c = 0; c * 2 + 1 < k.feetDistance_nominal.size() && static_cast<int>(c) < gait_cycles; ++c)
        for (int j = 0; j < 2; ++j) {
          o[c * 8 + 0 + j] = k.feetDistance_nominal[c * 2 + j];
          o[c * 8 + 2 + j] = k.cogSpeed_nominal[c * 2 + j];
          o[c * 8 + 4 + j] = k.feetDistance_centroid[c * 2 + j];
          o[c * 8 + 6 + j] = k.cogSpeed_centroid[c * 2 + j];
        }
    }
    if (out_centroid) {
      // the centroid message is never cleared between service calls (FP.cpp:591 clears only the nominal one), but
      // this planner object served exactly one call
      const foothold_planner_msgs::GlobalFootholds& cen = planner.centroidGlobalFootholdsMsg_;
      for (int64_t i = 0; i < R; ++i) std::memset(out_centroid + q * R + i, 0, sizeof(fhp_foothold));
      for (size_t i = 0; i < cen.footholds.size() && static_cast<int64_t>(i) < R; ++i) {
        fhp_foothold& o = out_centroid[q * R + i];
        o.x = cen.footholds[i].point.x;
        o.y = cen.footholds[i].point.y;
        o.z = cen.footholds[i].point.z;
        o.foot_id = cen.footholds[i].foot_id;
        o.gait_cycle_id = cen.footholds[i].gait_cycle_id;
      }
    }
  }
  return FHP_OK;
}

}  // extern "C"
