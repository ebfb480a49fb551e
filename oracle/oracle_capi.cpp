// ORACLE — TEST INFRASTRUCTURE ONLY (see gm_core.hpp header).  PARITY UNPINNED.
// extern "C" surface over the CPU restatement so tests/ and bench.py's cpu_baseline leg can drive it with the
// same POD structs as the product's C-ABI (include/fhp.h is included for the struct layouts only).
#include <atomic>
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <thread>

#include "../include/fhp.h"
#include "fp_oracle.hpp"

namespace {

// test hook: run the heading expressions even for heading == 0 (they must reduce bit-exactly to the reference there)
std::atomic<int> g_force_heading{0};

gm::GridMap make_map(const fhp_map_desc* d, const float* const* layers) {
  gm::GridMap m;
  m.length = gm::Vec2{d->length_x, d->length_y};
  m.position = gm::Vec2{d->pos_x, d->pos_y};
  m.resolution = d->resolution;
  m.size = gm::Idx2{d->rows, d->cols};
  m.startIndex = gm::Idx2{d->start_row, d->start_col};
  for (int i = 0; i < d->n_layers; ++i) m.layers[d->layer_names[i]] = layers[i];
  return m;
}

ref::Planner make_planner(const fhp_params& fp, const fhp_query& q) {
  ref::Planner pl;
  pl.p.footRadius = fp.foot_radius;
  pl.p.searchRadius = fp.search_radius;
  pl.p.stepLength = fp.step_length;
  pl.p.skew = fp.skew;
  pl.p.defaultFootholdThreshold = fp.thr_default;
  pl.p.candidateFootholdThreshold = fp.thr_candidate;
  pl.p.kinLength = fp.kin_length;
  pl.p.kinWidth = fp.kin_width;
  pl.p.kinL1 = fp.kin_l1;
  pl.p.rfFirst = fp.rf_first != 0;
  pl.p.h = fp.h;
  pl.p.driftPerCycle = fp.drift_per_cycle;
  pl.p.spiralMode = fp.spiral_mode == FHP_SPIRAL_FAITHFUL ? ref::SPIRAL_FAITHFUL : ref::SPIRAL_WRAP_CORRECT;
  pl.p.heading = fp.heading;
  pl.p.forceHeadingPath = g_force_heading.load() != 0;
  if (fp.gait == FHP_GAIT_CUSTOM) {
    pl.p.useCustomBias = true;
    for (int f = 0; f < 4; ++f) {
      pl.p.customBias[f][0] = fp.bias[f][0];
      pl.p.customBias[f][1] = fp.bias[f][1];
    }
  } else if (fp.gait == FHP_GAIT_WALK) {
    // extension (no reference code): lateral-sequence walk as a bias table; x phase offsets of
    // RF +3/8, RH +1/8, LH -3/8, LF -1/8 of the step length on the +-L/2, +-W/2 hip rectangle.
    const float lengthBase = fp.kin_length;
    const float widthBase = fp.kin_width + fp.kin_l1 * 2;
    const double sl = fp.step_length;
    pl.p.useCustomBias = true;
    pl.p.customBias[0][0] = 0.5 * lengthBase + 0.375 * sl;   pl.p.customBias[0][1] = -0.5 * widthBase;
    pl.p.customBias[1][0] = -0.5 * lengthBase + 0.125 * sl;  pl.p.customBias[1][1] = -0.5 * widthBase;
    pl.p.customBias[2][0] = -0.5 * lengthBase - 0.375 * sl;  pl.p.customBias[2][1] = 0.5 * widthBase;
    pl.p.customBias[3][0] = 0.5 * lengthBase - 0.125 * sl;   pl.p.customBias[3][1] = 0.5 * widthBase;
  }
  pl.initialPose[0] = q.x;
  pl.initialPose[1] = q.y;
  pl.initialPose[2] = q.z;
  pl.initialize();
  return pl;
}

void to_row(const ref::Foothold& f, fhp_foothold* o) {
  std::memset(o, 0, sizeof(*o));
  o->x = f.point.x;
  o->y = f.point.y;
  o->z = f.point.z;
  o->row = f.row;
  o->col = f.col;
  o->foot_id = f.foot_id;
  o->gait_cycle_id = f.gait_cycle_id;
  o->kind = f.kind;
}

void plan_one(const gm::GridMap& map, const fhp_params* params, const fhp_query& q, int gait_cycles,
              fhp_foothold* out_nom, fhp_foothold* out_cen, fhp_query_status* st, fhp_foothold* out_def = nullptr,
              double* out_kpi = nullptr) {
  ref::Planner pl = make_planner(params[q.param_id], q);
  ref::PlanResult r = pl.globalFootholdPlan(map, static_cast<uint8_t>(gait_cycles));
  for (size_t i = 0; i < r.footholds.size(); ++i) to_row(r.footholds[i], out_nom + i);
  if (out_cen)
    for (size_t i = 0; i < r.centroidFootholds.size(); ++i) to_row(r.centroidFootholds[i], out_cen + i);
  if (out_def)
    for (size_t i = 0; i < r.defaultFootholds.size(); ++i) to_row(r.defaultFootholds[i], out_def + i);
  if (out_kpi)  // per successful cycle: feetDistance nominal[2], cogSpeed nominal[2], feetDistance centroid[2], cogSpeed centroid[2]
    for (size_t c = 0; c * 2 < r.feetDistanceNominal.size(); ++c)
      for (int k = 0; k < 2; ++k) {
        out_kpi[c * 8 + 0 + k] = r.feetDistanceNominal[c * 2 + k];
        out_kpi[c * 8 + 2 + k] = r.cogSpeedNominal[c * 2 + k];
        out_kpi[c * 8 + 4 + k] = r.feetDistanceCentroid[c * 2 + k];
        out_kpi[c * 8 + 6 + k] = r.cogSpeedCentroid[c * 2 + k];
      }
  st->success = r.success ? 1 : 0;
  st->gait_cycles = r.gait_cycles;
  st->gait_cycles_succeed = r.gait_cycles_succeed;
  st->reserved = r.optSubmapWouldFail ? FHP_STATUS_OPT_SUBMAP_FAIL : 0;
  st->n_rows = static_cast<int32_t>(r.footholds.size());
}

}  // namespace

extern "C" {

void oracle_force_heading_path(int32_t on) { g_force_heading.store(on); }

// Whole plans for n_queries (same contract as fhp_plan).  n_threads <= 1: serial.
int oracle_plan(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params, int32_t n_params,
                const fhp_query* queries, int64_t n_queries, int32_t gait_cycles, fhp_foothold* out_nominal,
                fhp_foothold* out_centroid, fhp_query_status* status, int32_t n_threads) {
  if (!desc || !layers || !params || !queries || !out_nominal || !status || gait_cycles < 0 || gait_cycles > 255)
    return FHP_ERR_INVALID_ARG;
  for (int64_t q = 0; q < n_queries; ++q)
    if (queries[q].param_id < 0 || queries[q].param_id >= n_params) return FHP_ERR_INVALID_ARG;
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability") || !map.exists("elevation")) return FHP_ERR_MISSING_LAYER;
  const int64_t R = FHP_ROWS_PER_QUERY(gait_cycles);
  auto work = [&](int64_t q) {
    plan_one(map, params, queries[q], gait_cycles, out_nominal + q * R, out_centroid ? out_centroid + q * R : nullptr,
             status + q);
  };
  if (n_threads <= 1) {
    for (int64_t q = 0; q < n_queries; ++q) work(q);
  } else {
    std::atomic<int64_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
      th.emplace_back([&]() {
        for (;;) {
          int64_t q0 = next.fetch_add(16);
          if (q0 >= n_queries) break;
          for (int64_t q = q0; q < q0 + 16 && q < n_queries; ++q) work(q);
        }
      });
    for (auto& t : th) t.join();
  }
  return FHP_OK;
}

// Whole plans including the "default" chain rows and the KPIs: out_default like out_nominal;
// out_kpi: n_queries * gait_cycles * 8 doubles, entry c of a query = its c-th SUCCESSFUL cycle.
int oracle_plan_ex(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params, int32_t n_params,
                   const fhp_query* queries, int64_t n_queries, int32_t gait_cycles, fhp_foothold* out_nominal,
                   fhp_foothold* out_centroid, fhp_foothold* out_default, double* out_kpi, fhp_query_status* status) {
  if (!desc || !layers || !params || !queries || !out_nominal || !status || gait_cycles < 0 || gait_cycles > 255)
    return FHP_ERR_INVALID_ARG;
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability") || !map.exists("elevation")) return FHP_ERR_MISSING_LAYER;
  const int64_t R = FHP_ROWS_PER_QUERY(gait_cycles);
  for (int64_t q = 0; q < n_queries; ++q) {
    if (queries[q].param_id < 0 || queries[q].param_id >= n_params) return FHP_ERR_INVALID_ARG;
    plan_one(map, params, queries[q], gait_cycles, out_nominal + q * R, out_centroid ? out_centroid + q * R : nullptr,
             status + q, out_default ? out_default + q * R : nullptr,
             out_kpi ? out_kpi + q * static_cast<int64_t>(gait_cycles) * 8 : nullptr);
  }
  return FHP_OK;
}

// Latency of ONE plan_global_footholds call: `repeats` timed calls of the same query, milliseconds
// each into out_ms.  leg_threads != 0: the reference's threading (four std::threads per cycle, FP.cpp:863-909).
int oracle_plan_latency(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params,
                        const fhp_query* query, int32_t gait_cycles, int32_t repeats, int32_t leg_threads, double* out_ms,
                        fhp_query_status* status) {
  if (!desc || !layers || !params || !query || !out_ms || gait_cycles < 0 || gait_cycles > 255) return FHP_ERR_INVALID_ARG;
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability") || !map.exists("elevation")) return FHP_ERR_MISSING_LAYER;
  for (int r = 0; r < repeats; ++r) {
    const auto t0 = std::chrono::steady_clock::now();
    ref::Planner pl = make_planner(params[0], *query);
    pl.legThreads = leg_threads != 0;
    ref::PlanResult res = pl.globalFootholdPlan(map, static_cast<uint8_t>(gait_cycles));
    const auto t1 = std::chrono::steady_clock::now();
    out_ms[r] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (status) {
      status->success = res.success ? 1 : 0;
      status->gait_cycles = res.gait_cycles;
      status->gait_cycles_succeed = res.gait_cycles_succeed;
      status->reserved = res.optSubmapWouldFail ? FHP_STATUS_OPT_SUBMAP_FAIL : 0;
      status->n_rows = static_cast<int32_t>(res.footholds.size());
    }
  }
  return FHP_OK;
}

// Algorithmic-bytes instrumentation: per query, the number of unique map-cell BYTES (any layer)
// the reference algorithm reads per (cycle, leg) foothold, summed over the query's cycles*4 footholds, and
// the number of footholds whose spiral search ran.
int oracle_plan_touch_bytes(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params,
                            int32_t n_params, const fhp_query* queries, int64_t n_queries, int32_t gait_cycles,
                            int64_t* unique_cell_bytes_per_query, int64_t* spiral_taken_per_query,
                            int32_t with_centroid_height) {
  if (!desc || !layers || !params || !queries || !unique_cell_bytes_per_query) return FHP_ERR_INVALID_ARG;
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability") || !map.exists("elevation")) return FHP_ERR_MISSING_LAYER;
  for (int64_t q = 0; q < n_queries; ++q) {
    if (queries[q].param_id < 0 || queries[q].param_id >= n_params) return FHP_ERR_INVALID_ARG;
    ref::Planner pl = make_planner(params[queries[q].param_id], queries[q]);
    pl.countTouched = true;
    pl.centroidHeight = with_centroid_height != 0;
    pl.globalFootholdPlan(map, static_cast<uint8_t>(gait_cycles));
    unique_cell_bytes_per_query[q] = pl.touchedCells * 4;
    if (spiral_taken_per_query) spiral_taken_per_query[q] = pl.spiralTaken;
  }
  return FHP_OK;
}

// Spiral statistics for kernel design: hist[0..255] = searches that ended on their k-th visited cell, hist[256] =
// searches that found no cell (257 entries, accumulated over all queries).
int oracle_spiral_rank_hist(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params,
                            int32_t n_params, const fhp_query* queries, int64_t n_queries, int32_t gait_cycles,
                            int64_t* hist257) {
  if (!desc || !layers || !params || !queries || !hist257) return FHP_ERR_INVALID_ARG;
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability") || !map.exists("elevation")) return FHP_ERR_MISSING_LAYER;
  for (int64_t q = 0; q < n_queries; ++q) {
    if (queries[q].param_id < 0 || queries[q].param_id >= n_params) return FHP_ERR_INVALID_ARG;
    ref::Planner pl = make_planner(params[queries[q].param_id], queries[q]);
    pl.rankHist = hist257;
    pl.globalFootholdPlan(map, static_cast<uint8_t>(gait_cycles));
  }
  return FHP_OK;
}

// ---- fine-grained members (same contracts as the fhp_check_* entry points) ---------------------------------
int oracle_check_foothold(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params,
                          const double* centres_xy, const double* rect_centres_xy, int64_t n, fhp_foothold* out,
                          uint8_t* valid_out) {
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability") || !map.exists("elevation")) return FHP_ERR_MISSING_LAYER;
  fhp_query q0{0, 0, 0, 0, 0};
  ref::Planner pl = make_planner(*params, q0);
  for (int64_t i = 0; i < n; ++i) {
    gm::Vec2 c{centres_xy[2 * i], centres_xy[2 * i + 1]};
    ref::Point rc;
    rc.x = rect_centres_xy[2 * i];
    rc.y = rect_centres_xy[2 * i + 1];
    gm::Polygon poly = pl.searchPolygon(rc);
    ref::Foothold fh;
    bool ok = false;
    pl.checkFoothold(map, c, pl.p.footRadius, pl.p.searchRadius, poly, fh.point, ok, &fh.kind);
    ref::Planner::fillIndex(map, fh);
    to_row(fh, out + i);
    valid_out[i] = ok ? 1 : 0;
  }
  return FHP_OK;
}

int oracle_check_centroid(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params,
                          const double* points_xy, int64_t n, fhp_foothold* out, int8_t* case_out) {
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability") || !map.exists("elevation")) return FHP_ERR_MISSING_LAYER;
  fhp_query q0{0, 0, 0, 0, 0};
  ref::Planner pl = make_planner(*params, q0);
  for (int64_t i = 0; i < n; ++i) {
    ref::Point pt;
    pt.x = points_xy[2 * i];
    pt.y = points_xy[2 * i + 1];
    ref::Foothold fh;
    fh.kind = ref::KIND_DEFAULT;
    int c = -1;
    pl.checkFootholdUseCentroidMethod(map, pt, fh.point, &c);
    ref::Planner::fillIndex(map, fh);
    to_row(fh, out + i);
    case_out[i] = static_cast<int8_t>(c);
  }
  return FHP_OK;
}

int oracle_mean_height(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params,
                       const double* centres_xy, int64_t n, float* z_out) {
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("elevation")) return FHP_ERR_MISSING_LAYER;
  fhp_query q0{0, 0, 0, 0, 0};
  ref::Planner pl = make_planner(*params, q0);
  for (int64_t i = 0; i < n; ++i)
    z_out[i] = pl.getFootholdMeanHeight(map, gm::Vec2{centres_xy[2 * i], centres_xy[2 * i + 1]}, pl.p.footRadius,
                                        pl.p.h);
  return FHP_OK;
}

int oracle_check_default(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params,
                         const double* centres_xy, int64_t n, uint8_t* valid_out) {
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability")) return FHP_ERR_MISSING_LAYER;
  fhp_query q0{0, 0, 0, 0, 0};
  ref::Planner pl = make_planner(*params, q0);
  for (int64_t i = 0; i < n; ++i)
    valid_out[i] =
        pl.checkDefaultFoothold(map, gm::Vec2{centres_xy[2 * i], centres_xy[2 * i + 1]}, pl.p.footRadius) ? 1 : 0;
  return FHP_OK;
}

// checkCandidateFoothold (FP.cpp:2085-2114): the spiral search alone.  out = chosen cell centre (xy; z = 0) when valid.
int oracle_check_candidate(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params,
                           const double* centres_xy, const double* rect_centres_xy, int64_t n, fhp_foothold* out,
                           uint8_t* valid_out) {
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability")) return FHP_ERR_MISSING_LAYER;
  fhp_query q0{0, 0, 0, 0, 0};
  ref::Planner pl = make_planner(*params, q0);
  for (int64_t i = 0; i < n; ++i) {
    gm::Vec2 c{centres_xy[2 * i], centres_xy[2 * i + 1]};
    ref::Point rc;
    rc.x = rect_centres_xy[2 * i];
    rc.y = rect_centres_xy[2 * i + 1];
    gm::Polygon poly = pl.searchPolygon(rc);
    ref::Foothold fh;
    fh.kind = ref::KIND_CANDIDATE;
    const bool ok = pl.checkCandidateFoothold(map, c, pl.p.footRadius, pl.p.searchRadius, poly, fh.point);
    if (ok) ref::Planner::fillIndex(map, fh); else fh.row = fh.col = -1;
    to_row(fh, out + i);
    valid_out[i] = ok ? 1 : 0;
  }
  return FHP_OK;
}

// checkCirclePolygonFoothold (FP.cpp:2117-2163) at arbitrary foot centres against the rectangle around rect centres.
int oracle_check_circle_polygon(const fhp_map_desc* desc, const float* const* layers, const fhp_params* params,
                                const double* foot_centres_xy, const double* rect_centres_xy, int64_t n,
                                uint8_t* valid_out) {
  gm::GridMap map = make_map(desc, layers);
  if (!map.exists("traversability")) return FHP_ERR_MISSING_LAYER;
  fhp_query q0{0, 0, 0, 0, 0};
  ref::Planner pl = make_planner(*params, q0);
  for (int64_t i = 0; i < n; ++i) {
    ref::Point rc;
    rc.x = rect_centres_xy[2 * i];
    rc.y = rect_centres_xy[2 * i + 1];
    gm::Polygon poly = pl.searchPolygon(rc);
    valid_out[i] = pl.checkCirclePolygonFoothold(map, gm::Vec2{foot_centres_xy[2 * i], foot_centres_xy[2 * i + 1]},
                                                 pl.p.footRadius, poly) ? 1 : 0;
  }
  return FHP_OK;
}

// getPolygonCenter (FP.cpp:2421-2463), getDefaultFootholdNext via getFootholdSearchGridMap (:2166-2250) and
// getSearchPolygon (:2496-2517) for one set of four current feet: out_center_xy[2], out_next_xy[4][2],
// out_poly_xy[4][4][2] (vertex order of the reference).
int oracle_search_geometry(const fhp_params* params, const fhp_query* q, double adjusted_y, const double* cur_xyz,
                           double* out_center_xy, double* out_next_xy, double* out_poly_xy) {
  ref::Planner pl = make_planner(*params, *q);
  pl.ajustedPoseY = adjusted_y;
  ref::Point cur[4], next[4];
  for (int f = 0; f < 4; ++f) {
    cur[f].x = cur_xyz[3 * f];
    cur[f].y = cur_xyz[3 * f + 1];
    cur[f].z = cur_xyz[3 * f + 2];
  }
  ref::Point c = pl.getPolygonCenter(cur[0], cur[1], cur[2], cur[3]);
  out_center_xy[0] = c.x;
  out_center_xy[1] = c.y;
  gm::Polygon poly[4];
  pl.getFootholdSearchGridMap(cur, next, poly);
  for (int f = 0; f < 4; ++f) {
    out_next_xy[2 * f] = next[f].x;
    out_next_xy[2 * f + 1] = next[f].y;
    for (int v = 0; v < 4; ++v) {
      out_poly_xy[(f * 4 + v) * 2] = poly[f].vertices[v].x;
      out_poly_xy[(f * 4 + v) * 2 + 1] = poly[f].vertices[v].y;
    }
  }
  return FHP_OK;
}

// ---- grid_map_core probes (unit tests of the restatement itself) -------------------------------------------
// Buffer indices visited by SpiralIterator(map, centre, radius), in order.  Returns the count (<= max_out).
int64_t oracle_spiral_cells(const fhp_map_desc* desc, double cx, double cy, double radius, int32_t wrap_correct,
                            int32_t* out_idx, int64_t max_out) {
  static const float dummy = 0.f;
  const float* l[1] = {&dummy};
  fhp_map_desc d = *desc;
  d.n_layers = 0;
  gm::GridMap map = make_map(&d, l);
  int64_t n = 0;
  for (gm::SpiralIterator it(map, gm::Vec2{cx, cy}, radius, wrap_correct != 0); !it.isPastEnd(); ++it) {
    if (n < max_out) {
      gm::Idx2 i = *it;
      out_idx[2 * n] = i.i0;
      out_idx[2 * n + 1] = i.i1;
    }
    ++n;
  }
  return n;
}
// Buffer indices visited by CircleIterator(map, centre, radius), in order.
int64_t oracle_circle_cells(const fhp_map_desc* desc, double cx, double cy, double radius, int32_t* out_idx,
                            int64_t max_out) {
  static const float dummy = 0.f;
  const float* l[1] = {&dummy};
  fhp_map_desc d = *desc;
  d.n_layers = 0;
  gm::GridMap map = make_map(&d, l);
  int64_t n = 0;
  for (gm::CircleIterator it(map, gm::Vec2{cx, cy}, radius); !it.isPastEnd(); ++it) {
    if (n < max_out) {
      gm::Idx2 i = *it;
      out_idx[2 * n] = i.i0;
      out_idx[2 * n + 1] = i.i1;
    }
    ++n;
  }
  return n;
}
// getSubmapInformation: out = {tl_row, tl_col, size_rows, size_cols}; pos_len = {pos_x,pos_y,len_x,len_y}
int oracle_submap_info(const fhp_map_desc* desc, double px, double py, double lx, double ly, int32_t* out,
                       double* pos_len) {
  gm::SubmapInfo info;
  bool ok = gm::getSubmapInformation(info, gm::Vec2{px, py}, gm::Vec2{lx, ly}, gm::Vec2{desc->length_x, desc->length_y},
                                     gm::Vec2{desc->pos_x, desc->pos_y}, desc->resolution,
                                     gm::Idx2{desc->rows, desc->cols}, gm::Idx2{desc->start_row, desc->start_col});
  out[0] = info.topLeftIndex.i0;
  out[1] = info.topLeftIndex.i1;
  out[2] = info.size.i0;
  out[3] = info.size.i1;
  pos_len[0] = info.position.x;
  pos_len[1] = info.position.y;
  pos_len[2] = info.length.x;
  pos_len[3] = info.length.y;
  return ok ? 1 : 0;
}
int oracle_get_index(const fhp_map_desc* desc, double px, double py, int32_t* out) {
  gm::Idx2 idx;
  bool ok = gm::getIndexFromPosition(idx, gm::Vec2{px, py}, gm::Vec2{desc->length_x, desc->length_y},
                                     gm::Vec2{desc->pos_x, desc->pos_y}, desc->resolution,
                                     gm::Idx2{desc->rows, desc->cols}, gm::Idx2{desc->start_row, desc->start_col});
  out[0] = idx.i0;
  out[1] = idx.i1;
  return ok ? 1 : 0;
}
int oracle_get_position(const fhp_map_desc* desc, int32_t i0, int32_t i1, double* out) {
  gm::Vec2 p{0, 0};
  bool ok = gm::getPositionFromIndex(p, gm::Idx2{i0, i1}, gm::Vec2{desc->length_x, desc->length_y},
                                     gm::Vec2{desc->pos_x, desc->pos_y}, desc->resolution,
                                     gm::Idx2{desc->rows, desc->cols}, gm::Idx2{desc->start_row, desc->start_col});
  out[0] = p.x;
  out[1] = p.y;
  return ok ? 1 : 0;
}
int oracle_polygon_inside(const double* verts_xy, int32_t n_verts, double px, double py) {
  gm::Polygon poly;
  for (int i = 0; i < n_verts; ++i) poly.addVertex(gm::Vec2{verts_xy[2 * i], verts_xy[2 * i + 1]});
  return poly.isInside(gm::Vec2{px, py}) ? 1 : 0;
}
// Stance, bias table and first-gait positions for a parameter set (rows RF,RH,LH,LF; cols x,y).
void oracle_initial_geometry(const fhp_params* params, const fhp_query* q, double* initial_xy, double* bias_xy,
                             double* first_gait_xy) {
  ref::Planner pl = make_planner(*params, *q);
  ref::Point cur[4];
  pl.setFirstGait(pl.initial, cur);
  for (int f = 0; f < 4; ++f) {
    initial_xy[2 * f] = pl.initial[f].x;
    initial_xy[2 * f + 1] = pl.initial[f].y;
    bias_xy[2 * f] = pl.bias[f].x;
    bias_xy[2 * f + 1] = pl.bias[f].y;
    first_gait_xy[2 * f] = cur[f].x;
    first_gait_xy[2 * f + 1] = cur[f].y;
  }
}

}  // extern "C"
