// ORACLE — TEST INFRASTRUCTURE ONLY (see gm_core.hpp header).
// PARITY: the reference has no tests or fixtures of its own.  This restatement is pinned against the reference's
// OWN FootholdPlanner.cpp compiled here (oracle/ref_shim -> oracle/_ref, tests/test_reference_build.py,
// tests/golden/ref_*.npz): bit-exact responses of both chains on every completed service call.  What stays
// unpinned is the grid_map_core layer underneath both (gm_core.hpp: third party, not vendored, restated).
//
// CPU restatement of the hot path of lukechencqu/quadrupedal_foothold_planner,
// foothold_planner/src/FootholdPlanner.cpp ("FP.cpp"), with the reference's float/double member types
// (foothold_planner/include/foothold_planner/FootholdPlanner.hpp:609-697).  Each function cites the FP.cpp
// lines it follows.  Maps are passed by const-ref (the reference passes by value; semantically identical).
//
// Policies for undefined / quirky reference behaviour:
//  C.1 the centroid row scan reads one column past the window (FP.cpp:1720,1733-1736) -> that phantom cell is
//      treated as "not below threshold";
//  C.2 checkFootholdUseCentroidMethod has no return on its main paths (FP.cpp:1997) -> returns true; when no
//      case matches or the submap fails the result stays (0,0,0) (FP.cpp:810-813);
//  C.3 SpiralIterator ignores the circular-buffer start index -> two modes, `faithful` and `wrap_correct`;
//  the opt (NLopt) chain, the "default" chain, KPIs, markers and logs are out of scope.
#pragma once
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "gm_core.hpp"

namespace ref {

struct Point {  // geometry_msgs::Point (default-constructed = 0,0,0)
  double x = 0, y = 0, z = 0;
};

enum Kind : uint8_t { KIND_STANCE = 0, KIND_DEFAULT = 1, KIND_CANDIDATE = 2 };
enum SpiralMode : int { SPIRAL_WRAP_CORRECT = 0, SPIRAL_FAITHFUL = 1 };

// rosparam-loaded members (FP.cpp:249-311) with the reference's C++ types.
struct Params {
  float footRadius = 0.02f;                  // FP.hpp footRadius_ (float)
  float searchRadius = 0.1f;                 // searchRadius_ (float)
  float stepLength = 0.18f;                  // stepLength_ (float)
  float defaultFootholdThreshold = 0.9f;     // float
  float candidateFootholdThreshold = 0.7f;   // float
  float kinLength = 0.4387f, kinWidth = 0.175f, kinL1 = 0.037f;  // LaikagoKinematics (float)
  float skew = 0.04f;                        // isos_.skew (float)
  bool rfFirst = false;                      // RF_FIRST_
  double h = 0.01;                           // h_ (FP.cpp:336)
  double driftPerCycle = -0.007;             // FP.cpp:1578
  // extension (not in the reference): explicit per-leg bias table replaces the trot trapezoid table
  bool useCustomBias = false;
  double customBias[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  int spiralMode = SPIRAL_WRAP_CORRECT;
  // extension (not in the reference, which always advances along +x, FP.cpp:2199-2202): heading of the body in the
  // map frame (radians).  heading == 0 runs the restated reference code unchanged; any other value (or
  // forceHeadingPath, which the tests use to prove the general expressions reduce BIT-EXACTLY to the reference at
  // theta = 0) runs the rotated expressions marked "heading" below.  Definition:
  //   u = (cos, sin) forward, n = (-sin, cos) lateral;  rot(v) = (v.x*cos - v.y*sin, v.x*sin + v.y*cos)
  //   stance, bias table and the four rectangle vertex offsets are rot() of the reference's values;
  //   setFirstGait: cur = init - (stepLength/2) u;
  //   next centre = ((c.u) + stepLength) u + ((p0.n) + adjY) n  (terms with an exactly zero sin are dropped, see
  //   headingNextCenter) — the forward coordinate comes from the stance
  //   centroid c, the lateral one is pinned to the initial pose p0 plus the drift, exactly as FP.cpp:2199-2201 does
  //   for u = +x;  the centroid method keeps its axis-aligned getSubmap window (grid_map cuts no rotated submaps).
  double heading = 0.0;
  bool forceHeadingPath = false;
};

struct Foothold {  // foothold_planner_msgs/Foothold + diagnostics
  Point point;
  uint8_t foot_id = 0;        // RF=0, RH=1, LH=2, LF=3 (FP.cpp:685-699)
  uint8_t gait_cycle_id = 0;
  uint8_t kind = KIND_STANCE;
  int row = -1, col = -1;     // buffer index of the cell containing point.xy (-1,-1 if outside the map)
};

struct PlanResult {  // foothold_planner_msgs/GlobalFootholds
  bool success = false;
  uint8_t gait_cycles = 0;
  uint8_t gait_cycles_succeed = 0;
  std::vector<Foothold> footholds;           // nominal chain = the service response (FP.cpp:1588)
  std::vector<Foothold> centroidFootholds;   // centroid chain rows (published on global_footholds_centroid)
  std::vector<Foothold> defaultFootholds;    // "default" chain rows (globalFootholdsResult_.defaultFootholds, FP.cpp:1344-1348)
  // KPIs per successful cycle, two values each (FP.cpp:1357-1368, :1422-1433): getHipDistance / getCogSpeed
  std::vector<double> feetDistanceNominal, cogSpeedNominal, feetDistanceCentroid, cogSpeedCentroid;
  // at gait cycle 0 the opt (NLopt) chain still holds the setFirstGait state (FP.cpp:583-588), so its
  // gait-cycle submap request (getGaitCycleSearchGridMap FP.cpp:2307-2350 with rect = isos_.length x isos_.width,
  // FP.cpp:391-394) is a pure function of the query; when that getSubmap fails the reference aborts the whole service
  // call (`return false`, FP.cpp:931-934) before anything of cycle 0 is appended.  Later cycles depend on NLopt output.
  bool optSubmapWouldFail = false;
};

class Planner {
 public:
  Params p;
  double initialPose[3] = {0, 0, 0};  // initialPose_ (FP.cpp:293-295)

  // derived in initialize() (FP.cpp:317-421)
  float lengthBase = 0, widthBase = 0;
  Point initial[4];  // RF, RH, LH, LF initial stance
  Point bias[4];     // RF, RH, LH, LF default bias (trapezoid vertices relative to the stance centroid)
  double ajustedPoseY = 0;
  // heading extension: derived in initialize()
  bool useHeading = false;
  double ch = 1.0, sh = 0.0;   // cos / sin of the heading
  gm::Vec2 rectOff[4];         // rot() of the rectangle's vertex offsets, vertex order of getSearchPolygon
  // instrumentation: when countTouched is set, globalFootholdPlan accumulates the number of UNIQUE
  // map cells (any layer) read per (cycle, leg) foothold = centroid task + nominal task of that leg.
  bool countTouched = false;
  // response-only variant: skip the centroid chain's height lookups (their z never reaches the
  // plan_global_footholds response, FP.cpp:1588) — used only to count the algorithmic bytes of that variant.
  bool centroidHeight = true;
  // timing variant: run the four nominal-chain leg searches of a cycle on four std::threads and
  // join them, as the reference does (FP.cpp:863-909).  Same results; never combined with countTouched.
  bool legThreads = false;
  int64_t touchedCells = 0;
  int64_t spiralTaken = 0;  // footholds whose default-disc test failed (spiral search executed)
  // instrumentation for kernel design (scratch statistics, never part of a result): when set, rankHist[k] counts
  // spiral searches that ended on their k-th visited cell (k capped at 255) and rankHist[256] those that found none
  mutable int64_t* rankHist = nullptr;

  void initialize() {
    // FP.cpp:340-341 (float arithmetic)
    lengthBase = p.kinLength;
    widthBase = p.kinWidth + p.kinL1 * 2;
    // FP.cpp:350-378
    initial[0].x = lengthBase * 0.5;   initial[0].y = -widthBase * 0.5;  initial[0].z = 0;   // RF
    initial[1].x = -lengthBase * 0.5;  initial[1].y = -widthBase * 0.5;  initial[1].z = 0;   // RH
    initial[2].x = -lengthBase * 0.5;  initial[2].y = widthBase * 0.5;   initial[2].z = 0;   // LH
    initial[3].x = lengthBase * 0.5;   initial[3].y = widthBase * 0.5;   initial[3].z = 0;   // LF
    for (int f = 0; f < 4; ++f) {
      initial[f].x += initialPose[0];
      initial[f].y += initialPose[1];
      initial[f].z += initialPose[2];
    }
    ajustedPoseY = 0.0;  // FP.cpp:381
    // FP.cpp:403-421
    if (p.rfFirst) {
      bias[0].x = 0.5 * lengthBase + p.skew;   bias[0].y = -0.5 * widthBase;
      bias[1].x = -0.5 * lengthBase - p.skew;  bias[1].y = -0.5 * widthBase;
      bias[2].x = -0.5 * lengthBase + p.skew;  bias[2].y = 0.5 * widthBase;
      bias[3].x = 0.5 * lengthBase - p.skew;   bias[3].y = 0.5 * widthBase;
    } else {
      bias[0].x = 0.5 * lengthBase - p.skew;   bias[0].y = -0.5 * widthBase;
      bias[1].x = -0.5 * lengthBase + p.skew;  bias[1].y = -0.5 * widthBase;
      bias[2].x = -0.5 * lengthBase - p.skew;  bias[2].y = 0.5 * widthBase;
      bias[3].x = 0.5 * lengthBase + p.skew;   bias[3].y = 0.5 * widthBase;
    }
    if (p.useCustomBias)
      for (int f = 0; f < 4; ++f) {
        bias[f].x = p.customBias[f][0];
        bias[f].y = p.customBias[f][1];
      }
    // ---- heading extension: rotate stance, bias table and rectangle vertex offsets
    useHeading = (p.heading != 0.0) || p.forceHeadingPath;
    ch = std::cos(p.heading);
    sh = std::sin(p.heading);
    if (useHeading) {
      const double ox[4] = {lengthBase * 0.5, -lengthBase * 0.5, -lengthBase * 0.5, lengthBase * 0.5};
      const double oy[4] = {-widthBase * 0.5, -widthBase * 0.5, widthBase * 0.5, widthBase * 0.5};
      for (int f = 0; f < 4; ++f) {
        initial[f].x = (ox[f] * ch - oy[f] * sh) + initialPose[0];
        initial[f].y = (ox[f] * sh + oy[f] * ch) + initialPose[1];
        initial[f].z = 0.0 + initialPose[2];
        const double bx = bias[f].x, by = bias[f].y;
        bias[f].x = bx * ch - by * sh;
        bias[f].y = bx * sh + by * ch;
      }
      const double R = p.searchRadius, hR = 0.5 * p.searchRadius;  // float promoted (FP.cpp:2502-2511)
      const double vx[4] = {R, R, -R, -R}, vy[4] = {hR, -hR, -hR, hR};
      for (int k = 0; k < 4; ++k) {
        rectOff[k].x = vx[k] * ch - vy[k] * sh;
        rectOff[k].y = vx[k] * sh + vy[k] * ch;
      }
    }
  }

  // FP.cpp:2679-2699
  void setFirstGait(const Point init[4], Point cur[4]) const {
    for (int f = 0; f < 4; ++f) {
      cur[f] = init[f];
      if (useHeading) {  // heading: half a step back along u
        const double hs = p.stepLength / 2;
        cur[f].x = init[f].x - hs * ch;
        cur[f].y = init[f].y - hs * sh;
        continue;
      }
      cur[f].x = init[f].x - p.stepLength / 2;  // float/int -> float, promoted
    }
  }

  static double mulz(double a, double t) { return t == 0.0 ? 0.0 : a * t; }
  // heading: the next stance centre from the current stance centroid (see Params::heading)
  Point headingNextCenter(const Point& polygonCenter) const {
    Point n;
    // mulz: a term whose trigonometric factor is exactly zero is dropped (not multiplied), so theta = 0 reproduces
    // FP.cpp:2199-2201 even when the stance centroid is not finite (a degenerate stance polygon divides by a zero area)
    const double fwd = (polygonCenter.x * ch + mulz(polygonCenter.y, sh)) + p.stepLength;
    const double lat = (initialPose[1] * ch - mulz(initialPose[0], sh)) + ajustedPoseY;
    n.x = fwd * ch - mulz(lat, sh);
    n.y = mulz(fwd, sh) + lat * ch;
    n.z = polygonCenter.z;
    return n;
  }
  // heading: getSearchPolygon with rotated vertex offsets (same vertex order)
  gm::Polygon headingSearchPolygon(const Point& center) const {
    gm::Polygon poly;
    for (int k = 0; k < 4; ++k) poly.addVertex(gm::Vec2{center.x + rectOff[k].x, center.y + rectOff[k].y});
    return poly;
  }
  gm::Polygon searchPolygon(const Point& center) const {
    return useHeading ? headingSearchPolygon(center) : getSearchPolygon(center, p.searchRadius);
  }

  // FP.cpp:2421-2463
  static Point getPolygonCenter(const Point& RF, const Point& RH, const Point& LH, const Point& LF) {
    const int n = 4;
    double x1, y1, x2, y2, x3 = 0, y3 = 0;
    double sum_x = 0, sum_y = 0, sum_s = 0;
    x1 = RF.x;  y1 = RF.y;
    x2 = RH.x;  y2 = RH.y;
    for (int i = 1; i <= n - 2; i++) {
      switch (i) {
        case 1: x3 = LH.x; y3 = LH.y; break;
        case 2: x3 = LF.x; y3 = LF.y; break;
      }
      double s = ((x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)) / 2.0;
      sum_x += (x1 + x2 + x3) * s;
      sum_y += (y1 + y2 + y3) * s;
      sum_s += s;
      x2 = x3;
      y2 = y3;
    }
    Point c;
    c.x = sum_x / sum_s / 3.0;
    c.y = sum_y / sum_s / 3.0;
    c.z = (RF.z + RH.z + LH.z + LF.z) / 4.0;
    return c;
  }

  // FP.cpp:2411-2418 (z left at 0)
  static Point getDefaultFootholdNext(const Point& feetCenterNext, const Point& defaultBias) {
    Point f;
    f.x = feetCenterNext.x + defaultBias.x;
    f.y = feetCenterNext.y + defaultBias.y;
    return f;
  }

  // FP.cpp:2496-2517
  static gm::Polygon getSearchPolygon(const Point& center, float radius) {
    gm::Polygon poly;
    gm::Vec2 leftup, rightup, rightdown, leftdown;
    leftup.x = center.x + radius;     leftup.y = center.y + 0.5 * radius;
    rightup.x = center.x + radius;    rightup.y = center.y - 0.5 * radius;
    rightdown.x = center.x - radius;  rightdown.y = center.y - 0.5 * radius;
    leftdown.x = center.x - radius;   leftdown.y = center.y + 0.5 * radius;
    poly.addVertex(leftup);
    poly.addVertex(rightup);
    poly.addVertex(rightdown);
    poly.addVertex(leftdown);
    return poly;
  }

  // FP.cpp:2166-2250 (markers/publishers dropped)
  void getFootholdSearchGridMap(const Point cur[4], Point next[4], gm::Polygon poly[4]) const {
    Point polygonCenter = getPolygonCenter(cur[0], cur[1], cur[2], cur[3]);
    Point polygonCenterNext;
    polygonCenterNext.x = polygonCenter.x + p.stepLength;         // :2199
    polygonCenterNext.y = initialPose[1] + ajustedPoseY;          // :2201
    polygonCenterNext.z = polygonCenter.z;
    if (useHeading) polygonCenterNext = headingNextCenter(polygonCenter);
    for (int f = 0; f < 4; ++f) next[f] = getDefaultFootholdNext(polygonCenterNext, bias[f]);  // :2207-2213
    for (int f = 0; f < 4; ++f) poly[f] = searchPolygon(next[f]);                              // :2235-2244
  }

  // FP.cpp:2520-2554
  float getFootholdMeanHeight(const gm::GridMap& gridmap, const gm::Vec2& center, float radius, double h) const {
    float iHeight = 0.0, meanHeight = 0.0;
    int i = 0;
    for (gm::CircleIterator it(gridmap, center, radius); !it.isPastEnd(); ++it) {
      if (gridmap.isValid(*it, "elevation")) {
        iHeight = gridmap.at("elevation", *it);
      } else {
        iHeight = 0.0;
      }
      if (iHeight < 10) {
        i++;
        meanHeight = meanHeight + iHeight;
      }
    }
    if (i != 0) {
      meanHeight = meanHeight / i;
    } else {
      meanHeight = iHeight;
    }
    return (meanHeight + h);  // float + double -> double, narrowed to the float return type
  }

  // FP.cpp:2039-2082
  bool checkDefaultFoothold(const gm::GridMap& gridmap, const gm::Vec2& center, float footRadius) const {
    bool validation = false;
    bool isNAN = false;
    for (gm::CircleIterator it(gridmap, center, footRadius); !it.isPastEnd(); ++it) {
      if (gridmap.isValid(*it, "traversability")) {
        if (gridmap.at("traversability", *it) < p.defaultFootholdThreshold) {
          validation = false;
          isNAN = true;
          break;
        }
        if (false == isNAN) validation = true;
      }
      validation = true;  // :2078
    }
    return validation;
  }

  // FP.cpp:2117-2163
  bool checkCirclePolygonFoothold(const gm::GridMap& gridmap, const gm::Vec2& center, float footRadius,
                                  const gm::Polygon& polygon) const {
    bool validation = false;
    bool isNAN = false;
    for (gm::CircleIterator it(gridmap, center, footRadius); !it.isPastEnd(); ++it) {
      if (gridmap.isValid(*it, "traversability")) {
        gm::Vec2 pos{0, 0};
        gridmap.getPosition(*it, pos);
        if (gridmap.at("traversability", *it) < p.candidateFootholdThreshold || false == polygon.isInside(pos)) {
          validation = false;
          isNAN = true;
          break;
        }
        if (false == isNAN) validation = true;
      }
      validation = true;  // :2159
    }
    return validation;
  }

  // FP.cpp:2085-2114
  bool checkCandidateFoothold(const gm::GridMap& gridmap, const gm::Vec2& spiralCenter, float footRadius,
                              float searchRadius, const gm::Polygon& polygon, Point& footholdResult,
                              gm::Idx2* chosen = nullptr) const {
    bool validation = false;
    int visitedCells = 0;
    for (gm::SpiralIterator it(gridmap, spiralCenter, searchRadius, p.spiralMode == SPIRAL_WRAP_CORRECT);
         !it.isPastEnd(); ++it, ++visitedCells) {
      gm::Vec2 footCenter;
      gm::Idx2 idx = *it;
      gridmap.getPosition(idx, footCenter);
      validation = checkCirclePolygonFoothold(gridmap, footCenter, footRadius, polygon);
      if (validation) {
        gm::Vec2 pos{0, 0};
        gridmap.getPosition(idx, pos);
        footholdResult.x = pos.x;
        footholdResult.y = pos.y;
        if (chosen) *chosen = idx;
        break;
      }
    }
    if (rankHist) ++rankHist[validation ? (visitedCells < 255 ? visitedCells : 255) : 256];
    return validation;
  }

  // FP.cpp:2001-2036
  void checkFoothold(const gm::GridMap& gridmap, const gm::Vec2& center, float footRadius, float searchRadius,
                     const gm::Polygon& polygon, Point& footholdResult, bool& footholdValidation,
                     uint8_t* kind = nullptr) const {
    bool defaultFootholdIsvalid = checkDefaultFoothold(gridmap, center, footRadius);
    footholdValidation = defaultFootholdIsvalid;
    footholdResult.x = center.x;
    footholdResult.y = center.y;
    if (kind) *kind = KIND_DEFAULT;
    bool candidateFootholdIsvalid = false;
    if (!defaultFootholdIsvalid) {
      candidateFootholdIsvalid =
          checkCandidateFoothold(gridmap, center, footRadius, searchRadius, polygon, footholdResult);
      footholdValidation = candidateFootholdIsvalid;
      if (kind) *kind = KIND_CANDIDATE;
    }
    if (defaultFootholdIsvalid | candidateFootholdIsvalid) {
      footholdResult.z = getFootholdMeanHeight(gridmap, center, footRadius, p.h);  // at `center` (:2029)
    }
  }

  // FP.cpp:1605-1997.  `fullMap` plays the role of the member gridmap_ used for the height lookups.
  bool checkFootholdUseCentroidMethod(const gm::GridMap& gridmap, const Point& defaultFoothold,
                                      Point& resultFoothold, int* caseOut = nullptr) const {
    if (caseOut) *caseOut = -1;
    gm::Vec2 rect;
    rect.x = p.searchRadius * 2;  // float * int -> float, widened (:1616)
    rect.y = p.searchRadius;
    bool isSuccess;
    gm::Vec2 pos{defaultFoothold.x, defaultFoothold.y};
    gm::GridMap map = gridmap.getSubmap(pos, rect, isSuccess);  // :1627
    if (!isSuccess) return false;                               // :1628-1631

    // :1649-1658 whole-window test (NaN compares false => passes)
    bool wholeRegionValid = false;
    const float* data = map.layers.at("traversability");
    const int nCells = map.size.i0 * map.size.i1;
    for (int i = 0; i < nCells; ++i) {
      if (data[i] < p.defaultFootholdThreshold) {
        wholeRegionValid = false;
        break;
      }
      wholeRegionValid = true;
    }

    int minRow, maxRow;
    int topRow, bottomRow;
    int rightCol;
    gm::Idx2 size = map.size;
    topRow = 0;
    bottomRow = size.i0 - 1;
    rightCol = size.i1 - 1;

    if (wholeRegionValid) {  // :1684-1689
      if (caseOut) *caseOut = 0;
      if (centroidHeight) resultFoothold.z = getFootholdMeanHeight(gridmap, pos, p.footRadius, p.h);
      resultFoothold.x = pos.x;
      resultFoothold.y = pos.y;
    } else {
      minRow = 0;
      maxRow = 0;
      gm::Idx2 startIndex{0, 0};
      gm::Idx2 endIndex{0, size.i1};  // :1720 one past the last column
      int k = 0;
      for (int j = 0; j < (bottomRow + 1); ++j) {
        startIndex.i0 = j;
        endIndex.i0 = j;
        int i = 0;
        for (gm::LineIterator it(map, startIndex, endIndex); !it.isPastEnd(); ++it) {
          if ((*it).i1 >= size.i1) continue;  // policy C.1: phantom out-of-bounds cell is not below threshold
          if (map.at("traversability", *it) < p.defaultFootholdThreshold) ++i;  // :1736
        }
        if (i > ((rightCol + 1) * 0.5)) {  // :1743
          if (k == 0) minRow = startIndex.i0;
          maxRow = startIndex.i0;
          ++k;
        }
      }

      gm::Vec2 newRegionCentroid{0, 0};
      gm::Idx2 newIndex{0, 0};
      bool matched = false;
      if (minRow == topRow && maxRow != bottomRow) {  // case 1 (:1777)
        int newRow = static_cast<int>(std::floor((maxRow + bottomRow + 1) * 0.5));
        int newCol = static_cast<int>(std::floor((rightCol + 1) * 0.5));
        newIndex = gm::Idx2{newRow, newCol};
        matched = true;
        if (caseOut) *caseOut = 1;
      } else if (minRow != topRow && maxRow != bottomRow) {  // case 2 (:1843)
        if ((minRow - topRow) >= (bottomRow - maxRow)) {
          int newRow = static_cast<int>(std::ceil(minRow * 0.5));
          int newCol = static_cast<int>(std::floor((rightCol + 0) * 0.5));
          newIndex = gm::Idx2{newRow, newCol};
        } else {
          int newRow = static_cast<int>(std::floor((maxRow + bottomRow) * 0.5));
          int newCol = static_cast<int>(std::floor((rightCol + 0) * 0.5));
          newIndex = gm::Idx2{newRow, newCol};
        }
        matched = true;
        if (caseOut) *caseOut = 2;
      } else if (minRow != topRow && maxRow == bottomRow) {  // case 3 (:1944)
        int newRow = static_cast<int>(std::ceil(minRow * 0.5));
        int newCol = static_cast<int>(std::floor((rightCol + 0) * 0.5));
        newIndex = gm::Idx2{newRow, newCol};
        matched = true;
        if (caseOut) *caseOut = 3;
      }
      if (matched) {
        // :1816 / :1917 / :1982 — a failed getPosition only logs; the (uninitialised upstream, zero here)
        // position is still used.  newIndex is always inside the window for the three cases above.
        map.getPosition(newIndex, newRegionCentroid);
        if (centroidHeight) resultFoothold.z = getFootholdMeanHeight(gridmap, newRegionCentroid, p.footRadius, p.h);
        resultFoothold.x = newRegionCentroid.x;
        resultFoothold.y = newRegionCentroid.y;
      } else if (caseOut) {
        *caseOut = 4;  // minRow == topRow && maxRow == bottomRow: no branch, result stays (0,0,0)
      }
    }
    return true;  // policy C.2
  }

  // The success predicate of the opt chain's gait-cycle submap at cycle 0 (FP.cpp:916-934 -> :2307-2350): centre =
  // (stance centroid x + stepLength, initialPose y + adjusted y) (:2333-2336; heading extension: the rotated next
  // centre), rect = (isos_.length, isos_.width) with the reference's float/double mix (FP.hpp:677-703, FP.cpp:384-394).
  bool optGaitSubmapOk(const gm::GridMap& gridmap, const Point cur[4]) const {
    Point c = getPolygonCenter(cur[0], cur[1], cur[2], cur[3]);
    Point n;
    n.x = c.x + p.stepLength;
    n.y = initialPose[1] + ajustedPoseY;
    if (useHeading) n = headingNextCenter(c);
    const double longEdge = lengthBase + p.skew * 2;  // float arithmetic, stored in a double member
    const double shortEdge = widthBase;
    const double rectLength = p.searchRadius * 2, rectWidth = p.searchRadius;  // footSearchRect_ (float arithmetic)
    gm::SubmapInfo info;
    return gm::getSubmapInformation(info, gm::Vec2{n.x, n.y}, gm::Vec2{longEdge + rectLength, shortEdge + rectWidth},
                                    gridmap.length, gridmap.position, gridmap.resolution, gridmap.size, gridmap.startIndex);
  }

  // FP.cpp:2253-2304 — un-searched default footholds of the next cycle (z = mean height at the default position)
  void getDefaultFootholds(const gm::GridMap& gridmap, const Point cur[4], Point next[4]) const {
    Point polygonCenter = getPolygonCenter(cur[0], cur[1], cur[2], cur[3]);
    Point polygonCenterNext;
    polygonCenterNext.x = polygonCenter.x + p.stepLength;   // :2270
    polygonCenterNext.y = initialPose[1] + ajustedPoseY;    // :2272
    polygonCenterNext.z = polygonCenter.z;
    if (useHeading) polygonCenterNext = headingNextCenter(polygonCenter);
    for (int f = 0; f < 4; ++f) {
      next[f] = getDefaultFootholdNext(polygonCenterNext, bias[f]);
      next[f].z = getFootholdMeanHeight(gridmap, gm::Vec2{next[f].x, next[f].y}, p.footRadius, p.h);  // :2289-2301
    }
  }
  // FP.cpp:2571-2584
  static void getHipDistance(const Point r[4], std::vector<double>& feetDistance) {
    feetDistance.push_back(r[0].x - r[2].x);  // RF - LH
    feetDistance.push_back(r[3].x - r[1].x);  // LF - RH
  }
  // FP.cpp:2587-2623 (gaitCycle_ = 1.0, FP.cpp:332)
  void getCogSpeed(const Point r[4], const Point cur[4], std::vector<double>& cogSpeed) const {
    double feetCenter1, feetCenter2, feetCenter3;
    if (p.rfFirst) {
      feetCenter1 = (cur[0].x + cur[2].x) / 2;
      feetCenter2 = (r[3].x + r[1].x) / 2;
      feetCenter3 = (r[0].x + r[2].x) / 2;
    } else {
      feetCenter1 = (cur[3].x + cur[1].x) / 2;
      feetCenter2 = (r[0].x + r[2].x) / 2;
      feetCenter3 = (r[3].x + r[1].x) / 2;
    }
    const double gaitCycle = 1.0;
    cogSpeed.push_back((feetCenter2 - feetCenter1) / (0.5 * gaitCycle));
    cogSpeed.push_back((feetCenter3 - feetCenter2) / (0.5 * gaitCycle));
  }

  static void fillIndex(const gm::GridMap& map, Foothold& fh) {
    gm::Idx2 idx;
    if (map.getIndex(gm::Vec2{fh.point.x, fh.point.y}, idx)) {
      fh.row = idx.i0;
      fh.col = idx.i1;
    } else {
      fh.row = fh.col = -1;
    }
  }

  // FP.cpp:539-1602 — the plan_global_footholds service body restricted to the centroid + nominal chains.
  PlanResult globalFootholdPlan(const gm::GridMap& gridmap, uint8_t gaitCycles) {
    PlanResult res;
    Point nomCur[4], cenCur[4], defCur[4];
    setFirstGait(initial, defCur);  // :562
    setFirstGait(initial, cenCur);  // :569
    setFirstGait(initial, nomCur);  // :576
    res.gait_cycles = gaitCycles;   // :592-593
    res.gait_cycles_succeed = 0;    // :682
    res.success = false;            // :683
    for (int f = 0; f < 4; ++f) {   // :685-699
      Foothold fh;
      fh.point = initial[f];
      fh.foot_id = static_cast<uint8_t>(f);
      fh.gait_cycle_id = 0;
      fh.kind = KIND_STANCE;
      fillIndex(gridmap, fh);
      res.footholds.push_back(fh);
      res.centroidFootholds.push_back(fh);  // :713-727
      res.defaultFootholds.push_back(fh);   // :666-671
    }
    ajustedPoseY = 0.0;  // :759
    if (gaitCycles > 0) res.optSubmapWouldFail = !optGaitSubmapOk(gridmap, nomCur);
    const int gaitCycleNum = gaitCycles;
    for (int g = 0; g < gaitCycleNum; ++g) {  // :762
      // --- default chain (:774-781)
      Point defNext[4];
      getDefaultFootholds(gridmap, defCur, defNext);
      // --- centroid chain (:788-821)
      Point cenNext[4];
      gm::Polygon cenPoly[4];
      getFootholdSearchGridMap(cenCur, cenNext, cenPoly);
      Point cenRes[4];  // default-constructed (0,0,0) (:810-813)
      gm::TouchLog legLog[4];
      for (int f = 0; f < 4; ++f) {
        if (countTouched) gridmap.touch = &legLog[f];
        checkFootholdUseCentroidMethod(gridmap, cenNext[f], cenRes[f]);
      }
      // --- nominal chain (:830-910)
      Point nomNext[4];
      gm::Polygon nomPoly[4];
      getFootholdSearchGridMap(nomCur, nomNext, nomPoly);
      Point nomRes[4];
      bool ok[4];
      uint8_t kind[4];
      if (legThreads && !countTouched) {  // :863-909: one std::thread per leg, joined before the AND
        std::thread th[4];
        for (int f = 0; f < 4; ++f)
          th[f] = std::thread([&, f]() {
            gm::Vec2 center{cenNext[f].x, cenNext[f].y};
            checkFoothold(gridmap, center, p.footRadius, p.searchRadius, nomPoly[f], nomRes[f], ok[f], &kind[f]);
          });
        for (int f = 0; f < 4; ++f) th[f].join();
      } else
      for (int f = 0; f < 4; ++f) {
        gm::Vec2 center{cenNext[f].x, cenNext[f].y};  // centre from the CENTROID chain (:861-862)
        if (countTouched) gridmap.touch = &legLog[f];
        checkFoothold(gridmap, center, p.footRadius, p.searchRadius, nomPoly[f], nomRes[f], ok[f], &kind[f]);
        if (kind[f] == KIND_CANDIDATE) ++spiralTaken;
      }
      if (countTouched) {
        gridmap.touch = nullptr;
        for (int f = 0; f < 4; ++f) touchedCells += static_cast<int64_t>(legLog[f].cells.size());
      }
      bool footholdValidation = ok[0] & ok[1] & ok[2] & ok[3];  // :1323
      if (footholdValidation) {
        res.gait_cycles_succeed = static_cast<uint8_t>(g + 1);  // :1379
        res.success = true;                                     // :1380
        for (int f = 0; f < 4; ++f) {                           // :1382-1396
          Foothold fh;
          fh.point = nomRes[f];
          fh.foot_id = static_cast<uint8_t>(f);
          fh.gait_cycle_id = static_cast<uint8_t>(g);
          fh.kind = kind[f];
          fillIndex(gridmap, fh);
          res.footholds.push_back(fh);
          Foothold fc;                                          // :1448-1462
          fc.point = cenRes[f];
          fc.foot_id = static_cast<uint8_t>(f);
          fc.gait_cycle_id = static_cast<uint8_t>(g);
          fc.kind = KIND_DEFAULT;
          fillIndex(gridmap, fc);
          res.centroidFootholds.push_back(fc);
        }
        for (int f = 0; f < 4; ++f) {                          // :1338-1348
          Foothold fd;
          fd.point = defNext[f];
          fd.foot_id = static_cast<uint8_t>(f);
          fd.gait_cycle_id = static_cast<uint8_t>(g);
          fd.kind = KIND_DEFAULT;
          fillIndex(gridmap, fd);
          res.defaultFootholds.push_back(fd);
          defCur[f] = defNext[f];
        }
        getHipDistance(nomRes, res.feetDistanceNominal);        // :1357
        getCogSpeed(nomRes, nomCur, res.cogSpeedNominal);       // :1366 (before the chain advances)
        getHipDistance(cenRes, res.feetDistanceCentroid);       // :1422
        getCogSpeed(cenRes, cenCur, res.cogSpeedCentroid);      // :1431
        for (int f = 0; f < 4; ++f) nomCur[f] = nomRes[f];  // :1413-1416
        for (int f = 0; f < 4; ++f) cenCur[f] = cenRes[f];  // :1480-1483
      } else {
        res.success = false;  // :1574 (no break, :1575)
      }
      ajustedPoseY += p.driftPerCycle;  // :1578 (reference constant -0.007)
    }
    return res;
  }
};

}  // namespace ref
