// libfhp device code (sm_100a).  One LANE owns one query (or one fine-grained task); the 32 lanes of a warp
// cooperate only to stage each lane's search window as bit-planes in shared memory.
//
// Why this shape: the per-foothold work is ~100 map cells and a few hundred double-precision
// scalar operations that must reproduce grid_map_core's arithmetic bit for bit.  Lane-per-query makes that
// scalar restatement cost 1/32 of a warp instruction per query; the only wide, coalescable part — reading the
// (2*HW+1)^2 traversability window around the search centre — is done by all 32 lanes for one query at a time
// (rows of the column-major layer are contiguous, so a warp load covers whole columns of the window) and
// reduced with __ballot_sync into three bit-planes per window column:
//     D: cell < defaultFootholdThreshold      (raw compare; NaN -> 0)          FP.cpp:1653, :1736, :2057
//     C: cell < candidateFootholdThreshold                                      FP.cpp:2138
//     F: isfinite(cell)                        GridMap::isValid                 FP.cpp:2055, :2132
// Every decision of the reference (centroid row scan, foot-disc test, spiral candidates, polygon test) then
// becomes bit arithmetic on those planes in the owning lane.
//
// All double arithmetic follows the operand order of grid_map_core / FootholdPlanner.cpp; this translation unit
// is compiled with -fmad=false so no multiply-add is ever contracted (x86-64 reference builds have no FMA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fhp_types.h"

namespace fhp {

// ------------------------------------------------------------------------------------------------ grid_map math
// boundPositionToRange (GridMapMath.cpp) for one axis
__device__ __forceinline__ double bound_axis(double p, double len, double mp, double half) {
  double shifted = (p - mp) + half;
  double eps = 10.0 * 2.220446049250313e-16;
  if (fabs(p) > 1.0) eps *= fabs(p);
  if (shifted <= 0) {
    shifted = eps;
  } else if (shifted >= len) {
    shifted = len - eps;
  }
  return (shifted + mp) - half;
}
// getIndexFromPosition without the start-index wrap: (int)(-(((p - 0.5*len) - mapPos)/res))
__device__ __forceinline__ int raw_index_axis(double p, double half, double mp, double res) {
  double iv = ((p - half) - mp) / res;
  return static_cast<int>(-iv);
}
// checkIfPositionWithinMap for one axis
__device__ __forceinline__ bool within_axis(double p, double len, double mp, double half) {
  double t = -((p - mp) - half);
  return t >= 0.0 && t < len;
}
// getPositionFromIndex (unwrapped index): (mapPos + offset) + res * (double)(-u)
__device__ __forceinline__ double pos_axis(double fc, double res, int u) { return fc + res * static_cast<double>(-u); }

__device__ __forceinline__ int wrap_add(int u, int start, int n) {  // unwrapped (in range) -> buffer index
  int b = u + start;
  return b >= n ? b - n : b;
}
__device__ __forceinline__ bool in_map(const MapDev& M, int u0, int u1) {
  return u0 >= 0 && u1 >= 0 && u0 < M.rows && u1 < M.cols;
}
__device__ __forceinline__ float load_cell(const MapDev& M, const float* layer, int u0, int u1) {
  int b0 = wrap_add(u0, M.start0, M.rows), b1 = wrap_add(u1, M.start1, M.cols);
  return __ldg(layer + static_cast<size_t>(b1) * M.rows + b0);
}
// getIndex(position) as the reference's diagnostics use it: buffer index, or (-1,-1) outside the map
__device__ __forceinline__ void cell_of_position(const MapDev& M, double x, double y, int& row, int& col) {
  bool ok = within_axis(x, M.lenx, M.posx, M.halfx) && within_axis(y, M.leny, M.posy, M.halfy);
  if (!ok) {
    row = col = -1;
    return;
  }
  int u0 = raw_index_axis(x, M.halfx, M.posx, M.res), u1 = raw_index_axis(y, M.halfy, M.posy, M.res);
  if (!in_map(M, u0, u1)) {  // unreachable for finite geometry; keep the contract total
    row = col = -1;
    return;
  }
  row = wrap_add(u0, M.start0, M.rows);
  col = wrap_add(u1, M.start1, M.cols);
}

// ------------------------------------------------------------------------------------------------ bit planes
struct Planes {  // view of one query's staged window in shared memory
  const uint32_t* w;
  int pack, rpShift;
  __device__ __forceinline__ uint32_t col(int plane, int b) const {  // bits over window rows a for column b
    if (pack == 1) return w[b * 3 + plane];
    int k, sh;
    if (pack == 2) {
      k = b >> 1;
      sh = (b & 1) << 4;
      return (w[k * 3 + plane] >> sh) & 0xffffu;
    }
    k = b >> 2;
    sh = (b & 3) << 3;
    return (w[k * 3 + plane] >> sh) & 0xffu;
  }
  __device__ __forceinline__ bool bit(int plane, int a, int b) const { return (col(plane, b) >> a) & 1u; }
};
enum { PL_D = 0, PL_C = 1, PL_F = 2 };

// Warp-cooperative staging: for every lane j with has_window, read the WIN x WIN traversability window whose
// unwrapped origin is (w0,w1) of lane j and store its D/C/F bit-planes at smem[j*stride ...].
// Cells outside the map read as NaN (=> D=C=F=0), which is exactly "not visited" for every iterator.
__device__ __forceinline__ void stage_windows(const MapDev& M, const GeoDev& G, uint32_t* smemWarp, int lane,
                                              bool has_window, int w0, int w1, float thrD, float thrC) {
  const unsigned act = __ballot_sync(kFull, has_window);
  const int RP = 1 << G.rpShift;
  const int a = lane & (RP - 1);
  const int h = lane >> G.rpShift;
  __syncwarp();
  for (unsigned m = act; m; m &= m - 1) {
    const int j = __ffs(m) - 1;
    const int jw0 = __shfl_sync(kFull, w0, j);
    const int jw1 = __shfl_sync(kFull, w1, j);
    const float tD = __shfl_sync(kFull, thrD, j);
    const float tC = __shfl_sync(kFull, thrC, j);
    const int u0 = jw0 + a;
    const bool in0 = (a < G.WIN) && u0 >= 0 && u0 < M.rows;
    const int b0 = wrap_add(u0, M.start0, M.rows);
    uint32_t myD = 0, myC = 0, myF = 0;
    for (int k = 0; k < G.nIter; ++k) {
      const int b = k * G.pack + h;
      const int u1 = jw1 + b;
      const bool in = in0 && (b < G.WIN) && u1 >= 0 && u1 < M.cols;
      const int b1 = wrap_add(u1, M.start1, M.cols);
      float v = __int_as_float(0x7fc00000);
      if (in) v = __ldg(M.trav + static_cast<size_t>(b1) * M.rows + b0);
      const uint32_t bD = __ballot_sync(kFull, v < tD);
      const uint32_t bC = __ballot_sync(kFull, v < tC);
      const uint32_t bF = __ballot_sync(kFull, fabsf(v) < __int_as_float(0x7f800000));
      if (lane == k) {
        myD = bD;
        myC = bC;
        myF = bF;
      }
    }
    if (lane < G.nIter) {
      uint32_t* dst = smemWarp + j * G.stride + lane * 3;
      dst[0] = myD;
      dst[1] = myC;
      dst[2] = myF;
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ a9 mean height
// FootholdPlanner::getFootholdMeanHeight (FP.cpp:2520-2554) over grid_map::CircleIterator (row-major over the
// bounded window in unwrapped index space, cells with squared distance <= r^2).
__device__ __forceinline__ float mean_height(const MapDev& M, const ParDev& P, double cx, double cy) {
  const double tlx = bound_axis(cx + P.r, M.lenx, M.posx, M.halfx), tly = bound_axis(cy + P.r, M.leny, M.posy, M.halfy);
  const double brx = bound_axis(cx - P.r, M.lenx, M.posx, M.halfx), bry = bound_axis(cy - P.r, M.leny, M.posy, M.halfy);
  const int tl0 = raw_index_axis(tlx, M.halfx, M.posx, M.res), tl1 = raw_index_axis(tly, M.halfy, M.posy, M.res);
  const int br0 = raw_index_axis(brx, M.halfx, M.posx, M.res), br1 = raw_index_axis(bry, M.halfy, M.posy, M.res);
  float iHeight = 0.0f, meanHeight = 0.0f;
  int i = 0;
  for (int u0 = tl0; u0 <= br0; ++u0) {
    const double dx = pos_axis(M.fcx, M.res, u0) - cx;
    const double dx2 = dx * dx;
    for (int u1 = tl1; u1 <= br1; ++u1) {
      const double dy = pos_axis(M.fcy, M.res, u1) - cy;
      if (!(dx2 + dy * dy <= P.r2) || !in_map(M, u0, u1)) continue;
      const float e = load_cell(M, M.elev, u0, u1);
      iHeight = (fabsf(e) < __int_as_float(0x7f800000)) ? e : 0.0f;
      if (iHeight < 10.0f) {
        i++;
        meanHeight = meanHeight + iHeight;
      }
    }
  }
  if (i != 0) {
    meanHeight = meanHeight / static_cast<float>(i);
  } else {
    meanHeight = iHeight;
  }
  return static_cast<float>(static_cast<double>(meanHeight) + P.h);
}

// ------------------------------------------------------------------------------------------------ a4 centroid
struct CentroidOut {
  double x, y;  // result xy ((0,0) when the submap failed or no case matched)
  int ccase;    // 0 whole window valid, 1/2/3 reference cases, 4 no case matched, -1 submap failed
};
// FootholdPlanner::checkFootholdUseCentroidMethod (FP.cpp:1605-1997) on the staged planes of this lane.
// (w0,w1): unwrapped origin of the staged window.
__device__ __forceinline__ CentroidOut centroid_eval(const MapDev& M, const GeoDev& G, const ParDev& P, const Planes& pl,
                                                     int w0, int w1, double px, double py, unsigned& err) {
  CentroidOut o;
  o.x = 0.0;
  o.y = 0.0;
  o.ccase = -1;
  // GridMap::getSubmap -> getSubmapInformation (GridMapMath.cpp)
  const double tlx = bound_axis(px + 0.5 * P.rectLx, M.lenx, M.posx, M.halfx);
  const double tly = bound_axis(py + 0.5 * P.rectLy, M.leny, M.posy, M.halfy);
  if (!(within_axis(tlx, M.lenx, M.posx, M.halfx) && within_axis(tly, M.leny, M.posy, M.halfy))) return o;
  const int tl0 = raw_index_axis(tlx, M.halfx, M.posx, M.res), tl1 = raw_index_axis(tly, M.halfy, M.posy, M.res);
  const double brx = bound_axis(px - 0.5 * P.rectLx, M.lenx, M.posx, M.halfx);
  const double bry = bound_axis(py - 0.5 * P.rectLy, M.leny, M.posy, M.halfy);
  if (!(within_axis(brx, M.lenx, M.posx, M.halfx) && within_axis(bry, M.leny, M.posy, M.halfy))) return o;
  const int br0 = raw_index_axis(brx, M.halfx, M.posx, M.res), br1 = raw_index_axis(bry, M.halfy, M.posy, M.res);
  if (!in_map(M, tl0, tl1)) return o;  // getPositionFromIndex(topLeftCorner) range check
  const double cornx = pos_axis(M.fcx, M.res, tl0) + 0.5 * M.res;
  const double corny = pos_axis(M.fcy, M.res, tl1) + 0.5 * M.res;
  const int n0 = br0 - tl0 + 1, n1 = br1 - tl1 + 1;  // submap size (rows, cols)
  const double slx = static_cast<double>(n0) * M.res, sly = static_cast<double>(n1) * M.res;
  const double shx = 0.5 * slx, shy = 0.5 * sly;
  const double spx = cornx - shx, spy = corny - shy;  // submap position
  if (!(within_axis(px, slx, spx, shx) && within_axis(py, sly, spy, shy))) return o;  // requested centre in submap

  // window coordinates of the submap inside the staged planes
  const int a0 = tl0 - w0, a1 = br0 - w0, c0 = tl1 - w1, c1 = br1 - w1;
  if (a0 < 0 || c0 < 0 || a1 >= G.WIN || c1 >= G.WIN || n0 < 1 || n1 < 1 || n0 > 32) {
    err |= kErrWindow;
    return o;
  }
  const uint32_t rowmask = (n0 >= 32 ? 0xffffffffu : ((1u << n0) - 1u)) << a0;

  // whole-window test (FP.cpp:1649-1658) and per-row counts of cells below the threshold (:1722-1750) as a
  // bit-sliced vertical counter over the window columns (policy C.1: the phantom column is not counted)
  uint32_t any = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  for (int c = c0; c <= c1; ++c) {
    uint32_t x = pl.col(PL_D, c) & rowmask;
    any |= x;
    uint32_t t;
    t = s0 & x; s0 ^= x; x = t;
    t = s1 & x; s1 ^= x; x = t;
    t = s2 & x; s2 ^= x; x = t;
    t = s3 & x; s3 ^= x; x = t;
    s4 ^= x;
  }
  if (any == 0) {  // wholeRegionValid (:1684-1689)
    o.x = px;
    o.y = py;
    o.ccase = 0;
    return o;
  }
  // row is an "edge" when count > (rightCol+1)*0.5  <=>  count > floor(n1/2) for integers (:1743)
  const int T = n1 >> 1;
  uint32_t gt = 0, eq = 0xffffffffu;
  {
    const uint32_t sv[5] = {s0, s1, s2, s3, s4};
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      const uint32_t ti = ((T >> i) & 1) ? 0xffffffffu : 0u;
      gt |= eq & sv[i] & ~ti;
      eq &= ~(sv[i] ^ ti);
    }
  }
  const uint32_t edge = gt & rowmask;
  int minRow = 0, maxRow = 0;
  if (edge) {
    minRow = (__ffs(edge) - 1) - a0;
    maxRow = (31 - __clz(edge)) - a0;
  }
  const int topRow = 0, bottomRow = n0 - 1, rightCol = n1 - 1;
  int newRow, newCol;
  if (minRow == topRow && maxRow != bottomRow) {  // case 1 (:1777)
    newRow = (maxRow + bottomRow + 1) >> 1;      // floor((maxRow+bottomRow+1)*0.5)
    newCol = (rightCol + 1) >> 1;                // floor((rightCol+1)*0.5)
    o.ccase = 1;
  } else if (minRow != topRow && maxRow != bottomRow) {  // case 2 (:1843)
    if ((minRow - topRow) >= (bottomRow - maxRow)) {
      newRow = (minRow + 1) >> 1;  // ceil(minRow*0.5)
    } else {
      newRow = (maxRow + bottomRow) >> 1;  // floor((maxRow+bottomRow)*0.5)
    }
    newCol = rightCol >> 1;  // floor(rightCol*0.5)
    o.ccase = 2;
  } else if (minRow != topRow && maxRow == bottomRow) {  // case 3 (:1944)
    newRow = (minRow + 1) >> 1;
    newCol = rightCol >> 1;
    o.ccase = 3;
  } else {  // minRow == topRow && maxRow == bottomRow: no branch in the reference, result stays (0,0,0)
    o.ccase = 4;
    return o;
  }
  // map.getPosition(newIndex) on the submap (start index 0): (subPos + (0.5*subLen - 0.5*res)) + res*(-idx)
  o.x = (spx + (shx - 0.5 * M.res)) + M.res * static_cast<double>(-newRow);
  o.y = (spy + (shy - 0.5 * M.res)) + M.res * static_cast<double>(-newCol);
  return o;
}

// ------------------------------------------------------------------------------------------------ a5..a8 nominal
struct NominalOut {
  double x, y;
  bool ok;
  int kind;
  int row, col;  // buffer index of the chosen cell (or of the centre's cell), (-1,-1) outside the map
};

// One foot-disc candidate evaluated straight from global memory (used only for faithful-mode candidates that
// wrap to the far side of the circular buffer): checkCirclePolygonFoothold (FP.cpp:2117-2163).
__device__ __noinline__ bool candidate_slow(const MapDev& M, const GeoDev& G, const ParDev& P, int q0, int q1, double xlo,
                                            double xhi, double ylo, double yhi) {
  for (int s = 0; s < G.nSten; ++s) {
    const int u0 = q0 + G.sten[s][0], u1 = q1 + G.sten[s][1];
    if (!in_map(M, u0, u1)) continue;
    const float v = load_cell(M, M.trav, u0, u1);
    if (!(fabsf(v) < __int_as_float(0x7f800000))) continue;
    const double X = pos_axis(M.fcx, M.res, u0), Y = pos_axis(M.fcy, M.res, u1);
    const bool inside = (X >= xlo) && (X < xhi) && (Y >= ylo) && (Y < yhi);
    if (v < P.thrC || !inside) return false;
  }
  return true;
}

// FootholdPlanner::checkFoothold (FP.cpp:2001-2036) = checkDefaultFoothold (:2039-2082), then
// checkCandidateFoothold (:2085-2114) over grid_map::SpiralIterator with checkCirclePolygonFoothold (:2117-2163).
// (cx,cy): search centre; (rx,ry): centre of the search rectangle (getSearchPolygon, FP.cpp:2496-2517);
// (uc0,uc1): raw unwrapped index of the centre; (w0,w1) = (uc0-HW, uc1-HW): staged window origin.
template <bool DEFAULT_ONLY>
__device__ __forceinline__ NominalOut nominal_eval(const MapDev& M, const GeoDev& G, const ParDev& P, const Planes& pl,
                                                   int uc0, int uc1, double cx, double cy, double rx, double ry,
                                                   unsigned& err) {
  NominalOut o;
  o.x = cx;  // FP.cpp:2016-2017
  o.y = cy;
  o.ok = false;
  o.kind = FHP_KIND_DEFAULT;
  const int w0 = uc0 - G.HW, w1 = uc1 - G.HW;
  const bool centre_in_map = within_axis(cx, M.lenx, M.posx, M.halfx) && within_axis(cy, M.leny, M.posy, M.halfy) &&
                             in_map(M, uc0, uc1);
  o.row = centre_in_map ? wrap_add(uc0, M.start0, M.rows) : -1;
  o.col = centre_in_map ? wrap_add(uc1, M.start1, M.cols) : -1;

  // ---- checkDefaultFoothold: CircleIterator(centre, footRadius)
  {
    const double tlx = bound_axis(cx + P.r, M.lenx, M.posx, M.halfx), tly = bound_axis(cy + P.r, M.leny, M.posy, M.halfy);
    const double brx = bound_axis(cx - P.r, M.lenx, M.posx, M.halfx), bry = bound_axis(cy - P.r, M.leny, M.posy, M.halfy);
    const int tl0 = raw_index_axis(tlx, M.halfx, M.posx, M.res), tl1 = raw_index_axis(tly, M.halfy, M.posy, M.res);
    const int br0 = raw_index_axis(brx, M.halfx, M.posx, M.res), br1 = raw_index_axis(bry, M.halfy, M.posy, M.res);
    int visited = 0;
    bool below = false;
    for (int u0 = tl0; u0 <= br0; ++u0) {
      const double dx = pos_axis(M.fcx, M.res, u0) - cx;
      const double dx2 = dx * dx;
      for (int u1 = tl1; u1 <= br1; ++u1) {
        const double dy = pos_axis(M.fcy, M.res, u1) - cy;
        if (!(dx2 + dy * dy <= P.r2) || !in_map(M, u0, u1)) continue;
        ++visited;
        const int a = u0 - w0, b = u1 - w1;
        if (a < 0 || b < 0 || a >= G.WIN || b >= G.WIN) {
          err |= kErrWindow;
          continue;
        }
        const uint32_t bits = pl.col(PL_D, b) & pl.col(PL_F, b);
        below |= (bits >> a) & 1u;
      }
    }
    o.ok = (visited > 0) && !below;
  }
  if (o.ok || DEFAULT_ONLY) return o;

  // ---- checkCandidateFoothold: SpiralIterator(centre, searchRadius)
  o.kind = FHP_KIND_CANDIDATE;
  const double xhi = rx + P.R, xlo = rx - P.R, yhi = ry + P.halfR, ylo = ry - P.halfR;
  // Polygon::isInside for the axis-aligned rectangle == half-open box on the cell centre;
  // evaluate it once per window row / column with the reference's position arithmetic.
  uint32_t rowsBox = 0, colsBox = 0;
  for (int a = 0; a < G.WIN; ++a) {
    const double X = pos_axis(M.fcx, M.res, w0 + a);
    if (X >= xlo && X < xhi) rowsBox |= 1u << a;
    const double Y = pos_axis(M.fcy, M.res, w1 + a);
    if (Y >= ylo && Y < yhi) colsBox |= 1u << a;
  }
  const bool faithful = (P.spiralMode == FHP_SPIRAL_FAITHFUL) && !(M.start0 == 0 && M.start1 == 0);
  // faithful: upstream's centre index is the BUFFER index (getIndex), neighbours are taken in buffer space
  int cb0 = uc0, cb1 = uc1;
  if (faithful) {
    cb0 = uc0 + M.start0;
    cb1 = uc1 + M.start1;
    // wrapIndexToRange
    if (cb0 < 0) cb0 += ((-cb0 / M.rows) + 1) * M.rows;
    cb0 %= M.rows;
    if (cb1 < 0) cb1 += ((-cb1 / M.cols) + 1) * M.cols;
    cb1 %= M.cols;
  }
  for (int k = 0; k < G.nCand; ++k) {
    const int d0 = G.spiral[k][0], d1 = G.spiral[k][1];
    int q0, q1;  // unwrapped index of the candidate cell
    bool far = false;
    if (!faithful) {
      q0 = uc0 + d0;
      q1 = uc1 + d1;
      if (!in_map(M, q0, q1)) continue;  // checkIfIndexInRange(pointInMap)
      if (k >= G.lastTwoStart) {         // rings nRings-1, nRings: distance filter on the cell centre
        const double dx = pos_axis(M.fcx, M.res, q0) - cx, dy = pos_axis(M.fcy, M.res, q1) - cy;
        if (!(dx * dx + dy * dy <= P.R2)) continue;
      }
    } else {
      const int bi0 = cb0 + d0, bi1 = cb1 + d1;  // buffer index, NOT wrapped
      if (!in_map(M, bi0, bi1)) continue;
      if (k >= G.lastTwoStart) {  // upstream evaluates the position as if the start index were 0
        const double dx = pos_axis(M.fcx, M.res, bi0) - cx, dy = pos_axis(M.fcy, M.res, bi1) - cy;
        if (!(dx * dx + dy * dy <= P.R2)) continue;
      }
      q0 = bi0 - M.start0;
      if (q0 < 0) q0 += M.rows;
      q1 = bi1 - M.start1;
      if (q1 < 0) q1 += M.cols;
      far = (q0 != uc0 + d0) || (q1 != uc1 + d1);
    }
    bool good = true;
    if (far) {
      good = candidate_slow(M, G, P, q0, q1, xlo, xhi, ylo, yhi);
    } else {
      const int qa = d0 + G.HW, qb = d1 + G.HW;  // window coordinates of the candidate
      for (int s = 0; s < G.nSten && good; ++s) {
        const int a = qa + G.sten[s][0], b = qb + G.sten[s][1];
        // a,b are inside the window by construction (|d| <= nRings, |sten| <= ra); cells outside the map
        // have F = 0 and are skipped like the iterator skips them
        const uint32_t F = pl.col(PL_F, b), C = pl.col(PL_C, b);
        const uint32_t box = ((colsBox >> b) & 1u) ? rowsBox : 0u;
        const uint32_t bad = F & (C | ~box);
        good = !((bad >> a) & 1u);
      }
    }
    if (good) {
      o.ok = true;
      o.x = pos_axis(M.fcx, M.res, q0);  // gridmap.getPosition(index) (FP.cpp:2105-2107)
      o.y = pos_axis(M.fcy, M.res, q1);
      o.row = wrap_add(q0, M.start0, M.rows);
      o.col = wrap_add(q1, M.start1, M.cols);
      break;
    }
  }
  return o;
}

}  // namespace fhp
