// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or call it, and only as the checker / reported CPU baseline.
//
// PARITY UNPINNED for this file: the reference (lukechencqu/quadrupedal_foothold_planner) has no tests, fixtures or
// golden vectors, and this file restates a THIRD-PARTY dependency that is not vendored under
// /root/reference and not version-pinned there (the planner-level restatement in fp_oracle.hpp IS checked against
// the reference's own FootholdPlanner.cpp, which oracle/ref_shim compiles on top of this very file):
//     ANYbotics/grid_map -> grid_map_core  (BSD-3; era of the reference => release 1.6.2, ROS kinetic/melodic)
//     (foothold_planner/CMakeLists.txt:8-18 find_package(... grid_map_core ...), no version).
// What follows is a restatement of the published grid_map_core algorithms that the reference calls at
//     getSubmap        FootholdPlanner.cpp:1627, :2345        CircleIterator  :2048, :2126, :2529
//     GridMapIterator  :1651      LineIterator :1733          SpiralIterator  :2095
//     Polygon::isInside :2138     getPosition :1702,:1816,:2098,:2105,:2136   getIndex :1639-1641,:1703
//     isValid/at :2055-2057, :2132-2138, :2532-2533
// Every function names the upstream function it restates.  Arithmetic is double, operand order as
// upstream; compile with -ffp-contract=off (x86-64 ROS builds have no FMA contraction).
//
// No Eigen: 2-vectors are plain structs, a layer is a column-major float array (Eigen::MatrixXf
// layout: element (i0,i1) at i1*rows + i0; i0 <-> x, i1 <-> y).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <unordered_set>
#include <vector>

namespace gm {

struct Vec2 {  // grid_map::Position / Length / Vector (Eigen::Vector2d / Array2d)
  double x, y;
  double operator[](int i) const { return i == 0 ? x : y; }
  double& operator[](int i) { return i == 0 ? x : y; }
};
struct Idx2 {  // grid_map::Index / Size (Eigen::Array2i)
  int i0, i1;
  int operator[](int i) const { return i == 0 ? i0 : i1; }
  int& operator[](int i) { return i == 0 ? i0 : i1; }
  bool operator==(const Idx2& o) const { return i0 == o.i0 && i1 == o.i1; }
};

// ---------------------------------------------------------------- GridMapMath.cpp
// wrapIndexToRange(int&, int)
inline void wrapIndexToRange(int& index, int bufferSize) {
  if (index < 0) index += ((-index / bufferSize) + 1) * bufferSize;
  index = index % bufferSize;
}
inline void wrapIndexToRange(Idx2& index, const Idx2& bufferSize) {
  wrapIndexToRange(index.i0, bufferSize.i0);
  wrapIndexToRange(index.i1, bufferSize.i1);
}
// checkIfStartIndexAtDefaultPosition
inline bool startAtDefault(const Idx2& s) { return s.i0 == 0 && s.i1 == 0; }
// getIndexFromBufferIndex: buffer (storage) index -> unwrapped index
inline Idx2 getIndexFromBufferIndex(const Idx2& bufferIndex, const Idx2& bufferSize, const Idx2& start) {
  if (startAtDefault(start)) return bufferIndex;
  Idx2 index{bufferIndex.i0 - start.i0, bufferIndex.i1 - start.i1};
  wrapIndexToRange(index, bufferSize);
  return index;
}
// getBufferIndexFromIndex: unwrapped index -> buffer (storage) index
inline Idx2 getBufferIndexFromIndex(const Idx2& index, const Idx2& bufferSize, const Idx2& start) {
  if (startAtDefault(start)) return index;
  Idx2 b{index.i0 + start.i0, index.i1 + start.i1};
  wrapIndexToRange(b, bufferSize);
  return b;
}
// checkIfIndexInRange
inline bool checkIfIndexInRange(const Idx2& index, const Idx2& bufferSize) {
  return index.i0 >= 0 && index.i1 >= 0 && index.i0 < bufferSize.i0 && index.i1 < bufferSize.i1;
}
// getVectorToOrigin: 0.5 * mapLength
inline Vec2 getVectorToOrigin(const Vec2& mapLength) { return Vec2{0.5 * mapLength.x, 0.5 * mapLength.y}; }
// getVectorToFirstCell: vectorToOrigin - 0.5 * resolution
inline Vec2 getVectorToFirstCell(const Vec2& mapLength, double resolution) {
  Vec2 o = getVectorToOrigin(mapLength);
  return Vec2{o.x - 0.5 * resolution, o.y - 0.5 * resolution};
}
// checkIfPositionWithinMap: transformed = -(position - mapPosition - offset); 0 <= t < length
inline bool checkIfPositionWithinMap(const Vec2& position, const Vec2& mapLength, const Vec2& mapPosition) {
  Vec2 offset = getVectorToOrigin(mapLength);
  double tx = -((position.x - mapPosition.x) - offset.x);
  double ty = -((position.y - mapPosition.y) - offset.y);
  return tx >= 0.0 && ty >= 0.0 && tx < mapLength.x && ty < mapLength.y;
}
// getPositionFromIndex: position = mapPosition + offset + resolution * (-(unwrapped index))
inline bool getPositionFromIndex(Vec2& position, const Idx2& index, const Vec2& mapLength, const Vec2& mapPosition,
                                 double resolution, const Idx2& bufferSize, const Idx2& bufferStartIndex) {
  if (!checkIfIndexInRange(index, bufferSize)) return false;
  Vec2 offset = getVectorToFirstCell(mapLength, resolution);
  Idx2 u = getIndexFromBufferIndex(index, bufferSize, bufferStartIndex);
  double vx = static_cast<double>(-u.i0), vy = static_cast<double>(-u.i1);
  position.x = (mapPosition.x + offset.x) + resolution * vx;
  position.y = (mapPosition.y + offset.y) + resolution * vy;
  return true;
}
// double -> int of a NaN / out-of-int-range value is UB upstream (x86 cvttsd2si yields INT_MIN, i.e. "far out of
// range").  Policy (shared with the CUDA path): every such index becomes -2^30.
inline int toIndex(double v) {
  if (!(std::fabs(v) < 1073741824.0)) return -1073741824;
  return static_cast<int>(v);
}
// getIndexFromPosition: indexVector = (position - offset - mapPosition)/resolution; index = (int)(-indexVector)
// -> buffer index; returns false (index still written) when the position is outside the map.
inline bool getIndexFromPosition(Idx2& index, const Vec2& position, const Vec2& mapLength, const Vec2& mapPosition,
                                 double resolution, const Idx2& bufferSize, const Idx2& bufferStartIndex) {
  Vec2 offset = getVectorToOrigin(mapLength);
  double ivx = ((position.x - offset.x) - mapPosition.x) / resolution;
  double ivy = ((position.y - offset.y) - mapPosition.y) / resolution;
  Idx2 u{toIndex(-ivx), toIndex(-ivy)};
  index = getBufferIndexFromIndex(u, bufferSize, bufferStartIndex);
  return checkIfPositionWithinMap(position, mapLength, mapPosition);
}
// boundPositionToRange (always rewrites the position, even when it is inside)
inline void boundPositionToRange(Vec2& position, const Vec2& mapLength, const Vec2& mapPosition) {
  Vec2 vectorToOrigin = getVectorToOrigin(mapLength);
  Vec2 shifted{(position.x - mapPosition.x) + vectorToOrigin.x, (position.y - mapPosition.y) + vectorToOrigin.y};
  for (int i = 0; i < 2; i++) {
    double epsilon = 10.0 * DBL_EPSILON;
    if (std::fabs(position[i]) > 1.0) epsilon *= std::fabs(position[i]);
    if (shifted[i] <= 0) {
      shifted[i] = epsilon;
      continue;
    }
    if (shifted[i] >= mapLength[i]) {
      shifted[i] = mapLength[i] - epsilon;
      continue;
    }
  }
  position.x = (shifted.x + mapPosition.x) - vectorToOrigin.x;
  position.y = (shifted.y + mapPosition.y) - vectorToOrigin.y;
}
// getSubmapSizeFromCornerIndeces
inline Idx2 getSubmapSizeFromCornerIndeces(const Idx2& topLeft, const Idx2& bottomRight, const Idx2& bufferSize,
                                           const Idx2& start) {
  Idx2 tl = getIndexFromBufferIndex(topLeft, bufferSize, start);
  Idx2 br = getIndexFromBufferIndex(bottomRight, bufferSize, start);
  return Idx2{br.i0 - tl.i0 + 1, br.i1 - tl.i1 + 1};
}
// getSubmapInformation
struct SubmapInfo {
  Idx2 topLeftIndex;  // buffer index in the parent
  Idx2 size;
  Vec2 position;
  Vec2 length;
  Idx2 requestedIndexInSubmap;
};
inline bool getSubmapInformation(SubmapInfo& out, const Vec2& requestedPosition, const Vec2& requestedLength,
                                 const Vec2& mapLength, const Vec2& mapPosition, double resolution,
                                 const Idx2& bufferSize, const Idx2& bufferStartIndex) {
  // transform = -I: topLeft = p - transform*0.5*L = p + 0.5*L
  Vec2 topLeftPosition{requestedPosition.x + 0.5 * requestedLength.x, requestedPosition.y + 0.5 * requestedLength.y};
  boundPositionToRange(topLeftPosition, mapLength, mapPosition);
  if (!getIndexFromPosition(out.topLeftIndex, topLeftPosition, mapLength, mapPosition, resolution, bufferSize,
                            bufferStartIndex))
    return false;
  Idx2 topLeftIndex = getIndexFromBufferIndex(out.topLeftIndex, bufferSize, bufferStartIndex);

  Vec2 bottomRightPosition{requestedPosition.x - 0.5 * requestedLength.x,
                           requestedPosition.y - 0.5 * requestedLength.y};
  boundPositionToRange(bottomRightPosition, mapLength, mapPosition);
  Idx2 bottomRightIndex;
  if (!getIndexFromPosition(bottomRightIndex, bottomRightPosition, mapLength, mapPosition, resolution, bufferSize,
                            bufferStartIndex))
    return false;
  bottomRightIndex = getIndexFromBufferIndex(bottomRightIndex, bufferSize, bufferStartIndex);

  Vec2 topLeftCorner;
  if (!getPositionFromIndex(topLeftCorner, out.topLeftIndex, mapLength, mapPosition, resolution, bufferSize,
                            bufferStartIndex))
    return false;
  topLeftCorner.x += 0.5 * resolution;  // -= transform * 0.5*res
  topLeftCorner.y += 0.5 * resolution;

  out.size = Idx2{bottomRightIndex.i0 - topLeftIndex.i0 + 1, bottomRightIndex.i1 - topLeftIndex.i1 + 1};
  out.length = Vec2{static_cast<double>(out.size.i0) * resolution, static_cast<double>(out.size.i1) * resolution};
  Vec2 vectorToSubmapOrigin = getVectorToOrigin(out.length);
  out.position = Vec2{topLeftCorner.x - vectorToSubmapOrigin.x, topLeftCorner.y - vectorToSubmapOrigin.y};

  // index of the requested centre inside the submap (submap start index = 0)
  return getIndexFromPosition(out.requestedIndexInSubmap, requestedPosition, out.length, out.position, resolution,
                              out.size, Idx2{0, 0});
}

// ---------------------------------------------------------------- GridMap (the parts the planner touches)
// Unique-cell instrumentation (SURVEY §8d "algorithmic bytes"): when GridMap::touch is set, every at()
// records (layer id, linear buffer index).
struct TouchLog {
  std::unordered_set<uint64_t> cells;
};

class GridMap {
 public:
  Vec2 length{0, 0};
  Vec2 position{0, 0};
  double resolution = 0;
  Idx2 size{0, 0};
  Idx2 startIndex{0, 0};
  // layers: non-owning pointers for the parent map, owning storage for submaps
  std::map<std::string, const float*> layers;
  std::map<std::string, std::shared_ptr<std::vector<float>>> owned;
  mutable TouchLog* touch = nullptr;

  // GridMap::setGeometry(length, resolution, position)
  void setGeometry(const Vec2& len, double res, const Vec2& pos) {
    size.i0 = static_cast<int>(std::round(len.x / res));
    size.i1 = static_cast<int>(std::round(len.y / res));
    resolution = res;
    length = Vec2{static_cast<double>(size.i0) * res, static_cast<double>(size.i1) * res};
    position = pos;
    startIndex = Idx2{0, 0};
  }
  bool exists(const std::string& layer) const { return layers.count(layer) != 0; }
  // GridMap::at(layer, index): data_.at(layer)(index(0), index(1)); throws std::out_of_range on a missing layer
  float at(const std::string& layer, const Idx2& index) const {
    const float* d = layers.at(layer);
    size_t lin = static_cast<size_t>(index.i1) * static_cast<size_t>(size.i0) + static_cast<size_t>(index.i0);
    if (touch) {
      uint64_t lid = static_cast<uint64_t>(std::distance(layers.begin(), layers.find(layer)));
      touch->cells.insert((lid << 56) | static_cast<uint64_t>(lin));
    }
    return d[lin];
  }
  // GridMap::isValid(index, layer) = isfinite(at(layer,index))
  bool isValid(const Idx2& index, const std::string& layer) const { return std::isfinite(at(layer, index)); }
  // GridMap::getIndex / getPosition
  bool getIndex(const Vec2& p, Idx2& index) const {
    return getIndexFromPosition(index, p, length, position, resolution, size, startIndex);
  }
  bool getPosition(const Idx2& index, Vec2& p) const {
    return getPositionFromIndex(p, index, length, position, resolution, size, startIndex);
  }
  // GridMap::getSubmap(position, length, isSuccess): copies every layer, result has startIndex 0
  GridMap getSubmap(const Vec2& p, const Vec2& len, bool& isSuccess) const {
    GridMap sub;
    SubmapInfo info;
    isSuccess = getSubmapInformation(info, p, len, length, position, resolution, size, startIndex);
    if (!isSuccess) return sub;
    sub.setGeometry(info.length, resolution, info.position);
    if (sub.size.i0 != info.size.i0 || sub.size.i1 != info.size.i1) {  // cannot happen (round(n*res/res) == n)
      isSuccess = false;
      return GridMap();
    }
    Idx2 tlU = getIndexFromBufferIndex(info.topLeftIndex, size, startIndex);
    for (auto& kv : layers) {
      auto buf = std::make_shared<std::vector<float>>(static_cast<size_t>(sub.size.i0) * sub.size.i1);
      for (int c = 0; c < sub.size.i1; ++c)
        for (int r = 0; r < sub.size.i0; ++r) {
          Idx2 b = getBufferIndexFromIndex(Idx2{tlU.i0 + r, tlU.i1 + c}, size, startIndex);
          (*buf)[static_cast<size_t>(c) * sub.size.i0 + r] = at(kv.first, b);
        }
      sub.owned[kv.first] = buf;
      sub.layers[kv.first] = buf->data();
    }
    isSuccess = true;
    return sub;
  }
};

// ---------------------------------------------------------------- SubmapIterator (row-major over the window
// in unwrapped index space: i1 fastest, then i0) — incrementIndexForSubmap
class SubmapIterator {
 public:
  SubmapIterator(const GridMap& m, const Idx2& submapStartIndex, const Idx2& submapSize)
      : size_(m.size), startIndex_(m.startIndex), index_(submapStartIndex), submapSize_(submapSize),
        submapStartIndex_(submapStartIndex), submapIndex_{0, 0}, isPastEnd_(false) {}
  const Idx2& operator*() const { return index_; }
  SubmapIterator& operator++() {
    Idx2 t = submapIndex_;
    if (t.i1 + 1 < submapSize_.i1) {
      t.i1++;
    } else {
      t.i0++;
      t.i1 = 0;
    }
    if (!checkIfIndexInRange(t, submapSize_)) {
      isPastEnd_ = true;
      return *this;
    }
    Idx2 u = getIndexFromBufferIndex(submapStartIndex_, size_, startIndex_);
    index_ = getBufferIndexFromIndex(Idx2{u.i0 + t.i0, u.i1 + t.i1}, size_, startIndex_);
    submapIndex_ = t;
    return *this;
  }
  bool isPastEnd() const { return isPastEnd_; }

 private:
  Idx2 size_, startIndex_, index_, submapSize_, submapStartIndex_, submapIndex_;
  bool isPastEnd_;
};

// ---------------------------------------------------------------- CircleIterator
class CircleIterator {
 public:
  CircleIterator(const GridMap& m, const Vec2& center, double radius) : center_(center), radius_(radius) {
    radiusSquare_ = std::pow(radius_, 2);
    mapLength_ = m.length;
    mapPosition_ = m.position;
    resolution_ = m.resolution;
    bufferSize_ = m.size;
    bufferStartIndex_ = m.startIndex;
    Idx2 submapStartIndex, submapBufferSize;
    findSubmapParameters(center, radius, submapStartIndex, submapBufferSize);
    it_.reset(new SubmapIterator(m, submapStartIndex, submapBufferSize));
    if (!isInside()) ++(*this);
  }
  CircleIterator& operator++() {
    ++(*it_);
    if (it_->isPastEnd()) return *this;
    for (; !it_->isPastEnd(); ++(*it_))
      if (isInside()) break;
    return *this;
  }
  const Idx2& operator*() const { return *(*it_); }
  bool isPastEnd() const { return it_->isPastEnd(); }

 private:
  bool isInside() const {
    Vec2 p{0, 0};
    getPositionFromIndex(p, *(*it_), mapLength_, mapPosition_, resolution_, bufferSize_, bufferStartIndex_);
    double dx = p.x - center_.x, dy = p.y - center_.y;
    double squareNorm = dx * dx + dy * dy;
    return squareNorm <= radiusSquare_;
  }
  void findSubmapParameters(const Vec2& center, double radius, Idx2& startIndex, Idx2& bufferSize) const {
    Vec2 topLeft{center.x + radius, center.y + radius};
    Vec2 bottomRight{center.x - radius, center.y - radius};
    boundPositionToRange(topLeft, mapLength_, mapPosition_);
    boundPositionToRange(bottomRight, mapLength_, mapPosition_);
    getIndexFromPosition(startIndex, topLeft, mapLength_, mapPosition_, resolution_, bufferSize_, bufferStartIndex_);
    Idx2 endIndex;
    getIndexFromPosition(endIndex, bottomRight, mapLength_, mapPosition_, resolution_, bufferSize_,
                         bufferStartIndex_);
    bufferSize = getSubmapSizeFromCornerIndeces(startIndex, endIndex, bufferSize_, bufferStartIndex_);
  }
  Vec2 center_;
  double radius_, radiusSquare_;
  Vec2 mapLength_, mapPosition_;
  double resolution_;
  Idx2 bufferSize_, bufferStartIndex_;
  std::unique_ptr<SubmapIterator> it_;
};

// ---------------------------------------------------------------- SpiralIterator
// Upstream does NOT store the map's start index: ring neighbours are taken in raw buffer space (no wrap,
// clipped at the buffer edges) and the last-two-rings distance filter evaluates positions as if
// startIndex == 0.  `wrapCorrect` selects the repaired behaviour: walk unwrapped indices,
// filter with true positions, yield wrapped buffer indices.  With startIndex == (0,0) both coincide.
// Policy for upstream UB: operator++ generates only ONE ring after the current ring empties; if that ring
// is empty too (every cell clipped/filtered) upstream dereferences an empty vector.  Here ring generation
// continues until a non-empty ring or the end (same loop as upstream's constructor).
inline int signum(int v) { return (v > 0) - (v < 0); }
class SpiralIterator {
 public:
  SpiralIterator(const GridMap& m, const Vec2& center, double radius, bool wrapCorrect = false)
      : center_(center), radius_(radius), distance_(0), wrapCorrect_(wrapCorrect) {
    radiusSquare_ = radius_ * radius_;
    mapLength_ = m.length;
    mapPosition_ = m.position;
    resolution_ = m.resolution;
    bufferSize_ = m.size;
    startIndex_ = m.startIndex;
    m.getIndex(center_, indexCenter_);  // upstream: BUFFER index (wrapped into range when startIndex != 0)
    if (wrapCorrect_)                   // repaired mode: the raw unwrapped index (no wrap, may be out of range,
      getIndexFromPosition(indexCenter_, center_, mapLength_, mapPosition_, resolution_, bufferSize_,
                           Idx2{0, 0});  // exactly what upstream computes when startIndex == (0,0))
    nRings_ = static_cast<unsigned int>(std::ceil(radius_ / resolution_));
    if (checkIfIndexInRange(indexCenter_, bufferSize_))
      pointsRing_.push_back(indexCenter_);
    else
      while (pointsRing_.empty() && !isPastEnd()) generateRing();
  }
  SpiralIterator& operator++() {
    pointsRing_.pop_back();
    while (pointsRing_.empty() && !isPastEnd()) generateRing();  // upstream: `if` (see policy above)
    return *this;
  }
  Idx2 operator*() const {
    if (wrapCorrect_) return getBufferIndexFromIndex(pointsRing_.back(), bufferSize_, startIndex_);
    return pointsRing_.back();
  }
  bool isPastEnd() const { return distance_ == nRings_ && pointsRing_.empty(); }
  unsigned int nRings() const { return nRings_; }

 private:
  bool isInside(const Idx2& index) const {
    Vec2 p{0, 0};
    getPositionFromIndex(p, index, mapLength_, mapPosition_, resolution_, bufferSize_, Idx2{0, 0});
    double dx = p.x - center_.x, dy = p.y - center_.y;
    return (dx * dx + dy * dy) <= radiusSquare_;
  }
  void generateRing() {
    distance_++;
    Idx2 point{static_cast<int>(distance_), 0};
    Idx2 pointInMap, normal;
    do {
      pointInMap.i0 = point.i0 + indexCenter_.i0;
      pointInMap.i1 = point.i1 + indexCenter_.i1;
      if (checkIfIndexInRange(pointInMap, bufferSize_)) {
        if (distance_ == nRings_ || distance_ == nRings_ - 1) {
          if (isInside(pointInMap)) pointsRing_.push_back(pointInMap);
        } else {
          pointsRing_.push_back(pointInMap);
        }
      }
      normal.i0 = -signum(point.i1);
      normal.i1 = signum(point.i0);
      if (normal.i0 != 0 &&
          static_cast<int>(std::sqrt(static_cast<double>((point.i0 + normal.i0) * (point.i0 + normal.i0) +
                                                         point.i1 * point.i1))) == static_cast<int>(distance_))
        point.i0 += normal.i0;
      else if (normal.i1 != 0 &&
               static_cast<int>(std::sqrt(static_cast<double>(point.i0 * point.i0 + (point.i1 + normal.i1) *
                                                                                        (point.i1 + normal.i1)))) ==
                   static_cast<int>(distance_))
        point.i1 += normal.i1;
      else {
        point.i0 += normal.i0;
        point.i1 += normal.i1;
      }
    } while (point.i0 != static_cast<int>(distance_) || point.i1 != 0);
  }
  Vec2 center_;
  Idx2 indexCenter_;
  double radius_, radiusSquare_;
  unsigned int nRings_, distance_;
  std::vector<Idx2> pointsRing_;
  Vec2 mapLength_, mapPosition_;
  double resolution_;
  Idx2 bufferSize_, startIndex_;
  bool wrapCorrect_;
};

// ---------------------------------------------------------------- LineIterator (Bresenham in unwrapped
// index space; NO range check on `end`) — the reference's row scan runs it one column past the window.
class LineIterator {
 public:
  LineIterator(const GridMap& m, const Idx2& start, const Idx2& end) {
    bufferSize_ = m.size;
    bufferStartIndex_ = m.startIndex;
    iCell_ = 0;
    index_ = start;
    Idx2 s = getIndexFromBufferIndex(start, bufferSize_, bufferStartIndex_);
    Idx2 e = getIndexFromBufferIndex(end, bufferSize_, bufferStartIndex_);
    Idx2 delta{std::abs(e.i0 - s.i0), std::abs(e.i1 - s.i1)};
    increment1_.i0 = increment2_.i0 = (e.i0 >= s.i0) ? 1 : -1;
    increment1_.i1 = increment2_.i1 = (e.i1 >= s.i1) ? 1 : -1;
    if (delta.i0 >= delta.i1) {
      increment1_.i0 = 0;
      increment2_.i1 = 0;
      denominator_ = delta.i0;
      numerator_ = delta.i0 / 2;
      numeratorAdd_ = delta.i1;
      nCells_ = delta.i0 + 1;
    } else {
      increment2_.i0 = 0;
      increment1_.i1 = 0;
      denominator_ = delta.i1;
      numerator_ = delta.i1 / 2;
      numeratorAdd_ = delta.i0;
      nCells_ = delta.i1 + 1;
    }
  }
  const Idx2& operator*() const { return index_; }
  LineIterator& operator++() {
    numerator_ += numeratorAdd_;
    if (numerator_ >= denominator_) {
      numerator_ -= denominator_;
      Idx2 u = getIndexFromBufferIndex(index_, bufferSize_, bufferStartIndex_);
      index_ = getBufferIndexFromIndex(Idx2{u.i0 + increment1_.i0, u.i1 + increment1_.i1}, bufferSize_,
                                       bufferStartIndex_);
    }
    Idx2 u = getIndexFromBufferIndex(index_, bufferSize_, bufferStartIndex_);
    index_ = getBufferIndexFromIndex(Idx2{u.i0 + increment2_.i0, u.i1 + increment2_.i1}, bufferSize_,
                                     bufferStartIndex_);
    ++iCell_;
    return *this;
  }
  bool isPastEnd() const { return iCell_ >= nCells_; }

 private:
  Idx2 index_, increment1_, increment2_, bufferSize_, bufferStartIndex_;
  int iCell_, nCells_, numerator_, denominator_, numeratorAdd_;
};

// ---------------------------------------------------------------- Polygon::isInside (W. R. Franklin PNPOLY)
class Polygon {
 public:
  std::vector<Vec2> vertices;
  void addVertex(const Vec2& v) { vertices.push_back(v); }
  bool isInside(const Vec2& point) const {
    int cross = 0;
    for (int i = 0, j = static_cast<int>(vertices.size()) - 1; i < static_cast<int>(vertices.size()); j = i++) {
      if (((vertices[i].y > point.y) != (vertices[j].y > point.y)) &&
          (point.x < (vertices[j].x - vertices[i].x) * (point.y - vertices[i].y) / (vertices[j].y - vertices[i].y) +
                         vertices[i].x)) {
        cross++;
      }
    }
    return static_cast<bool>(cross % 2);
  }
};

}  // namespace gm
