// ORACLE / TEST INFRASTRUCTURE ONLY — header SHIM so that the reference's own FootholdPlanner.cpp can be compiled in
// place (from /root/reference, never copied) without ROS / grid_map / Eigen.  It exposes the slice of the
// grid_map_core + Eigen API the reference touches, implemented on the restatement in oracle/gm_core.hpp.
// grid_map_core itself is a third-party dependency that is NOT vendored in the reference; see gm_core.hpp.
#pragma once
#include <cmath>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../gm_core.hpp"

namespace Eigen {

struct Vector2d {
  double v[2];
  Vector2d() { v[0] = v[1] = 0; }
  Vector2d(double a, double b) { v[0] = a; v[1] = b; }
  double& x() { return v[0]; }
  double& y() { return v[1]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  const Vector2d& transpose() const { return *this; }
};
typedef Vector2d Array2d;
inline std::ostream& operator<<(std::ostream& o, const Vector2d& a) { return o << a.v[0] << " " << a.v[1]; }

struct Array2i {
  int v[2];
  Array2i() { v[0] = v[1] = 0; }
  Array2i(int a, int b) { v[0] = a; v[1] = b; }
  int& x() { return v[0]; }
  int& y() { return v[1]; }
  int x() const { return v[0]; }
  int y() const { return v[1]; }
  int& operator()(int i) { return v[i]; }
  int operator()(int i) const { return v[i]; }
  int& operator()(int i, int) { return v[i]; }  // (row, 0) on a 2x1 array
  int operator()(int i, int) const { return v[i]; }
  int& operator[](int i) { return v[i]; }
  int operator[](int i) const { return v[i]; }
  const Array2i& transpose() const { return *this; }
};
inline std::ostream& operator<<(std::ostream& o, const Array2i& a) { return o << a.v[0] << " " << a.v[1]; }

struct Vector3d {
  double v[3];
  Vector3d() { v[0] = v[1] = v[2] = 0; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};

class VectorXd {
 public:
  std::vector<double> d;
  VectorXd() {}
  explicit VectorXd(int n) : d(n, 0.0) {}
  int size() const { return static_cast<int>(d.size()); }
  double& operator()(int i) { return d[i]; }
  double operator()(int i) const { return d[i]; }
  double& operator[](int i) { return d[i]; }
  double operator[](int i) const { return d[i]; }
  struct Comma {
    VectorXd* p;
    int i;
    Comma& operator,(double x) {
      p->d[i++] = x;
      return *this;
    }
  };
  Comma operator<<(double x) {
    d[0] = x;
    Comma c;
    c.p = this;
    c.i = 1;
    return c;
  }
};

template <typename T>
class MatrixX {
 public:
  int r, c;
  std::vector<T> d;  // column-major
  MatrixX() : r(0), c(0) {}
  MatrixX(int rows, int cols) : r(rows), c(cols), d(static_cast<size_t>(rows) * cols, T()) {}
  int rows() const { return r; }
  int cols() const { return c; }
  T& operator()(int i, int j) { return d[static_cast<size_t>(j) * r + i]; }
  const T& operator()(int i, int j) const { return d[static_cast<size_t>(j) * r + i]; }
  T& operator()(int i) { return d[i]; }
  const T& operator()(int i) const { return d[i]; }
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
};
typedef MatrixX<int> MatrixXi;
typedef MatrixX<float> MatrixXf;

}  // namespace Eigen

namespace grid_map {

typedef Eigen::MatrixXf Matrix;
typedef Eigen::Vector2d Position;
typedef Eigen::Vector2d Vector;
typedef Eigen::Array2i Index;
typedef Eigen::Array2i Size;
typedef Eigen::Array2d Length;

inline gm::Vec2 toGm(const Position& p) { return gm::Vec2{p.x(), p.y()}; }
inline gm::Idx2 toGm(const Index& i) { return gm::Idx2{i.x(), i.y()}; }

// grid_map::GridMap: owning layers (Eigen::MatrixXf, column-major), geometry, circular-buffer start index.
class GridMap {
 public:
  GridMap() : resolution_(0) {}
  // --- test-driver side (what GridMapRosConverter::fromMessage would fill)
  void setRaw(int rows, int cols, double res, double lenx, double leny, double px, double py, int s0, int s1) {
    size_ = Size(rows, cols);
    resolution_ = res;
    length_ = Length(lenx, leny);
    position_ = Position(px, py);
    startIndex_ = Index(s0, s1);
  }
  void add(const std::string& layer, const float* colmajor) {
    Matrix m(size_.x(), size_.y());
    std::copy(colmajor, colmajor + m.d.size(), m.d.begin());
    data_[layer] = m;
    layers_.push_back(layer);
  }
  // --- grid_map API used by the reference
  const Length& getLength() const { return length_; }
  const Position& getPosition() const { return position_; }
  double getResolution() const { return resolution_; }
  const Size& getSize() const { return size_; }
  const Index& getStartIndex() const { return startIndex_; }
  const std::string& getFrameId() const { return frameId_; }
  void setFrameId(const std::string& f) { frameId_ = f; }
  const std::vector<std::string>& getLayers() const { return layers_; }
  bool exists(const std::string& layer) const { return data_.count(layer) != 0; }
  Matrix& operator[](const std::string& layer) { return data_.at(layer); }
  const Matrix& operator[](const std::string& layer) const { return data_.at(layer); }
  // Eigen (NDEBUG) would read whatever lies past the buffer for an out-of-range index; the reference does exactly
  // that once per row in its centroid scan (LineIterator end index = size(1), FootholdPlanner.cpp:1720,1733-1736).
  // Deterministic stand-in for that heap garbage, matching policy C.1 of the restatement: NaN ("not below").
  float& at(const std::string& layer, const Index& index) {
    Matrix& m = data_.at(layer);
    if (index(0) < 0 || index(1) < 0 || index(0) >= m.rows() || index(1) >= m.cols()) {
      static float oob;
      oob = std::nanf("");
      return oob;
    }
    return m(index(0), index(1));
  }
  float at(const std::string& layer, const Index& index) const {
    const Matrix& m = data_.at(layer);
    if (index(0) < 0 || index(1) < 0 || index(0) >= m.rows() || index(1) >= m.cols()) return std::nanf("");
    return m(index(0), index(1));
  }
  float& atPosition(const std::string& layer, const Position& p) {
    Index i;
    if (!getIndex(p, i)) throw std::out_of_range("GridMap::atPosition(...) : Position is out of range.");
    return at(layer, i);
  }
  bool isValid(const Index& index, const std::string& layer) const { return std::isfinite(at(layer, index)); }
  bool getIndex(const Position& p, Index& index) const {
    gm::Idx2 i;
    bool ok = gm::getIndexFromPosition(i, toGm(p), toGm(length_), toGm(position_), resolution_, toGm(size_), toGm(startIndex_));
    index = Index(i.i0, i.i1);
    return ok;
  }
  bool getPosition(const Index& index, Position& p) const {
    gm::Vec2 q{p.x(), p.y()};
    bool ok = gm::getPositionFromIndex(q, toGm(index), toGm(length_), toGm(position_), resolution_, toGm(size_), toGm(startIndex_));
    if (ok) p = Position(q.x, q.y);
    return ok;
  }
  // GridMap::getSubmap(position, length, isSuccess)
  GridMap getSubmap(const Position& p, const Length& len, bool& isSuccess) const {
    GridMap sub;
    gm::SubmapInfo info;
    isSuccess = gm::getSubmapInformation(info, toGm(p), toGm(len), toGm(length_), toGm(position_), resolution_,
                                         toGm(size_), toGm(startIndex_));
    if (!isSuccess) return sub;
    // setGeometry(SubmapGeometry): size = round(length/res)
    const int r = static_cast<int>(std::round(info.length.x / resolution_));
    const int c = static_cast<int>(std::round(info.length.y / resolution_));
    sub.size_ = Size(r, c);
    sub.resolution_ = resolution_;
    sub.length_ = Length(static_cast<double>(r) * resolution_, static_cast<double>(c) * resolution_);
    sub.position_ = Position(info.position.x, info.position.y);
    sub.startIndex_ = Index(0, 0);
    sub.frameId_ = frameId_;
    gm::Idx2 tlU = gm::getIndexFromBufferIndex(info.topLeftIndex, toGm(size_), toGm(startIndex_));
    for (const std::string& name : layers_) {
      Matrix m(r, c);
      const Matrix& src = data_.at(name);
      for (int j = 0; j < c; ++j)
        for (int i = 0; i < r; ++i) {
          gm::Idx2 b = gm::getBufferIndexFromIndex(gm::Idx2{tlU.i0 + i, tlU.i1 + j}, toGm(size_), toGm(startIndex_));
          m(i, j) = src(b.i0, b.i1);
        }
      sub.data_[name] = m;
      sub.layers_.push_back(name);
    }
    isSuccess = true;
    return sub;
  }
  gm::GridMap view() const {  // geometry-only view for the iterator restatements
    gm::GridMap g;
    g.length = toGm(length_);
    g.position = toGm(position_);
    g.resolution = resolution_;
    g.size = toGm(size_);
    g.startIndex = toGm(startIndex_);
    return g;
  }

 private:
  std::map<std::string, Matrix> data_;
  std::vector<std::string> layers_;
  Length length_;
  Position position_;
  double resolution_;
  Size size_;
  Index startIndex_;
  std::string frameId_;
};

// Iterators: thin adapters over the restatements in gm_core.hpp (operator* yields grid_map::Index)
class GridMapIterator {
 public:
  explicit GridMapIterator(const GridMap& m) : rows_(m.getSize()(0)), n_(m.getSize()(0) * m.getSize()(1)), lin_(0) {}
  GridMapIterator& operator++() {
    ++lin_;
    return *this;
  }
  bool isPastEnd() const { return lin_ >= n_; }
  int getLinearIndex() const { return lin_; }
  Index operator*() const { return Index(lin_ % rows_, lin_ / rows_); }

 private:
  int rows_, n_, lin_;
};
class CircleIterator {
 public:
  CircleIterator(const GridMap& m, const Position& c, double radius) : it_(m.view(), toGm(c), radius) {}
  CircleIterator& operator++() {
    ++it_;
    return *this;
  }
  bool isPastEnd() const { return it_.isPastEnd(); }
  Index operator*() const {
    const gm::Idx2& i = *it_;
    return Index(i.i0, i.i1);
  }

 private:
  gm::CircleIterator it_;
};
class SpiralIterator {
 public:
  SpiralIterator(const GridMap& m, const Position& c, double radius) : it_(m.view(), toGm(c), radius, wrapCorrect()) {}
  SpiralIterator& operator++() {
    ++it_;
    return *this;
  }
  bool isPastEnd() const { return it_.isPastEnd(); }
  Index operator*() const {
    gm::Idx2 i = *it_;
    return Index(i.i0, i.i1);
  }
  static bool& wrapCorrect() {  // test-driver switch: false = upstream behaviour (ignores the start index)
    static bool w = false;
    return w;
  }

 private:
  gm::SpiralIterator it_;
};
class LineIterator {
 public:
  LineIterator(const GridMap& m, const Index& s, const Index& e) : it_(m.view(), toGm(s), toGm(e)) {}
  LineIterator& operator++() {
    ++it_;
    return *this;
  }
  bool isPastEnd() const { return it_.isPastEnd(); }
  Index operator*() const {
    const gm::Idx2& i = *it_;
    return Index(i.i0, i.i1);
  }

 private:
  gm::LineIterator it_;
};

class Polygon {
 public:
  void addVertex(const Position& v) { p_.addVertex(toGm(v)); }
  void setFrameId(const std::string& f) { frameId_ = f; }
  const std::string& getFrameId() const { return frameId_; }
  bool isInside(const Position& p) const { return p_.isInside(toGm(p)); }
  Position getCentroid() const {  // Polygon::getCentroid (shoelace); only printed by the reference
    const std::vector<gm::Vec2>& v = p_.vertices;
    double a = 0, cx = 0, cy = 0;
    for (size_t i = 0; i < v.size(); ++i) {
      const gm::Vec2& p0 = v[i];
      const gm::Vec2& p1 = v[(i + 1) % v.size()];
      double cr = p0.x * p1.y - p1.x * p0.y;
      a += cr;
      cx += (p0.x + p1.x) * cr;
      cy += (p0.y + p1.y) * cr;
    }
    a *= 0.5;
    return Position(cx / (6.0 * a), cy / (6.0 * a));
  }
  const std::vector<gm::Vec2>& vertices() const { return p_.vertices; }

 private:
  gm::Polygon p_;
  std::string frameId_;
};

}  // namespace grid_map
