// ORACLE / TEST INFRASTRUCTURE ONLY — minimal stand-in for roscpp so the reference's FootholdPlanner.cpp compiles
// in place.  Parameters come from a process-wide table filled by the test driver; publishers/subscribers are no-ops.
#pragma once
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace ros {

struct Time {
  double t;
  Time() : t(0) {}
  static Time now() { return Time(); }
  double toSec() const { return t; }
};
struct Duration {
  double d;
  Duration() : d(0) {}
  explicit Duration(double s) : d(s) {}
  bool sleep() const { return true; }
};

class Publisher {
 public:
  template <typename T>
  void publish(const T&) const {}
};
class Subscriber {};
class ServiceServer {};

struct ParamTable {
  std::map<std::string, double> num;
  std::map<std::string, std::string> str;
  static ParamTable& get() {
    static ParamTable t;
    return t;
  }
};

class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <typename T>
  bool param(const std::string& name, T& out, const T& def) const {
    std::map<std::string, double>::const_iterator it = ParamTable::get().num.find(name);
    if (it == ParamTable::get().num.end()) {
      out = def;
      return false;
    }
    out = static_cast<T>(it->second);
    return true;
  }
  bool param(const std::string& name, std::string& out, const std::string& def) const {
    std::map<std::string, std::string>::const_iterator it = ParamTable::get().str.find(name);
    if (it == ParamTable::get().str.end()) {
      out = def;
      return false;
    }
    out = it->second;
    return true;
  }
  template <typename M, typename C>
  Subscriber subscribe(const std::string&, unsigned, void (C::*)(M), C*) { return Subscriber(); }
  template <typename M>
  Publisher advertise(const std::string&, unsigned, bool = false) { return Publisher(); }
  template <typename C, typename Req, typename Res>
  ServiceServer advertiseService(const std::string&, bool (C::*)(Req&, Res&), C*) { return ServiceServer(); }
  void shutdown() {}
};

inline void init(int&, char**, const std::string&) {}
inline bool ok() { return true; }

}  // namespace ros

// logging: swallowed (the reference prints a lot; the driver also redirects stdout)
#define ROS_INFO(...) do { if (0) printf(__VA_ARGS__); } while (0)
#define ROS_WARN(...) do { if (0) printf(__VA_ARGS__); } while (0)
#define ROS_ERROR(...) do { if (0) printf(__VA_ARGS__); } while (0)
#define ROS_FATAL(...) do { if (0) printf(__VA_ARGS__); } while (0)
#define ROS_DEBUG(...) do { if (0) printf(__VA_ARGS__); } while (0)
#define ROS_INFO_STREAM(x) do { if (0) { std::stringstream ss__; ss__ << x; } } while (0)
#define ROS_WARN_STREAM(x) do { if (0) { std::stringstream ss__; ss__ << x; } } while (0)
#define ROS_ERROR_STREAM(x) do { if (0) { std::stringstream ss__; ss__ << x; } } while (0)
