// ORACLE / TEST INFRASTRUCTURE ONLY — stand-in for NLopt (absent here).  The "opt" chain of the reference
// (FootholdPlanner.cpp:913-1319) is outside the accelerated path and its result is never part of the
// plan_global_footholds response; this stub leaves the start point unchanged so the chain keeps running.
#pragma once
#include <stdexcept>
#include <vector>
namespace nlopt {
enum result { FAILURE = -1, SUCCESS = 1, STOPVAL_REACHED = 2, FTOL_REACHED = 3, XTOL_REACHED = 4 };
enum algorithm { LN_COBYLA = 25, LN_BOBYQA = 34, LD_MMA = 24, NUM_ALGORITHMS };
typedef double (*vfunc)(const std::vector<double>& x, std::vector<double>& grad, void* data);
class opt {
 public:
  opt() : n_(0) {}
  opt(const char*, unsigned n) : n_(n) {}
  opt(algorithm, unsigned n) : n_(n) {}
  void set_lower_bounds(const std::vector<double>&) {}
  void set_upper_bounds(const std::vector<double>&) {}
  void set_min_objective(vfunc f, void* d) { f_ = f; d_ = d; }
  void set_max_objective(vfunc f, void* d) { f_ = f; d_ = d; }
  void add_inequality_constraint(vfunc, void*, double = 0) {}
  void add_equality_constraint(vfunc, void*, double = 0) {}
  void remove_inequality_constraints() {}
  void set_xtol_rel(double) {}
  void set_ftol_rel(double) {}
  void set_maxeval(int) {}
  void get_initial_step(const std::vector<double>& x, std::vector<double>& dx) const { dx.assign(x.size(), 1.0); }
  void set_initial_step(const std::vector<double>&) {}
  result optimize(std::vector<double>& x, double& minf) {
    std::vector<double> g;
    minf = f_ ? f_(x, g, d_) : 0.0;
    return XTOL_REACHED;
  }

 private:
  unsigned n_;
  vfunc f_ = 0;
  void* d_ = 0;
};
}  // namespace nlopt
