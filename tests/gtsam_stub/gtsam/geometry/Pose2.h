// TYPE STUB for tests only (tests/test_cpp_host.py): just enough of gtsam's names for include/ndt2d_gtsam.hpp to compile
// and for the test to read back what the glue handed over. It is NOT GTSAM, solves nothing and ships with nothing.
#pragma once
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>
namespace gtsam {
using Key = std::uint64_t;
struct Matrix3 {
    double m[9] = {};
    double &operator()(int i, int j) { return m[3 * i + j]; }
    double operator()(int i, int j) const { return m[3 * i + j]; }
};
class Pose2 {
  public:
    Pose2(double x, double y, double theta) : x_(x), y_(y), t_(theta) {}
    double x() const { return x_; }
    double y() const { return y_; }
    double theta() const { return t_; }
  private:
    double x_, y_, t_;
};
} // namespace gtsam
