// Compiled by tests/test_cpp_host.py against tests/gtsam_stub (a type stub, not GTSAM): include/ndt2d_gtsam.hpp parses,
// the factor carries the measured pose, and the information matrix is the Hessian rotated into the measurement's frame.
#include <cmath>
#include <cstdio>

#include "ndt2d_gtsam.hpp"

#ifndef NDT2D_HAVE_GTSAM
#error "the gtsam stub was not found on the include path"
#endif

int main()
{
    ndt2d::Result r{};
    r.pose[0] = 1.0; r.pose[1] = 2.0; r.pose[2] = M_PI / 2;      // a 90 degree relative pose
    const double H[9] = {100.0, 0.0, 3.0, 0.0, 1.0, 5.0, 3.0, 5.0, 50.0}; // stiff along the target frame's x axis
    for (int i = 0; i < 9; ++i) r.hessian[i] = H[i];
    r.status = NDT2D_CONVERGED;
    gtsam::NonlinearFactorGraph g;
    if (!ndt2d::addBetweenFactor(g, 7, 8, r) || g.size() != 1) return 1;
    const auto &f = *g.factors[0];
    const gtsam::Matrix3 &I = f.model->information;
    // local x axis = target y axis at 90 degrees: the stiff direction must move to the local y axis, the couplings follow
    const double exp[9] = {1.0, 0.0, 5.0, 0.0, 100.0, -3.0, 5.0, -3.0, 50.0};
    for (int i = 0; i < 9; ++i)
        if (std::fabs(I.m[i] - exp[i]) > 1e-12) { std::printf("entry %d: %g != %g\n", i, I.m[i], exp[i]); return 2; }
    if (f.key1 != 7 || f.key2 != 8 || f.measured.x() != 1.0 || f.measured.theta() != r.pose[2]) return 3;
    r.status = NDT2D_MAX_ITERATIONS;
    if (ndt2d::addBetweenFactor(g, 8, 9, r) || g.size() != 1) return 4;   // not converged: no factor
    std::puts("gtsam glue ok");
    return 0;
}
