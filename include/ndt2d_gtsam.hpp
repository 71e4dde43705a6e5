// ndt2d_gtsam.hpp — glue from ndt2d results to GTSAM factors.
//
// Compiles to nothing unless GTSAM's headers are on the include path: GTSAM, Eigen and Boost are absent from
// the build container, so there it is compiled against a type stub only (tests/gtsam_stub/, tests/test_cpp_host.py:
// the header parses, the factor receives the rotated information matrix). The sparse incremental
// iSAM2 back end stays on the host in GTSAM, as BASELINE.json's north_star requires; nothing here replaces it.
// Reference glue code replaced: none citable (/root/reference/README.md:1 is the whole mount).
#ifndef NDT2D_GTSAM_HPP
#define NDT2D_GTSAM_HPP

#include "ndt2d.hpp"

#if defined(__has_include)
#if __has_include(<gtsam/geometry/Pose2.h>) && __has_include(<gtsam/slam/BetweenFactor.h>)
#define NDT2D_HAVE_GTSAM 1
#endif
#endif

#ifdef NDT2D_HAVE_GTSAM
#include <gtsam/geometry/Pose2.h>
#include <gtsam/linear/NoiseModel.h>
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/slam/BetweenFactor.h>

namespace ndt2d {

inline gtsam::Pose2 toPose2(const Result &r) { return gtsam::Pose2(r.pose[0], r.pose[1], r.pose[2]); }

// The Hessian of f = -score at the optimum as the information matrix of the relative pose, in the frame the factor
// uses: BetweenFactor<Pose2> measures the error in the local (measurement) frame, the matcher's Hessian is in the target
// frame, so it is rotated by J = blockdiag(R(theta), 1) (informationInLocalFrame in ndt2d.hpp). `scale` maps the NDT score
// scale to the caller's noise scale (1 keeps it as returned).
inline gtsam::SharedNoiseModel toNoiseModel(const Result &r, double scale = 1.0)
{
    double info[9];
    informationInLocalFrame(r, info);
    gtsam::Matrix3 H;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) H(i, j) = scale * info[3 * i + j];
    return gtsam::noiseModel::Gaussian::Information(H);
}

// Adds BetweenFactor<Pose2>(from, to, measured pose, information from the Hessian) when the align converged.
inline bool addBetweenFactor(gtsam::NonlinearFactorGraph &graph, gtsam::Key from, gtsam::Key to, const Result &r,
                             double scale = 1.0)
{
    if (r.status != NDT2D_CONVERGED) return false;
    graph.emplace_shared<gtsam::BetweenFactor<gtsam::Pose2>>(from, to, toPose2(r), toNoiseModel(r, scale));
    return true;
}

} // namespace ndt2d
#endif // NDT2D_HAVE_GTSAM

#endif // NDT2D_GTSAM_HPP
