// Host-side example (C++17): how a 2D SLAM front end drives the matcher through include/ndt2d.hpp.
//   1. scan-to-scan odometry: setTarget(previous scan), align(current scan, motion prior)
//   2. the returned Hessian is the information matrix of the relative-pose factor
//   3. loop-closure check: sweep a lattice of hypotheses around a candidate, refine the best
// Build: g++ -std=c++17 -Iinclude examples/odometry_demo.cpp -Lgtsam_ndt_b200 -lndt2d -Wl,-rpath,$PWD/gtsam_ndt_b200
// The world is a 20 x 12 m room with one pillar; scans are exact ray casts (no noise) so the demo is
// self-contained. GTSAM itself is not needed here: the factor construction is shown as a comment.
#include <cmath>
#include <cstdio>
#include <vector>

#include "ndt2d.hpp"

namespace {

double ray_box(double px, double py, double dx, double dy, double x0, double y0, double x1, double y1, bool inside)
{
    double tx0 = (x0 - px) / dx, tx1 = (x1 - px) / dx, ty0 = (y0 - py) / dy, ty1 = (y1 - py) / dy;
    double tn = std::fmax(std::fmin(tx0, tx1), std::fmin(ty0, ty1)), tf = std::fmin(std::fmax(tx0, tx1), std::fmax(ty0, ty1));
    if (inside) return tf;
    return (tn <= tf && tn > 0) ? tn : 1e30;
}

std::vector<ndt2d::Point2f> cast_scan(const ndt2d::Pose2d &p, int beams)
{
    std::vector<ndt2d::Point2f> out;
    for (int i = 0; i < beams; ++i) {
        double a = -M_PI + 2.0 * M_PI * i / beams, w = p.theta + a;
        double dx = std::cos(w), dy = std::sin(w);
        double r = std::fmin(ray_box(p.x, p.y, dx, dy, -10, -6, 10, 6, true), ray_box(p.x, p.y, dx, dy, 2, 1, 3.5, 2.5, false));
        out.push_back({static_cast<float>(r * std::cos(a)), static_cast<float>(r * std::sin(a))});
    }
    return out;
}

} // namespace

int main()
{
    try {
        ndt2d::Matcher ndt(0);
        ndt.setResolutions({2.0f, 1.0f, 0.5f});
        std::vector<ndt2d::Pose2d> truth;
        for (int k = 0; k < 6; ++k) truth.push_back({-4.0 + 0.12 * k, -1.0 + 0.05 * k, 0.10 + 0.02 * k});
        ndt2d::Pose2d est = truth[0];
        auto prev = cast_scan(truth[0], 720);
        for (size_t k = 1; k < truth.size(); ++k) {
            auto cur = cast_scan(truth[k], 720);
            ndt.setTarget(prev);                                   // "set target scan"
            ndt2d::Result r = ndt.align(cur, ndt2d::Pose2d{});     // align(scan, initial pose) -> pose, score, Hessian
            // GTSAM side (not compiled here):
            //   auto info  = gtsam::Matrix3(Eigen::Map<const Eigen::Matrix<double,3,3,Eigen::RowMajor>>(r.hessian));
            //   auto model = gtsam::noiseModel::Gaussian::Information(info);
            //   graph.emplace_shared<gtsam::BetweenFactor<gtsam::Pose2>>(X(k-1), X(k), gtsam::Pose2(r.pose[0], r.pose[1], r.pose[2]), model);
            double c = std::cos(est.theta), s = std::sin(est.theta);
            est = {est.x + c * r.pose[0] - s * r.pose[1], est.y + s * r.pose[0] + c * r.pose[1], est.theta + r.pose[2]};
            std::printf("step %zu: rel (%.4f %.4f %.5f) status %d iters %d score %.1f | est (%.3f %.3f %.4f) truth (%.3f %.3f %.4f)\n", k,
                        r.pose[0], r.pose[1], r.pose[2], r.status, r.iterations, r.score, est.x, est.y, est.theta, truth[k].x,
                        truth[k].y, truth[k].theta);
            prev = cur;
        }
        // loop-closure candidate: search +-1 m / +-10 deg around a rough guess against the first scan's frame
        ndt.setTarget(cast_scan(truth[0], 720));
        auto scan = cast_scan(truth[5], 720);
        std::vector<float> hyp;
        for (int ix = -5; ix <= 5; ++ix)
            for (int iy = -5; iy <= 5; ++iy)
                for (int it = -5; it <= 5; ++it) {
                    hyp.push_back(0.2f * ix); hyp.push_back(0.2f * iy); hyp.push_back(0.035f * it);
                }
        auto best = ndt.relocalize(scan, hyp, 3, 0);
        double c0 = std::cos(truth[0].theta), s0 = std::sin(truth[0].theta), dx = truth[5].x - truth[0].x, dy = truth[5].y - truth[0].y;
        std::printf("loop closure: best (%.3f %.3f %.4f) score %.1f | truth (%.3f %.3f %.4f)\n", best[0].pose[0], best[0].pose[1],
                    best[0].pose[2], best[0].score, c0 * dx + s0 * dy, -s0 * dx + c0 * dy, truth[5].theta - truth[0].theta);
        double err = std::hypot(est.x - truth.back().x, est.y - truth.back().y);
        std::printf("final odometry error %.4f m\n", err);
        return err < 0.05 ? 0 : 1;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ndt2d demo failed: %s\n", e.what());
        return 2;
    }
}
