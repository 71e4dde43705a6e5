// SLAM front end (C++17 host code over include/ndt2d.hpp): the part of BASELINE configs[4] that exists without GTSAM.
//   GPU NDT odometry (scan-to-scan, pyramid) -> relative-pose factors with the returned Hessian as information
//   GPU loop-closure search (hypothesis sweep + top-k refinement, ndt2d_relocalize) -> loop-closure factors
//   pose graph written in g2o format (VERTEX_SE2 / EDGE_SE2): gtsam::readG2o() + ISAM2::update() consume it as is.
// The sparse incremental iSAM2 back end stays on the host in GTSAM (north_star); GTSAM, Eigen and Boost are absent
// from this container, so nothing here solves the graph: the program only checks its own factors against the truth.
// Build: g++ -std=c++17 -Iinclude examples/slam_frontend.cpp -Lgtsam_ndt_b200 -lndt2d -Wl,-rpath,$PWD/gtsam_ndt_b200
// Run:   examples/slam_frontend [out.g2o]
// Reference front end replaced: none citable (/root/reference/README.md:1 is the whole mount).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ndt2d.hpp"

namespace {

struct Box { double x0, y0, x1, y1; };

// room 24 x 16 m with four obstacles; the robot drives a closed loop around the middle one
const Box kRoom{-12, -8, 12, 8};
const Box kObstacles[] = {{-2, -1.5, 2, 1.5}, {-9, 4, -7, 6}, {7, -6, 9.5, -4.5}, {5, 4.5, 6, 7}};

double ray_box(double px, double py, double dx, double dy, const Box &b, bool inside)
{
    double tx0 = (b.x0 - px) / dx, tx1 = (b.x1 - px) / dx, ty0 = (b.y0 - py) / dy, ty1 = (b.y1 - py) / dy;
    double tn = std::fmax(std::fmin(tx0, tx1), std::fmin(ty0, ty1)), tf = std::fmin(std::fmax(tx0, tx1), std::fmax(ty0, ty1));
    if (inside) return tf;
    return (tn <= tf && tn > 0) ? tn : 1e30;
}

// deterministic noise: SplitMix64 -> sum of four uniforms (variance 1/3), scaled to sigma
std::uint64_t splitmix(std::uint64_t &s)
{
    std::uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
double noise(std::uint64_t &s, double sigma)
{
    double u = 0;
    for (int i = 0; i < 4; ++i) u += (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    return u * std::sqrt(3.0) * sigma;
}

std::vector<ndt2d::Point2f> cast_scan(const ndt2d::Pose2d &p, int beams, std::uint64_t seed)
{
    std::vector<ndt2d::Point2f> out;
    out.reserve(beams);
    for (int i = 0; i < beams; ++i) {
        double a = -0.75 * M_PI + 1.5 * M_PI * i / (beams - 1), w = p.theta + a;   // 270 degree field of view
        double dx = std::cos(w), dy = std::sin(w);
        double r = ray_box(p.x, p.y, dx, dy, kRoom, true);
        for (const Box &b : kObstacles) r = std::fmin(r, ray_box(p.x, p.y, dx, dy, b, false));
        r += noise(seed, 0.01);
        out.push_back({static_cast<float>(r * std::cos(a)), static_cast<float>(r * std::sin(a))});
    }
    return out;
}

ndt2d::Pose2d compose(const ndt2d::Pose2d &a, const ndt2d::Pose2d &b)   // a (+) b
{
    double c = std::cos(a.theta), s = std::sin(a.theta);
    return {a.x + c * b.x - s * b.y, a.y + s * b.x + c * b.y, a.theta + b.theta};
}
ndt2d::Pose2d between(const ndt2d::Pose2d &a, const ndt2d::Pose2d &b)   // a^-1 (+) b
{
    double c = std::cos(a.theta), s = std::sin(a.theta), dx = b.x - a.x, dy = b.y - a.y;
    double dth = std::remainder(b.theta - a.theta, 2.0 * M_PI);
    return {c * dx + s * dy, -s * dx + c * dy, dth};
}

struct Edge { int i, j; ndt2d::Pose2d z; double info[6]; bool loop; };

// The matcher's pose is the scan frame expressed in the target frame, so with target = scan i and source = scan j the
// result IS the relative pose z_ij. Its Hessian (of f = -score, order x, y, theta) is taken in the target frame; EDGE_SE2
// (like gtsam::BetweenFactor<Pose2>) measures the error in the local frame of the measurement, hence the rotation
// J^T H J, J = blockdiag(R(theta), 1), done by ndt2d::informationInLocalFrame.
Edge make_edge(int i, int j, const ndt2d::Result &r, bool loop)
{
    Edge e{i, j, {r.pose[0], r.pose[1], r.pose[2]}, {}, loop};
    double H[9];
    ndt2d::informationInLocalFrame(r, H);
    e.info[0] = H[0]; e.info[1] = H[1]; e.info[2] = H[2];
    e.info[3] = H[4]; e.info[4] = H[5]; e.info[5] = H[8];
    return e;
}

} // namespace

int main(int argc, char **argv)
{
    const std::string out_path = argc > 1 ? argv[1] : "slam_frontend.g2o";
    try {
        const int N = 96, BEAMS = 720;
        ndt2d::Matcher ndt(0);
        ndt.setResolutions({2.0f, 1.0f, 0.5f});

        // truth: an ellipse around the central obstacle, heading along the tangent; the last pose closes the loop
        std::vector<ndt2d::Pose2d> truth;
        for (int k = 0; k <= N; ++k) {
            double a = 2.0 * M_PI * k / N;
            truth.push_back({7.0 * std::cos(a), 4.5 * std::sin(a), std::atan2(4.5 * std::cos(a), -7.0 * std::sin(a))});
        }
        std::vector<std::vector<ndt2d::Point2f>> scans;
        for (int k = 0; k <= N; ++k) scans.push_back(cast_scan(truth[k], BEAMS, 1000 + k));

        // 1. odometry: target = scan k-1, source = scan k, prior = the previous relative motion (constant velocity)
        std::vector<Edge> edges;
        std::vector<ndt2d::Pose2d> est{truth[0]};
        std::vector<ndt2d::Pose2d> priors;
        std::vector<ndt2d::Result> sequential;
        ndt2d::Pose2d prior{};
        double worst_rel = 0;
        int bad = 0;
        ndt.setTarget(scans[0]);                         // warm-up: the handle's buffers are allocated on first use
        (void)ndt.align(scans[1], prior);
        const auto t_odo = std::chrono::steady_clock::now();
        for (int k = 1; k <= N; ++k) {
            ndt.setTarget(scans[k - 1]);
            ndt2d::Result r = ndt.align(scans[k], prior);
            priors.push_back(prior);
            sequential.push_back(r);
            if (r.status != NDT2D_CONVERGED) ++bad;
            edges.push_back(make_edge(k - 1, k, r, false));
            prior = edges.back().z;
            est.push_back(compose(est.back(), prior));
            ndt2d::Pose2d t = between(truth[k - 1], truth[k]);
            worst_rel = std::fmax(worst_rel, std::hypot(t.x - prior.x, t.y - prior.y));
        }
        const double odo_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_odo).count();
        double drift = std::hypot(est[N].x - truth[N].x, est[N].y - truth[N].y);

        // 1b. the same odometry as ONE batched call (a log processed offline, or loop-closure candidate pairs): every
        //     target's grid is built and every pair aligned by the GPU in one go; the results are the sequential ones, bit for bit
        std::vector<ndt2d::Point2f> packed;
        std::vector<std::int64_t> offsets{0};
        std::vector<std::pair<std::int32_t, std::int32_t>> pairs;
        for (int k = 0; k <= N; ++k) {
            packed.insert(packed.end(), scans[k].begin(), scans[k].end());
            offsets.push_back(static_cast<std::int64_t>(packed.size()));
            if (k > 0) pairs.emplace_back(k - 1, k);
        }
        ndt.alignPairs(packed, offsets, pairs, priors);   // warm-up (allocations)
        const auto t_batch = std::chrono::steady_clock::now();
        std::vector<ndt2d::Result> batched = ndt.alignPairs(packed, offsets, pairs, priors);
        const double batch_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_batch).count();
        int differing = 0;
        for (int k = 0; k < N; ++k)
            if (std::memcmp(&batched[k], &sequential[k], sizeof(ndt2d::Result)) != 0) ++differing;

        // 2. loop closure: the last scan against the first; the search lattice is centred on the drifted estimate
        ndt.setTarget(scans[0]);
        ndt2d::Pose2d guess = between(est[0], est[N]);
        std::vector<float> hyp;
        for (int ix = -10; ix <= 10; ++ix)
            for (int iy = -10; iy <= 10; ++iy)
                for (int it = -6; it <= 6; ++it) {
                    hyp.push_back(static_cast<float>(guess.x + 0.1 * ix));
                    hyp.push_back(static_cast<float>(guess.y + 0.1 * iy));
                    hyp.push_back(static_cast<float>(guess.theta + 0.02 * it));
                }
        const auto t_loop = std::chrono::steady_clock::now();
        std::vector<ndt2d::Result> refined = ndt.relocalize(scans[N], hyp, 4, 1);   // sweep on the 1 m level, refine the 4 best
        const double loop_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_loop).count();
        const ndt2d::Result *best = nullptr;
        for (const ndt2d::Result &r : refined)
            if (r.status == NDT2D_CONVERGED && (!best || r.score > best->score)) best = &r;
        double loop_err = 1e9;
        if (best) {
            edges.push_back(make_edge(0, N, *best, true));
            ndt2d::Pose2d t = between(truth[0], truth[N]);
            loop_err = std::hypot(t.x - best->pose[0], t.y - best->pose[1]);
        }

        // 3. the pose graph, g2o text format: initial values = dead-reckoned odometry
        FILE *f = std::fopen(out_path.c_str(), "w");
        if (!f) throw std::runtime_error("cannot open " + out_path);
        for (int k = 0; k <= N; ++k) std::fprintf(f, "VERTEX_SE2 %d %.9f %.9f %.9f\n", k, est[k].x, est[k].y, est[k].theta);
        for (const Edge &e : edges)
            std::fprintf(f, "EDGE_SE2 %d %d %.9f %.9f %.9f %.6f %.6f %.6f %.6f %.6f %.6f\n", e.i, e.j, e.z.x, e.z.y, e.z.theta, e.info[0],
                         e.info[1], e.info[2], e.info[3], e.info[4], e.info[5]);
        std::fclose(f);

        std::printf("odometry: %d edges, %d not converged, worst relative error %.4f m, dead-reckoning drift %.4f m after %d poses\n", N, bad,
                    worst_rel, drift, N);
        std::printf("loop closure 0-%d: %s, error vs truth %.4f m\n", N, best ? "accepted" : "none", loop_err);
        std::printf("timing: %.3f ms per odometry step (set target scan + 3-level align, host buffers), %.3f ms for the loop-closure search "
                    "(%zu hypotheses + 4 refinements)\n", odo_ms / N, loop_ms, hyp.size() / 3);
        std::printf("batched odometry: %d pairs in %.3f ms through alignPairs (host buffers), %d results differ from the sequential run\n", N,
                    batch_ms, differing);
        std::printf("pose graph: %d vertices, %zu edges -> %s\n", N + 1, edges.size(), out_path.c_str());
        bool ok = bad == 0 && worst_rel < 0.03 && best && loop_err < 0.03 && differing == 0;
        return ok ? 0 : 1;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "slam_frontend failed: %s\n", e.what());
        return 2;
    }
}
