// TYPE STUB for tests only, see geometry/Pose2.h
#pragma once
#include <gtsam/slam/BetweenFactor.h>
namespace gtsam {
struct NonlinearFactorGraph {
    std::vector<std::shared_ptr<BetweenFactor<Pose2>>> factors;
    template <class F, class... A>
    void emplace_shared(A &&...a) { factors.push_back(std::make_shared<F>(std::forward<A>(a)...)); }
    size_t size() const { return factors.size(); }
};
} // namespace gtsam
