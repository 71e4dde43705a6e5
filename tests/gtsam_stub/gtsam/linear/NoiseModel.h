// TYPE STUB for tests only, see geometry/Pose2.h
#pragma once
#include <gtsam/geometry/Pose2.h>
namespace gtsam {
namespace noiseModel {
struct Base {
    Matrix3 information;
};
struct Gaussian : Base {
    static std::shared_ptr<Gaussian> Information(const Matrix3 &M)
    {
        auto g = std::make_shared<Gaussian>();
        g->information = M;
        return g;
    }
};
} // namespace noiseModel
using SharedNoiseModel = std::shared_ptr<noiseModel::Base>;
} // namespace gtsam
