// TYPE STUB for tests only, see geometry/Pose2.h
#pragma once
#include <gtsam/linear/NoiseModel.h>
namespace gtsam {
template <class T>
struct BetweenFactor {
    Key key1, key2;
    T measured;
    SharedNoiseModel model;
    BetweenFactor(Key a, Key b, const T &z, const SharedNoiseModel &m) : key1(a), key2(b), measured(z), model(m) {}
};
} // namespace gtsam
