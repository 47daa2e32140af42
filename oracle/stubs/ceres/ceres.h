// TEST INFRASTRUCTURE ONLY (oracle build): planner_utils.h uses ceres::abs / ceres::isinf on doubles.
#pragma once
#include <cmath>
namespace ceres
{
    inline double abs(double x) { return std::fabs(x); }
    inline bool isinf(double x) { return std::isinf(x); }
}
