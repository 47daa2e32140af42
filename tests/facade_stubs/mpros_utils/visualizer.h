// TEST INFRASTRUCTURE ONLY: signature-only stand-in for mpros_utils/visualizer.h (RViz publishers, out of scope), with the
// methods main_planner.cpp calls (visualizer.h:94-373) so that the node's main() can be syntax-checked against the facade.
#pragma once
#include <string>
#include <vector>

#include <Eigen/Core>

#include "mpros_utils/pathpoint.h"
#include "ros/ros.h"
class Visualizer
{
public:
    Visualizer() = default;
    explicit Visualizer(ros::NodeHandle &) {}
    void publishPath(const std::vector<std::vector<double>> &) {}
    void publishPath(std::vector<PathPoint>) {}
    void publishFootprints(const std::vector<PathPoint> &) {}
    void publishSmoothedPath(std::vector<PathPoint>) {}
    void publishTrajectory(std::vector<PathPoint>) {}
    template <typename T>
    void publish(const T &, std::string) {}
    void PublishSearchedTree(const std::vector<Eigen::Vector4d> &) {}
    void publishPathTangent(const std::vector<PathPoint> &) {}
};
