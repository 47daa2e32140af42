// TEST INFRASTRUCTURE ONLY (oracle build): shadows the reference's RViz publisher wrapper
// (mpros_utils/visualizer.h) with an empty class; HybridAstar only stores a pointer to it
// (hybrid_a_star.cpp:68).
#pragma once
#include "ros/ros.h"
#include "nav_msgs/Path.h"
#include "nav_msgs/OccupancyGrid.h"
class Visualizer
{
public:
    Visualizer() = default;
    explicit Visualizer(ros::NodeHandle &) {}
};
