// Drop-in for mpros_navigation_map/include/mpros_navigation_map/nav_map.h: put hybrid-astar-planner_b200/host/compat ahead of the
// reference's include directories and build with -DMPROS_FACADE_ROS; `NavMap` then names the B200-backed class
// (host/mpros_facade.h), with the reference's public methods and data members (nav_map.h:17-547).
#ifndef MPROS_NAVIGATION_MAP__NAV_MAP_H
#define MPROS_NAVIGATION_MAP__NAV_MAP_H
#ifndef MPROS_FACADE_ROS
#error "the compat headers need -DMPROS_FACADE_ROS (real ros::NodeHandle / nav_msgs / Eigen types)"
#endif
#include "tf2/utils.h"
#include "tf2_geometry_msgs/tf2_geometry_msgs.h"

#include "mpros_facade.h"
class NavMap : public mpros_b200::NavMap
{
public:
    using mpros_b200::NavMap::NavMap;
};
#endif
