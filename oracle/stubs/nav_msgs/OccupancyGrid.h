// TEST INFRASTRUCTURE ONLY (oracle build): stand-in for nav_msgs/OccupancyGrid.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include "geometry_msgs/msgs.h"
namespace nav_msgs
{
    struct MapMetaData
    {
        ros::Time map_load_time;
        float resolution = 0; // float32 in the ROS message definition
        uint32_t width = 0, height = 0;
        geometry_msgs::Pose origin;
    };
    struct OccupancyGrid
    {
        std_msgs::Header header;
        MapMetaData info;
        std::vector<int8_t> data;
        typedef std::shared_ptr<OccupancyGrid> Ptr;
        typedef std::shared_ptr<const OccupancyGrid> ConstPtr;
    };
}
