// TEST INFRASTRUCTURE ONLY (oracle build): stand-in for nav_msgs/Path.
#pragma once
#include <vector>
#include "geometry_msgs/msgs.h"
namespace nav_msgs
{
    struct Path
    {
        std_msgs::Header header;
        std::vector<geometry_msgs::PoseStamped> poses;
    };
}
