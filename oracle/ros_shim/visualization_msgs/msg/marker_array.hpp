// Stand-in for visualization_msgs/msg/MarkerArray.msg: TEST INFRASTRUCTURE ONLY.
#pragma once
#include <vector>
#include <visualization_msgs/msg/marker.hpp>
namespace visualization_msgs
{
namespace msg
{
struct MarkerArray
{
    std::vector<Marker> markers;
};
} // namespace msg
} // namespace visualization_msgs
