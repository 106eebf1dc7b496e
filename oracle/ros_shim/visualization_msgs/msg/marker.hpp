// Stand-in for visualization_msgs/msg/Marker.msg (the fields the reference sets, src/conversions.hpp:72-120):
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <builtin_interfaces/msg/time.hpp>
#include <cstdint>
#include <geometry_msgs/msg/point.hpp>
#include <std_msgs/msg/header.hpp>
#include <string>
#include <vector>
namespace visualization_msgs
{
namespace msg
{
struct Marker
{
    static constexpr std::int32_t ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5;
    static constexpr std::int32_t ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3;
    std_msgs::msg::Header header;
    std::string ns;
    std::int32_t id{0};
    std::int32_t type{0};
    std::int32_t action{0};
    geometry_msgs::msg::Pose pose;
    geometry_msgs::msg::Vector3 scale;
    std_msgs::msg::ColorRGBA color;
    builtin_interfaces::msg::Duration lifetime;
    bool frame_locked{false};
    std::vector<geometry_msgs::msg::Point> points;
    std::vector<std_msgs::msg::ColorRGBA> colors;
};
} // namespace msg
} // namespace visualization_msgs
