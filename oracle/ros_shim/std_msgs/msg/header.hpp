// Stand-in for std_msgs/msg/Header.msg and ColorRGBA.msg: TEST INFRASTRUCTURE ONLY (see builtin_interfaces/msg/time.hpp).
#pragma once
#include <builtin_interfaces/msg/time.hpp>
#include <string>
namespace std_msgs
{
namespace msg
{
struct Header
{
    builtin_interfaces::msg::Time stamp;
    std::string frame_id;
};
struct ColorRGBA
{
    float r{0.0F}, g{0.0F}, b{0.0F}, a{0.0F};
};
} // namespace msg
} // namespace std_msgs
