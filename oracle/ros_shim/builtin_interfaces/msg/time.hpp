// Minimal stand-in for ROS 2 message headers: TEST INFRASTRUCTURE ONLY (oracle build of the reference node).
// ROS 2 is not installed in this image. Field names, types and order follow the published .msg definitions
// (builtin_interfaces/msg/Time.msg, Duration.msg).
#pragma once
#include <cstdint>
namespace builtin_interfaces
{
namespace msg
{
struct Time
{
    std::int32_t sec{0};
    std::uint32_t nanosec{0U};
};
struct Duration
{
    std::int32_t sec{0};
    std::uint32_t nanosec{0U};
};
} // namespace msg
} // namespace builtin_interfaces
