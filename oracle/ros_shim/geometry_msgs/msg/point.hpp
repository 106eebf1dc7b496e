// Stand-in for geometry_msgs/msg/{Point,Quaternion,Pose,Vector3}.msg: TEST INFRASTRUCTURE ONLY.
#pragma once
namespace geometry_msgs
{
namespace msg
{
struct Point
{
    double x{0.0}, y{0.0}, z{0.0};
};
struct Quaternion
{
    double x{0.0}, y{0.0}, z{0.0}, w{1.0};
};
struct Pose
{
    Point position;
    Quaternion orientation;
};
struct Vector3
{
    double x{0.0}, y{0.0}, z{0.0};
};
} // namespace msg
} // namespace geometry_msgs
