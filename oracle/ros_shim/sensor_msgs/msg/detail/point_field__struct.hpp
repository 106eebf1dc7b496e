#pragma once
#include <sensor_msgs/msg/point_cloud2.hpp>
