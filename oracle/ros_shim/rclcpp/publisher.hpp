#pragma once
#include <rclcpp/rclcpp.hpp>
