// stand-in for <pcl/impl/point_types.hpp> (src/conversions.cpp:26). TEST INFRASTRUCTURE ONLY.
#pragma once
#include <pcl/point_types.h>
