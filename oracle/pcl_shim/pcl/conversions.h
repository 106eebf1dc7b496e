// Minimal stand-in for <pcl/conversions.h> / <pcl/PCLPointField.h>: TEST INFRASTRUCTURE ONLY.
// The reference only uses the datatype enumerators (src/conversions.hpp:49, conversions.cpp:96-99); values follow PCL's.
#pragma once
#include <cstdint>
namespace pcl
{
struct PCLPointField
{
    enum PointFieldTypes
    {
        INT8 = 1,
        UINT8 = 2,
        INT16 = 3,
        UINT16 = 4,
        INT32 = 5,
        UINT32 = 6,
        FLOAT32 = 7,
        FLOAT64 = 8
    };
};
} // namespace pcl
