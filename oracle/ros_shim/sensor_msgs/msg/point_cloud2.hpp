// Stand-in for sensor_msgs/msg/{PointField,PointCloud2}.msg: TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
#include <std_msgs/msg/header.hpp>
#include <string>
#include <vector>
namespace sensor_msgs
{
namespace msg
{
struct PointField
{
    static constexpr std::uint8_t INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8;
    std::string name;
    std::uint32_t offset{0U};
    std::uint8_t datatype{0U};
    std::uint32_t count{0U};
};
struct PointCloud2
{
    std_msgs::msg::Header header;
    std::uint32_t height{0U};
    std::uint32_t width{0U};
    std::vector<PointField> fields;
    bool is_bigendian{false};
    std::uint32_t point_step{0U};
    std::uint32_t row_step{0U};
    std::vector<std::uint8_t> data;
    bool is_dense{false};
};
} // namespace msg
} // namespace sensor_msgs
