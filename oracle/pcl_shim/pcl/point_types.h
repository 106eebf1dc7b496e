// Minimal stand-in for <pcl/point_types.h>: TEST INFRASTRUCTURE ONLY (oracle build + host-shell tests).
// Real PCL is not installed in this image. Layouts follow PCL's published ones: every point is
// 16-byte aligned with x,y,z,(pad=1.0f) in the first 16 bytes; PointXYZ is 16 B, the rest 32 B.
// Only the members the reference's hot path and its caller touch are provided
// (reference: src/processor.cpp:152-163, src/clustering.cpp:56-61, src/segmentation.cpp:137-144).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace pcl
{
struct alignas(16) PointXYZ
{
    float x{0.0F}, y{0.0F}, z{0.0F}, _w{1.0F};
    PointXYZ() = default;
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

struct alignas(16) PointXYZI
{
    float x{0.0F}, y{0.0F}, z{0.0F}, _w{1.0F};
    float intensity{0.0F};
    float _pad[3]{0.0F, 0.0F, 0.0F};
    PointXYZI() = default;
    PointXYZI(float x_, float y_, float z_, float i_ = 0.0F) : x(x_), y(y_), z(z_), intensity(i_) {}
};

struct alignas(16) PointXYZL
{
    float x{0.0F}, y{0.0F}, z{0.0F}, _w{1.0F};
    std::uint32_t label{0U};
    std::uint32_t _pad[3]{0U, 0U, 0U};
    PointXYZL() = default;
    PointXYZL(float x_, float y_, float z_, std::uint32_t l_ = 0U) : x(x_), y(y_), z(z_), label(l_) {}
};

// b, g, r, a share their 4 bytes with the packed float `rgb` (PCL_ADD_RGB); the reference takes offsetof(..., rgb)
// (src/conversions.cpp:96-99).
struct alignas(16) PointXYZRGB
{
    float x{0.0F}, y{0.0F}, z{0.0F}, _w{1.0F};
    union
    {
        struct
        {
            std::uint8_t b, g, r, a;
        };
        float rgb;
        std::uint32_t rgba;
    };
    std::uint32_t _pad[3]{0U, 0U, 0U};
    PointXYZRGB() : b(0), g(0), r(0), a(255) {}
    PointXYZRGB(float x_, float y_, float z_, std::uint8_t r_ = 0, std::uint8_t g_ = 0, std::uint8_t b_ = 0)
        : x(x_), y(y_), z(z_), b(b_), g(g_), r(r_), a(255)
    {
    }
};

struct alignas(16) PointXYZRGBL
{
    float x{0.0F}, y{0.0F}, z{0.0F}, _w{1.0F};
    union
    {
        struct
        {
            std::uint8_t b, g, r, a;
        };
        float rgb;
        std::uint32_t rgba;
    };
    std::uint32_t label{0U};
    std::uint32_t _pad[2]{0U, 0U};
    PointXYZRGBL() : b(0), g(0), r(0), a(255) {}
    PointXYZRGBL(float x_, float y_, float z_, std::uint8_t r_ = 0, std::uint8_t g_ = 0, std::uint8_t b_ = 0,
                 std::uint32_t l_ = 0U)
        : x(x_), y(y_), z(z_), b(b_), g(g_), r(r_), a(255), label(l_)
    {
    }
};

static_assert(sizeof(PointXYZ) == 16, "PointXYZ layout");
static_assert(sizeof(PointXYZI) == 32, "PointXYZI layout");
static_assert(sizeof(PointXYZL) == 32, "PointXYZL layout");
static_assert(sizeof(PointXYZRGB) == 32, "PointXYZRGB layout");
static_assert(sizeof(PointXYZRGBL) == 32, "PointXYZRGBL layout");
} // namespace pcl
