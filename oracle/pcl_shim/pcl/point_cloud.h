// Minimal stand-in for <pcl/point_cloud.h>: TEST INFRASTRUCTURE ONLY. See point_types.h.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

namespace pcl
{
template <typename PointT> class PointCloud
{
  public:
    using VectorType = std::vector<PointT>;
    using iterator = typename VectorType::iterator;
    using const_iterator = typename VectorType::const_iterator;

    VectorType points;
    std::uint32_t width{0U};
    std::uint32_t height{1U};
    bool is_dense{true};

    std::size_t size() const noexcept { return points.size(); }
    bool empty() const noexcept { return points.empty(); }
    void reserve(std::size_t n) { points.reserve(n); }
    void resize(std::size_t n)
    {
        points.resize(n);
        width = static_cast<std::uint32_t>(n);
        height = 1U;
    }
    void clear()
    {
        points.clear();
        width = 0U;
        height = 0U;
    }
    void push_back(const PointT &p)
    {
        points.push_back(p);
        width = static_cast<std::uint32_t>(points.size());
        height = 1U;
    }
    template <typename... Args> PointT &emplace_back(Args &&...args)
    {
        points.emplace_back(std::forward<Args>(args)...);
        width = static_cast<std::uint32_t>(points.size());
        height = 1U;
        return points.back();
    }
    PointT &at(std::size_t i) { return points.at(i); }
    const PointT &at(std::size_t i) const { return points.at(i); }
    PointT &operator[](std::size_t i) { return points[i]; }
    const PointT &operator[](std::size_t i) const { return points[i]; }
    iterator begin() noexcept { return points.begin(); }
    iterator end() noexcept { return points.end(); }
    const_iterator begin() const noexcept { return points.begin(); }
    const_iterator end() const noexcept { return points.end(); }
    const_iterator cbegin() const noexcept { return points.cbegin(); }
    const_iterator cend() const noexcept { return points.cend(); }
};
} // namespace pcl
