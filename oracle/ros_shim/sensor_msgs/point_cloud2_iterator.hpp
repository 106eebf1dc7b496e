// Stand-in for sensor_msgs/point_cloud2_iterator.hpp: TEST INFRASTRUCTURE ONLY. A const iterator over one named field
// of a PointCloud2 (advances by point_step, ends after width * height points), as the reference uses it
// (src/conversions.cpp:72-85).
#pragma once
#include <cstring>
#include <sensor_msgs/msg/point_cloud2.hpp>
#include <stdexcept>
#include <string>
namespace sensor_msgs
{
template <typename T> class PointCloud2ConstIterator
{
  public:
    PointCloud2ConstIterator(const msg::PointCloud2 &cloud, const std::string &field)
    {
        std::uint32_t offset = 0U;
        bool found = false;
        for (const auto &f : cloud.fields)
            if (f.name == field)
            {
                offset = f.offset;
                found = true;
            }
        if (!found)
            throw std::runtime_error("Field " + field + " does not exist");
        step_ = cloud.point_step;
        ptr_ = cloud.data.data() + offset;
        end_ = ptr_ + static_cast<std::size_t>(cloud.width) * cloud.height * step_;
    }
    PointCloud2ConstIterator end() const
    {
        PointCloud2ConstIterator e(*this);
        e.ptr_ = end_;
        return e;
    }
    bool operator!=(const PointCloud2ConstIterator &o) const { return ptr_ != o.ptr_; }
    PointCloud2ConstIterator &operator++()
    {
        ptr_ += step_;
        return *this;
    }
    T operator*() const
    {
        T v;
        std::memcpy(&v, ptr_, sizeof(T));
        return v;
    }

  private:
    const std::uint8_t *ptr_{nullptr};
    const std::uint8_t *end_{nullptr};
    std::uint32_t step_{0U};
};
} // namespace sensor_msgs
