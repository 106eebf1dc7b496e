// Minimal stand-in for rclcpp: TEST INFRASTRUCTURE ONLY. Enough of Node / QoS / Publisher / Subscription for the
// UNMODIFIED reference node src/processor.cpp to compile and run in-process without DDS: a subscription keeps its
// callback in a process-wide registry keyed by topic, a publisher hands every message to a per-topic sink, and
// rclcpp::spin() calls the harness (oracle/ref_node_wrap.cpp), which feeds messages to the "pointcloud" callback the way
// the executor would, one at a time on the calling thread (processor.cpp:93-94, 279).
#pragma once
#include <any>
#include <chrono>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <typeindex>
#include <utility>

namespace rclcpp
{
enum class LivelinessPolicy
{
    SystemDefault,
    Automatic,
    ManualByTopic
};

class QoS
{
  public:
    explicit QoS(std::size_t depth) : depth_(depth) {}
    QoS &keep_last(std::size_t depth)
    {
        depth_ = depth;
        return *this;
    }
    QoS &reliable() { return *this; }
    QoS &durability_volatile() { return *this; }
    QoS &liveliness(LivelinessPolicy) { return *this; }
    template <typename D> QoS &liveliness_lease_duration(D) { return *this; }
    template <typename D> QoS &deadline(D) { return *this; }

  private:
    std::size_t depth_;
};

namespace shim
{
// topic -> type-erased std::function<void(const T&)>
inline std::map<std::string, std::any> &callbacks()
{
    static std::map<std::string, std::any> m;
    return m;
}
inline std::map<std::string, std::any> &sinks()
{
    static std::map<std::string, std::any> m;
    return m;
}
template <typename T> void set_sink(const std::string &topic, std::function<void(const T &)> f) { sinks()[topic] = std::move(f); }
template <typename T> void deliver(const std::string &topic, const T &msg) // what the executor does with a received message
{
    std::any_cast<std::function<void(const T &)> &>(callbacks().at(topic))(msg);
}
inline std::function<void()> &spin_hook()
{
    static std::function<void()> f;
    return f;
}
} // namespace shim

template <typename T> class Publisher
{
  public:
    using SharedPtr = std::shared_ptr<Publisher<T>>;
    explicit Publisher(std::string topic) : topic_(std::move(topic)) {}
    void publish(const T &msg)
    {
        auto it = shim::sinks().find(topic_);
        if (it != shim::sinks().end())
            std::any_cast<std::function<void(const T &)> &>(it->second)(msg);
    }

  private:
    std::string topic_;
};

template <typename T> class Subscription
{
  public:
    using SharedPtr = std::shared_ptr<Subscription<T>>;
};

struct Logger
{
};

class Node
{
  public:
    explicit Node(const std::string &name) : name_(name) {}
    virtual ~Node() = default;
    template <typename T, typename CallbackT>
    typename Subscription<T>::SharedPtr create_subscription(const std::string &topic, const QoS &, CallbackT &&callback)
    {
        shim::callbacks()[topic] = std::function<void(const T &)>(std::forward<CallbackT>(callback));
        return std::make_shared<Subscription<T>>();
    }
    template <typename T> typename Publisher<T>::SharedPtr create_publisher(const std::string &topic, const QoS &)
    {
        return std::make_shared<Publisher<T>>(topic);
    }
    Logger get_logger() const { return Logger{}; }

  private:
    std::string name_;
};

inline void init(int, const char *const *) {}
inline void install_signal_handlers() {}
inline void shutdown() {}
inline void spin(std::shared_ptr<Node> node)
{
    if (shim::spin_hook())
        shim::spin_hook()();
    (void)node;
}
} // namespace rclcpp

namespace rclcpp
{
namespace shim
{
template <typename... Args> inline void log_sink(const Logger &, Args &&...) {} // the node's per-stage timing lines
} // namespace shim
} // namespace rclcpp
#define RCLCPP_INFO(logger, ...) ::rclcpp::shim::log_sink(logger, __VA_ARGS__)
