// TEST INFRASTRUCTURE ONLY (oracle build): minimal stand-in for <ros/ros.h> so the
// reference's unmodified nav_map.cpp / rs_curve.cpp / hybrid_a_star.cpp compile headless.
// Nothing here carries arithmetic; parameters come from a process-wide string->double map.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace ros
{
    struct Time
    {
        uint32_t sec = 0, nsec = 0;
        Time() = default;
        Time(uint32_t s, uint32_t n) : sec(s), nsec(n) {}
        uint64_t toNSec() const { return (uint64_t)sec * 1000000000ull + nsec; }
        static Time now() { return Time(1, 0); }
    };
    struct Subscriber
    {
    };
    struct Publisher
    {
        template <typename M>
        void publish(const M &) const {}
    };
    // parameter server: flat map of fully-qualified key -> double (bools/ints are stored as doubles)
    inline std::map<std::string, double> &param_store()
    {
        static std::map<std::string, double> store;
        return store;
    }
    class NodeHandle
    {
    public:
        bool ok() const { return true; }
        template <typename T>
        bool param(const std::string &key, T &var, const T &dflt) const
        {
            auto it = param_store().find(key);
            if (it == param_store().end())
            {
                var = dflt;
                return false;
            }
            var = static_cast<T>(it->second);
            return true;
        }
        template <typename... A>
        Subscriber subscribe(A &&...) { return Subscriber(); }
        template <typename M, typename... A> // subscribe<Msg>(topic, queue, callback)
        Subscriber subscribe(const std::string &, A &&...) { return Subscriber(); }
        template <typename M, typename... A>
        Publisher advertise(A &&...) { return Publisher(); }
    };
    inline void init(int &, char **, const std::string &) {}
}

#define MPROS_STUB_LOG(...) \
    do                      \
    {                       \
        if (0)              \
            std::printf(__VA_ARGS__); \
    } while (0)
#define ROS_DEBUG(...) MPROS_STUB_LOG(__VA_ARGS__)
#define ROS_INFO(...) MPROS_STUB_LOG(__VA_ARGS__)
#define ROS_WARN(...) MPROS_STUB_LOG(__VA_ARGS__)
#define ROS_ERROR(...) MPROS_STUB_LOG(__VA_ARGS__)
#define MPROS_STUB_STREAM(x)       \
    do                             \
    {                              \
        if (0)                     \
        {                          \
            std::stringstream ss_; \
            ss_ << x;              \
        }                          \
    } while (0)
#define ROS_DEBUG_STREAM(x) MPROS_STUB_STREAM(x)
#define ROS_INFO_STREAM(x) MPROS_STUB_STREAM(x)
#define ROS_WARN_STREAM(x) MPROS_STUB_STREAM(x)
#define ROS_ERROR_STREAM(x) MPROS_STUB_STREAM(x)
