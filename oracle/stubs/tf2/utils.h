// TEST INFRASTRUCTURE ONLY (oracle build): the subset of tf2 the reference touches
// (Quaternion, setRPY, fromMsg/toMsg, getYaw). getYaw follows tf2/impl/utils.h.
#pragma once
#include <cmath>
#include "geometry_msgs/msgs.h"
namespace tf2
{
    class Quaternion
    {
    public:
        Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
        Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
        void setRPY(double roll, double pitch, double yaw)
        {
            double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
            double cy = std::cos(hy), sy = std::sin(hy);
            double cp = std::cos(hp), sp = std::sin(hp);
            double cr = std::cos(hr), sr = std::sin(hr);
            x_ = sr * cp * cy - cr * sp * sy;
            y_ = cr * sp * cy + sr * cp * sy;
            z_ = cr * cp * sy - sr * sp * cy;
            w_ = cr * cp * cy + sr * sp * sy;
        }
        double x() const { return x_; }
        double y() const { return y_; }
        double z() const { return z_; }
        double w() const { return w_; }

    private:
        double x_, y_, z_, w_;
    };
    inline void fromMsg(const geometry_msgs::Quaternion &in, Quaternion &out)
    {
        out = Quaternion(in.x, in.y, in.z, in.w);
    }
    inline geometry_msgs::Quaternion toMsg(const Quaternion &in)
    {
        geometry_msgs::Quaternion q;
        q.x = in.x();
        q.y = in.y();
        q.z = in.z();
        q.w = in.w();
        return q;
    }
    inline double getYaw(const Quaternion &q)
    {
        double sqx = q.x() * q.x(), sqy = q.y() * q.y(), sqz = q.z() * q.z(), sqw = q.w() * q.w();
        double sarg = -2 * (q.x() * q.z() - q.w() * q.y()) / (sqx + sqy + sqz + sqw);
        if (sarg <= -0.99999)
            return -2 * std::atan2(q.y(), q.x());
        if (sarg >= 0.99999)
            return 2 * std::atan2(q.y(), q.x());
        return std::atan2(2 * (q.x() * q.y() + q.w() * q.z()), sqw + sqx - sqy - sqz);
    }
    inline double getYaw(const geometry_msgs::Quaternion &m)
    {
        Quaternion q;
        fromMsg(m, q);
        return getYaw(q);
    }
}
