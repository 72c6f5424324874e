// oracle/ref_build/stubs/ssf_ref_stubs.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Minimal stand-ins for the ROS / PCL / tf / ceres / gtsam headers that the reference's include/header.h pulls in, so that
// the reference's src/frameFeature.cpp compiles FROM WHERE IT LIES (/root/reference, unmodified, never copied) with plain g++
// into oracle/_ref/libframe_feature_ref.so (recipe: oracle/ref_build/Makefile).  Only what frameFeature.cpp touches has a
// body: the point-cloud container, the message, the publisher registry the harness reads the node's output from.  The
// numerics of the node (float atan / sqrt overloads from <math.h>, the 11-tap curvature, the greedy selection) are the
// reference's own compiled code; nothing here computes.
#pragma once
#include <math.h>   // tf/LinearMath/Scalar.h includes <math.h> in a real ROS build: the float overloads of atan / sqrt are global
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

// ---- PCL point-type macros (pcl/point_types.h): x, y, z + padding float, 16-byte aligned
#define PCL_ADD_POINT4D float x; float y; float z; float ssf_pad_;
#define PCL_ADD_INTENSITY float intensity
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 __attribute__((aligned(16)))
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)

namespace ros {
struct Time {
    uint32_t sec = 0, nsec = 0;
};
}  // namespace ros

namespace sensor_msgs {
struct Header {
    uint32_t seq = 0;
    ros::Time stamp;
    std::string frame_id;
};
// the message as the stubs carry it: N points x (x, y, z, intensity)
struct PointCloud2 {
    Header header;
    std::vector<float> xyzi;
};
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
}  // namespace sensor_msgs

namespace pcl {
struct PointXYZI {
    float x = 0, y = 0, z = 0, intensity = 0;
};
template <typename T>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<T>> Ptr;
    std::vector<T> points;
    void clear() { points.clear(); }
    size_t size() const { return points.size(); }
    void push_back(const T& p) { points.push_back(p); }
    T& operator[](size_t i) { return points[i]; }
    const T& operator[](size_t i) const { return points[i]; }
};
// the node filters into a temporary it never publishes (src/frameFeature.cpp:128-131): nothing to reproduce
template <typename T>
class VoxelGrid {
public:
    void setLeafSize(float, float, float) {}
    void setInputCloud(const typename PointCloud<T>::Ptr&) {}
    void filter(PointCloud<T>&) {}
};
template <typename T>
void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<T>& cloud) {
    cloud.points.resize(msg.xyzi.size() / 4);
    for (size_t i = 0; i < cloud.points.size(); ++i) {
        cloud.points[i].x = msg.xyzi[4 * i];
        cloud.points[i].y = msg.xyzi[4 * i + 1];
        cloud.points[i].z = msg.xyzi[4 * i + 2];
        cloud.points[i].intensity = msg.xyzi[4 * i + 3];
    }
}
template <typename T>
void toROSMsg(const PointCloud<T>& cloud, sensor_msgs::PointCloud2& msg) {
    msg.xyzi.resize(cloud.points.size() * 4);
    for (size_t i = 0; i < cloud.points.size(); ++i) {
        msg.xyzi[4 * i] = cloud.points[i].x;
        msg.xyzi[4 * i + 1] = cloud.points[i].y;
        msg.xyzi[4 * i + 2] = cloud.points[i].z;
        msg.xyzi[4 * i + 3] = cloud.points[i].intensity;
    }
}
}  // namespace pcl

namespace ssf_ref {
// what the node published last, per topic, and the callback it subscribed (read by the harness)
inline std::map<std::string, sensor_msgs::PointCloud2>& published() {
    static std::map<std::string, sensor_msgs::PointCloud2> m;
    return m;
}
typedef void (*CloudHandler)(const sensor_msgs::PointCloud2ConstPtr&);
inline CloudHandler& handler() {
    static CloudHandler h = nullptr;
    return h;
}
}  // namespace ssf_ref

namespace ros {
inline void init(int&, char**, const char*) {}
inline void spin() {}
class Subscriber {};
class Publisher {
public:
    std::string topic;
    void publish(const sensor_msgs::PointCloud2& m) const { ssf_ref::published()[topic] = m; }
};
class NodeHandle {
public:
    template <typename M>
    Subscriber subscribe(const std::string&, int, void (*cb)(const std::shared_ptr<const M>&)) {
        ssf_ref::handler() = cb;
        return Subscriber();
    }
    template <typename M>
    Publisher advertise(const std::string& topic, int) {
        Publisher p;
        p.topic = topic;
        return p;
    }
};
}  // namespace ros

#define ROS_INFO(...) do {} while (0)
