/* liodom facade — type layer (replaces include/liodom/defs.h:35-39 of the reference).
 *
 * The reference builds on pcl::PointXYZI / pcl::PointCloud / Eigen / ros::NodeHandle.  None of
 * those libraries exists in this image, so this header provides layout-compatible stand-ins in
 * namespace liodom with the member names the reference code touches (points, width, height,
 * header, at(col,row), matrix(), translation(), ...).  A build that HAS ROS/PCL/Eigen defines
 * LIODOM_FACADE_USE_PCL and gets the original typedefs instead. */
#ifndef INCLUDE_LIODOM_DEFS_H
#define INCLUDE_LIODOM_DEFS_H

#include <chrono>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#ifdef LIODOM_FACADE_USE_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <ros/ros.h>
#include <std_msgs/Header.h>
#include <sensor_msgs/PointCloud2.h>
namespace liodom {
typedef pcl::PointXYZI Point;
typedef pcl::PointCloud<Point> PointCloud;
typedef std::chrono::high_resolution_clock Clock;
typedef Eigen::Isometry3d Isometry3d;
typedef Eigen::Quaterniond Quaterniond;
typedef Eigen::Matrix4d Matrix4d;
typedef std_msgs::Header Header;
typedef ros::NodeHandle NodeHandle;
typedef sensor_msgs::PointCloud2 PointCloud2;
typedef sensor_msgs::PointField PointField;
namespace detail {
inline void set_stamp(Header& h, double t) { h.stamp.fromSec(t); }
}
}  // namespace liodom
#else

namespace liodom {

typedef std::chrono::high_resolution_clock Clock;

struct Time {
  double secs = 0.0;
  double toSec() const { return secs; }
};
struct Header {          // std_msgs::Header
  uint32_t seq = 0;
  Time stamp;
  std::string frame_id;
};

/* pcl::PointXYZI memory layout: 4 floats (x, y, z, pad) + intensity + 3 pad = 32 bytes, 16-aligned. */
struct alignas(16) Point {
  float x = 0.f, y = 0.f, z = 0.f, _pad0 = 1.f;
  float intensity = 0.f, _pad1[3] = {0.f, 0.f, 0.f};
};
static_assert(sizeof(Point) == 32, "liodom::Point must match pcl::PointXYZI");

/* sensor_msgs/PointField + PointCloud2: what lidarClb / mapClb receive (src/liodom_node.cc:40-64). */
struct PointField {
  enum { INT8 = 1, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 };
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 1;
};
struct PointCloud2 {
  typedef std::shared_ptr<PointCloud2> Ptr;
  typedef std::shared_ptr<const PointCloud2> ConstPtr;
  Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};

struct PointCloud {     // the part of pcl::PointCloud<PointXYZI> the reference uses
  typedef std::shared_ptr<PointCloud> Ptr;
  typedef std::shared_ptr<const PointCloud> ConstPtr;
  std::vector<Point> points;
  /* Set by fromROSMsgDeferred: the message bytes travel to the device as they are and the split
   * kernel reads the fields in place; `points` then stays empty until someone asks for them. */
  PointCloud2::ConstPtr raw;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  Header header;
  size_t size() const { return raw ? (size_t)raw->width * raw->height : points.size(); }
  bool empty() const { return size() == 0; }
  void clear() { points.clear(); width = height = 0; }
  void push_back(const Point& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  const Point& at(int column, int row) const { return points[(size_t)row * width + column]; }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = (uint32_t)points.size(); height = 1;
    return *this;
  }
};

struct Quaterniond {     // Eigen::Quaterniond: (w, x, y, z) constructor and accessors
  double qx = 0, qy = 0, qz = 0, qw = 1;
  Quaterniond() {}
  Quaterniond(double w, double x, double y, double z) : qx(x), qy(y), qz(z), qw(w) {}
  double x() const { return qx; } double y() const { return qy; } double z() const { return qz; } double w() const { return qw; }
};

struct Matrix4d {        // row-major 4x4
  double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  double& operator()(int r, int c) { return m[r * 4 + c]; }
  double operator()(int r, int c) const { return m[r * 4 + c]; }
};

struct Isometry3d {      // the part of Eigen::Isometry3d the reference uses
  Matrix4d mat;
  static Isometry3d Identity() { return Isometry3d(); }
  Matrix4d& matrix() { return mat; }
  const Matrix4d& matrix() const { return mat; }
  double operator()(int r, int c) const { return mat(r, c); }
  Isometry3d operator*(const Isometry3d& o) const {
    Isometry3d r;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) r.mat(i, j) = mat(i, 0) * o.mat(0, j) + mat(i, 1) * o.mat(1, j) + mat(i, 2) * o.mat(2, j);
      r.mat(i, 3) = (mat(i, 0) * o.mat(0, 3) + mat(i, 1) * o.mat(1, 3) + mat(i, 2) * o.mat(2, 3)) + mat(i, 3);
    }
    return r;
  }
  Isometry3d inverse() const {
    Isometry3d r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.mat(i, j) = mat(j, i);
    for (int i = 0; i < 3; ++i) r.mat(i, 3) = (-r.mat(i, 0)) * mat(0, 3) + (-r.mat(i, 1)) * mat(1, 3) + (-r.mat(i, 2)) * mat(2, 3);
    return r;
  }
};

/* Parameter source standing in for ros::NodeHandle's private-parameter lookup
 * (nh.param(name, out, default), src/params.cc:40-108). */
class NodeHandle {
 public:
  NodeHandle() {}
  template <typename T> void setParam(const std::string& name, const T& v) { std::ostringstream s; s << v; kv_[name] = s.str(); }
  void setParam(const std::string& name, bool v) { kv_[name] = v ? "1" : "0"; }
  template <typename T> bool param(const std::string& name, T& out, const T& def) const {
    auto it = kv_.find(name);
    if (it == kv_.end()) { out = def; return false; }
    std::istringstream s(it->second); s >> out;
    return true;
  }
  bool param(const std::string& name, bool& out, const bool& def) const {
    auto it = kv_.find(name);
    if (it == kv_.end()) { out = def; return false; }
    out = (it->second == "1" || it->second == "true" || it->second == "True");
    return true;
  }
  bool param(const std::string& name, std::string& out, const std::string& def) const {
    auto it = kv_.find(name);
    if (it == kv_.end()) { out = def; return false; }
    out = it->second;
    return true;
  }
 private:
  std::map<std::string, std::string> kv_;
};

namespace detail {
inline void set_stamp(Header& h, double t) { h.stamp.secs = t; }
}
}  // namespace liodom
#endif  // LIODOM_FACADE_USE_PCL

namespace liodom {

/* nav_msgs/Odometry + geometry_msgs/TwistStamped content filled by LaserOdometer::publishOdom
 * (src/laser_odometry.cc:395-446); the TF broadcast carries the same position / orientation. */
struct Odometry {
  Header header;               // frame_id = fixed_frame, stamp = the scan's
  std::string child_frame_id;  // base_frame
  Quaterniond orientation;     // of base_link in the fixed frame
  double position[3] = {0, 0, 0};
  double twist_linear[3] = {0, 0, 0};
  double twist_angular[3] = {0, 0, 0};
};

/* Only the API common to the stand-ins and to Eigen is used on poses: matrix()(i, j), Identity(), operator*. */
namespace detail {
inline void pose_to16(const Isometry3d& T, double* m) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i * 4 + j] = T.matrix()(i, j); }
inline Isometry3d pose_from16(const double* m) {
  Isometry3d T = Isometry3d::Identity();
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T.matrix()(i, j) = m[i * 4 + j];
  return T;
}
}  // namespace detail

}  // namespace liodom
#endif  // INCLUDE_LIODOM_DEFS_H
