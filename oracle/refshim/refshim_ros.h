// refshim — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Stand-ins for the roscpp / message / tf types named by the reference's sources, so that
// src/{feature_extractor,laser_odometry,params,shared_data,stats}.cc compile UNMODIFIED.  No ROS
// master, no transport: parameters come from a process-wide table, published messages land in a
// process-wide "bus" keyed by topic, the static base->laser transform comes from a table, and
// ros::Time::now() is the wall clock unless the driver freezes it.  tf's LinearMath (Quaternion,
// Matrix3x3 setRotation / getRPY / setRPY / getRotation) is restated from its published algorithm.
#pragma once
// (the real ros/ros.h pulls these in transitively; the reference relies on that)
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <typeindex>
#include <vector>

namespace refshim {

struct ParamTable {
  std::mutex mu;
  std::map<std::string, std::string> values;   // everything kept as text
  static ParamTable& get() { static ParamTable t; return t; }
};

struct BusMessage { std::shared_ptr<void> msg; std::type_index type = std::type_index(typeid(void)); long seq = 0; };
struct Bus {
  std::mutex mu;
  std::map<std::string, BusMessage> last;   // latest message per topic
  std::function<void(const std::string&)> on_publish;   // driver hook, called in the publisher's thread
  static Bus& get() { static Bus b; return b; }
  template <typename M> void put(const std::string& topic, const M& m) {
    {
      std::lock_guard<std::mutex> lk(mu);
      BusMessage& bm = last[topic];
      bm.msg = std::make_shared<M>(m); bm.type = std::type_index(typeid(M)); bm.seq++;
    }
    if (on_publish) on_publish(topic);
  }
  template <typename M> bool peek(const std::string& topic, M* out, long* seq) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = last.find(topic);
    if (it == last.end() || it->second.type != std::type_index(typeid(M))) return false;
    if (out) *out = *static_cast<M*>(it->second.msg.get());
    if (seq) *seq = it->second.seq;
    return true;
  }
};

// node-level plumbing used by the unmodified src/liodom_node.cc / liodom_mapping_node.cc (drop-in build of the facade)
struct Subscribers {
  std::mutex mu;
  std::map<std::string, std::function<void(const std::shared_ptr<const void>&)> > table;
  static Subscribers& get() { static Subscribers s; return s; }
  template <typename M> bool deliver(const std::string& topic, const std::shared_ptr<const M>& msg) {
    std::function<void(const std::shared_ptr<const void>&)> f;
    { std::lock_guard<std::mutex> lk(mu); auto it = table.find(topic); if (it == table.end()) return false; f = it->second; }
    f(std::static_pointer_cast<const void>(msg));
    return true;
  }
};
struct SpinHook { std::function<void()> player; static SpinHook& get() { static SpinHook h; return h; } };

struct ClockState { bool frozen = false; double now = 0.0; int log_level = 1; long warnings = 0; static ClockState& get() { static ClockState c; return c; } };

}  // namespace refshim

#define REFSHIM_LOG(level, tag, ...)                                                             \
  do {                                                                                           \
    if ((level) >= 3) refshim::ClockState::get().warnings++;                                     \
    if ((level) >= refshim::ClockState::get().log_level + 2) { std::fprintf(stderr, "[ref " tag "] "); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } \
  } while (0)
#define ROS_DEBUG(...) REFSHIM_LOG(1, "debug", __VA_ARGS__)
#define ROS_INFO(...) REFSHIM_LOG(2, "info", __VA_ARGS__)
#define ROS_WARN(...) REFSHIM_LOG(3, "warn", __VA_ARGS__)
#define ROS_ERROR(...) REFSHIM_LOG(4, "error", __VA_ARGS__)
#define ROS_INFO_STREAM(args) do { if (2 >= refshim::ClockState::get().log_level + 2) { std::ostringstream ss_; ss_ << args; std::fprintf(stderr, "[ref info] %s\n", ss_.str().c_str()); } } while (0)
#define ROS_ERROR_ONCE(...) do { static bool hit_ = false; if (!hit_) { hit_ = true; REFSHIM_LOG(4, "error", __VA_ARGS__); } } while (0)

namespace ros {

struct Duration {
  double d;
  explicit Duration(double s = 0.0) : d(s) {}
  double toSec() const { return d; }
  bool sleep() const { std::this_thread::sleep_for(std::chrono::duration<double>(d)); return true; }
};
struct WallDuration { double d; explicit WallDuration(double s = 0.0) : d(s) {} double toSec() const { return d; } };
struct WallTime {
  double t = 0.0;
  static WallTime now() { WallTime w; w.t = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); return w; }
  WallDuration operator-(const WallTime& o) const { return WallDuration(t - o.t); }
};
struct WallTimerEvent {};
struct WallTimer {};   // never fires in the shim: the harness drives the node message by message
inline void init(int&, char**, const std::string&) {}
inline bool ok() { return true; }
inline void spin() { if (refshim::SpinHook::get().player) refshim::SpinHook::get().player(); }
struct Time {
  uint32_t sec = 0, nsec = 0;
  Time() {}
  explicit Time(double t) { fromSec(t); }
  Time& fromSec(double t) { sec = (uint32_t)std::floor(t); nsec = (uint32_t)std::llround((t - sec) * 1e9); if (nsec >= 1000000000u) { sec++; nsec -= 1000000000u; } return *this; }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  static Time now() {
    refshim::ClockState& c = refshim::ClockState::get();
    if (c.frozen) return Time(c.now);
    return Time(std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count());
  }
};

class Publisher {
 public:
  Publisher() {}
  explicit Publisher(const std::string& topic) : topic_(topic) {}
  template <typename M> void publish(const M& m) const { refshim::Bus::get().put(topic_, m); }
  std::string getTopic() const { return topic_; }
  uint32_t getNumSubscribers() const { return 1; }   // someone always listens in the harness
 private:
  std::string topic_;
};

class Subscriber {};

class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string& ns) : ns_(ns) {}
  template <typename T> bool param(const std::string& name, T& var, const T& def) const {
    std::string text;
    if (!lookup(name, &text)) { var = def; return false; }
    parse(text, &var);
    return true;
  }
  template <typename M> Publisher advertise(const std::string& topic, uint32_t /*queue*/, bool /*latch*/ = false) const { return Publisher(topic); }
  template <typename M> Subscriber subscribe(const std::string& topic, uint32_t /*queue*/, void (*fp)(const std::shared_ptr<M const>&)) const {
    refshim::Subscribers& s = refshim::Subscribers::get();
    std::lock_guard<std::mutex> lk(s.mu);
    s.table[topic] = [fp](const std::shared_ptr<const void>& m) { fp(std::static_pointer_cast<const M>(m)); };
    return Subscriber();
  }
  WallTimer createWallTimer(const WallDuration&, void (*)(const WallTimerEvent&)) const { return WallTimer(); }
 private:
  static bool lookup(const std::string& name, std::string* text) {
    refshim::ParamTable& t = refshim::ParamTable::get();
    std::lock_guard<std::mutex> lk(t.mu);
    auto it = t.values.find(name);
    if (it == t.values.end()) return false;
    *text = it->second;
    return true;
  }
  static void parse(const std::string& s, double* v) { *v = std::stod(s); }
  static void parse(const std::string& s, int* v) { *v = std::stoi(s); }
  static void parse(const std::string& s, bool* v) { *v = (s == "1" || s == "true" || s == "True"); }
  static void parse(const std::string& s, std::string* v) { *v = s; }
  std::string ns_;
};

}  // namespace ros

namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
}

namespace sensor_msgs {
struct PointField {
  enum { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
  std::string name; uint32_t offset = 0; uint8_t datatype = 0; uint32_t count = 0;
};
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
};
typedef std::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
struct Quat { double x = 0, y = 0, z = 0, w = 1; };
struct Imu { std_msgs::Header header; Quat orientation; };
typedef std::shared_ptr<Imu const> ImuConstPtr;
}  // namespace sensor_msgs

namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {}; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; double covariance[36] = {}; };
struct TwistStamped { std_msgs::Header header; Twist twist; };
}  // namespace geometry_msgs

namespace nav_msgs {
struct Odometry { std_msgs::Header header; std::string child_frame_id; geometry_msgs::PoseWithCovariance pose; geometry_msgs::TwistWithCovariance twist; };
}

namespace tf {

typedef double tfScalar;
struct TransformException : public std::runtime_error { explicit TransformException(const std::string& w) : std::runtime_error(w) {} };

class Vector3 {
 public:
  Vector3() : v_{0, 0, 0} {}
  Vector3(double x, double y, double z) : v_{x, y, z} {}
  double x() const { return v_[0]; } double y() const { return v_[1]; } double z() const { return v_[2]; }
 private:
  double v_[3];
};

class Quaternion {
 public:
  Quaternion() : q_{0, 0, 0, 1} {}
  Quaternion(double x, double y, double z, double w) : q_{x, y, z, w} {}
  double x() const { return q_[0]; } double y() const { return q_[1]; } double z() const { return q_[2]; } double w() const { return q_[3]; }
  double length2() const { return q_[0] * q_[0] + q_[1] * q_[1] + q_[2] * q_[2] + q_[3] * q_[3]; }
 private:
  double q_[4];
};

// tf/LinearMath/Matrix3x3.h
class Matrix3x3 {
 public:
  Matrix3x3() : m_{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  explicit Matrix3x3(const Quaternion& q) { setRotation(q); }
  void setRotation(const Quaternion& q) {
    const double d = q.length2();
    const double s = 2.0 / d;
    const double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
    const double wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
    const double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
    const double yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
    m_[0] = 1.0 - (yy + zz); m_[1] = xy - wz; m_[2] = xz + wy;
    m_[3] = xy + wz; m_[4] = 1.0 - (xx + zz); m_[5] = yz - wx;
    m_[6] = xz - wy; m_[7] = yz + wx; m_[8] = 1.0 - (xx + yy);
  }
  // getEulerYPR, solution 1
  void getRPY(double& roll, double& pitch, double& yaw, unsigned int /*solution_number*/ = 1) const {
    const double kPi = 3.14159265358979323846;
    if (std::fabs(m_[6]) >= 1.0) {
      yaw = 0.0;
      if (m_[6] < 0.0) { const double delta = std::atan2(m_[1], m_[2]); pitch = kPi / 2.0; roll = delta; }
      else { const double delta = std::atan2(-m_[1], -m_[2]); pitch = -kPi / 2.0; roll = delta; }
    } else {
      pitch = -std::asin(m_[6]);
      roll = std::atan2(m_[7] / std::cos(pitch), m_[8] / std::cos(pitch));
      yaw = std::atan2(m_[3] / std::cos(pitch), m_[0] / std::cos(pitch));
    }
  }
  // setEulerYPR(yaw, pitch, roll)
  void setRPY(double roll, double pitch, double yaw) {
    const double ci = std::cos(roll), cj = std::cos(pitch), ch = std::cos(yaw);
    const double si = std::sin(roll), sj = std::sin(pitch), sh = std::sin(yaw);
    const double cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
    m_[0] = cj * ch; m_[1] = sj * sc - cs; m_[2] = sj * cc + ss;
    m_[3] = cj * sh; m_[4] = sj * ss + cc; m_[5] = sj * cs - sc;
    m_[6] = -sj; m_[7] = cj * si; m_[8] = cj * ci;
  }
  void getRotation(Quaternion& q) const {
    const double trace = m_[0] + m_[4] + m_[8];
    double t[4];
    if (trace > 0.0) {
      double s = std::sqrt(trace + 1.0);
      t[3] = s * 0.5;
      s = 0.5 / s;
      t[0] = (m_[7] - m_[5]) * s; t[1] = (m_[2] - m_[6]) * s; t[2] = (m_[3] - m_[1]) * s;
    } else {
      const int i = m_[0] < m_[4] ? (m_[4] < m_[8] ? 2 : 1) : (m_[0] < m_[8] ? 2 : 0);
      const int j = (i + 1) % 3, k = (i + 2) % 3;
      double s = std::sqrt(m_[i * 3 + i] - m_[j * 3 + j] - m_[k * 3 + k] + 1.0);
      t[i] = s * 0.5;
      s = 0.5 / s;
      t[3] = (m_[k * 3 + j] - m_[j * 3 + k]) * s;
      t[j] = (m_[j * 3 + i] + m_[i * 3 + j]) * s;
      t[k] = (m_[k * 3 + i] + m_[i * 3 + k]) * s;
    }
    q = Quaternion(t[0], t[1], t[2], t[3]);
  }
 private:
  double m_[9];   // row-major
};

class Transform {
 public:
  void setOrigin(const Vector3& o) { origin_ = o; }
  void setRotation(const Quaternion& q) { rot_ = q; }
  const Vector3& getOrigin() const { return origin_; }
  Quaternion getRotation() const { return rot_; }
 private:
  Vector3 origin_;
  Quaternion rot_;
};

class StampedTransform : public Transform {
 public:
  ros::Time stamp_;
  std::string frame_id_, child_frame_id_;
  StampedTransform() {}
  StampedTransform(const Transform& t, const ros::Time& stamp, const std::string& frame, const std::string& child)
      : Transform(t), stamp_(stamp), frame_id_(frame), child_frame_id_(child) {}
};

// static transforms registered by the driver: (target_frame, source_frame) -> transform
struct StaticTransforms {
  std::mutex mu;
  std::map<std::pair<std::string, std::string>, Transform> table;
  static StaticTransforms& get() { static StaticTransforms s; return s; }
};

class TransformListener {
 public:
  bool waitForTransform(const std::string& target, const std::string& source, const ros::Time&, const ros::Duration&) const {
    StaticTransforms& s = StaticTransforms::get();
    std::lock_guard<std::mutex> lk(s.mu);
    return s.table.count(std::make_pair(target, source)) != 0;
  }
  void lookupTransform(const std::string& target, const std::string& source, const ros::Time& t, StampedTransform& out) const {
    StaticTransforms& s = StaticTransforms::get();
    std::lock_guard<std::mutex> lk(s.mu);
    auto it = s.table.find(std::make_pair(target, source));
    if (it == s.table.end()) throw TransformException("\"" + target + "\" passed to lookupTransform argument target_frame does not exist.");
    out = StampedTransform(it->second, t, target, source);
  }
};

class TransformBroadcaster {
 public:
  void sendTransform(const StampedTransform& t) {
    {
      StaticTransforms& s = StaticTransforms::get();
      std::lock_guard<std::mutex> lk(s.mu);
      s.table[std::make_pair(t.frame_id_, t.child_frame_id_)] = t;
    }
    refshim::Bus::get().put(std::string("/tf"), t);
  }
};

}  // namespace tf
