// Drop-in harness — TEST INFRASTRUCTURE ONLY.
//
// Proves the boundary of SURVEY.md §8(b): the reference's OWN node sources, src/liodom_node.cc and
// src/liodom_mapping_node.cc, are compiled UNMODIFIED (only `main` is renamed on the command line) against THIS
// repo's facade headers (include/liodom/*.h, built with -DLIODOM_FACADE_USE_PCL) and linked with the facade
// (liodom_b200/host/facade.cc -> CUDA library).  ROS / PCL / Eigen / tf types come from the API shim in
// oracle/refshim/ (none of those libraries is installed here).  ros::spin() hands control to the player below,
// which delivers synthetic messages to the node's own subscriber callbacks (lidarClb) one by one.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "refshim/refshim_ros.h"
#include "refshim/refshim_pcl.h"
#include "refshim/refshim_eigen.h"

#ifdef DROPIN_MAPPING
int liodom_mapping_node_main(int argc, char** argv);   // src/liodom_mapping_node.cc:108-150
#else
int liodom_node_main(int argc, char** argv);           // src/liodom_node.cc:72-121
#endif

namespace {

void set_params(const char* kv) {   // "name=value;name=value"
  refshim::ParamTable& t = refshim::ParamTable::get();
  std::lock_guard<std::mutex> lk(t.mu);
  t.values.clear();
  std::string s(kv ? kv : "");
  size_t p0 = 0;
  while (p0 < s.size()) {
    size_t p1 = s.find(';', p0);
    if (p1 == std::string::npos) p1 = s.size();
    const std::string item = s.substr(p0, p1 - p0);
    const size_t eq = item.find('=');
    if (eq != std::string::npos) t.values[item.substr(0, eq)] = item.substr(eq + 1);
    p0 = p1 + 1;
  }
}

std::shared_ptr<sensor_msgs::PointCloud2> make_msg(const float* pts, int n, int stride_f, int width, int height, double stamp, uint32_t seq, const char* frame) {
  auto m = std::make_shared<sensor_msgs::PointCloud2>();
  m->header.seq = seq; m->header.stamp.fromSec(stamp); m->header.frame_id = frame;
  m->width = width > 0 ? (uint32_t)width : (uint32_t)n; m->height = height > 0 ? (uint32_t)height : 1u;
  const char* names[4] = {"x", "y", "z", "intensity"};
  for (int k = 0; k < 4; ++k) { sensor_msgs::PointField f; f.name = names[k]; f.offset = 4u * k; f.datatype = sensor_msgs::PointField::FLOAT32; f.count = 1; m->fields.push_back(f); }
  m->point_step = 16; m->row_step = 16 * m->width; m->is_dense = true; m->is_bigendian = false;
  m->data.resize((size_t)n * 16);
  for (int i = 0; i < n; ++i) std::memcpy(m->data.data() + (size_t)i * 16, pts + (size_t)i * stride_f, 16);
  return m;
}

template <typename M> long seq_of(const std::string& topic) { long s = 0; M m; return refshim::Bus::get().peek<M>(topic, &m, &s) ? s : 0; }

bool wait_for(const std::function<bool()>& done, double seconds) {
  const auto t0 = std::chrono::steady_clock::now();
  while (!done()) {
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) return false;
    std::this_thread::sleep_for(std::chrono::microseconds(200));
  }
  return true;
}

}  // namespace

extern "C" {

#ifndef DROPIN_MAPPING
// The reference's liodom_node main() over a sequence of scans.  odom_out[n x 13]: what publishOdom put into
// nav_msgs/Odometry (orientation x,y,z,w, position, twist linear, twist angular); nedges_out[n]: width of "edges".
// Returns the number of frames that produced an odometry message.
int dropin_node_run(const float* pts, const int32_t* npts, int nframes, int stride_f, int width, int height, double dt,
                    const char* params_kv, double* odom_out, int32_t* nedges_out) {
  set_params(params_kv);
  {
    tf::Transform ident;   // the launch files publish base_link -> velo_link statically (launch/liodom.launch)
    tf::StaticTransforms& s = tf::StaticTransforms::get();
    std::lock_guard<std::mutex> lk(s.mu);
    s.table.clear();
    s.table[std::make_pair(std::string("velo_link"), std::string("base_link"))] = ident;
  }
  int produced = 0;
  refshim::SpinHook::get().player = [&]() {
    size_t off = 0;
    for (int f = 0; f < nframes; ++f) {
      const long odom0 = seq_of<nav_msgs::Odometry>("odom"), edges0 = seq_of<sensor_msgs::PointCloud2>("edges");
      auto msg = make_msg(pts + off * stride_f, npts[f], stride_f, width, height, 1000.0 + dt * f, (uint32_t)f, "velo_link");
      off += (size_t)npts[f];
      if (!refshim::Subscribers::get().deliver<sensor_msgs::PointCloud2>("points", msg)) return;   // lidarClb
      if (!wait_for([&] { return seq_of<nav_msgs::Odometry>("odom") > odom0 && seq_of<sensor_msgs::PointCloud2>("edges") > edges0; }, 60.0)) return;
      nav_msgs::Odometry m; long s;
      refshim::Bus::get().peek("odom", &m, &s);
      const double v[13] = {m.pose.pose.orientation.x, m.pose.pose.orientation.y, m.pose.pose.orientation.z, m.pose.pose.orientation.w,
                            m.pose.pose.position.x, m.pose.pose.position.y, m.pose.pose.position.z,
                            m.twist.twist.linear.x, m.twist.twist.linear.y, m.twist.twist.linear.z,
                            m.twist.twist.angular.x, m.twist.twist.angular.y, m.twist.twist.angular.z};
      std::memcpy(odom_out + 13 * (size_t)f, v, sizeof(v));
      sensor_msgs::PointCloud2 em;
      refshim::Bus::get().peek("edges", &em, &s);
      if (nedges_out) nedges_out[f] = (int32_t)em.width;
      ++produced;
    }
  };
  char arg0[] = "liodom_node";
  char* argv[] = {arg0, nullptr};
  liodom_node_main(1, argv);
  refshim::SpinHook::get().player = nullptr;
  return produced;
}
#else
// The reference's liodom_mapping_node main() over a sequence of edge clouds + poses (the TF the odometry node would
// have broadcast).  local_sizes_out[n]: width of "map_local" after every message; map_out: the last "map" message
// (x,y,z,intensity), at most map_cap points; returns its size, or -1.
int dropin_mapping_run(const float* pts, const int32_t* npts, int nframes, const double* poses16, const char* params_kv,
                       int32_t* local_sizes_out, float* last_local_out, int local_cap, float* map_out, int map_cap) {
  set_params(params_kv);
  int last_map = -1;
  refshim::SpinHook::get().player = [&]() {
    size_t off = 0;
    for (int f = 0; f < nframes; ++f) {
      const double* T = poses16 + 16 * (size_t)f;
      Eigen::Matrix3d R;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = T[i * 4 + j];
      const Eigen::Quaterniond q(R);
      tf::Transform t;
      t.setOrigin(tf::Vector3(T[3], T[7], T[11]));
      t.setRotation(tf::Quaternion(q.x(), q.y(), q.z(), q.w()));
      {
        tf::StaticTransforms& s = tf::StaticTransforms::get();
        std::lock_guard<std::mutex> lk(s.mu);
        s.table[std::make_pair(std::string("world"), std::string("base_link"))] = t;
      }
      auto msg = make_msg(pts + off * 4, npts[f], 4, 0, 0, 1000.0 + 0.1 * f, (uint32_t)f, "base_link");
      off += (size_t)npts[f];
      if (!refshim::Subscribers::get().deliver<sensor_msgs::PointCloud2>("points", msg)) return;   // lidarClb (synchronous)
      sensor_msgs::PointCloud2 lm; long s;
      if (refshim::Bus::get().peek("map_local", &lm, &s)) {
        if (local_sizes_out) local_sizes_out[f] = (int32_t)lm.width;
        if (f == nframes - 1 && last_local_out && (int)lm.width <= local_cap)
          for (uint32_t i = 0; i < lm.width; ++i) { std::memcpy(last_local_out + 4 * i, lm.data.data() + (size_t)i * lm.point_step, 12); std::memcpy(last_local_out + 4 * i + 3, lm.data.data() + (size_t)i * lm.point_step + 16, 4); }
      }
      sensor_msgs::PointCloud2 mm;
      if (f == nframes - 1 && refshim::Bus::get().peek("map", &mm, &s)) {
        last_map = (int)mm.width;
        if (map_out && last_map <= map_cap)
          for (uint32_t i = 0; i < mm.width; ++i) { std::memcpy(map_out + 4 * i, mm.data.data() + (size_t)i * mm.point_step, 12); std::memcpy(map_out + 4 * i + 3, mm.data.data() + (size_t)i * mm.point_step + 16, 4); }
      }
    }
  };
  char arg0[] = "liodom_mapping";
  char* argv[] = {arg0, nullptr};
  liodom_mapping_node_main(1, argv);
  refshim::SpinHook::get().player = nullptr;
  return last_map;
}
#endif

void dropin_set_log_level(int level) { refshim::ClockState::get().log_level = level; }

}  // extern "C"
