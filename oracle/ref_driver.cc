// oracle/_ref driver — TEST INFRASTRUCTURE ONLY.
//
// C entry points around the reference's OWN translation units (/root/reference/src/{feature_extractor,
// laser_odometry,map,params,shared_data,stats}.cc compiled unmodified against oracle/refshim/), used by
// tests/ to pin the restated oracle (oracle/liodom_oracle.cc) against the reference's object code.
// The worker functors are driven synchronously: a publish hook on the shim's message bus clears the
// `running` flag, so FeatureExtractor::operator() / LaserOdometer::operator() process exactly one
// queued item in the caller's thread and return (same code path as src/liodom_node.cc:85-91, minus
// the threads).  Private helpers are reached with the `#define private public` include trick (class
// layout is unaffected).
// every standard header the reference's headers pull in is included BEFORE the access-specifier trick
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <queue>
#include <sstream>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "refshim/refshim_ros.h"
#include "refshim/refshim_pcl.h"
#include "refshim/refshim_eigen.h"
#include "refshim/ceres/ceres.h"
#include <omp.h>
#include <numeric>
#include <unordered_map>

#define private public
#define protected public
#include <liodom/feature_extractor.h>
#include <liodom/laser_odometry.h>
#include <liodom/map.h>
#undef private
#undef protected

using liodom::Point;
using liodom::PointCloud;

namespace {

PointCloud::Ptr cloud_from(const float* pts, int n, int stride_f, int width, int height) {
  PointCloud::Ptr pc(new PointCloud);
  pc->points.resize(n);
  for (int i = 0; i < n; ++i) {
    Point& p = pc->points[i];
    p.x = pts[(size_t)i * stride_f]; p.y = pts[(size_t)i * stride_f + 1]; p.z = pts[(size_t)i * stride_f + 2];
    p.intensity = pts[(size_t)i * stride_f + (stride_f >= 8 ? 4 : 3)];
  }
  if (width > 0 && height > 0) { pc->width = width; pc->height = height; }
  else { pc->width = n; pc->height = 1; }
  pc->is_dense = true;   // pcl::fromROSMsg copies the message's is_dense; the synthetic scans are dense
  return pc;
}

void cloud_to(const PointCloud& pc, float* out) {
  for (size_t i = 0; i < pc.points.size(); ++i) {
    out[4 * i] = pc.points[i].x; out[4 * i + 1] = pc.points[i].y; out[4 * i + 2] = pc.points[i].z; out[4 * i + 3] = pc.points[i].intensity;
  }
}

Eigen::Isometry3d iso_from(const double* T16) {
  Eigen::Isometry3d I;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) I.matrix()(i, j) = T16[i * 4 + j];
  return I;
}
void iso_to(const Eigen::Isometry3d& I, double* T16) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T16[i * 4 + j] = I.matrix()(i, j); }

// Run a worker functor until it has published on `stop_topic` (one queued item), then return.
template <typename Worker>
bool run_one(Worker& w, const std::string& stop_topic) {
  std::atomic<bool> running(true);
  bool fired = false;
  refshim::Bus::get().on_publish = [&](const std::string& topic) { if (topic == stop_topic) { fired = true; running = false; } };
  w(running);
  refshim::Bus::get().on_publish = nullptr;
  return fired;
}

struct RefOdom {
  liodom::LaserOdometer* lodom;
  double stamp = 1000.0;
};

}  // namespace

extern "C" {

// ---- parameters (ros::NodeHandle::param table -> Params::readParams, src/params.cc:37-110) ----------
void ref_set_param(const char* name, const char* value) {
  refshim::ParamTable& t = refshim::ParamTable::get();
  std::lock_guard<std::mutex> lk(t.mu);
  t.values[name] = value;
}
void ref_clear_params() {
  refshim::ParamTable& t = refshim::ParamTable::get();
  std::lock_guard<std::mutex> lk(t.mu);
  t.values.clear();
}
void ref_read_params() { liodom::Params::getInstance()->readParams(ros::NodeHandle("~")); }
// min_range, max_range, lidar_type, scan_lines, scan_regions, edges_per_region, min_points_per_scan, local_map_size,
// save_results, use_imu, filter_local_map, mapping, publish_tf
void ref_get_params(double* out13) {
  liodom::Params* p = liodom::Params::getInstance();
  const double v[13] = {p->min_range_, p->max_range_, (double)p->lidar_type_, (double)p->scan_lines_, (double)p->scan_regions_,
                        (double)p->edges_per_region_, (double)p->min_points_per_scan_, (double)p->local_map_size_, (double)p->save_results_,
                        (double)p->use_imu_, (double)p->filter_local_map_, (double)p->mapping_, (double)p->publish_tf_};
  std::memcpy(out13, v, sizeof(v));
}
void ref_set_log_level(int level) { refshim::ClockState::get().log_level = level; }
long ref_warning_count() { return refshim::ClockState::get().warnings; }
void ref_freeze_clock(int frozen, double now) { refshim::ClockState::get().frozen = frozen != 0; refshim::ClockState::get().now = now; }
void ref_set_static_tf(const char* target, const char* source, const double* xyz, const double* q_xyzw) {
  tf::Transform t;
  t.setOrigin(tf::Vector3(xyz[0], xyz[1], xyz[2]));
  t.setRotation(tf::Quaternion(q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]));
  tf::StaticTransforms& s = tf::StaticTransforms::get();
  std::lock_guard<std::mutex> lk(s.mu);
  s.table[std::make_pair(std::string(target), std::string(source))] = t;
}
void ref_clear_static_tf() { tf::StaticTransforms& s = tf::StaticTransforms::get(); std::lock_guard<std::mutex> lk(s.mu); s.table.clear(); }

// ---- FeatureExtractor (src/feature_extractor.cc) -----------------------------------------------------
void* ref_fext_create() { return new liodom::FeatureExtractor(ros::NodeHandle("~")); }
void ref_fext_destroy(void* h) { delete static_cast<liodom::FeatureExtractor*>(h); }

int ref_is_valid_point(void* h, double x, double y, double z, double* dist) {
  return static_cast<liodom::FeatureExtractor*>(h)->isValidPoint(x, y, z, dist) ? 1 : 0;
}

// splitPointCloud: rings_xyzi = the scan_lines_ clouds back to back, ring_offsets[scan_lines_+1]
int ref_split(void* h, const float* pts, int n, int stride_f, int width, int height, float* rings_xyzi, int32_t* ring_offsets) {
  liodom::FeatureExtractor* fe = static_cast<liodom::FeatureExtractor*>(h);
  std::vector<PointCloud::Ptr> scans;
  fe->splitPointCloud(cloud_from(pts, n, stride_f, width, height), scans);
  int off = 0;
  for (size_t r = 0; r < scans.size(); ++r) {
    ring_offsets[r] = off;
    cloud_to(*scans[r], rings_xyzi + 4 * (size_t)off);
    off += (int)scans[r]->points.size();
  }
  ring_offsets[scans.size()] = off;
  return off;
}

// extractFeatures on given rings
int ref_extract(void* h, const float* rings_xyzi, const int32_t* ring_offsets, float* edges_xyzi, int cap) {
  liodom::FeatureExtractor* fe = static_cast<liodom::FeatureExtractor*>(h);
  const int L = fe->params->scan_lines_;
  std::vector<PointCloud::Ptr> scans;
  for (int r = 0; r < L; ++r) scans.push_back(cloud_from(rings_xyzi + 4 * (size_t)ring_offsets[r], ring_offsets[r + 1] - ring_offsets[r], 4, 0, 0));
  PointCloud::Ptr edges(new PointCloud);
  fe->extractFeatures(scans, edges);
  const int E = (int)edges->points.size();
  if (E <= cap) cloud_to(*edges, edges_xyzi);
  return E;
}

// The worker functor itself: pushPointCloud -> FeatureExtractor::operator() -> popFeatures
// (src/liodom_node.cc:40-55, src/feature_extractor.cc:42-82).
int ref_fext_process(void* h, const float* pts, int n, int stride_f, int width, int height, float* edges_xyzi, int cap) {
  liodom::FeatureExtractor* fe = static_cast<liodom::FeatureExtractor*>(h);
  std_msgs::Header hd; hd.frame_id = "velo_link"; hd.stamp.fromSec(1000.0);
  liodom::SharedData::getInstance()->pushPointCloud(cloud_from(pts, n, stride_f, width, height), hd);
  if (!run_one(*fe, "edges")) return -1;
  PointCloud::Ptr feats; std_msgs::Header fh;
  if (!liodom::SharedData::getInstance()->popFeatures(feats, fh)) return -2;
  const int E = (int)feats->points.size();
  if (E <= cap) cloud_to(*feats, edges_xyzi);
  return E;
}

// ---- LocalMapManager (src/laser_odometry.cc:24-69) ------------------------------------------------------
void* ref_lmap_create(int max_frames) { return new liodom::LocalMapManager((size_t)max_frames); }
void ref_lmap_destroy(void* h) { delete static_cast<liodom::LocalMapManager*>(h); }
void ref_lmap_add(void* h, const float* xyzi, int n) { static_cast<liodom::LocalMapManager*>(h)->addPointCloud(cloud_from(xyzi, n, 4, 0, 0)); }
int ref_lmap_size(void* h) { PointCloud::Ptr m; static_cast<liodom::LocalMapManager*>(h)->getLocalMap(m); return (int)m->points.size(); }
int ref_lmap_frames(void* h) { PointCloud::Ptr m; return (int)static_cast<liodom::LocalMapManager*>(h)->getLocalMap(m); }
void ref_lmap_get(void* h, float* xyzi) { PointCloud::Ptr m; static_cast<liodom::LocalMapManager*>(h)->getLocalMap(m); cloud_to(*m, xyzi); }
void ref_lmap_set_max_frames(void* h, int n) { static_cast<liodom::LocalMapManager*>(h)->setMaxFrames((size_t)n); }

// ---- Point2LineFactor (include/liodom/factors.hpp:64-121) through the autodiff cost function ------------
// r[3]; Jq[3x4] w.r.t. (x,y,z,w); Jt[3x3]; Jlocal[3x6] = [Jq * d Plus / d delta | Jt]
void ref_factor(const double* c, const double* a, const double* b, double min_range, double max_range, const double* q, const double* t,
                double* r, double* Jq, double* Jt, double* Jlocal) {
  ceres::CostFunction* cf = liodom::Point2LineFactor::create(Eigen::Vector3d(c[0], c[1], c[2]), Eigen::Vector3d(a[0], a[1], a[2]),
                                                             Eigen::Vector3d(b[0], b[1], b[2]), min_range, max_range);
  const double* params[2] = {q, t};
  double jq[12], jt[9];
  double* jac[2] = {jq, jt};
  cf->Evaluate(params, r, jac);
  if (Jq) std::memcpy(Jq, jq, sizeof(jq));
  if (Jt) std::memcpy(Jt, jt, sizeof(jt));
  if (Jlocal) {
    ceres::EigenQuaternionParameterization par;
    double P[12];
    par.ComputeJacobian(q, P);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += jq[i * 4 + k] * P[k * 3 + j]; Jlocal[i * 6 + j] = s; }
      for (int j = 0; j < 3; ++j) Jlocal[i * 6 + 3 + j] = jt[i * 3 + j];
    }
  }
  delete cf;
}
// residual only (the double instantiation of the functor)
void ref_factor_residual(const double* c, const double* a, const double* b, double min_range, double max_range, const double* q, const double* t, double* r) {
  liodom::Point2LineFactor f(Eigen::Vector3d(c[0], c[1], c[2]), Eigen::Vector3d(a[0], a[1], a[2]), Eigen::Vector3d(b[0], b[1], b[2]), min_range, max_range);
  f(q, t, r);
}

// ---- LaserOdometer (src/laser_odometry.cc:71-446) ----------------------------------------------------------
void* ref_odom_create() {
  RefOdom* o = new RefOdom;
  o->lodom = new liodom::LaserOdometer(ros::NodeHandle("~"));
  return o;
}
void ref_odom_destroy(void* h) { RefOdom* o = static_cast<RefOdom*>(h); delete o->lodom; delete o; }

// One popFeatures() iteration of LaserOdometer::operator() (:100-272).  Returns 0, or -1 if the worker
// returned without publishing (the TF lookup of the first frame failed, :115-119).
int ref_odom_process(void* h, const float* edges_xyzi, int n, double stamp, double* pose16) {
  RefOdom* o = static_cast<RefOdom*>(h);
  std_msgs::Header hd; hd.frame_id = "velo_link"; hd.stamp.fromSec(stamp);
  liodom::SharedData::getInstance()->pushFeatures(cloud_from(edges_xyzi, n, 4, 0, 0), hd);
  const std::string last_topic = liodom::Params::getInstance()->publish_tf_ ? "/tf" : "twist";
  if (!run_one(*o->lodom, last_topic)) return -1;
  if (pose16) iso_to(o->lodom->odom_, pose16);
  return 0;
}
void ref_odom_get_pose(void* h, double* odom16, double* prev16) { RefOdom* o = static_cast<RefOdom*>(h); iso_to(o->lodom->odom_, odom16); iso_to(o->lodom->prev_odom_, prev16); }
void ref_odom_set_pose(void* h, const double* odom16, const double* prev16) {
  RefOdom* o = static_cast<RefOdom*>(h); o->lodom->odom_ = iso_from(odom16); o->lodom->prev_odom_ = iso_from(prev16);
}
int ref_odom_window_size(void* h) { return (int)static_cast<RefOdom*>(h)->lodom->lmap_manager.total_points_->points.size(); }
int ref_odom_window_frames(void* h) { return (int)static_cast<RefOdom*>(h)->lodom->lmap_manager.nframes_; }
void ref_odom_get_window(void* h, float* xyzi) { cloud_to(*static_cast<RefOdom*>(h)->lodom->lmap_manager.total_points_, xyzi); }
void ref_odom_set_window(void* h, const float* xyzi, const int32_t* frame_sizes, int nframes) {
  liodom::LocalMapManager& m = static_cast<RefOdom*>(h)->lodom->lmap_manager;
  int total = 0;
  std::queue<size_t> sizes;
  for (int k = 0; k < nframes; ++k) { sizes.push((size_t)frame_sizes[k]); total += frame_sizes[k]; }
  m.total_points_ = cloud_from(xyzi, total, 4, 0, 0);
  m.sizes_ = sizes;
  m.nframes_ = (size_t)nframes;
}
// the last messages publishOdom produced (:395-446): orientation x,y,z,w, position, twist linear, twist angular
int ref_odom_last_msg(double* out13) {
  nav_msgs::Odometry m; long seq = 0;
  if (!refshim::Bus::get().peek("odom", &m, &seq)) return -1;
  const double v[13] = {m.pose.pose.orientation.x, m.pose.pose.orientation.y, m.pose.pose.orientation.z, m.pose.pose.orientation.w,
                        m.pose.pose.position.x, m.pose.pose.position.y, m.pose.pose.position.z,
                        m.twist.twist.linear.x, m.twist.twist.linear.y, m.twist.twist.linear.z,
                        m.twist.twist.angular.x, m.twist.twist.angular.y, m.twist.twist.angular.z};
  std::memcpy(out13, v, sizeof(v));
  return (int)seq;
}
// SharedData::setLocalMap / setLastIMUOri (src/shared_data.cc:91-117), what mapClb / imuClb do (src/liodom_node.cc:57-70)
void ref_set_received_map(const float* xyzi, int n) { liodom::SharedData::getInstance()->setLocalMap(cloud_from(xyzi, n, 4, 0, 0)); }
void ref_set_imu(const double* q_xyzw) { Eigen::Quaterniond q(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]); liodom::SharedData::getInstance()->setLastIMUOri(q); }

// Whole node pipeline on a sequence: lidarClb -> FeatureExtractor worker -> LaserOdometer worker, frame by frame.
// poses_out[nframes*16]; nedges_out[nframes].  Returns frames processed.
int ref_run_sequence(const float* pts, const int32_t* npts, int nframes, int stride_f, int width, int height, double dt,
                     double* poses_out, int32_t* nedges_out) {
  liodom::FeatureExtractor fext{ros::NodeHandle("~")};
  liodom::LaserOdometer lodom{ros::NodeHandle("~")};
  liodom::SharedData* sdata = liodom::SharedData::getInstance();
  const std::string last_topic = liodom::Params::getInstance()->publish_tf_ ? "/tf" : "twist";
  size_t off = 0;
  for (int f = 0; f < nframes; ++f) {
    std_msgs::Header hd; hd.frame_id = "velo_link"; hd.stamp.fromSec(1000.0 + dt * f); hd.seq = (uint32_t)f;
    sdata->pushPointCloud(cloud_from(pts + off * stride_f, npts[f], stride_f, width, height), hd);
    off += (size_t)npts[f];
    if (!run_one(fext, "edges")) return f;
    sensor_msgs::PointCloud2 em; long seq;
    refshim::Bus::get().peek("edges", &em, &seq);
    if (nedges_out) nedges_out[f] = (int32_t)em.width;
    if (!run_one(lodom, last_topic)) return f;
    iso_to(lodom.odom_, poses_out + (size_t)f * 16);
  }
  return nframes;
}

// ---- Map (src/map.cc) -----------------------------------------------------------------------------------
void* ref_map_create(double xy, double z, double res) { return new liodom::Map(xy, z, res); }
void ref_map_destroy(void* h) { delete static_cast<liodom::Map*>(h); }
void ref_map_update(void* h, const float* xyzi, int n, const double* T16) { static_cast<liodom::Map*>(h)->updateMap(cloud_from(xyzi, n, 4, 0, 0), iso_from(T16)); }
int ref_map_size(void* h) { return (int)static_cast<liodom::Map*>(h)->getMap()->points.size(); }
void ref_map_get(void* h, float* xyzi) { cloud_to(*static_cast<liodom::Map*>(h)->getMap(), xyzi); }
int ref_map_num_cells(void* h) { return (int)static_cast<liodom::Map*>(h)->cells_vector_.size(); }
// keys3[ncells*3], counts[ncells] in creation order (cells_vector_)
void ref_map_cells(void* h, int32_t* keys3, int32_t* counts) {
  liodom::Map* m = static_cast<liodom::Map*>(h);
  std::map<liodom::Cell*, liodom::HashKey> inv;
  for (auto& kv : m->cells_) inv[kv.second] = kv.first;
  for (size_t i = 0; i < m->cells_vector_.size(); ++i) {
    const liodom::HashKey& k = inv[m->cells_vector_[i]];
    keys3[3 * i] = k.x; keys3[3 * i + 1] = k.y; keys3[3 * i + 2] = k.z;
    counts[i] = (int32_t)m->cells_vector_[i]->getPoints()->points.size();
  }
}
int ref_map_get_local(void* h, const double* T16, int cells_xy, int cells_z, float* xyzi, int cap) {
  PointCloud::Ptr pc = static_cast<liodom::Map*>(h)->getLocalMap(iso_from(T16), cells_xy, cells_z);
  const int n = (int)pc->points.size();
  if (xyzi && n <= cap) cloud_to(*pc, xyzi);
  return n;
}
double ref_map_entropy(void* h) { return static_cast<liodom::Map*>(h)->getMapEntropy(); }

// ---- Stats (src/stats.cc) -----------------------------------------------------------------------------------
void ref_stats_add_pose(const double* T16) { Eigen::Matrix4d M; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M(i, j) = T16[i * 4 + j]; liodom::Stats::getInstance()->addPose(M); }
void ref_stats_add_nfeats(long n) { liodom::Stats::getInstance()->addNumOfFeats((size_t)n); }
void ref_stats_add_times_ms(double fext_ms, double lodom_ms) {
  liodom::Stats* s = liodom::Stats::getInstance();
  const liodom::Clock::time_point t0 = liodom::Clock::now();
  s->addFeatureExtractionTime(t0, t0 + std::chrono::microseconds((long)(fext_ms * 1000)));
  s->addLaserOdometryTime(t0, t0 + std::chrono::microseconds((long)(lodom_ms * 1000)));
  s->startFrame(t0); s->stopFrame(t0 + std::chrono::microseconds((long)((fext_ms + lodom_ms) * 1000)));
}
void ref_stats_clear() {
  liodom::Stats* s = liodom::Stats::getInstance();
  s->poses_.clear(); s->feat_extr_.clear(); s->laser_odom_.clear(); s->num_of_features_.clear(); s->frame_times_.clear();
  while (!s->start_times_.empty()) s->start_times_.pop();
}
void ref_stats_write(const char* dir) { liodom::Stats::getInstance()->writeResults(dir); }

}  // extern "C"
